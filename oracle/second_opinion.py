"""A SECOND, independent restatement of resolv's narrowphase and LineTest — in plain Python floats.

TEST INFRASTRUCTURE ONLY (same rule as the rest of oracle/): nothing under resolv_b200/ may import this.

Why it exists: the reference (resolv v0.8.0, Go) cannot run in this image and ships no golden vectors, so the C++ oracle
(oracle/resolv_oracle.cpp) is "parity unpinned". This file is a separate line-by-line transcription of the same Go
sources, made without looking at the C++ oracle, in a different language with different data structures (tuples and
lists, no SoA, no scratch reuse). Python floats are IEEE-754 doubles, `*` and `+` are never fused, `math.sqrt` is
correctly rounded — the same arithmetic Go/amd64 does. tests/test_oracle_second_opinion.py demands that the two
restatements agree BIT FOR BIT on seeded random shapes; a transcription slip in either one shows up as a mismatch.
It narrows "unpinned" to "two independent readings of the source agree"; only parity/go run against real resolv pins it.

Go stdlib behaviour restated here: math.Min (NaN / signed zero), sort.Slice (pdqsort_func, full),
math.Pow(x, 2) == x*x for finite non-overflowing x.
"""
import math

MAXF = 1.7976931348623157e308  # math.MaxFloat64


# ---------------------------------------------------------------- vector.go
def v_sub(a, b):                      # vector.go:72-77
    return (a[0] - b[0], a[1] - b[1])


def v_add(a, b):                      # vector.go:54-58
    return (a[0] + b[0], a[1] + b[1])


def v_scale(a, s):                    # vector.go:272-276
    return (a[0] * s, a[1] * s)


def v_invert(a):                      # vector.go:104-108
    return (-a[0], -a[1])


def v_mag(a):                         # vector.go:111-113
    return math.sqrt(a[0] * a[0] + a[1] * a[1])


def v_dot(a, b):                      # vector.go:286-288
    return a[0] * b[0] + a[1] * b[1]


def v_dist2(a, b):                    # vector.go:152-156
    x = a[0] - b[0]
    y = a[1] - b[1]
    return x * x + y * y


def v_unit(a):                        # vector.go:167-175
    l = v_mag(a)
    if l < 1e-8 or l == 1:
        return a
    return (a[0] / l, a[1] / l)


def go_min(x, y):                     # Go math.Min: NaN wins, Min(-0, +0) = -0
    if math.isinf(x) and x < 0 or math.isinf(y) and y < 0:
        return -math.inf
    if x != x or y != y:
        return math.nan
    if x == 0 and x == y:
        return x if math.copysign(1.0, x) < 0 else y
    return x if x < y else y


def go_max(x, y):                     # Go math.Max
    if math.isinf(x) and x > 0 or math.isinf(y) and y > 0:
        return math.inf
    if x != x or y != y:
        return math.nan
    if x == 0 and x == y:
        return y if math.copysign(1.0, x) < 0 else x
    return x if x > y else y


def go_sort_slice(data, less):
    """sort.Slice(data, less): pdqsort_func of Go >= 1.19 (sort/zsortfunc.go; resolv's go.mod says go 1.20), restated from
    the published algorithm, separately from the C++ oracle's restatement: insertion sort up to 12 items, else
    pattern-defeating quicksort (choosePivot with ninther from 50 items, partialInsertionSort, partitionEqual for runs of
    duplicates, breakPatterns with an xorshift seeded by the length, heapsort when the limit runs out). In place."""
    n = len(data)

    def swap(i, j):
        data[i], data[j] = data[j], data[i]

    def lt(i, j):
        return less(data[i], data[j])

    def insertion(a, b):
        for i in range(a + 1, b):
            j = i
            while j > a and lt(j, j - 1):
                swap(j, j - 1)
                j -= 1

    def sift_down(lo, hi, first):
        root = lo
        while True:
            child = 2 * root + 1
            if child >= hi:
                return
            if child + 1 < hi and lt(first + child, first + child + 1):
                child += 1
            if not lt(first + root, first + child):
                return
            swap(first + root, first + child)
            root = child

    def heap_sort(a, b):
        first, lo, hi = a, 0, b - a
        for i in range((hi - 1) // 2, -1, -1):
            sift_down(i, hi, first)
        for i in range(hi - 1, -1, -1):
            swap(first, first + i)
            sift_down(lo, i, first)

    def order2(a, b, sw):
        if lt(b, a):
            sw[0] += 1
            return b, a
        return a, b

    def median(a, b, c, sw):
        a, b = order2(a, b, sw)
        b, c = order2(b, c, sw)
        a, b = order2(a, b, sw)
        return b

    def choose_pivot(a, b):
        l = b - a
        sw = [0]
        i, j, k = a + l // 4 * 1, a + l // 4 * 2, a + l // 4 * 3
        if l >= 8:
            if l >= 50:
                i = median(i - 1, i, i + 1, sw)
                j = median(j - 1, j, j + 1, sw)
                k = median(k - 1, k, k + 1, sw)
            j = median(i, j, k, sw)
        if sw[0] == 0:
            return j, 1          # increasingHint
        if sw[0] == 12:
            return j, 2          # decreasingHint
        return j, 0

    def reverse_range(a, b):
        i, j = a, b - 1
        while i < j:
            swap(i, j)
            i += 1
            j -= 1

    def partial_insertion(a, b):
        i = a + 1
        for _ in range(5):
            while i < b and not lt(i, i - 1):
                i += 1
            if i == b:
                return True
            if b - a < 50:
                return False
            swap(i, i - 1)
            if i - a >= 2:
                j = i - 1
                while j >= 1:
                    if not lt(j, j - 1):
                        break
                    swap(j, j - 1)
                    j -= 1
            if b - i >= 2:
                j = i + 1
                while j < b:
                    if not lt(j, j - 1):
                        break
                    swap(j, j - 1)
                    j += 1
        return False

    def break_patterns(a, b):
        length = b - a
        if length >= 8:
            rnd = length
            modulus = 1 << length.bit_length()
            idx = a + (length // 4) * 2 - 1
            for i in range(3):
                rnd ^= (rnd << 13) & 0xFFFFFFFFFFFFFFFF
                rnd ^= rnd >> 7
                rnd ^= (rnd << 17) & 0xFFFFFFFFFFFFFFFF
                other = rnd & (modulus - 1)
                if other >= length:
                    other -= length
                swap(idx - 1 + i, a + other)

    def partition_equal(a, b, pivot):
        swap(a, pivot)
        i, j = a + 1, b - 1
        while True:
            while i <= j and not lt(a, i):
                i += 1
            while i <= j and lt(a, j):
                j -= 1
            if i > j:
                break
            swap(i, j)
            i += 1
            j -= 1
        return i

    def partition(a, b, pivot):
        swap(a, pivot)
        i, j = a + 1, b - 1
        while i <= j and lt(i, a):
            i += 1
        while i <= j and not lt(j, a):
            j -= 1
        if i > j:
            swap(j, a)
            return j, True
        swap(i, j)
        i += 1
        j -= 1
        while True:
            while i <= j and lt(i, a):
                i += 1
            while i <= j and not lt(j, a):
                j -= 1
            if i > j:
                break
            swap(i, j)
            i += 1
            j -= 1
        swap(j, a)
        return j, False

    def pdq(a, b, limit):
        was_balanced = was_partitioned = True
        while True:
            length = b - a
            if length <= 12:
                insertion(a, b)
                return
            if limit == 0:
                heap_sort(a, b)
                return
            if not was_balanced:
                break_patterns(a, b)
                limit -= 1
            pivot, hint = choose_pivot(a, b)
            if hint == 2:
                reverse_range(a, b)
                pivot = (b - 1) - (pivot - a)
                hint = 1
            if was_balanced and was_partitioned and hint == 1:
                if partial_insertion(a, b):
                    return
            if a > 0 and not lt(a - 1, pivot):
                a = partition_equal(a, b, pivot)
                continue
            mid, already = partition(a, b, pivot)
            was_partitioned = already
            left, right = mid - a, b - mid
            thr = length // 8
            if left < right:
                was_balanced = left >= thr
                pdq(a, mid, limit)
                a = mid + 1
            else:
                was_balanced = right >= thr
                pdq(mid + 1, b, limit)
                b = mid

    pdq(0, n, n.bit_length())


def go_sort_small(items, less):       # kept name: every sort.Slice call of the reference goes through here
    go_sort_slice(items, less)


# ---------------------------------------------------------------- utils.go: Bounds
def b_move(b, x, y):                  # utils.go:137-143
    return ((b[0][0] + x, b[0][1] + y), (b[1][0] + x, b[1][1] + y))


def b_center(b):                      # utils.go:100-102
    return v_add(b[0], v_scale(v_sub(b[1], b[0]), 0.5))


def b_is_intersecting(b, o):          # utils.go:150-175 (IsEmpty subtracts Min.X from Max.Y: kept)
    if o[1][0] < b[0][0] or o[0][0] > b[1][0] or o[1][1] < b[0][1] or o[0][1] > b[1][1]:
        mn, mx = (0.0, 0.0), (0.0, 0.0)
    else:
        mx = (go_min(o[1][0], b[1][0]), go_min(o[1][1], b[1][1]))
        mn = (go_max(o[0][0], b[0][0]), go_max(o[0][1], b[0][1]))
    empty = (mx[0] - mn[0] == 0) and (mx[1] - mn[0] == 0)
    return not empty


def overlap(pa, pb):                  # utils.go:75-77
    return go_min(pa[1] - pb[0], pb[1] - pa[0])


# ---------------------------------------------------------------- shapes
class Circle:                         # circle.go:4-17
    def __init__(self, x, y, r):
        self.pos = (float(x), float(y))
        self.r = float(r)

    def bounds(self):                 # circle.go:32-38
        return ((self.pos[0] - self.r, self.pos[1] - self.r), (self.pos[0] + self.r, self.pos[1] + self.r))

    def project(self, axis):          # circle.go:40-53
        axis = v_unit(axis)
        pc = v_dot(axis, self.pos)
        pr = v_mag(axis) * self.r
        lo = pc - pr
        hi = pc + pr
        if lo > hi:
            lo, hi = hi, lo
        return (lo, hi)


class Polygon:                        # convexPolygon.go:12-46, scale 1, rotation 0
    def __init__(self, x, y, pts, closed=True):
        self.pos = (float(x), float(y))
        self.points = [(float(pts[i]), float(pts[i + 1])) for i in range(0, len(pts), 2)]
        self.closed = closed
        self.update_bounds()

    def transformed(self):            # convexPolygon.go:144-154 (scale (1,1): x*1 is exact; rotation 0 skipped)
        return [(p[0] * 1.0 + self.pos[0], p[1] * 1.0 + self.pos[1]) for p in self.points]

    def update_bounds(self):          # convexPolygon.go:163-197
        t = self.transformed()
        tlx, tly = t[0]
        brx, bry = t[0]
        for p in t:
            if p[0] < tlx:
                tlx = p[0]
            elif p[0] > brx:
                brx = p[0]
            if p[1] < tly:
                tly = p[1]
            elif p[1] > bry:
                bry = p[1]
        self.local = b_move(((tlx, tly), (brx, bry)), -self.pos[0], -self.pos[1])

    def bounds(self):                 # convexPolygon.go:158-161
        return b_move(self.local, self.pos[0], self.pos[1])

    def center(self):                 # convexPolygon.go:200-214
        return b_center(self.bounds())

    def lines(self):                  # convexPolygon.go:117-141
        v = self.transformed()
        out = []
        for i in range(len(v)):
            start, end = v[i], v[0]
            if i < len(v) - 1:
                end = v[i + 1]
            elif (not self.closed) or len(self.points) <= 2:
                break
            out.append((start, end))
        return out

    def project(self, axis):          # convexPolygon.go:217-231
        axis = v_unit(axis)
        v = self.transformed()
        lo = v_dot(axis, v[0])
        hi = lo
        for i in range(1, len(v)):
            p = v_dot(axis, v[i])
            if p < lo:
                lo = p
            elif p > hi:
                hi = p
        return (lo, hi)

    def sat_axes(self):               # convexPolygon.go:234-242
        return [line_normal(l) for l in self.lines()]


def rectangle(x, y, w, h):            # convexPolygon.go:616-632
    hw = w / 2
    hh = h / 2
    return Polygon(x, y, [-hw, -hh, hw, -hh, hw, hh, -hw, hh])


def line_shape(x1, y1, x2, y2):       # convexPolygon.go:671-684
    cx = x1 + ((x2 - x1) / 2)
    cy = y1 + ((y2 - y1) / 2)
    return Polygon(cx, cy, [cx - x1, cy - y1, cx - x2, cy - y2])


# ---------------------------------------------------------------- collidingLine
def line_vector(l):                   # convexPolygon.go:709-711
    return v_unit(v_sub(l[1], l[0]))


def line_normal(l):                   # convexPolygon.go:704-707
    v = line_vector(l)
    return v_unit((v[1], -v[0]))


def line_x_line(l, o):                # convexPolygon.go:714-744
    (sx, sy), (ex, ey) = l
    (osx, osy), (oex, oey) = o
    det = (ex - sx) * (oey - osy) - (oex - osx) * (ey - sy)
    if det != 0:
        lam = (((sy - osy) * (oex - osx)) - ((sx - osx) * (oey - osy)) + 1) / det
        gam = (((sy - osy) * (ex - sx)) - ((sx - osx) * (ey - sy)) + 1) / det
        if (0 < lam and lam < 1) and (0 < gam and gam < 1):
            dx = ex - sx
            dy = ey - sy
            return (sx + (lam * dx), sy + (lam * dy))
    return None


def line_x_circle(l, c):              # convexPolygon.go:747-789
    pts = []
    ls = v_sub(l[0], c.pos)
    le = v_sub(l[1], c.pos)
    d = v_sub(le, ls)
    a = d[0] * d[0] + d[1] * d[1]
    b = 2 * ((d[0] * ls[0]) + (d[1] * ls[1]))
    cc = (ls[0] * ls[0]) + (ls[1] * ls[1]) - (c.r * c.r)
    det = b * b - (4 * a * cc)
    if det < 0:
        pass
    elif det == 0:
        t = _div(-b, 2 * a)
        if t >= 0 and t <= 1:
            pts.append((l[0][0] + t * d[0], l[0][1] + t * d[1]))
    else:
        t = _div(-b + math.sqrt(det), 2 * a)
        if t >= 0 and t <= 1:
            pts.append((l[0][0] + t * d[0], l[0][1] + t * d[1]))
        t = _div(-b - math.sqrt(det), 2 * a)
        if t >= 0 and t <= 1:
            pts.append((l[0][0] + t * d[0], l[0][1] + t * d[1]))
    return pts


def _div(x, y):                       # Go float division: x/0 is ±Inf or NaN, never a panic
    if y == 0:
        if x == 0 or x != x:
            return math.nan
        return math.copysign(math.inf, x) * math.copysign(1.0, y)
    return x / y


# ---------------------------------------------------------------- calculateMTV (convexPolygon.go:418-514)
def calculate_mtv(cp, other):
    smallest = (MAXF, 0.0)
    if isinstance(other, Polygon):
        for axes in (cp.sat_axes(), other.sat_axes()):
            for axis in axes:
                ov = overlap(cp.project(axis), other.project(axis))
                if ov <= 0:
                    return None
                if v_mag(smallest) > ov:
                    smallest = v_scale(axis, ov)
        if v_dot(v_sub(cp.center(), other.center()), smallest) < 0:
            smallest = v_invert(smallest)
    else:
        verts = list(cp.transformed())
        center = other.pos
        go_sort_small(verts, lambda p, q: v_mag(v_sub(p, center)) < v_mag(v_sub(q, center)))
        axis = (center[0] - verts[0][0], center[1] - verts[0][1])
        ov = overlap(cp.project(axis), other.project(axis))
        if ov <= 0:
            return None
        smallest = v_scale(v_unit(axis), ov)
        for axis in cp.sat_axes():
            ov = overlap(cp.project(axis), other.project(axis))
            if ov <= 0:
                return None
            if v_mag(smallest) > ov:
                smallest = v_scale(axis, ov)
        if v_dot(v_sub(cp.center(), other.pos), smallest) < 0:
            smallest = v_invert(smallest)
    delta = smallest
    if v_dot(v_sub(other.pos, cp.pos), delta) > 0:
        delta = v_invert(delta)
    return delta


# ---------------------------------------------------------------- the four pair tests (shape.go:318-479)
def intersection(a, b):
    """a.Intersection(b) -> (mtv, [(point, normal), ...], sorted: bool). `sorted` is False when the contact order could
    not be restated (more than 12 contacts: pdqsort) — compare those as multisets."""
    mtv = (0.0, 0.0)
    cons = []
    ordered = True
    if isinstance(a, Circle) and isinstance(b, Circle):          # shape.go:399-436
        if not b_is_intersecting(a.bounds(), b.bounds()):
            return mtv, cons, ordered
        dx = b.pos[0] - a.pos[0]
        dy = b.pos[1] - a.pos[1]
        d = math.sqrt(dx * dx + dy * dy)
        if d > a.r + b.r or d < abs(a.r - b.r) or d == 0:
            return mtv, cons, ordered
        aa = (a.r * a.r - b.r * b.r + d * d) / (2 * d)
        h = math.sqrt(a.r * a.r - aa * aa) if a.r * a.r - aa * aa >= 0 else math.nan
        x2 = a.pos[0] + aa * (b.pos[0] - a.pos[0]) / d
        y2 = a.pos[1] + aa * (b.pos[1] - a.pos[1]) / d
        pts = [(x2 + h * (b.pos[1] - a.pos[1]) / d, y2 - h * (b.pos[0] - a.pos[0]) / d),
               (x2 - h * (b.pos[1] - a.pos[1]) / d, y2 + h * (b.pos[0] - a.pos[0]) / d)]
        cons = [(p, v_unit(v_sub(p, a.pos))) for p in pts]
        m = (a.pos[0] - b.pos[0], a.pos[1] - b.pos[1])
        dist = v_mag(m)
        mtv = v_scale(v_unit(m), a.r + b.r - dist)
        return mtv, cons, ordered
    if isinstance(a, Circle):                                    # shape.go:318-354 circle -> convex
        if not b_is_intersecting(b.bounds(), a.bounds()):
            return mtv, cons, ordered
        for l in b.lines():
            for p in line_x_circle(l, a):
                cons.append((p, line_normal(l)))
        if cons:
            m = calculate_mtv(b, a)
            if m is not None:
                mtv = v_invert(m)
        return mtv, cons, ordered
    if isinstance(b, Circle):                                    # shape.go:356-397 convex -> circle
        if not b_is_intersecting(a.bounds(), b.bounds()):
            return mtv, cons, ordered
        for l in a.lines():
            for p in line_x_circle(l, b):
                cons.append((p, v_unit(v_sub(p, b.pos))))
        if cons:
            go_sort_small(cons, lambda s, t: v_dist2(s[0], b.pos) < v_dist2(t[0], b.pos))
            m = calculate_mtv(a, b)
            if m is not None:
                mtv = m
        return mtv, cons, ordered
    if not b_is_intersecting(a.bounds(), b.bounds()):            # shape.go:438-479 convex -> convex
        return mtv, cons, ordered
    for ol in b.lines():
        for l in a.lines():
            p = line_x_line(l, ol)
            if p is not None:
                cons.append((p, line_normal(ol)))
    if cons:
        center = a.center()
        go_sort_small(cons, lambda s, t: v_dist2(s[0], center) < v_dist2(t[0], center))
        m = calculate_mtv(a, b)
        if m is not None:
            mtv = m
    return mtv, cons, ordered


# ---------------------------------------------------------------- LineTest against ONE shape (utils.go:276-370)
def linetest_one(start, end, shape):
    """The IntersectionSet LineTest builds for `shape`, or None: (mtv, center, [(point, normal)...] sorted, sort key)."""
    vu = v_unit(v_sub(end, start))
    st = v_sub(start, v_scale(vu, 0.01))
    line = (st, end)
    cons = []
    if isinstance(shape, Circle):
        for p in line_x_circle(line, shape):
            cons.append((p, v_unit(v_sub(p, shape.pos))))
    else:
        for ol in shape.lines():
            p = line_x_line(line, ol)
            if p is not None:
                cons.append((p, line_normal(ol)))
    if not cons:
        return None
    c = (0.0, 0.0)
    for p, _ in cons:
        c = v_add(c, p)
    c = (c[0] / float(len(cons)), c[1] / float(len(cons)))
    go_sort_small(cons, lambda s, t: v_dist2(s[0], start) < v_dist2(t[0], start))
    mtv = v_sub(v_sub(cons[0][0], start), v_scale(vu, 0.01))
    key = v_dist2(cons[0][0], st)
    return mtv, c, cons, key


# ---------------------------------------------------------------- broadphase: Space, Cell, CellSelection, IntersectionTest
class Space:
    """space.go:7-46 with cell.go's per-cell lists kept as a dict of Python lists (append / swap-remove), so the
    candidate ORDER after motion — which depends on the register/unregister history — is reproduced too."""

    def __init__(self, space_w, space_h, cell_w, cell_h):          # space.go:16-30
        self.cw, self.ch = cell_w, cell_h
        self.w = int(math.ceil(float(space_w) / float(cell_w)))
        self.h = int(math.ceil(float(space_h) / float(cell_h)))
        self.cells = {}                                             # (cx, cy) -> [shape, ...]; only in-grid keys
        self.shapes = []

    def to_cell_space(self, b):                                     # utils.go:85-95
        return (int(math.floor(b[0][0] / float(self.cw))), int(math.floor(b[0][1] / float(self.ch))),
                int(math.floor(b[1][0] / float(self.cw))), int(math.floor(b[1][1] / float(self.ch))))

    def cell(self, cx, cy):                                         # space.go:134-142 (nil outside the grid)
        if 0 <= cy < self.h and 0 <= cx < self.w:
            return self.cells.setdefault((cx, cy), [])
        return None

    def add(self, shape, tags=0):                                   # space.go:33-46
        shape.tags = tags
        shape.touching = []
        shape.id = len(self.shapes)
        self.update(shape)
        self.shapes.append(shape)
        return shape.id

    def update(self, shape):                                        # shape.go:183-214, 239-242
        for c in shape.touching:                                    # cell.go:26-38: swap-remove
            for i, o in enumerate(c):
                if o is shape:
                    c[i] = c[-1]
                    c.pop()
                    break
        shape.touching = []
        cx, cy, ex, ey = self.to_cell_space(shape.bounds())
        for y in range(cy, ey + 1):
            for x in range(cx, ex + 1):
                c = self.cell(x, y)
                if c is not None:
                    if not any(o is shape for o in c):              # cell.go:19-23
                        c.append(shape)
                    shape.touching.append(c)

    def set_position(self, shape, x, y):                            # shape.go:115-119
        shape.pos = (float(x), float(y))
        self.update(shape)

    def candidates(self, shape, margin, use_any=False, any_mask=0, none_mask=0):
        """SelectTouchingCells(margin).FilterShapes()[.ByTags(any)][.NotByTags(none)] walked as IntersectionTest walks
        it (shape.go:220-237, cell.go:86-121, shapefilter.go:26-61, shape.go:277-288): [(other, distance²)] in
        generation order."""
        cx, cy, ex, ey = self.to_cell_space(shape.bounds())
        cx -= margin
        cy -= margin
        ex += margin
        ey += margin
        seen = []
        out = []
        for y in range(cy, ey + 1):
            for x in range(cx, ex + 1):
                c = self.cell(x, y)
                if c is None:
                    continue
                for s in c:
                    if s is shape:
                        continue
                    if s.id in seen:
                        continue
                    seen.append(s.id)                               # filtered-out shapes enter the set as well
                    if use_any and not ((s.tags & any_mask) > 0):   # tags.go:29-31
                        continue
                    if (s.tags & none_mask) > 0:
                        continue
                    out.append((s, v_dist2(s.pos, shape.pos)))      # shape.go:179-181: other.DistanceSquaredTo(self)
        return out

    def intersection_test(self, shape, margin, **filters):
        """shape.go:273-316 with OnIntersect returning true: the non-empty sets in callback order, or in an order
        flagged as unspecified when the candidate sort is pdqsort territory (> 12 candidates) AND has exact ties."""
        cand = self.candidates(shape, margin, **filters)
        defined = True
        go_sort_small(cand, lambda p, q: p[1] < q[1])
        sets = []
        for other, d in cand:
            mtv, cons, ordered = intersection(shape, other)
            if cons:
                sets.append((other.id, d, mtv, cons, ordered))
        return sets, defined

    def filter_cells(self, bounds):
        """space.FilterCells(bounds).FilterShapes() (space.go:181-195, cell.go:73-121): shapes in walk order."""
        cx, cy, ex, ey = self.to_cell_space(bounds)
        seen, out = [], []
        for y in range(cy, ey + 1):
            for x in range(cx, ex + 1):
                c = self.cell(x, y)
                if c is None:
                    continue
                for s in c:
                    if s.id in seen:
                        continue
                    seen.append(s.id)
                    out.append(s)
        return out


def linetest(start, end, against):
    """utils.go:276-370 with OnIntersect returning true: [(other id, mtv, center, contacts)] in callback order, and
    whether that order is defined by the restated part of sort.Slice (n <= 12, or no exact key ties)."""
    sets = []
    for shape in against:
        r = linetest_one(start, end, shape)
        if r is not None:
            sets.append((shape.id,) + r)
    go_sort_small(sets, lambda p, q: p[4] < q[4])
    return sets, True


# ---------------------------------------------------------------- the parity kit's dump, from this restatement
def _unhex(tok):
    import struct
    return struct.unpack("<d", struct.pack("<Q", int(tok, 16)))[0]


def _hx(x):
    import struct
    return "%016x" % struct.unpack("<Q", struct.pack("<d", float(x)))[0]


def dump_world(path):
    """Python twin of parity/go/dump/main.go running on this restatement instead of the real package: reads a
    parity/worlds/*.world file (own parser) and returns (dump text, set of line numbers whose ORDER is unspecified
    because Go's pdqsort would have to break exact ties among more than 12 items)."""
    with open(path) as fh:
        rows = [l.split() for l in fh.read().splitlines() if l and not l.startswith("#")]
    assert rows[0] == ["RESOLVWORLD", "1"]
    it = iter(rows[1:])
    name = next(it)[1]
    sw, sh, cw, ch = (int(t) for t in next(it)[1:5])
    margin = int(next(it)[1])
    n = int(next(it)[1])
    sp = Space(sw, sh, cw, ch)
    recs = []
    for _ in range(n):
        t = next(it)
        pos0 = (_unhex(t[8]), _unhex(t[9]))
        pos = (_unhex(t[10]), _unhex(t[11]))
        if t[1] == "C":
            s = Circle(pos[0], pos[1], _unhex(t[12]))
        else:
            nv = int(t[13])
            s = Polygon(pos0[0], pos0[1], [_unhex(x) for x in t[14:14 + 2 * nv]], closed=(t[3] == "1"))
            s.pos = pos                       # SetPosition before Space.Add: no cells yet, cached bounds kept
        sp.add(s, tags=int(t[5], 16))
        recs.append(dict(shape=s, tester=t[2] == "1", use_any=t[4] == "1", any_mask=int(t[6], 16), none_mask=int(t[7], 16)))
    frames = int(next(it)[1])
    frame_pos = []
    for f in range(1, frames):
        assert next(it) == ["pos", str(f)]
        frame_pos.append([tuple(_unhex(x) for x in next(it)[:2]) for _ in range(n)])
    m = int(next(it)[1])
    rays = [[_unhex(x) for x in next(it)[1:5]] for _ in range(m)]

    out = [f"RESOLVDUMP 1 {name}"]
    loose = set()

    def set_text(other, mtv, cons, center=None):
        s = f" {other} {_hx(mtv[0])} {_hx(mtv[1])}"
        if center is not None:
            s += f" {_hx(center[0])} {_hx(center[1])}"
        s += f" {len(cons)}"
        for p, q in cons:
            s += f" {_hx(p[0])} {_hx(p[1])} {_hx(q[0])} {_hx(q[1])}"
        return s

    for f in range(frames):
        if f:
            for i, p in enumerate(frame_pos[f - 1]):
                sp.set_position(recs[i]["shape"], p[0], p[1])
        out.append(f"frame {f}")
        for cy in range(sp.h):
            for cx in range(sp.w):
                c = sp.cells.get((cx, cy))
                if c:
                    out.append(f"cell {cy} {cx} " + " ".join(str(s.id) for s in c))
        for a, rec in enumerate(recs):
            if not rec["tester"]:
                continue
            flt = dict(use_any=rec["use_any"], any_mask=rec["any_mask"], none_mask=rec["none_mask"])
            cands = sp.candidates(rec["shape"], margin, **flt)
            out.append(f"test {a} {len(cands)}" + "".join(f" {o.id}" for o, _ in cands))
            sets, defined = sp.intersection_test(rec["shape"], margin, **flt)
            for k, (other, _, mtv, cons, ordered) in enumerate(sets):
                if not (defined and ordered):
                    loose.add(len(out))
                out.append(f"hit {a} {k}" + set_text(other, mtv, cons))
    for mode in ("all", "rect"):
        for r, ray in enumerate(rays):
            start, end = (ray[0], ray[1]), (ray[2], ray[3])
            if mode == "all":
                against = sp.shapes
            else:
                vu = v_unit(v_sub(end, start))
                pulled = v_sub(start, v_scale(vu, 0.01))
                against = sp.filter_cells(((go_min(pulled[0], end[0]), go_min(pulled[1], end[1])),
                                           (go_max(pulled[0], end[0]), go_max(pulled[1], end[1]))))
            out.append(f"ray {r} {mode}")
            sets, defined = linetest(start, end, against)
            for k, (other, mtv, center, cons, _key) in enumerate(sets):
                if not defined:
                    loose.add(len(out))
                out.append(f"rhit {r} {k}" + set_text(other, mtv, cons, center))
    return "\n".join(out) + "\n", loose


def shape_linetest(space, cp, vec, start_offset=(0.0, 0.0), include_all=False, margin=1, **filters):
    """ConvexPolygon.ShapeLineTest (convexPolygon.go:322-415) with TestAgainst = cp.SelectTouchingCells(margin)
    .FilterShapes()[.ByTags][.NotByTags] and no `Lines` selection. The reference walks its vertex SET in Go-map order
    (random); like the C++ oracle, the vertices are taken in insertion order here — the definition the host layer
    adopts. Returns [(other id, mtv, center, contacts)] sorted by |MTV|^2 (insertion sort: <= 12 sets)."""
    verts = []

    def add(v):                       # Set[Vector]: Go map keys compare with ==
        if not any(o[0] == v[0] and o[1] == v[1] for o in verts):
            verts.append(v)

    if not include_all:
        for l in cp.lines():
            if v_dot(line_normal(l), vec) < 0.01:
                continue
            v = v_invert(v_scale(line_vector(l), 0.5))
            add(v_sub(l[0], v))
            add(v_sub(l[1], v_invert(v)))
    else:
        for v in cp.transformed():
            add(v)
    vu = v_unit(vec)
    results = []
    for p in verts:
        start = v_sub(p, v_add(vu, start_offset))
        against = [o for o, _ in space.candidates(cp, margin, **filters)]   # callingShape is excluded by the selection
        sets, _ = linetest(start, v_add(p, vec), against)
        for other, mtv, center, cons, _key in sets:
            for r in results:
                if r[0] == other:
                    r[3].extend(cons)
                    if mtv[0] * mtv[0] + mtv[1] * mtv[1] < r[1][0] * r[1][0] + r[1][1] * r[1][1]:
                        r[1] = mtv
                    break
            else:
                results.append([other, mtv, center, list(cons)])
    go_sort_small(results, lambda a, b: a[1][0] * a[1][0] + a[1][1] * a[1][1] < b[1][0] * b[1][0] + b[1][1] * b[1][1])
    return results
