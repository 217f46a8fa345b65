// resolv_oracle.cpp — CPU restatement of SolarLune/resolv's collision hot path.
//
// *** TEST INFRASTRUCTURE ONLY. ***  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load this library. The
// product path (resolv_b200/csrc) never links, imports or calls it.
//
// *** PARITY UNPINNED. ***  The reference ships no tests, golden vectors or
// fixtures for this path and no Go toolchain exists in this
// image, so this restatement has been pinned only against the hand-derived
// known-answer tests of SURVEY.md §8c (tests/test_oracle_kat.py) and against
// independent property checks (closed-form candidate predicate, brute force),
// and cross-checked bit for bit against a second, separately written restatement
// (oracle/second_opinion.py, tests/test_oracle_second_opinion.py). Two readings
// of the source that agree are still not the reference: only parity/go run
// against real resolv pins it.
//
// Every function cites the reference file:line (relative to /root/reference)
// whose algorithm it follows. Data structures mirror the reference on purpose
// (per-cell shape vectors with linear-scan register / swap-remove unregister,
// row-major cell walk with linear-scan ID dedupe, per-candidate distance sort,
// per-call vertex re-transformation) so that it is also an honest CPU baseline.
//
// Arithmetic: IEEE-754 binary64, no FMA contraction (build with
// -ffp-contract=off -fno-fast-math), matching Go's default amd64 code
// generation. Go stdlib behaviours restated here (source not under
// /root/reference; Go >= 1.20 per go.mod:3):
//   math.Sqrt / Floor / Ceil / Abs  -> IEEE-exact C equivalents
//   math.Pow(x, 2)                  -> x*x (exact for normal doubles)
//   math.Min / math.Max             -> go_min / go_max below (NaN, +-0 rules)
//   math.Sin / math.Cos             -> libm (NOT bit-equal to Go; all parity
//                                      worlds use rotation == 0, pre-rotated verts)
//   sort.Slice                      -> go_sort_slice below: insertion sort for
//                                      n <= 12, pdqsort_func otherwise (restated
//                                      from memory of Go >= 1.19 sort/zsortfunc.go,
//                                      unverified here; only matters for exact
//                                      ties with n > 12)

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

namespace orc {

// ------------------------------------------------------------------ Go stdlib

// Go math.Min (src/math/dim.go): -Inf wins, NaN propagates, Min(-0,+0) = -0.
static inline double go_min(double x, double y) {
    if (std::isinf(x) && x < 0) return x;
    if (std::isinf(y) && y < 0) return y;
    if (std::isnan(x) || std::isnan(y)) return NAN;
    if (x == 0 && x == y) return std::signbit(x) ? x : y;
    return x < y ? x : y;
}

// Go math.Max (src/math/dim.go): +Inf wins, NaN propagates, Max(-0,+0) = +0.
static inline double go_max(double x, double y) {
    if (std::isinf(x) && x > 0) return x;
    if (std::isinf(y) && y > 0) return y;
    if (std::isnan(x) || std::isnan(y)) return NAN;
    if (x == 0 && x == y) return std::signbit(x) ? y : x;
    return x > y ? x : y;
}

// Go sort.Slice -> pdqsort_func over a (less, swap) pair. See header note.
struct LessSwap {
    std::function<bool(int, int)> less;
    std::function<void(int, int)> swap;
};

static void insertion_sort(const LessSwap& d, int a, int b) {
    for (int i = a + 1; i < b; i++)
        for (int j = i; j > a && d.less(j, j - 1); j--) d.swap(j, j - 1);
}

static void sift_down(const LessSwap& d, int lo, int hi, int first) {
    int root = lo;
    for (;;) {
        int child = 2 * root + 1;
        if (child >= hi) return;
        if (child + 1 < hi && d.less(first + child, first + child + 1)) child++;
        if (!d.less(first + root, first + child)) return;
        d.swap(first + root, first + child);
        root = child;
    }
}

static void heap_sort(const LessSwap& d, int a, int b) {
    int first = a, lo = 0, hi = b - a;
    for (int i = (hi - 1) / 2; i >= 0; i--) sift_down(d, i, hi, first);
    for (int i = hi - 1; i >= 0; i--) {
        d.swap(first, first + i);
        sift_down(d, lo, i, first);
    }
}

static inline int bits_len(unsigned long long v) {
    int n = 0;
    while (v) { n++; v >>= 1; }
    return n;
}

enum SortedHint { kUnknown, kIncreasing, kDecreasing };

static void order2(const LessSwap& d, int& a, int& b, int& swaps) {
    if (d.less(b, a)) { swaps++; std::swap(a, b); }
}
static int median3(const LessSwap& d, int a, int b, int c, int& swaps) {
    order2(d, a, b, swaps);
    order2(d, b, c, swaps);
    order2(d, a, b, swaps);
    return b;
}
static int median_adjacent(const LessSwap& d, int a, int& swaps) { return median3(d, a - 1, a, a + 1, swaps); }

static void choose_pivot(const LessSwap& d, int a, int b, int& pivot, SortedHint& hint) {
    const int shortestNinther = 50, maxSwaps = 4 * 3;
    int l = b - a, swaps = 0;
    int i = a + l / 4 * 1, j = a + l / 4 * 2, k = a + l / 4 * 3;
    if (l >= 8) {
        if (l >= shortestNinther) {
            i = median_adjacent(d, i, swaps);
            j = median_adjacent(d, j, swaps);
            k = median_adjacent(d, k, swaps);
        }
        j = median3(d, i, j, k, swaps);
    }
    pivot = j;
    hint = swaps == 0 ? kIncreasing : (swaps == maxSwaps ? kDecreasing : kUnknown);
}

static void reverse_range(const LessSwap& d, int a, int b) {
    int i = a, j = b - 1;
    while (i < j) { d.swap(i, j); i++; j--; }
}

static bool partial_insertion_sort(const LessSwap& d, int a, int b) {
    const int maxSteps = 5, shortestShifting = 50;
    int i = a + 1;
    for (int step = 0; step < maxSteps; step++) {
        while (i < b && !d.less(i, i - 1)) i++;
        if (i == b) return true;
        if (b - a < shortestShifting) return false;
        d.swap(i, i - 1);
        if (i - a >= 2)
            for (int j = i - 1; j >= 1; j--) {
                if (!d.less(j, j - 1)) break;
                d.swap(j, j - 1);
            }
        if (b - i >= 2)
            for (int j = i + 1; j < b; j++) {
                if (!d.less(j, j - 1)) break;
                d.swap(j, j - 1);
            }
    }
    return false;
}

static void break_patterns(const LessSwap& d, int a, int b) {
    int length = b - a;
    if (length >= 8) {
        uint64_t r = (uint64_t)length;
        unsigned modulus = 1u << bits_len((unsigned)length);
        int idx = a + (length / 4) * 2 - 1;
        for (int i = 0; i < 3; i++) {
            r ^= r << 13; r ^= r >> 7; r ^= r << 17;
            int other = (int)((unsigned)r & (modulus - 1));
            if (other >= length) other -= length;
            d.swap(idx - 1 + i, a + other);
        }
    }
}

static int partition_equal(const LessSwap& d, int a, int b, int pivot) {
    d.swap(a, pivot);
    int i = a + 1, j = b - 1;
    for (;;) {
        while (i <= j && !d.less(a, i)) i++;
        while (i <= j && d.less(a, j)) j--;
        if (i > j) break;
        d.swap(i, j);
        i++; j--;
    }
    return i;
}

static int partition(const LessSwap& d, int a, int b, int pivot, bool& already) {
    d.swap(a, pivot);
    int i = a + 1, j = b - 1;
    while (i <= j && d.less(i, a)) i++;
    while (i <= j && !d.less(j, a)) j--;
    if (i > j) { d.swap(j, a); already = true; return j; }
    d.swap(i, j);
    i++; j--;
    for (;;) {
        while (i <= j && d.less(i, a)) i++;
        while (i <= j && !d.less(j, a)) j--;
        if (i > j) break;
        d.swap(i, j);
        i++; j--;
    }
    d.swap(j, a);
    already = false;
    return j;
}

static void pdqsort(const LessSwap& d, int a, int b, int limit) {
    const int maxInsertion = 12;
    bool wasBalanced = true, wasPartitioned = true;
    for (;;) {
        int length = b - a;
        if (length <= maxInsertion) { insertion_sort(d, a, b); return; }
        if (limit == 0) { heap_sort(d, a, b); return; }
        if (!wasBalanced) { break_patterns(d, a, b); limit--; }
        int pivot; SortedHint hint;
        choose_pivot(d, a, b, pivot, hint);
        if (hint == kDecreasing) {
            reverse_range(d, a, b);
            pivot = (b - 1) - (pivot - a);
            hint = kIncreasing;
        }
        if (wasBalanced && wasPartitioned && hint == kIncreasing)
            if (partial_insertion_sort(d, a, b)) return;
        if (a > 0 && !d.less(a - 1, pivot)) {
            a = partition_equal(d, a, b, pivot);
            continue;
        }
        bool already;
        int mid = partition(d, a, b, pivot, already);
        wasPartitioned = already;
        int leftLen = mid - a, rightLen = b - mid, balanceThreshold = length / 8;
        if (leftLen < rightLen) {
            wasBalanced = leftLen >= balanceThreshold;
            pdqsort(d, a, mid, limit);
            a = mid + 1;
        } else {
            wasBalanced = rightLen >= balanceThreshold;
            pdqsort(d, mid + 1, b, limit);
            b = mid;
        }
    }
}

template <class T, class Less>
static void go_sort_slice(std::vector<T>& v, Less less) {
    int n = (int)v.size();
    if (n <= 12) {  // fast path == insertion_sort without std::function overhead
        for (int i = 1; i < n; i++)
            for (int j = i; j > 0 && less(v[j], v[j - 1]); j--) std::swap(v[j], v[j - 1]);
        return;
    }
    LessSwap d{[&](int i, int j) { return less(v[i], v[j]); }, [&](int i, int j) { std::swap(v[i], v[j]); }};
    pdqsort(d, 0, n, bits_len((unsigned)n));
}

// ------------------------------------------------------------------ vector.go

struct Vec { double x = 0, y = 0; };

static inline Vec vadd(Vec a, Vec b) { return {a.x + b.x, a.y + b.y}; }     // vector.go:54-58
static inline Vec vsub(Vec a, Vec b) { return {a.x - b.x, a.y - b.y}; }     // vector.go:73-77
static inline Vec vinv(Vec a) { return {-a.x, -a.y}; }                      // vector.go:105-109
static inline double vmag(Vec a) { return std::sqrt(a.x * a.x + a.y * a.y); }  // vector.go:112-114
static inline double vmag2(Vec a) { return a.x * a.x + a.y * a.y; }         // vector.go:117-119
static inline double vdist2(Vec a, Vec b) {                                 // vector.go:152-156
    a.x -= b.x; a.y -= b.y;
    return a.x * a.x + a.y * a.y;
}
static inline Vec vunit(Vec a) {                                            // vector.go:167-175
    double l = vmag(a);
    if (l < 1e-8 || l == 1) return a;
    return {a.x / l, a.y / l};
}
static inline Vec vrot(Vec a, double angle) {                               // vector.go:246-252
    double x = a.x, y = a.y;
    return {x * std::cos(angle) - y * std::sin(angle), x * std::sin(angle) + y * std::cos(angle)};
}
static inline Vec vscale(Vec a, double s) { return {a.x * s, a.y * s}; }    // vector.go:273-277
static inline double vdot(Vec a, Vec b) { return a.x * b.x + a.y * b.y; }   // vector.go:287-289

// ------------------------------------------------------------------ utils.go

struct Space;

struct Projection { double min, max; };
static inline double overlap(Projection p, Projection o) {                  // utils.go:75-77
    return go_min(p.max - o.min, o.max - p.min);
}

struct Bounds {
    Vec min, max;
    const Space* space = nullptr;
};

struct Shape;

struct Cell {                                                               // cell.go:4-7
    int x, y;
    std::vector<Shape*> shapes;
    bool contains(Shape* s) const {                                         // cell.go:41-48
        for (Shape* o : shapes) if (o == s) return true;
        return false;
    }
    void reg(Shape* s) { if (!contains(s)) shapes.push_back(s); }           // cell.go:19-23
    void unreg(Shape* s) {                                                  // cell.go:26-38
        for (size_t i = 0; i < shapes.size(); i++)
            if (shapes[i] == s) {
                shapes[i] = shapes.back();
                shapes.pop_back();
                break;
            }
    }
};

struct Space {                                                              // space.go:7-11
    std::vector<std::vector<Cell*>> cells;  // [y][x]
    std::vector<Shape*> shapes;
    int cellWidth, cellHeight;

    Cell* cell(long long cx, long long cy) const {                          // space.go:135-142
        if (cy >= 0 && cy < (long long)cells.size() && cx >= 0 && cx < (long long)cells[cy].size())
            return cells[cy][cx];
        return nullptr;
    }
    int widthInCells() const { return cells.empty() ? 0 : (int)cells[0].size(); }
    int heightInCells() const { return (int)cells.size(); }
};

static inline void to_cell_space(const Bounds& b, long long out[4]) {       // utils.go:90-98
    out[0] = (long long)std::floor(b.min.x / (double)b.space->cellWidth);
    out[1] = (long long)std::floor(b.min.y / (double)b.space->cellHeight);
    out[2] = (long long)std::floor(b.max.x / (double)b.space->cellWidth);
    out[3] = (long long)std::floor(b.max.y / (double)b.space->cellHeight);
}

static inline Vec bounds_center(const Bounds& b) {                          // utils.go:101-103
    return vadd(b.min, vscale(vsub(b.max, b.min), 0.5));
}

static inline Bounds bounds_move(Bounds b, double x, double y) {            // utils.go:132-138
    b.min.x += x; b.min.y += y; b.max.x += x; b.max.y += y;
    return b;
}

static inline Bounds bounds_intersection(const Bounds& b, const Bounds& o) {  // utils.go:152-168
    Bounds ov;
    if (o.max.x < b.min.x || o.min.x > b.max.x || o.max.y < b.min.y || o.min.y > b.max.y) return ov;
    ov.max.x = go_min(o.max.x, b.max.x);
    ov.min.x = go_max(o.min.x, b.min.x);
    ov.max.y = go_min(o.max.y, b.max.y);
    ov.min.y = go_max(o.min.y, b.min.y);
    return ov;
}
static inline bool bounds_is_empty(const Bounds& b) {                       // utils.go:171-173 (Min.X typo kept)
    return b.max.x - b.min.x == 0 && b.max.y - b.min.x == 0;
}
static inline bool bounds_is_intersecting(const Bounds& b, const Bounds& o) {  // utils.go:146-149
    return !bounds_is_empty(bounds_intersection(b, o));
}

// ------------------------------------------------------------------ contactSet.go

struct Intersection { Vec point, normal; };                                 // contactSet.go:8-11
struct IntersectionSet {                                                    // contactSet.go:15-20
    std::vector<Intersection> intersections;
    Vec center, mtv;
    Shape* other = nullptr;
    bool empty() const { return intersections.empty(); }                    // contactSet.go:107-109
};

// ------------------------------------------------------------------ shapes

enum Kind { kCircle = 0, kPolygon = 1 };

struct Line {                                                               // convexPolygon.go:690-692
    Vec start, end;
    Vec vector() const { return vunit(vsub(end, start)); }                  // convexPolygon.go:710-712
    Vec normal() const {                                                    // convexPolygon.go:705-708
        Vec v = vector();
        return vunit(Vec{v.y, -v.x});
    }
};

struct Shape {                                                              // shape.go:54-62, circle.go:4-7, convexPolygon.go:12-20
    Kind kind;
    Vec pos;
    Space* space = nullptr;
    std::vector<Cell*> touching;
    uint64_t tags = 0;
    uint32_t id = 0;
    // circle
    double radius = 0;
    // polygon
    Vec scale{1, 1};
    double rotation = 0;
    std::vector<Vec> points;
    bool closed = true;
    Bounds lbounds;  // cached *local* bounds (convexPolygon.go:196)

    std::vector<Vec> transformed() const {                                  // convexPolygon.go:144-154
        std::vector<Vec> t;
        for (const Vec& pt : points) {
            Vec p{pt.x * scale.x, pt.y * scale.y};
            if (rotation != 0) p = vrot(p, -rotation);
            t.push_back(Vec{p.x + pos.x, p.y + pos.y});
        }
        return t;
    }

    void update_bounds() {                                                  // convexPolygon.go:163-197
        std::vector<Vec> t = transformed();
        Vec tl = t[0], br = t[0];
        for (size_t i = 0; i < t.size(); i++) {
            Vec p = t[i];
            if (p.x < tl.x) tl.x = p.x; else if (p.x > br.x) br.x = p.x;
            if (p.y < tl.y) tl.y = p.y; else if (p.y > br.y) br.y = p.y;
        }
        Bounds b; b.min = tl; b.max = br; b.space = space;
        lbounds = bounds_move(b, -pos.x, -pos.y);
    }

    Bounds bounds() const {
        if (kind == kCircle) {                                              // circle.go:33-39
            Bounds b;
            b.min = Vec{pos.x - radius, pos.y - radius};
            b.max = Vec{pos.x + radius, pos.y + radius};
            b.space = space;
            return b;
        }
        Bounds b = lbounds;                                                 // convexPolygon.go:158-161
        b.space = space;
        return bounds_move(b, pos.x, pos.y);
    }

    Vec center() const { return bounds_center(bounds()); }                  // convexPolygon.go:200-215

    std::vector<Line> lines() const {                                       // convexPolygon.go:117-141
        std::vector<Line> out;
        std::vector<Vec> v = transformed();
        for (size_t i = 0; i < v.size(); i++) {
            Vec s = v[i], e = v[0];
            if (i + 1 < v.size()) e = v[i + 1];
            else if (!closed || points.size() <= 2) break;
            out.push_back(Line{s, e});
        }
        return out;
    }

    std::vector<Vec> sat_axes() const {                                     // convexPolygon.go:235-243
        std::vector<Vec> axes;
        for (const Line& l : lines()) axes.push_back(l.normal());
        return axes;
    }

    Projection project(Vec axis) const {
        if (kind == kCircle) {                                              // circle.go:41-54
            axis = vunit(axis);
            double pc = vdot(axis, pos);
            double pr = vmag(axis) * radius;
            double mn = pc - pr, mx = pc + pr;
            if (mn > mx) std::swap(mn, mx);
            return {mn, mx};
        }
        axis = vunit(axis);                                                 // convexPolygon.go:218-232
        std::vector<Vec> v = transformed();
        double mn = vdot(axis, v[0]), mx = mn;
        for (size_t i = 1; i < v.size(); i++) {
            double p = vdot(axis, v[i]);
            if (p < mn) mn = p; else if (p > mx) mx = p;
        }
        return {mn, mx};
    }

    void remove_from_touching() {                                           // shape.go:183-189
        for (Cell* c : touching) c->unreg(this);
        touching.clear();
    }
    void add_to_touching() {                                                // shape.go:191-214
        if (!space) return;
        long long r[4];
        to_cell_space(bounds(), r);
        for (long long y = r[1]; y <= r[3]; y++)
            for (long long x = r[0]; x <= r[2]; x++) {
                Cell* c = space->cell(x, y);
                if (c) { c->reg(this); touching.push_back(c); }
            }
    }
    void update() { remove_from_touching(); add_to_touching(); }            // shape.go:239-242
    void set_position(double x, double y) { pos.x = x; pos.y = y; update(); }  // shape.go:115-119
    void move(double x, double y) { pos.x += x; pos.y += y; update(); }     // shape.go:98-102
};

// convexPolygon.go:715-744
static inline bool line_x_line(const Line& line, const Line& other, Vec& out) {
    double det = (line.end.x - line.start.x) * (other.end.y - other.start.y) -
                 (other.end.x - other.start.x) * (line.end.y - line.start.y);
    if (det != 0) {
        double lambda = (((line.start.y - other.start.y) * (other.end.x - other.start.x)) -
                         ((line.start.x - other.start.x) * (other.end.y - other.start.y)) + 1) / det;
        double gamma = (((line.start.y - other.start.y) * (line.end.x - line.start.x)) -
                        ((line.start.x - other.start.x) * (line.end.y - line.start.y)) + 1) / det;
        if ((0 < lambda && lambda < 1) && (0 < gamma && gamma < 1)) {
            double dx = line.end.x - line.start.x, dy = line.end.y - line.start.y;
            out = Vec{line.start.x + (lambda * dx), line.start.y + (lambda * dy)};
            return true;
        }
    }
    return false;
}

// convexPolygon.go:747-789
static inline int line_x_circle(const Line& line, const Shape& circle, Vec out[2]) {
    int n = 0;
    Vec cp = circle.pos;
    Vec ls = vsub(line.start, cp), le = vsub(line.end, cp), diff = vsub(le, ls);
    double a = diff.x * diff.x + diff.y * diff.y;
    double b = 2 * ((diff.x * ls.x) + (diff.y * ls.y));
    double c = (ls.x * ls.x) + (ls.y * ls.y) - (circle.radius * circle.radius);
    double det = b * b - (4 * a * c);
    if (det < 0) {
    } else if (det == 0) {
        double t = -b / (2 * a);
        if (t >= 0 && t <= 1) out[n++] = Vec{line.start.x + t * diff.x, line.start.y + t * diff.y};
    } else {
        double t = (-b + std::sqrt(det)) / (2 * a);
        if (t >= 0 && t <= 1) out[n++] = Vec{line.start.x + t * diff.x, line.start.y + t * diff.y};
        t = (-b - std::sqrt(det)) / (2 * a);
        if (t >= 0 && t <= 1) out[n++] = Vec{line.start.x + t * diff.x, line.start.y + t * diff.y};
    }
    return n;
}

// convexPolygon.go:418-514
static bool calculate_mtv(const Shape& cp, const Shape& other, Vec& out) {
    Vec delta{0, 0};
    Vec smallest{DBL_MAX, 0};
    if (other.kind == kPolygon) {
        for (const Vec& axis : cp.sat_axes()) {
            double ov = overlap(cp.project(axis), other.project(axis));
            if (ov <= 0) return false;
            if (vmag(smallest) > ov) smallest = vscale(axis, ov);
        }
        for (const Vec& axis : other.sat_axes()) {
            double ov = overlap(cp.project(axis), other.project(axis));
            if (ov <= 0) return false;
            if (vmag(smallest) > ov) smallest = vscale(axis, ov);
        }
        if (vdot(vsub(cp.center(), other.center()), smallest) < 0) smallest = vinv(smallest);
    } else {
        std::vector<Vec> verts = cp.transformed();
        Vec center = other.pos;
        go_sort_slice(verts, [&](const Vec& a, const Vec& b) { return vmag(vsub(a, center)) < vmag(vsub(b, center)); });
        Vec axis{center.x - verts[0].x, center.y - verts[0].y};
        double ov = overlap(cp.project(axis), other.project(axis));
        if (ov <= 0) return false;
        smallest = vscale(vunit(axis), ov);
        for (const Vec& ax : cp.sat_axes()) {
            double o2 = overlap(cp.project(ax), other.project(ax));
            if (o2 <= 0) return false;
            if (vmag(smallest) > o2) smallest = vscale(ax, o2);
        }
        if (vdot(vsub(cp.center(), other.pos), smallest) < 0) smallest = vinv(smallest);
    }
    delta.x = smallest.x;
    delta.y = smallest.y;
    Vec pointing = vsub(other.pos, cp.pos);
    if (vdot(pointing, delta) > 0) delta = vinv(delta);
    out = delta;
    return true;
}

// shape.go:318-354
static IntersectionSet circle_convex_test(Shape& circle, Shape& convex) {
    IntersectionSet set;
    if (!bounds_is_intersecting(convex.bounds(), circle.bounds())) return set;
    for (const Line& line : convex.lines()) {
        Vec pts[2];
        int n = line_x_circle(line, circle, pts);
        for (int i = 0; i < n; i++) set.intersections.push_back({pts[i], line.normal()});
    }
    if (!set.empty()) {
        set.other = &convex;
        Vec mtv;
        if (calculate_mtv(convex, circle, mtv)) set.mtv = vinv(mtv);
    }
    return set;
}

// shape.go:356-397
static IntersectionSet convex_circle_test(Shape& convex, Shape& circle) {
    IntersectionSet set;
    if (!bounds_is_intersecting(convex.bounds(), circle.bounds())) return set;
    for (const Line& line : convex.lines()) {
        Vec pts[2];
        int n = line_x_circle(line, circle, pts);
        for (int i = 0; i < n; i++) set.intersections.push_back({pts[i], vunit(vsub(pts[i], circle.pos))});
    }
    if (!set.empty()) {
        set.other = &circle;
        Vec c = circle.pos;
        go_sort_slice(set.intersections,
                      [&](const Intersection& a, const Intersection& b) { return vdist2(a.point, c) < vdist2(b.point, c); });
        Vec mtv;
        if (calculate_mtv(convex, circle, mtv)) set.mtv = mtv;
    }
    return set;
}

// shape.go:399-436
static IntersectionSet circle_circle_test(Shape& A, Shape& B) {
    IntersectionSet set;
    if (!bounds_is_intersecting(A.bounds(), B.bounds())) return set;
    double dx = B.pos.x - A.pos.x, dy = B.pos.y - A.pos.y;
    double d = std::sqrt(dx * dx + dy * dy);
    if (d > A.radius + B.radius || d < std::fabs(A.radius - B.radius) || d == 0) return set;
    double a = (A.radius * A.radius - B.radius * B.radius + d * d) / (2 * d);
    double h = std::sqrt(A.radius * A.radius - a * a);
    double x2 = A.pos.x + a * (B.pos.x - A.pos.x) / d;
    double y2 = A.pos.y + a * (B.pos.y - A.pos.y) / d;
    set.intersections.push_back({Vec{x2 + h * (B.pos.y - A.pos.y) / d, y2 - h * (B.pos.x - A.pos.x) / d}, Vec{}});
    set.intersections.push_back({Vec{x2 - h * (B.pos.y - A.pos.y) / d, y2 + h * (B.pos.x - A.pos.x) / d}, Vec{}});
    for (auto& it : set.intersections) it.normal = vunit(vsub(it.point, A.pos));
    set.mtv = Vec{A.pos.x - B.pos.x, A.pos.y - B.pos.y};
    double dist = vmag(set.mtv);
    set.mtv = vscale(vunit(set.mtv), A.radius + B.radius - dist);
    set.other = &B;
    return set;
}

// shape.go:438-479
static IntersectionSet convex_convex_test(Shape& A, Shape& B) {
    IntersectionSet set;
    if (!bounds_is_intersecting(A.bounds(), B.bounds())) return set;
    for (const Line& ol : B.lines())
        for (const Line& l : A.lines()) {
            Vec p;
            if (line_x_line(l, ol, p)) set.intersections.push_back({p, ol.normal()});
        }
    if (!set.empty()) {
        set.other = &B;
        Vec c = A.center();
        go_sort_slice(set.intersections,
                      [&](const Intersection& a, const Intersection& b) { return vdist2(a.point, c) < vdist2(b.point, c); });
        Vec mtv;
        if (calculate_mtv(A, B, mtv)) set.mtv = mtv;
    }
    return set;
}

// circle.go:69-82, convexPolygon.go:288-300
static IntersectionSet intersection(Shape& a, Shape& b) {
    if (a.kind == kCircle) return b.kind == kPolygon ? circle_convex_test(a, b) : circle_circle_test(a, b);
    return b.kind == kPolygon ? convex_convex_test(a, b) : convex_circle_test(a, b);
}

// ------------------------------------------------------------------ filters / iteration

// shapefilter.go:17-61 restricted to what the hot path evaluates on device:
// one optional ByTags(any) and one merged NotByTags(none) predicate. (Chained
// NotByTags masks OR together exactly; further ByTags / closure filters are
// applied host-side on the compacted buffer in the product.)
struct TagFilter {
    bool useAny = false;
    uint64_t any = 0, none = 0;
    bool pass(uint64_t t) const {
        if (useAny && !((t & any) > 0)) return false;                        // tags.go:29-31, shapefilter.go:47-52
        if ((t & none) > 0) return false;                                   // shapefilter.go:56-61
        return true;
    }
};

struct CellSelection {                                                      // cell.go:66-70
    long long sx, sy, ex, ey;
    const Space* space;
    Shape* excludeSelf;
};

// shape.go:220-237
static CellSelection select_touching_cells(Shape& s, int margin) {
    long long r[4];
    Bounds b = s.bounds();
    to_cell_space(b, r);
    return CellSelection{r[0] - margin, r[1] - margin, r[2] + margin, r[3] + margin, s.space, &s};
}

struct Scratch {  // the reference's package-level scratch slices (shape.go:265, utils.go:255,272), per thread here
    std::vector<uint32_t> idset;
    struct Possible { Shape* shape; double dist; };
    std::vector<Possible> possible;
    std::vector<IntersectionSet> lineSets;
};

// cell.go:86-121 — row-major walk, skip self, linear-scan ID dedupe.
template <class F>
static void cell_selection_for_each(const CellSelection& c, Scratch& sc, F fn) {
    sc.idset.clear();
    for (long long y = c.sy; y <= c.ey; y++)
        for (long long x = c.sx; x <= c.ex; x++) {
            Cell* cell = c.space->cell(x, y);
            if (!cell) continue;
            for (Shape* s : cell->shapes) {
                if (s == c.excludeSelf) continue;
                bool seen = false;
                for (uint32_t v : sc.idset) if (v == s->id) { seen = true; break; }  // utils.go:246-253
                if (seen) continue;
                if (!fn(s)) break;
                sc.idset.push_back(s->id);
            }
        }
}

// shapefilter.go:26-43 over a CellSelection
template <class F>
static void filter_for_each(const CellSelection& c, const TagFilter& f, Scratch& sc, F fn) {
    cell_selection_for_each(c, sc, [&](Shape* s) {
        if (!f.pass(s->tags)) return true;
        fn(s);
        return true;
    });
}

// shape.go:273-316. `gathered` (oracle instrumentation, not in the reference)
// receives the candidate list in generation order before the distance sort.
template <class OnIntersect, class OnGather>
static bool intersection_test(Shape& self, const CellSelection& sel, const TagFilter& f, Scratch& sc,
                              OnIntersect onIntersect, OnGather gathered) {
    sc.possible.clear();
    filter_for_each(sel, f, sc, [&](Shape* other) {
        if (other == &self) return;
        sc.possible.push_back({other, vdist2(other->pos, self.pos)});       // shape.go:179-181
    });
    gathered(sc.possible);
    go_sort_slice(sc.possible, [](const Scratch::Possible& a, const Scratch::Possible& b) { return a.dist < b.dist; });
    bool collided = false;
    for (const auto& p : sc.possible) {
        IntersectionSet res = intersection(self, *p.shape);
        if (!res.empty()) {
            collided = true;
            if (!onIntersect(res, p.dist)) break;
        }
    }
    return collided;
}

// utils.go:276-370. The iterator is abstracted as a callable `forEach(fn)`.
template <class ForEach, class OnIntersect>
static bool line_test(Vec S, Vec E, Shape* callingShape, ForEach forEach, Scratch& sc, OnIntersect onIntersect) {
    double castMargin = 0.01;
    Vec vu = vunit(vsub(E, S));
    Vec start = vsub(S, vscale(vu, castMargin));
    Line line{start, E};
    sc.lineSets.clear();
    forEach([&](Shape* other) {
        if (other == callingShape) return;
        IntersectionSet cs;
        if (other->kind == kCircle) {
            Vec pts[2];
            int n = line_x_circle(line, *other, pts);
            for (int i = 0; i < n; i++) cs.intersections.push_back({pts[i], vunit(vsub(pts[i], other->pos))});
        } else {
            for (const Line& ol : other->lines()) {
                Vec p;
                if (line_x_line(line, ol, p)) cs.intersections.push_back({p, ol.normal()});
            }
        }
        if (!cs.intersections.empty()) {
            cs.other = other;
            for (const auto& c : cs.intersections) cs.center = vadd(cs.center, c.point);
            cs.center.x /= (double)cs.intersections.size();
            cs.center.y /= (double)cs.intersections.size();
            go_sort_slice(cs.intersections,
                          [&](const Intersection& a, const Intersection& b) { return vdist2(a.point, S) < vdist2(b.point, S); });
            cs.mtv = vsub(vsub(cs.intersections[0].point, S), vscale(vu, castMargin));
            sc.lineSets.push_back(cs);
        }
    });
    go_sort_slice(sc.lineSets, [&](const IntersectionSet& a, const IntersectionSet& b) {
        return vdist2(a.intersections[0].point, line.start) < vdist2(b.intersections[0].point, line.start);
    });
    int n = (int)sc.lineSets.size();
    for (int i = 0; i < n; i++)
        if (!onIntersect(sc.lineSets[i], i, n)) break;
    return n > 0;
}

}  // namespace orc

// =====================================================================================
// C API for the ctypes binding (oracle/oracle.py). Shapes are addressed by their index in
// Space.shapes (== insertion order == shape ID within the space).
// =====================================================================================

using namespace orc;

struct OrcWorld {
    Space space;
    std::vector<Shape*> all;  // owned
    // last frame / query results
    std::vector<uint32_t> candA, candB, candRank;
    std::vector<double> candDist;
    std::vector<uint32_t> hitA, hitB, hitFirst, hitCount, hitOrder;
    std::vector<double> hitMtv, hitDist;
    std::vector<double> contacts;  // (px,py,nx,ny) per contact
    std::vector<double> hitCenter;
    ~OrcWorld() {
        for (Shape* s : all) delete s;
        for (auto& row : space.cells) for (Cell* c : row) delete c;
    }
};

extern "C" {

// space.go:16-30,113-131
void* orc_space_new(int sw, int sh, int cw, int ch) {
    OrcWorld* w = new OrcWorld;
    w->space.cellWidth = cw;
    w->space.cellHeight = ch;
    int W = (int)std::ceil((double)sw / (double)cw), H = (int)std::ceil((double)sh / (double)ch);
    for (int y = 0; y < H; y++) {
        w->space.cells.emplace_back();
        for (int x = 0; x < W; x++) w->space.cells[y].push_back(new Cell{x, y, {}});
    }
    return w;
}
void orc_space_free(void* p) { delete (OrcWorld*)p; }
int orc_width_cells(void* p) { return ((OrcWorld*)p)->space.widthInCells(); }
int orc_height_cells(void* p) { return ((OrcWorld*)p)->space.heightInCells(); }
int orc_num_shapes(void* p) { return (int)((OrcWorld*)p)->all.size(); }

static int add_shape(OrcWorld* w, Shape* s) {                               // space.go:33-46
    s->id = (uint32_t)w->all.size();
    s->space = &w->space;
    s->update();
    w->space.shapes.push_back(s);
    w->all.push_back(s);
    return (int)s->id;
}

// circle.go:10-17 + space.go:33-46
int orc_add_circle(void* p, double x, double y, double r, uint64_t tags) {
    Shape* s = new Shape;
    s->kind = kCircle; s->pos = Vec{x, y}; s->radius = r; s->tags = tags;
    return add_shape((OrcWorld*)p, s);
}

// convexPolygon.go:28-46,92-107 (AddPoints -> updateBounds) + space.go:33-46
int orc_add_polygon(void* p, double x, double y, const double* pts, int npts, int closed, uint64_t tags) {
    Shape* s = new Shape;
    s->kind = kPolygon; s->pos = Vec{x, y}; s->tags = tags; s->closed = closed != 0;
    for (int i = 0; i < npts; i++) s->points.push_back(Vec{pts[2 * i], pts[2 * i + 1]});
    s->update_bounds();
    return add_shape((OrcWorld*)p, s);
}

// convexPolygon.go:616-632 (NewRectangle)
int orc_add_rectangle(void* p, double x, double y, double w, double h, uint64_t tags) {
    double hw = w / 2, hh = h / 2;
    double pts[8] = {-hw, -hh, hw, -hh, hw, hh, -hw, hh};
    return orc_add_polygon(p, x, y, pts, 4, 1, tags);
}

// convexPolygon.go:638-643 (NewRectangleFromTopLeft): construct, then Move(w/2, h/2) *before* Add.
int orc_add_rectangle_top_left(void* p, double x, double y, double w, double h, uint64_t tags) {
    OrcWorld* wd = (OrcWorld*)p;
    double hw = w / 2, hh = h / 2;
    Shape* s = new Shape;
    s->kind = kPolygon; s->pos = Vec{x, y}; s->tags = tags; s->closed = true;
    s->points = {Vec{-hw, -hh}, Vec{hw, -hh}, Vec{hw, hh}, Vec{-hw, hh}};
    s->update_bounds();
    s->pos.x += w / 2; s->pos.y += h / 2;
    return add_shape(wd, s);
}

// convexPolygon.go:671-684 (NewLine): two mirrored points, stays open because len(Points) <= 2.
int orc_add_line(void* p, double x1, double y1, double x2, double y2, uint64_t tags) {
    double cx = x1 + ((x2 - x1) / 2), cy = y1 + ((y2 - y1) / 2);
    double pts[4] = {cx - x1, cy - y1, cx - x2, cy - y2};
    return orc_add_polygon(p, cx, cy, pts, 2, 1, tags);
}

// Bulk construction for generated worlds. Polygon i is built at pos0 (NewConvexPolygon -> AddPoints ->
// updateBounds, convexPolygon.go:28-46,92-107), moved to pos while still outside any Space (like
// NewRectangleFromTopLeft's Move, convexPolygon.go:638-643), then Space.Add()ed (space.go:33-46).
void orc_add_bulk(void* p, int n, const uint8_t* kind, const double* pos0, const double* pos, const double* radius,
                  const uint32_t* vert_off, const uint32_t* nv, const double* verts, const uint8_t* closed,
                  const uint64_t* tags) {
    OrcWorld* w = (OrcWorld*)p;
    for (int i = 0; i < n; i++) {
        Shape* s = new Shape;
        s->tags = tags ? tags[i] : 0;
        if (kind[i] == 0) {
            s->kind = kCircle; s->pos = Vec{pos[2 * i], pos[2 * i + 1]}; s->radius = radius[i];
        } else {
            s->kind = kPolygon; s->pos = Vec{pos0[2 * i], pos0[2 * i + 1]}; s->closed = closed[i] != 0;
            for (uint32_t k = 0; k < nv[i]; k++) s->points.push_back(Vec{verts[2 * (vert_off[i] + k)], verts[2 * (vert_off[i] + k) + 1]});
            s->update_bounds();
            s->pos = Vec{pos[2 * i], pos[2 * i + 1]};
        }
        add_shape(w, s);
    }
}

void orc_set_position(void* p, int i, double x, double y) { ((OrcWorld*)p)->all[i]->set_position(x, y); }
void orc_move(void* p, int i, double dx, double dy) { ((OrcWorld*)p)->all[i]->move(dx, dy); }
void orc_set_tags(void* p, int i, uint64_t t) { ((OrcWorld*)p)->all[i]->tags = t; }

// Bulk SetPosition for all shapes (the per-frame motion phase of the two-phase harness).
void orc_set_positions(void* p, const double* xy) {
    OrcWorld* w = (OrcWorld*)p;
    for (size_t i = 0; i < w->all.size(); i++) w->all[i]->set_position(xy[2 * i], xy[2 * i + 1]);
}

void orc_get_position(void* p, int i, double* out) {
    Shape* s = ((OrcWorld*)p)->all[i];
    out[0] = s->pos.x; out[1] = s->pos.y;
}
// cached local bounds (convexPolygon.go:196) or (-r,-r,r,r) for circles
void orc_get_local_bounds(void* p, int i, double* out) {
    Shape* s = ((OrcWorld*)p)->all[i];
    if (s->kind == kCircle) { out[0] = -s->radius; out[1] = -s->radius; out[2] = s->radius; out[3] = s->radius; return; }
    out[0] = s->lbounds.min.x; out[1] = s->lbounds.min.y; out[2] = s->lbounds.max.x; out[3] = s->lbounds.max.y;
}
void orc_get_bounds(void* p, int i, double* out) {
    Bounds b = ((OrcWorld*)p)->all[i]->bounds();
    out[0] = b.min.x; out[1] = b.min.y; out[2] = b.max.x; out[3] = b.max.y;
}
void orc_get_cell_rect(void* p, int i, long long* out) {
    Bounds b = ((OrcWorld*)p)->all[i]->bounds();
    to_cell_space(b, out);
}

// Cell membership dump: counts[H*W] and, if entries != null, the shape indices per cell in the
// reference's in-cell order, concatenated row-major. Returns total registrations R.
long long orc_cells_dump(void* p, uint32_t* counts, uint32_t* entries, long long cap) {
    OrcWorld* w = (OrcWorld*)p;
    long long R = 0;
    int W = w->space.widthInCells(), H = w->space.heightInCells();
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            Cell* c = w->space.cells[y][x];
            if (counts) counts[(size_t)y * W + x] = (uint32_t)c->shapes.size();
            for (Shape* s : c->shapes) {
                if (entries && R < cap) entries[R] = s->id;
                R++;
            }
        }
    return R;
}

// One ordered pair test A.Intersection(B). out_mtv[2]; contacts (px,py,nx,ny)*cap. Returns #contacts.
int orc_intersection(void* p, int a, int b, double* out_mtv, double* out_contacts, int cap) {
    OrcWorld* w = (OrcWorld*)p;
    IntersectionSet s = intersection(*w->all[a], *w->all[b]);
    out_mtv[0] = s.mtv.x; out_mtv[1] = s.mtv.y;
    int n = (int)s.intersections.size();
    for (int i = 0; i < n && i < cap; i++) {
        out_contacts[4 * i + 0] = s.intersections[i].point.x;
        out_contacts[4 * i + 1] = s.intersections[i].point.y;
        out_contacts[4 * i + 2] = s.intersections[i].normal.x;
        out_contacts[4 * i + 3] = s.intersections[i].normal.y;
    }
    return n;
}

struct OrcFrameCounts {
    long long testers, candidates, hits, contacts;
    double mtv_sum_x, mtv_sum_y;  // accumulated like examples/worldBouncer.go:162 so timing runs keep the work live
    double seconds;
};

// The two-phase ("frozen") frame of SURVEY.md §3.2: every tester i with tester[i] != 0 runs
//   self.IntersectionTest{ TestAgainst: self.SelectTouchingCells(margin).FilterShapes()[.ByTags(any[i])][.NotByTags(none[i])],
//                          OnIntersect: record, return true }
// against the current (unmodified) world. record != 0 keeps candidates / hits / contacts for the
// parity tests; record == 0 only counts (timing mode). threads > 1 splits the tester range over
// std::threads with thread-local scratch (the reference itself is single-goroutine, shape.go:265).
void orc_frame(void* p, int margin, const uint8_t* tester, const uint8_t* use_any, const uint64_t* any,
               const uint64_t* none, int record, int threads, OrcFrameCounts* out) {
    OrcWorld* w = (OrcWorld*)p;
    int n = (int)w->all.size();
    if (threads < 1) threads = 1;
    if (record) threads = 1;
    w->candA.clear(); w->candB.clear(); w->candRank.clear(); w->candDist.clear();
    w->hitA.clear(); w->hitB.clear(); w->hitFirst.clear(); w->hitCount.clear(); w->hitOrder.clear();
    w->hitMtv.clear(); w->hitDist.clear(); w->contacts.clear();
    std::vector<OrcFrameCounts> part(threads);
    auto t0 = std::chrono::steady_clock::now();
    auto work = [&](int tid) {
        Scratch sc;
        OrcFrameCounts c{};
        int lo = (int)((long long)n * tid / threads), hi = (int)((long long)n * (tid + 1) / threads);
        for (int i = lo; i < hi; i++) {
            if (tester && !tester[i]) continue;
            Shape& self = *w->all[i];
            c.testers++;
            TagFilter f;
            if (use_any && use_any[i]) { f.useAny = true; f.any = any[i]; }
            if (none) f.none = none[i];
            CellSelection sel = select_touching_cells(self, margin);
            uint32_t order = 0;
            intersection_test(
                self, sel, f, sc,
                [&](const IntersectionSet& set, double dist) {
                    c.hits++;
                    c.contacts += (long long)set.intersections.size();
                    c.mtv_sum_x += set.mtv.x; c.mtv_sum_y += set.mtv.y;
                    if (record) {
                        w->hitA.push_back(self.id); w->hitB.push_back(set.other->id);
                        w->hitOrder.push_back(order);
                        w->hitFirst.push_back((uint32_t)(w->contacts.size() / 4));
                        w->hitCount.push_back((uint32_t)set.intersections.size());
                        w->hitMtv.push_back(set.mtv.x); w->hitMtv.push_back(set.mtv.y);
                        w->hitDist.push_back(dist);
                        for (const auto& it : set.intersections) {
                            w->contacts.push_back(it.point.x); w->contacts.push_back(it.point.y);
                            w->contacts.push_back(it.normal.x); w->contacts.push_back(it.normal.y);
                        }
                    }
                    order++;
                    return true;
                },
                [&](const std::vector<Scratch::Possible>& poss) {
                    c.candidates += (long long)poss.size();
                    if (record)
                        for (size_t k = 0; k < poss.size(); k++) {
                            w->candA.push_back(self.id); w->candB.push_back(poss[k].shape->id);
                            w->candRank.push_back((uint32_t)k); w->candDist.push_back(poss[k].dist);
                        }
                });
        }
        part[tid] = c;
    };
    if (threads == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < threads; t++) th.emplace_back(work, t);
        for (auto& t : th) t.join();
    }
    auto t1 = std::chrono::steady_clock::now();
    OrcFrameCounts tot{};
    for (auto& c : part) {
        tot.testers += c.testers; tot.candidates += c.candidates; tot.hits += c.hits; tot.contacts += c.contacts;
        tot.mtv_sum_x += c.mtv_sum_x; tot.mtv_sum_y += c.mtv_sum_y;
    }
    tot.seconds = std::chrono::duration<double>(t1 - t0).count();
    if (out) *out = tot;
}

// Config 1: the Bouncer demo's update loop headless (examples/worldBouncer.go:139-174), for
// `frames` frames over the dynamic shapes dyn[0..n_dyn) with velocities vel[2*i] (in/out):
//   Movement.Y += 0.25; Movement = Movement.ClampMagnitude(8); MoveVec(Movement);
//   IntersectionTest{SelectTouchingCells(1).FilterShapes(), OnIntersect: Movement = Movement.Reflect(normal0).Scale(0.9);
//                    totalMTV += MTV; return true};  MoveVec(totalMTV)
// sequential != 0: shape by shape, each seeing the moves of the ones before it (what the demo does).
// sequential == 0: the two-phase / frozen form (all moves, all tests against the frozen world, all MTV moves) — the
// form a batched frame (rsv_frame) computes.
void orc_bouncer_run(void* p, const int* dyn, int n_dyn, double* vel, int frames, int sequential, OrcFrameCounts* out) {
    OrcWorld* w = (OrcWorld*)p;
    Scratch sc;
    OrcFrameCounts c{};
    std::vector<Vec> total((size_t)n_dyn);
    auto t0 = std::chrono::steady_clock::now();
    auto pre = [&](int k) {
        Vec m{vel[2 * k], vel[2 * k + 1]};
        m.y += 0.25;
        if (vmag(m) > 8) m = vscale(vunit(m), 8);                  // vector.go:122-127
        vel[2 * k] = m.x; vel[2 * k + 1] = m.y;
        w->all[dyn[k]]->move(m.x, m.y);
    };
    auto test = [&](int k) {
        Shape& self = *w->all[dyn[k]];
        Vec tot{0, 0};
        c.testers++;
        intersection_test(
            self, select_touching_cells(self, 1), TagFilter{}, sc,
            [&](const IntersectionSet& set, double) {
                Vec m{vel[2 * k], vel[2 * k + 1]};
                Vec nn = vunit(set.intersections[0].normal);        // vector.go:197-200
                m = vscale(vsub(m, vscale(nn, 2 * vdot(nn, m))), 0.9);
                vel[2 * k] = m.x; vel[2 * k + 1] = m.y;
                tot = vadd(tot, set.mtv);
                c.hits++;
                c.contacts += (long long)set.intersections.size();
                c.mtv_sum_x += set.mtv.x; c.mtv_sum_y += set.mtv.y;
                return true;
            },
            [&](const std::vector<Scratch::Possible>& poss) { c.candidates += (long long)poss.size(); });
        return tot;
    };
    for (int f = 0; f < frames; f++) {
        if (sequential) {
            for (int k = 0; k < n_dyn; k++) {
                pre(k);
                Vec t = test(k);
                w->all[dyn[k]]->move(t.x, t.y);
            }
        } else {
            for (int k = 0; k < n_dyn; k++) pre(k);
            for (int k = 0; k < n_dyn; k++) total[(size_t)k] = test(k);
            for (int k = 0; k < n_dyn; k++) w->all[dyn[k]]->move(total[(size_t)k].x, total[(size_t)k].y);
        }
    }
    c.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (out) *out = c;
}

long long orc_num_candidates(void* p) { return (long long)((OrcWorld*)p)->candA.size(); }
long long orc_num_hits(void* p) { return (long long)((OrcWorld*)p)->hitA.size(); }
long long orc_num_contacts(void* p) { return (long long)((OrcWorld*)p)->contacts.size() / 4; }
const uint32_t* orc_cand_a(void* p) { return ((OrcWorld*)p)->candA.data(); }
const uint32_t* orc_cand_b(void* p) { return ((OrcWorld*)p)->candB.data(); }
const uint32_t* orc_cand_rank(void* p) { return ((OrcWorld*)p)->candRank.data(); }
const double* orc_cand_dist(void* p) { return ((OrcWorld*)p)->candDist.data(); }
const uint32_t* orc_hit_a(void* p) { return ((OrcWorld*)p)->hitA.data(); }
const uint32_t* orc_hit_b(void* p) { return ((OrcWorld*)p)->hitB.data(); }
const uint32_t* orc_hit_first(void* p) { return ((OrcWorld*)p)->hitFirst.data(); }
const uint32_t* orc_hit_count(void* p) { return ((OrcWorld*)p)->hitCount.data(); }
const uint32_t* orc_hit_order(void* p) { return ((OrcWorld*)p)->hitOrder.data(); }
const double* orc_hit_mtv(void* p) { return ((OrcWorld*)p)->hitMtv.data(); }
const double* orc_hit_dist(void* p) { return ((OrcWorld*)p)->hitDist.data(); }
const double* orc_contacts(void* p) { return ((OrcWorld*)p)->contacts.data(); }
const double* orc_hit_center(void* p) { return ((OrcWorld*)p)->hitCenter.data(); }

// Batched LineTest (utils.go:276-370) with TestAgainst = space.FilterShapes()[.ByTags][.NotByTags]
// (space.go:106-110: every shape in the Space's list, in insertion order; no broadphase).
// rays = (sx,sy,ex,ey)*n. Results land in the hit* / contacts arrays with hitA = ray index,
// hitOrder = callback index, hitDist = d2(first contact, pulled-back start), hitCenter = set.Center.
// mode 0: TestAgainst = space.FilterShapes() (brute force). mode 1: TestAgainst =
// space.FilterCells(Bounds{componentwise min/max of the cast line's end points}).FilterShapes() (space.go:181-195).
// mode 2: TestAgainst = the explicit ShapeCollection cand[cand_off[r] .. cand_off[r+1]) (shapefilter.go:193-202).
void orc_linetest_ex(void* p, const double* rays, int nrays, const uint8_t* use_any, const uint64_t* any,
                     const uint64_t* none, int record, int threads, OrcFrameCounts* out, int mode,
                     const uint32_t* cand, const uint64_t* cand_off);

void orc_linetest(void* p, const double* rays, int nrays, const uint8_t* use_any, const uint64_t* any,
                  const uint64_t* none, int record, int threads, OrcFrameCounts* out) {
    orc_linetest_ex(p, rays, nrays, use_any, any, none, record, threads, out, 0, nullptr, nullptr);
}

void orc_linetest_ex(void* p, const double* rays, int nrays, const uint8_t* use_any, const uint64_t* any,
                     const uint64_t* none, int record, int threads, OrcFrameCounts* out, int mode,
                     const uint32_t* cand, const uint64_t* cand_off) {
    OrcWorld* w = (OrcWorld*)p;
    if (threads < 1) threads = 1;
    if (record) threads = 1;
    w->hitA.clear(); w->hitB.clear(); w->hitFirst.clear(); w->hitCount.clear(); w->hitOrder.clear();
    w->hitMtv.clear(); w->hitDist.clear(); w->contacts.clear(); w->hitCenter.clear();
    std::vector<OrcFrameCounts> part(threads);
    auto t0 = std::chrono::steady_clock::now();
    auto work = [&](int tid) {
        Scratch sc;
        OrcFrameCounts c{};
        int lo = (int)((long long)nrays * tid / threads), hi = (int)((long long)nrays * (tid + 1) / threads);
        for (int r = lo; r < hi; r++) {
            Vec S{rays[4 * r], rays[4 * r + 1]}, E{rays[4 * r + 2], rays[4 * r + 3]};
            TagFilter f;
            if (use_any && use_any[r]) { f.useAny = true; f.any = any[r]; }
            if (none) f.none = none[r];
            c.testers++;
            Vec vu = vunit(vsub(E, S));
            Vec pulled = vsub(S, vscale(vu, 0.01));
            line_test(
                S, E, nullptr,
                [&](auto fn) {
                    if (mode == 1) {
                        Bounds b;                                           // Space.FilterCells (space.go:181-195)
                        b.min = Vec{pulled.x < E.x ? pulled.x : E.x, pulled.y < E.y ? pulled.y : E.y};
                        b.max = Vec{pulled.x < E.x ? E.x : pulled.x, pulled.y < E.y ? E.y : pulled.y};
                        b.space = &w->space;
                        long long r[4];
                        to_cell_space(b, r);
                        CellSelection sel{r[0], r[1], r[2], r[3], &w->space, nullptr};
                        Scratch walk;
                        filter_for_each(sel, f, walk, [&](Shape* s) { c.candidates++; fn(s); });
                        return;
                    }
                    if (mode == 2) {
                        for (uint64_t k = cand_off[r]; k < cand_off[r + 1]; k++) {
                            Shape* s = w->all[cand[k]];
                            if (!f.pass(s->tags)) continue;
                            c.candidates++;
                            fn(s);
                        }
                        return;
                    }
                    for (Shape* s : w->space.shapes) {                      // shapefilter.go:196-202
                        if (!f.pass(s->tags)) continue;
                        c.candidates++;
                        fn(s);
                    }
                },
                sc,
                [&](const IntersectionSet& set, int idx, int) {
                    c.hits++;
                    c.contacts += (long long)set.intersections.size();
                    c.mtv_sum_x += set.mtv.x; c.mtv_sum_y += set.mtv.y;
                    if (record) {
                        w->hitA.push_back((uint32_t)r); w->hitB.push_back(set.other->id);
                        w->hitOrder.push_back((uint32_t)idx);
                        w->hitFirst.push_back((uint32_t)(w->contacts.size() / 4));
                        w->hitCount.push_back((uint32_t)set.intersections.size());
                        w->hitMtv.push_back(set.mtv.x); w->hitMtv.push_back(set.mtv.y);
                        w->hitDist.push_back(vdist2(set.intersections[0].point, pulled));
                        w->hitCenter.push_back(set.center.x); w->hitCenter.push_back(set.center.y);
                        for (const auto& it : set.intersections) {
                            w->contacts.push_back(it.point.x); w->contacts.push_back(it.point.y);
                            w->contacts.push_back(it.normal.x); w->contacts.push_back(it.normal.y);
                        }
                    }
                    return true;
                });
        }
        part[tid] = c;
    };
    if (threads == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < threads; t++) th.emplace_back(work, t);
        for (auto& t : th) t.join();
    }
    auto t1 = std::chrono::steady_clock::now();
    OrcFrameCounts tot{};
    for (auto& c : part) {
        tot.testers += c.testers; tot.candidates += c.candidates; tot.hits += c.hits; tot.contacts += c.contacts;
        tot.mtv_sum_x += c.mtv_sum_x; tot.mtv_sum_y += c.mtv_sum_y;
    }
    tot.seconds = std::chrono::duration<double>(t1 - t0).count();
    if (out) *out = tot;
}

// ConvexPolygon.ShapeLineTest (convexPolygon.go:322-415) with TestAgainst =
// self.SelectTouchingCells(margin).FilterShapes()[.ByTags][.NotByTags] (examples/worldPlatformer.go:210-246).
// The reference iterates its vertex set in Go-map (random) order (convexPolygon.go:366); here the
// vertices are visited in insertion order, so only set-level parity of merged contacts is meaningful.
// settings.Lines is not modelled (empty == all lines). Results in hit* arrays (hitA = shape index).
int orc_shape_linetest(void* p, int shape, double vx, double vy, double offx, double offy, int include_all,
                       int margin, int use_any, uint64_t any, uint64_t none) {
    OrcWorld* w = (OrcWorld*)p;
    Shape& cp = *w->all[shape];
    w->hitA.clear(); w->hitB.clear(); w->hitFirst.clear(); w->hitCount.clear(); w->hitOrder.clear();
    w->hitMtv.clear(); w->hitDist.clear(); w->contacts.clear(); w->hitCenter.clear();
    Scratch sc, scWalk;
    std::vector<IntersectionSet> results;
    std::vector<Vec> verts;
    auto addv = [&](Vec v) {
        for (const Vec& o : verts) if (o.x == v.x && o.y == v.y) return;  // Set[Vector] semantics (utils.go:178-193)
        verts.push_back(v);
    };
    Vec V{vx, vy}, off{offx, offy};
    if (!include_all) {
        for (const Line& l : cp.lines()) {
            if (vdot(l.normal(), V) < 0.01) continue;                       // convexPolygon.go:347
            Vec v = vinv(vscale(l.vector(), 0.5));                          // convexPolygon.go:352
            addv(vsub(l.start, v));
            addv(vsub(l.end, vinv(v)));
        }
    } else {
        for (const Vec& v : cp.transformed()) addv(v);
    }
    Vec vu = vunit(V);
    TagFilter f; f.useAny = use_any != 0; f.any = any; f.none = none;
    CellSelection sel = select_touching_cells(cp, margin);
    for (const Vec& pt : verts) {
        Vec start = vsub(pt, vadd(vu, off));                                // convexPolygon.go:368
        line_test(
            start, vadd(pt, V), &cp,
            [&](auto fn) { filter_for_each(sel, f, scWalk, [&](Shape* s) { fn(s); }); }, sc,
            [&](const IntersectionSet& set, int, int) {
                for (auto& r : results)
                    if (r.other == set.other) {
                        r.intersections.insert(r.intersections.end(), set.intersections.begin(), set.intersections.end());
                        if (vmag2(set.mtv) < vmag2(r.mtv)) r.mtv = set.mtv;
                        return true;
                    }
                results.push_back(set);
                return true;
            });
    }
    go_sort_slice(results, [](const IntersectionSet& a, const IntersectionSet& b) { return vmag2(a.mtv) < vmag2(b.mtv); });
    for (size_t i = 0; i < results.size(); i++) {
        const IntersectionSet& set = results[i];
        w->hitA.push_back((uint32_t)shape); w->hitB.push_back(set.other->id);
        w->hitOrder.push_back((uint32_t)i);
        w->hitFirst.push_back((uint32_t)(w->contacts.size() / 4));
        w->hitCount.push_back((uint32_t)set.intersections.size());
        w->hitMtv.push_back(set.mtv.x); w->hitMtv.push_back(set.mtv.y);
        w->hitDist.push_back(vmag2(set.mtv));
        w->hitCenter.push_back(set.center.x); w->hitCenter.push_back(set.center.y);
        for (const auto& it : set.intersections) {
            w->contacts.push_back(it.point.x); w->contacts.push_back(it.point.y);
            w->contacts.push_back(it.normal.x); w->contacts.push_back(it.normal.y);
        }
    }
    return (int)results.size();
}

// sort.Slice restatement exposed for unit tests: sorts keys ascending, returns the permutation.
void orc_go_sort_f64(const double* keys, int n, int* perm) {
    struct KV { double k; int i; };
    std::vector<KV> v(n);
    for (int i = 0; i < n; i++) v[i] = KV{keys[i], i};
    go_sort_slice(v, [](const KV& a, const KV& b) { return a.k < b.k; });
    for (int i = 0; i < n; i++) perm[i] = v[i].i;
}

int orc_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
