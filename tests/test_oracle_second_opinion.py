"""Two independent restatements of the reference must agree bit for bit.

oracle/resolv_oracle.cpp (C++, mirrors the reference's data structures) and oracle/second_opinion.py (plain Python,
transcribed separately from the same Go sources) are run on the same seeded shapes. The reference cannot run here, so
this is the strongest cross-check available short of parity/go: a transcription slip in either restatement (operand
order, a swapped branch, a missing Unit(), the IsEmpty typo, Go's math.Min) breaks bit equality.
"""
import struct

import numpy as np
import pytest

from oracle import second_opinion as so
from oracle.oracle import OracleSpace


def bits(x):
    return struct.pack("<d", float(x))


def same(x, y):
    return bits(x) == bits(y) or (x != x and y != y)


def random_polygon(rng, cx, cy, nv, rad):
    # clockwise in screen coordinates like convexPolygon.go:23; irregular angles and radii that stay convex
    ang = np.sort(rng.uniform(0, 2 * np.pi, nv))
    return (cx, cy, [c for a in ang for c in (rad * np.cos(a), rad * np.sin(a))])


def make_shape(rng, kind, where, size):
    cx, cy = rng.uniform(*where, size=2)
    if kind == "circle":
        return ("circle", (cx, cy, rng.uniform(*size)))
    if kind == "rect":
        return ("rect", (cx, cy, rng.uniform(*size) * 2, rng.uniform(*size) * 2))
    if kind == "line":
        dx, dy = rng.uniform(-2 * size[1], 2 * size[1], size=2)
        return ("line", (cx, cy, cx + dx, cy + dy))
    if kind == "polyline":
        x, y, pts = random_polygon(rng, cx, cy, int(rng.integers(3, 7)), rng.uniform(*size))
        return ("polyline", (x, y, pts))
    x, y, pts = random_polygon(rng, cx, cy, int(rng.integers(3, 9)), rng.uniform(*size))
    return ("poly", (x, y, pts))


def add_both(space, spec):
    kind, a = spec
    if kind == "circle":
        return space.add_circle(*a), so.Circle(*a)
    if kind == "rect":
        return space.add_rectangle(*a), so.rectangle(*a)
    if kind == "line":
        s = so.line_shape(*a)
        return space.add_line(*a), s
    if kind == "polyline":
        return space.add_polygon(a[0], a[1], a[2], closed=False), so.Polygon(a[0], a[1], a[2], closed=False)
    return space.add_polygon(a[0], a[1], a[2]), so.Polygon(a[0], a[1], a[2])


def check_pair(space, ia, ib, sa, sb, stats):
    mtv, con = space.intersection(ia, ib)
    mtv2, con2, ordered = so.intersection(sa, sb)
    assert len(con) == len(con2), (ia, ib)
    assert same(mtv[0], mtv2[0]) and same(mtv[1], mtv2[1]), (ia, ib, mtv, mtv2)
    rows = [tuple(bits(v) for v in (p[0], p[1], n[0], n[1])) for p, n in con2]
    orows = [tuple(bits(v) for v in r) for r in con]
    if ordered:
        assert rows == orows, (ia, ib)
    else:
        assert sorted(rows) == sorted(orows), (ia, ib)
    stats["pairs"] += 1
    stats["hits"] += len(con) > 0
    stats["mtv"] += bool(mtv[0] != 0 or mtv[1] != 0)


@pytest.mark.parametrize("seed", range(4))
def test_pair_tests_agree_bit_for_bit(seed):
    rng = np.random.default_rng(8800 + seed)
    space = OracleSpace(640, 360, 16, 16)
    kinds = ["circle", "rect", "poly", "poly", "line", "polyline"]
    ids, shapes = [], []
    for i in range(90):
        i1, s1 = add_both(space, make_shape(rng, kinds[i % len(kinds)], (100, 180), (4, 16)))
        ids.append(i1)
        shapes.append(s1)
    # move a third of them so the cached local bounds (computed at the creation position) differ from fresh ones
    for k in range(0, len(ids), 3):
        x, y = rng.uniform(100, 180, size=2)
        space.set_position(ids[k], x, y)
        shapes[k].pos = (float(x), float(y))
    stats = dict(pairs=0, hits=0, mtv=0)
    for a in range(len(ids)):
        for b in range(len(ids)):
            if a != b:
                check_pair(space, ids[a], ids[b], shapes[a], shapes[b], stats)
    # the comparison has to bite: plenty of non-empty sets and non-zero MTVs among them
    assert stats["hits"] > 500 and stats["mtv"] > 400, stats


def test_exact_and_degenerate_configurations_agree():
    """Integer-aligned rectangles (ties, zero overlaps, shared edges and corners), concentric and tangent circles,
    identical shapes, the IsEmpty typo window (utils.go:171-173)."""
    space = OracleSpace(640, 360, 16, 16)
    specs = []
    for x in (0.0, 8.0, 16.0, 24.0):
        for y in (0.0, 8.0, 16.0):
            specs.append(("rect", (x, y, 16.0, 16.0)))
    specs += [("circle", (8.0, 8.0, 8.0)), ("circle", (8.0, 8.0, 8.0)), ("circle", (8.0, 8.0, 4.0)), ("circle", (24.0, 8.0, 8.0)),
              ("circle", (16.0, 16.0, 0.0)), ("rect", (8.0, 8.0, 0.0, 0.0)), ("line", (0.0, 0.0, 32.0, 0.0)),
              ("line", (0.0, 8.0, 32.0, 8.0)), ("line", (8.0, -8.0, 8.0, 40.0)), ("rect", (-4.0, 4.0, 8.0, 8.0)),
              ("rect", (4.0, 4.0, 8.0, 0.0))]
    both = [add_both(space, s) for s in specs]
    stats = dict(pairs=0, hits=0, mtv=0)
    for a, (ia, sa) in enumerate(both):
        for b, (ib, sb) in enumerate(both):
            if a != b:
                check_pair(space, ia, ib, sa, sb, stats)
    assert stats["hits"] > 100, stats


def test_linetest_sets_agree_bit_for_bit():
    rng = np.random.default_rng(8899)
    space = OracleSpace(640, 360, 16, 16)
    kinds = ["circle", "rect", "poly", "line", "polyline"]
    both = [add_both(space, make_shape(rng, kinds[i % len(kinds)], (60, 260), (4, 24))) for i in range(60)]
    rays = np.concatenate([rng.uniform(40, 280, size=(300, 2)), rng.uniform(40, 280, size=(300, 2))], axis=1)
    rays[:20, 3] = rays[:20, 1]            # horizontal
    rays[20:40, 2] = rays[20:40, 0]        # vertical
    rays[40:44, 2:] = rays[40:44, :2]      # zero length: Unit() of (0,0) stays (0,0)
    _, hits = space.linetest(rays, mode="all")
    got = {}
    for k in range(len(hits.a)):
        f, c = int(hits.first[k]), int(hits.count[k])
        got[(int(hits.a[k]), int(hits.b[k]))] = (hits.mtv[k], hits.center[k], hits.contacts[f:f + c], hits.dist[k])
    n_sets = 0
    for r, ray in enumerate(rays):
        for ib, sb in both:
            exp = so.linetest_one((ray[0], ray[1]), (ray[2], ray[3]), sb)
            have = got.get((r, ib))
            assert (exp is None) == (have is None), (r, ib)
            if exp is None:
                continue
            mtv, center, cons, key = exp
            assert same(mtv[0], have[0][0]) and same(mtv[1], have[0][1])
            assert same(center[0], have[1][0]) and same(center[1], have[1][1])
            assert len(cons) == len(have[2])
            for (p, n), row in zip(cons, have[2]):
                assert [bits(v) for v in (p[0], p[1], n[0], n[1])] == [bits(v) for v in row]
            assert same(key, have[3])
            n_sets += 1
    assert n_sets > 300, n_sets


def _world(seed, n=260):
    rng = np.random.default_rng(seed)
    space = OracleSpace(320, 240, 16, 16)
    sp2 = so.Space(320, 240, 16, 16)
    kinds = ["circle", "rect", "poly", "circle", "rect", "line", "polyline"]
    shapes = []
    for i in range(n):
        # some shapes straddle or leave the grid (Space.Cell returns nil there), some span many cells
        size = (2, 10) if i % 11 else (20, 45)
        spec = make_shape(rng, kinds[i % len(kinds)], (-20, 340), size)
        tag = 1 << int(rng.integers(0, 4))
        kind, a = spec
        if kind == "circle":
            ia = space.add_circle(*a, tags=tag)
        elif kind == "rect":
            ia = space.add_rectangle(*a, tags=tag)
        elif kind == "line":
            ia = space.add_line(*a, tags=tag)
        else:
            ia = space.add_polygon(a[0], a[1], a[2], closed=(kind != "polyline"), tags=tag)
        _, s2 = add_both(OracleSpace(16, 16, 16, 16), spec)
        assert sp2.add(s2, tags=tag) == ia
        shapes.append(s2)
    return rng, space, sp2, shapes


@pytest.mark.parametrize("seed,margin", [(8811, 1), (8812, 0), (8813, 2)])
def test_broadphase_and_intersection_test_agree_after_motion(seed, margin):
    """Cell membership, candidate generation order (which depends on the swap-remove history of every cell), distance
    keys, tag filters, callback order and every IntersectionSet — over three frames of motion."""
    rng, space, sp2, shapes = _world(seed)
    n = len(shapes)
    use_any = (rng.integers(0, 2, n)).astype(np.uint8)
    any_mask = rng.integers(1, 16, n).astype(np.uint64)
    none_mask = np.where(rng.integers(0, 4, n) == 0, 1 << rng.integers(0, 4, n), 0).astype(np.uint64)
    total_c = total_h = 0
    for frame in range(3):
        if frame:
            for i in rng.permutation(n)[: n // 2]:                  # sequential moves, in a scrambled order
                x, y = np.array(shapes[i].pos) + rng.uniform(-12, 12, size=2)
                space.set_position(int(i), x, y)
                sp2.set_position(shapes[i], x, y)
        fc, cands, hits = space.frame(margin=margin, use_any=use_any, any_mask=any_mask, none_mask=none_mask)
        k = h = 0
        for i, s in enumerate(shapes):
            flt = dict(use_any=bool(use_any[i]), any_mask=int(any_mask[i]), none_mask=int(none_mask[i]))
            mine = sp2.candidates(s, margin, **flt)
            for rank, (o, d) in enumerate(mine):
                assert (int(cands["a"][k]), int(cands["b"][k]), int(cands["rank"][k])) == (i, o.id, rank), (frame, i, rank)
                assert same(cands["dist"][k], d)
                k += 1
            sets, defined = sp2.intersection_test(s, margin, **flt)
            rows = []
            while h < len(hits.a) and int(hits.a[h]) == i:
                f, c = int(hits.first[h]), int(hits.count[h])
                rows.append((int(hits.b[h]), bits(hits.dist[h]), bits(hits.mtv[h][0]), bits(hits.mtv[h][1]),
                             [tuple(bits(v) for v in r) for r in hits.contacts[f:f + c]]))
                h += 1
            exp = [(o, bits(d), bits(m[0]), bits(m[1]), [tuple(bits(v) for v in (p[0], p[1], q[0], q[1])) for p, q in cons])
                   for o, d, m, cons, ordered in sets]
            if any(not ordered for *_, ordered in sets):
                rows = [(a, b, c, d, sorted(e)) for a, b, c, d, e in rows]
                exp = [(a, b, c, d, sorted(e)) for a, b, c, d, e in exp]
            if defined:
                assert rows == exp, (frame, i)
            else:
                assert sorted(rows) == sorted(exp), (frame, i)
        assert k == len(cands["a"]) and h == len(hits.a)
        total_c += k
        total_h += h
    assert total_c > 1000 and total_h > 200, (total_c, total_h)


def test_shape_linetest_agrees_bit_for_bit():
    """ConvexPolygon.ShapeLineTest (convexPolygon.go:322-415): leading-edge vertices, per-OtherShape merge, smallest
    |MTV| first — what the platformer calls (examples/worldPlatformer.go:221-246)."""
    rng, space, sp2, shapes = _world(8821, n=220)
    n_sets = n_calls = 0
    for i, s in enumerate(shapes):
        if not isinstance(s, so.Polygon):
            continue
        for vec, inc in (((0.0, 14.0), False), ((9.0, -5.0), False), ((-6.0, 0.0), True)):
            off = (0.0, 0.0) if i % 2 else (0.5, -0.25)
            use_any, any_mask = bool(i % 3 == 0), 0b0110
            got = space.shape_linetest(i, vec, start_offset=off, include_all=inc, margin=2, use_any=use_any, any_mask=any_mask)
            exp = so.shape_linetest(sp2, s, vec, start_offset=off, include_all=inc, margin=2, use_any=use_any, any_mask=any_mask)
            assert len(exp) == len(got.b), (i, vec)
            for k, (other, mtv, center, cons) in enumerate(exp):
                f, c = int(got.first[k]), int(got.count[k])
                assert int(got.b[k]) == other and same(got.mtv[k][0], mtv[0]) and same(got.mtv[k][1], mtv[1])
                assert same(got.center[k][0], center[0]) and same(got.center[k][1], center[1])
                assert [tuple(bits(v) for v in r) for r in got.contacts[f:f + c]] == \
                       [tuple(bits(v) for v in (p[0], p[1], q[0], q[1])) for p, q in cons]
            n_sets += len(exp)
            n_calls += 1
    assert n_calls > 200 and n_sets > 150, (n_calls, n_sets)


def test_the_two_restatements_of_gos_pdqsort_agree():
    """sort.Slice is pdqsort_func (unstable beyond 12 items): the C++ oracle and second_opinion.py each restate it from
    the published algorithm; they must produce the same permutation, ties included, for every input shape that steers
    it (random, few distinct keys, sorted, reversed, half sorted, constant; 0-400 items: insertion sort, ninther pivots,
    partialInsertionSort, partitionEqual, breakPatterns)."""
    from oracle.oracle import go_sort
    rng = np.random.default_rng(8831)
    beyond = 0
    for trial in range(1200):
        n = int(rng.integers(0, 400))
        mode = trial % 6
        if mode == 0:
            keys = rng.integers(0, 5, n).astype(float)
        elif mode == 1:
            keys = rng.random(n)
        elif mode == 2:
            keys = np.sort(rng.integers(0, 50, n)).astype(float)
        elif mode == 3:
            keys = np.sort(rng.random(n))[::-1].copy()
        elif mode == 4:
            keys = np.concatenate([np.sort(rng.random(n // 2)), rng.integers(0, 3, n - n // 2).astype(float)])
        else:
            keys = np.zeros(n)
        items = list(range(n))
        so.go_sort_slice(items, lambda i, j: keys[i] < keys[j])
        assert items == [int(x) for x in go_sort(keys)], (trial, n, mode)
        assert all(keys[items[i]] <= keys[items[i + 1]] for i in range(n - 1))
        beyond += n > 12
    assert beyond > 1000


def test_lattice_shapes_agree_bit_for_bit():
    """Small-integer coordinates: exact tangencies (the det == 0 branch of IntersectionPointsCircle), coincident and
    collinear edges (det == 0 in IntersectionPointsLine), unit-length axes (the l == 1 exit of Vector.Unit), zero
    overlaps, distance ties in every sort, shapes sharing vertices."""
    rng = np.random.default_rng(8841)
    space = OracleSpace(64, 64, 8, 8)
    both = []
    for i in range(70):
        kind = i % 5
        cx, cy = (float(v) for v in rng.integers(8, 40, size=2))
        if kind == 0:
            spec = ("circle", (cx, cy, float(rng.integers(1, 7))))
        elif kind == 1:
            spec = ("rect", (cx, cy, float(2 * rng.integers(1, 6)), float(2 * rng.integers(1, 6))))
        elif kind == 2:                                   # lattice triangle / quad, clockwise on screen (y down)
            r = float(rng.integers(2, 7))
            pts = [-r, -r, r, -r, r, r] if i % 2 else [0.0, -r, r, 0.0, 0.0, r, -r, 0.0]
            spec = ("poly", (cx, cy, pts))
        elif kind == 3:
            dx, dy = (float(v) for v in rng.integers(-8, 9, size=2))
            spec = ("line", (cx, cy, cx + dx, cy + dy))   # includes zero-length lines
        else:
            r = float(rng.integers(2, 6))
            spec = ("polyline", (cx, cy, [-r, 0.0, 0.0, -r, r, 0.0, r, r]))
        both.append(add_both(space, spec))
    stats = dict(pairs=0, hits=0, mtv=0)
    for a, (ia, sa) in enumerate(both):
        for b, (ib, sb) in enumerate(both):
            if a != b:
                check_pair(space, ia, ib, sa, sb, stats)
    assert stats["hits"] > 300 and stats["hits"] - stats["mtv"] > 20, stats   # (sets with a zero MTV: calculateMTV said "no overlap")
    # rays along lattice lines: through vertices, along edges, tangent to circles
    rays = np.array([[float(a), float(b), float(c), float(d)] for a, b, c, d in rng.integers(0, 48, size=(250, 4))])
    _, hits = space.linetest(rays, mode="all")
    got = {(int(hits.a[k]), int(hits.b[k])): k for k in range(len(hits.a))}
    n_sets = 0
    for r, ray in enumerate(rays):
        for ib, sb in both:
            exp = so.linetest_one((ray[0], ray[1]), (ray[2], ray[3]), sb)
            k = got.get((r, ib))
            assert (exp is None) == (k is None), (r, ib)
            if exp is None:
                continue
            mtv, center, cons, key = exp
            f, c = int(hits.first[k]), int(hits.count[k])
            assert same(mtv[0], hits.mtv[k][0]) and same(mtv[1], hits.mtv[k][1]) and same(key, hits.dist[k])
            assert same(center[0], hits.center[k][0]) and same(center[1], hits.center[k][1])
            assert [tuple(bits(v) for v in row) for row in hits.contacts[f:f + c]] == \
                   [tuple(bits(v) for v in (p[0], p[1], q[0], q[1])) for p, q in cons]
            n_sets += 1
    assert n_sets > 300, n_sets
