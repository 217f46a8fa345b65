"""Pins the CPU oracle against the hand-derived known-answer tests of SURVEY.md §8c (KAT1-6).

The reference ships no tests or golden vectors, so these KATs — computed in the survey by
transcribing the cited reference formulas into IEEE-double arithmetic — are the only external pin the oracle has
("parity unpinned", see oracle/resolv_oracle.cpp).
"""
import numpy as np
import pytest

from oracle.oracle import OracleSpace, go_sort

TOL = 1e-12


def _space():
    return OracleSpace(640, 360, 16, 16)


def test_kat1_rect_circle_readme():
    # readme.md:45-49; shape.go:356-397 (convex->circle), convexPolygon.go:747-789, 418-514
    s = _space()
    r = s.add_rectangle(200, 100, 32, 32)
    c = s.add_circle(200, 120, 8)
    mtv, con = s.intersection(r, c)
    assert mtv.tolist() == [0.0, -4.0]
    exp = np.array([[193.0717967697245, 116, -0.8660254037844384, -0.5000000000000006],
                    [206.9282032302755, 116, 0.8660254037844384, -0.5000000000000006]])
    assert con.shape == (2, 4) and np.abs(con - exp).max() < TOL
    # reversed: shape.go:318-354 — unsorted, edge normals, MTV = Invert(calculateMTV)
    mtv, con = s.intersection(c, r)
    assert mtv[0] == 0.0 and np.signbit(mtv[0]) and mtv[1] == 4.0  # (-0, 4): Invert of (0,-4)
    exp = np.array([[193.0717967697245, 116, 0, 1], [206.9282032302755, 116, 0, 1]])
    assert np.abs(con - exp).max() < TOL


def test_kat2_rect_rect_plus_one_fudge():
    # shape.go:438-479; convexPolygon.go:715-744 (the "+1" moves the reported points)
    s = _space()
    a = s.add_rectangle(0, 0, 10, 10)
    b = s.add_rectangle(8, 2, 10, 10)
    mtv, con = s.intersection(a, b)
    assert mtv.tolist() == [-2.0, 0.0]
    assert np.abs(con - np.array([[2.9, 5, -1, -0.0], [5, -3.1, 0, -1]])).max() < TOL
    mtv, con = s.intersection(b, a)
    assert mtv[0] == 2.0 and mtv[1] == 0.0 and np.signbit(mtv[1])
    assert np.abs(con - np.array([[5.1, -3, 1, -0.0], [3, 5.1, 0, 1]])).max() < TOL


def test_kat3_circle_circle():
    # shape.go:399-436
    s = _space()
    a = s.add_circle(0, 0, 5)
    b = s.add_circle(6, 0, 3)
    mtv, con = s.intersection(a, b)
    assert mtv.tolist() == [-2.0, 0.0]
    exp = np.array([[4.333333333333333, -2.494438257849295, 0.8666666666666666, -0.498887651569859],
                    [4.333333333333333, 2.494438257849295, 0.8666666666666666, 0.498887651569859]])
    assert np.abs(con - exp).max() < TOL


def test_kat4_containment_is_not_an_intersection():
    # shape.go:446-461: no edge crossing -> empty set, both orders
    s = _space()
    a = s.add_rectangle(100, 100, 4, 4)
    b = s.add_rectangle(100, 100, 20, 20)
    for x, y in ((a, b), (b, a)):
        mtv, con = s.intersection(x, y)
        assert len(con) == 0 and mtv.tolist() == [0.0, 0.0]
    # circle inside circle: d < |ra-rb| (shape.go:409)
    c1 = s.add_circle(300, 100, 10)
    c2 = s.add_circle(301, 100, 2)
    assert len(s.intersection(c1, c2)[1]) == 0 and len(s.intersection(c2, c1)[1]) == 0


def test_kat5_linetest():
    # utils.go:276-370
    s = _space()
    s.add_rectangle(10, 5, 4, 4)
    fc, h = s.linetest([[0, 5, 20, 5]])
    assert fc.hits == 1 and h.count.tolist() == [2]
    assert np.abs(h.contacts - np.array([[7.75, 5, -1, -0.0], [12.250000000000002, 5, 1, -0.0]])).max() < TOL
    assert np.abs(h.mtv - np.array([[7.74, 0.0]])).max() < TOL
    assert np.abs(h.center - np.array([[10.0, 5.0]])).max() < TOL


def test_kat6_grid_dims_and_wall_registration():
    # space.go:16-30; convexPolygon.go:638-643; shape.go:191-214; space.go:135-142 (OOB column skipped)
    s = _space()
    assert (s.width_cells, s.height_cells) == (40, 23)
    w = s.add_rectangle_top_left(0, 0, 640, 16)
    assert s.bounds(w).tolist() == [0, 0, 640, 16]
    assert s.cell_rect(w).tolist() == [0, 0, 40, 1]
    counts, entries = s.cells()
    grid = counts.reshape(23, 40)
    assert grid[:2].sum() == 80 and grid[2:].sum() == 0 and len(entries) == 80


def test_bounds_is_empty_typo_quirk():
    # utils.go:171-173: "empty" iff width == 0 and Max.Y - Min.X == 0. Two rects touching along x = 10
    # (zero-width overlap box) are NOT rejected unless ovMaxY == ovMinX.
    s = _space()
    a = s.add_rectangle(5, 5, 10, 10)      # [0,10]x[0,10]
    b = s.add_rectangle(15, 5, 10, 10)     # [10,20]x[0,10]: overlap box x:[10,10], y:[0,10] -> MaxY-MinX = 0 -> "empty"
    assert len(s.intersection(a, b)[1]) == 0
    c = s.add_rectangle(15, 6, 10, 10)     # [10,20]x[1,11]: overlap y:[1,10] -> MaxY-MinX = 0 again (10-10)
    assert len(s.intersection(a, c)[1]) == 0


def test_unregister_swap_remove_and_move():
    # cell.go:26-38 + shape.go:239-242
    s = _space()
    ids = [s.add_circle(24, 24, 4) for _ in range(3)]  # all in cell (1,1)
    counts, entries = s.cells()
    assert counts.reshape(23, 40)[1, 1] == 3 and entries.tolist() == ids
    s.set_position(ids[0], 100, 100)  # leaves cell (1,1): last swapped into slot 0
    counts, entries = s.cells()
    assert entries.tolist()[:2] == [ids[2], ids[1]]


def test_frame_candidates_sorted_and_self_excluded():
    s = _space()
    a = s.add_circle(100, 100, 8)
    b = s.add_circle(110, 100, 8)
    c = s.add_circle(104, 100, 8)
    fc, cand, hits = s.frame(margin=1)
    assert fc.testers == 3 and fc.candidates == 6
    # tester a sees b (d2=100) and c (d2=16); callbacks arrive nearest first (shape.go:291-293)
    sel = hits.a == a
    assert hits.b[sel].tolist() == [c, b] and hits.order[sel].tolist() == [0, 1]


def test_tag_filters_any_bit_semantics():
    # tags.go:29-31 Has == any bit; shapefilter.go:47-61
    s = _space()
    a = s.add_circle(100, 100, 8, tags=1)
    b = s.add_circle(104, 100, 8, tags=2)
    c = s.add_circle(96, 100, 8, tags=4 | 2)
    n = 3
    use_any = np.ones(n, np.uint8)
    anym = np.full(n, 2, np.uint64)
    fc, cand, hits = s.frame(1, use_any=use_any, any_mask=anym)
    assert sorted(cand["b"][cand["a"] == a].tolist()) == [b, c]
    fc, cand, hits = s.frame(1, use_any=use_any, any_mask=anym, none_mask=np.full(n, 4, np.uint64))
    assert cand["b"][cand["a"] == a].tolist() == [b]


@pytest.mark.parametrize("n", [0, 1, 5, 12, 13, 40, 200])
def test_go_sort_is_a_sort_and_stable_when_small(n):
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 5, size=n).astype(np.float64)  # many ties
    perm = go_sort(keys)
    assert sorted(perm.tolist()) == list(range(n))
    assert np.all(np.diff(keys[perm]) >= 0)
    if n <= 12:  # insertion sort regime: stable
        assert perm.tolist() == np.argsort(keys, kind="stable").tolist()
