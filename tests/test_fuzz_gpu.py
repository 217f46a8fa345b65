"""Randomised / adversarial parity sweep (GPU path vs oracle, through the C ABI): lattice-aligned shapes that touch
exactly, share edges and corners, tie in distance and hit every strict/non-strict comparison of the reference
(Bounds gate incl. the IsEmpty typo utils.go:171-173, det == 0 in convexPolygon.go:717-719, 0<lambda<1, d == 0 and
d == |ra-rb| in shape.go:409, l == 1 in vector.go:169), degenerate polygons, shapes far outside the grid."""
import numpy as np
import pytest

from resolv_b200 import space_for_world, worlds
from tests.util import OracleBatch, compare_frame, device_cells

pytestmark = pytest.mark.gpu


def _lattice_world(rs, n, size, cell, step):
    rows = []
    for _ in range(n):
        x = float(rs.integers(-2, size // step + 3) * step)
        y = float(rs.integers(-2, size // step + 3) * step)
        kind = rs.integers(0, 5)
        tag = int(1 << rs.integers(0, 3))
        if kind == 0:
            rows.append(dict(kind=0, pos=(x, y), radius=float(rs.integers(0, 4) * step / 2), tags=tag))   # radius 0 allowed
        elif kind == 1:
            w, h = float(rs.integers(0, 4) * step), float(rs.integers(0, 4) * step)                      # zero-size allowed
            rows.append(dict(kind=1, pos=(x, y), radius=0.0, verts=worlds.rect_verts(w, h), tags=tag))
        elif kind == 2:   # NewLine-style open 2-gon, axis-aligned or diagonal
            dx, dy = float(rs.integers(-3, 4) * step / 2), float(rs.integers(-3, 4) * step / 2)
            rows.append(dict(kind=1, pos=(x, y), radius=0.0, verts=np.array([[-dx, -dy], [dx, dy]]), tags=tag))
        elif kind == 3:   # right triangle on the lattice
            a, b = float(rs.integers(1, 4) * step), float(rs.integers(1, 4) * step)
            rows.append(dict(kind=1, pos=(x, y), radius=0.0, verts=np.array([[0.0, 0.0], [a, 0.0], [0.0, b]]), tags=tag))
        else:             # open polyline of 3 points
            rows.append(dict(kind=1, pos=(x, y), radius=0.0, closed=0,
                             verts=np.array([[-step, 0.0], [0.0, float(rs.integers(-2, 3) * step)], [step, 0.0]]), tags=tag))
    return worlds._from_rows(rows, (size, size), (cell, cell), "lattice").finalize()


@pytest.mark.parametrize("seed", range(6))
def test_lattice_worlds(seed):
    rs = np.random.default_rng(1000 + seed)
    w = _lattice_world(rs, 400, 256, 16, 8)
    if seed % 2:
        w.use_any[:] = rs.integers(0, 2, w.n)
        w.any_mask[:] = rs.integers(0, 8, w.n)
        w.none_mask[:] = rs.integers(0, 8, w.n) & rs.integers(0, 8, w.n)
    ob = OracleBatch(w)
    ds = space_for_world(w)
    try:
        for margin in (0, 1, 2):
            c, exact = compare_frame(ds, ob, margin)
            assert exact
        assert device_cells(ds, c.n_entries) == ob.cells()
        assert c.n_hits > 50
    finally:
        ds.close()


def test_many_tiny_random_worlds():
    rs = np.random.default_rng(7)
    total_hits = 0
    for trial in range(40):
        n = int(rs.integers(2, 60))
        size = int(rs.choice([64, 100, 257]))
        cell = int(rs.choice([7, 16, 33]))
        w = worlds.make_mixed(n, seed=int(rs.integers(1, 1 << 30)), space=size, cell=cell, filtered=bool(trial % 2))
        w.pos = w.pos * 1.4 - 0.2 * size          # part of the shapes off-grid
        w.pos0 = w.pos.copy()
        w.tester[:] = rs.integers(0, 2, n) | (rs.random(n) < 0.5)
        w.finalize()
        ob = OracleBatch(w)
        ds = space_for_world(w)
        try:
            c, exact = compare_frame(ds, ob, int(rs.integers(0, 4)))
            assert exact
            total_hits += c.n_hits
        finally:
            ds.close()
    assert total_hits > 100


def test_coincident_and_identical_shapes():
    # many copies of the same shape at the same place: d == 0 for circles (no hit, shape.go:409), all-edges-parallel
    # polygons (det == 0 everywhere -> no contacts), maximal in-cell occupancy (in-cell sort with many entries)
    rows = [dict(kind=0, pos=(50.0, 50.0), radius=10.0, tags=1) for _ in range(40)]
    rows += [dict(kind=1, pos=(120.0, 50.0), radius=0.0, verts=worlds.rect_verts(20.0, 20.0), tags=1) for _ in range(40)]
    rows += [dict(kind=0, pos=(50.0 + i, 120.0), radius=10.0, tags=1) for i in range(40)]
    w = worlds._from_rows(rows, (256, 256), (32, 32), "coincident").finalize()
    ob = OracleBatch(w)
    ds = space_for_world(w)
    try:
        c, exact = compare_frame(ds, ob, 1)
        assert exact and c.n_candidates >= 3 * 40 * 39
    finally:
        ds.close()
