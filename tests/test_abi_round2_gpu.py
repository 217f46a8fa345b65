"""Entry points added in round 2, through the C ABI, against the oracle / against the plain path:
rsv_test_pairs (explicit ordered pairs), device position pages + the upload stream, the sparse position update, and
two frames in flight on different position pages."""
import numpy as np
import pytest

from resolv_b200 import _abi, space_for_world, worlds
from tests.util import OracleBatch, bits

pytestmark = pytest.mark.gpu


def _world(n=6000, space=2400):
    return worlds.make_mixed(n, space=space, cell=32, filtered=False)


def test_explicit_pairs_equal_the_oracles_intersection():
    """rsv_test_pairs == A.Intersection(B) (circle.go:69-82, convexPolygon.go:288-300) for arbitrary ordered pairs —
    neighbours, far-apart shapes, a shape with itself — against the positions committed last, not the last frame's."""
    w = _world()
    ds = space_for_world(w)
    sp = OracleBatch(w).spaces[0]
    try:
        c = ds.frame(1, _abi.FRAME_DOWNLOAD, grow=True)
        hits, _ = ds.results()
        near = np.stack([hits["a"], hits["b"]], -1)[:3000].astype(np.uint32)          # gated pairs of the frame
        rs = np.random.default_rng(5)
        far = rs.integers(0, w.n, (2000, 2)).astype(np.uint32)                          # mostly disjoint, some a == b
        pairs = np.concatenate([near, far, near[:, ::-1]])
        # move everything AFTER the frame: the pair tests must see the new positions
        worlds.advance(w, 1)
        ds.set_positions(w.pos)
        sp.set_positions(w.pos)
        cp, ph, pc = ds.test_pairs(pairs)
        assert len(ph) == len(pairs) and cp.n_candidates == len(pairs)
        assert np.array_equal(ph["a"], pairs[:, 0]) and np.array_equal(ph["b"], pairs[:, 1])
        assert np.array_equal(ph["count_rank"] >> 8, np.arange(len(pairs)))
        n_hit = 0
        for k, (a, b) in enumerate(pairs):
            mtv, con = sp.intersection(int(a), int(b))
            n = int(ph["count_rank"][k] & 0x7f)
            assert n == len(con), (k, a, b, n, len(con))
            if n:
                n_hit += 1
                f = int(ph["first_contact"][k])
                dev = np.concatenate([pc["p"][f:f + n], pc["n"][f:f + n]], -1)
                assert np.array_equal(bits(dev), bits(con)) and np.array_equal(bits(ph["mtv"][k]), bits(mtv)), (k, a, b)
        assert n_hit > 500 and cp.n_hits == n_hit
        # the grid of the last frame is untouched: a ray query still works and the next frame is a normal frame
        ds.frame(1, _abi.FRAME_DOWNLOAD, grow=True)
    finally:
        ds.close()


def _frame_table(ds, margin, flags=0):
    c = ds.frame(margin, _abi.FRAME_DOWNLOAD | flags, grow=True)
    hits, con = ds.results()
    return c.as_dict(), hits.copy(), con.copy()


def _same_frames(x, y):
    assert x[0] == y[0]
    ox, oy = np.lexsort((x[1]["b"], x[1]["a"])), np.lexsort((y[1]["b"], y[1]["a"]))
    for f in ("a", "b", "count_rank"):
        assert np.array_equal(x[1][f][ox], y[1][f][oy]), f
    assert np.array_equal(bits(x[1]["mtv"][ox]), bits(y[1]["mtv"][oy]))
    n = x[1]["count_rank"] & 0x7f
    cx = np.concatenate([np.concatenate([x[2]["p"][f:f + k], x[2]["n"][f:f + k]], -1) for f, k in zip(x[1]["first_contact"][ox], n[ox])])
    cy = np.concatenate([np.concatenate([y[2]["p"][f:f + k], y[2]["n"][f:f + k]], -1) for f, k in zip(y[1]["first_contact"][oy], n[ox])])
    assert np.array_equal(bits(cx), bits(cy))


def test_position_pages_sparse_updates_and_frames_in_flight():
    w = _world()
    ds = space_for_world(w)
    try:
        sets = [w.pos.copy()]
        for f in range(3):
            worlds.advance(w, f + 1)
            sets.append(w.pos.copy())
        # reference: plain commit + frame per position set
        want = []
        for p in sets:
            ds.set_positions(p)
            want.append(_frame_table(ds, 1))
        # (1) device pages: selecting page k == having committed set k
        ds.device_pos_pages(sets)
        for k in (2, 0, 3, 1):
            ds.select_pos_page(k)
            _same_frames(_frame_table(ds, 1), want[k])
        # (2) upload stream: page 1 is overwritten with set 3 while frames on page 0 are in flight
        ds.select_pos_page(0)
        ds.frame_launch(1, _abi.FRAME_DOWNLOAD)
        ds.pos_page(1)[:2 * w.n] = sets[3].reshape(-1)
        ds.pos_page_upload(1, 1)
        ds.select_pos_page(1)
        ds.frame_launch(1, _abi.FRAME_DOWNLOAD)
        c0 = ds.frame_finish()
        t0 = (c0.as_dict(), ds.results()[0].copy(), ds.results()[1].copy())
        c1 = ds.frame_finish()
        t1 = (c1.as_dict(), ds.results()[0].copy(), ds.results()[1].copy())
        _same_frames(t0, want[0])
        _same_frames(t1, want[3])
        # (3) sparse update: move 500 shapes of page 0 to their set-2 positions == committing that mixed position set
        ds.select_pos_page(0)
        ds.set_positions(sets[0])
        rs = np.random.default_rng(11)
        moved = np.sort(rs.choice(w.n, 500, replace=False)).astype(np.uint32)
        slots, xy = ds.pos_list()
        slots[:500] = moved
        xy[:1000] = sets[2][moved].reshape(-1)
        ds.commit_pos_list(500)
        got = _frame_table(ds, 1)
        mixed = sets[0].copy()
        mixed[moved] = sets[2][moved]
        ds.set_positions(mixed)
        _same_frames(got, _frame_table(ds, 1))
    finally:
        ds.close()
