"""Parity of the CUDA path (through the C ABI) against the CPU oracle on the same seeded worlds.

Bar (BASELINE.json north_star): cell membership, candidate-pair set (with generation rank) and intersecting-pair
set bit-exact; MTV and contact points within 1e-9 absolute in float64. We additionally assert bit-identity of
MTV / contacts where it holds (it is expected everywhere: same operation order, IEEE div/sqrt, no FMA).
"""
import numpy as np
import pytest

from resolv_b200 import _abi, space_for_world, worlds
from tests.util import OracleBatch, compare_frame, device_cells

pytestmark = pytest.mark.gpu


def _check_world(w, frames=1, margin=None, expect_exact=True):
    margin = w.margin if margin is None else margin
    ob = OracleBatch(w)
    ds = space_for_world(w)  # default capacities; compare_frame grows them on RSV_ERR_CAPACITY
    try:
        for f in range(frames):
            if f:
                worlds.advance(w, f)
                ob.set_positions(w.pos)
                ds.set_positions(w.pos)
            # moved oracle (history-dependent in-cell order): sets; freshly built oracle: sets + generation rank
            c, exact = compare_frame(ds, ob, margin, check_rank=(f == 0))
            if f:
                c, exact2 = compare_frame(ds, OracleBatch(w), margin, check_rank=True)
                assert exact == exact2
            assert device_cells(ds, c.n_entries) == ob.cells(), "cell membership mismatch"
            if expect_exact:
                assert exact, "MTV/contacts within 1e-9 but not bit-identical"
        return c
    finally:
        ds.close()


def test_cfg1_bouncer_world():
    c = _check_world(worlds.make_bouncer(294), frames=3)
    assert c.n_hits > 0


def test_cfg2_rects_dense():
    # config-2 density (N / cells ~ 0.24) at a size the oracle finishes in a second
    w = worlds.make_rects(20000, space=9280, cell=32)
    c = _check_world(w, frames=2)
    assert c.n_hits > 1000


def test_cfg3_mixed_filtered():
    w = worlds.make_mixed(20000, space=9280, cell=32)
    c = _check_world(w, frames=2)
    assert c.n_hits > 500


def test_cfg3_mixed_crowded():
    # much denser than the configs: many contacts per shape, long candidate lists (> 12 -> pdqsort regime on the host)
    w = worlds.make_mixed(3000, space=800, cell=32, filtered=False)
    c = _check_world(w, frames=2)
    assert c.n_candidates / c.n_testers > 12


def test_cfg4_platformer_batch():
    w = worlds.make_platformer_batch(8)
    c = _check_world(w, frames=3)
    assert c.n_testers == 8 * 32 and c.n_hits > 0


@pytest.mark.parametrize("margin", [0, 1, 3])
def test_margins(margin):
    w = worlds.make_mixed(5000, space=4000, cell=32, filtered=False)
    _check_world(w, margin=margin)


def test_shapes_outside_and_straddling_the_grid():
    # out-of-grid cells are skipped (space.go:135-142): shapes fully outside are in no cell but still test
    # against in-grid cells their margin reaches; straddling shapes register only in-grid.
    w = worlds.make_mixed(4000, space=1024, cell=32, filtered=False)
    w.pos = w.pos * 1.3 - 150.0          # spills over every border
    w.pos0 = w.pos.copy()
    w.finalize()
    _check_world(w, frames=2)


def test_non_square_cells_and_ragged_grid():
    w = worlds.make_mixed(3000, space=1000, cell=32, filtered=False)   # 1000/24 and 1000/40 are not integers
    w.cell_w, w.cell_h = 24, 40
    _check_world(w)


def test_empty_and_single():
    w = worlds.make_rects(1, space=256, cell=32)
    c = _check_world(w)
    assert c.n_candidates == 0 and c.n_hits == 0
    w = worlds.make_rects(2, space=256, cell=32)
    w.pos[:] = [[100, 100], [110, 104]]
    w.pos0 = w.pos.copy()
    w.finalize()
    c = _check_world(w)
    assert c.n_hits == 2


def test_open_polylines_and_lines():
    # NewLine-style 2-point polygons (never closed, convexPolygon.go:129) and open 3+-point polylines
    w = worlds.make_mixed(2000, space=700, cell=32, filtered=False)
    poly = np.nonzero(w.kind == 1)[0]
    w.closed[poly[::3]] = 0
    two = poly[1::5]
    w.nv[two] = 2
    w.finalize()
    _check_world(w, frames=2)


def test_large_shapes_spanning_many_cells():
    rows = []
    worlds._append_top_left_rects(rows, 0.0, 0.0, 640.0, 16.0, 1)
    worlds._append_top_left_rects(rows, 0.0, 100.0, 640.0, 48.0, 1)
    rows.append(dict(kind=0, pos=(320.0, 200.0), radius=150.0, tags=1))
    rs = np.random.default_rng(5)
    for _ in range(200):
        x, y = rs.uniform(0, 640), rs.uniform(0, 360)
        rows.append(dict(kind=0, pos=(x, y), radius=rs.uniform(2, 12), tags=1))
    w = worlds._from_rows(rows, (640, 360), (16, 16), "big").finalize()
    _check_world(w)


def test_polygons_with_many_vertices_take_the_unstaged_path():
    rs = np.random.default_rng(7)
    rows = []
    for i in range(300):
        n = int(rs.integers(9, 20))
        ang = np.sort(rs.uniform(0, 2 * np.pi, n))
        r = rs.uniform(10, 30)
        rows.append(dict(kind=1, pos=(rs.uniform(0, 400), rs.uniform(0, 400)), radius=0.0,
                         verts=np.stack([np.cos(ang) * r, np.sin(ang) * r], -1), tags=1))
    for i in range(100):
        rows.append(dict(kind=0, pos=(rs.uniform(0, 400), rs.uniform(0, 400)), radius=rs.uniform(5, 20), tags=1))
    w = worlds._from_rows(rows, (400, 400), (32, 32), "bigpoly").finalize()
    _check_world(w)


def test_capacity_overflow_is_reported_not_silent():
    w = worlds.make_mixed(3000, space=800, cell=32, filtered=False)
    ds = space_for_world(w, max_pairs=64, max_contacts=64)
    try:
        with pytest.raises(_abi.RsvError) as e:
            ds.frame(1)
        assert e.value.code == _abi.RSV_ERR_CAPACITY
        ds.reserve(pairs=64 * w.n, contacts=16 * w.n)
        c = ds.frame(1)
        assert c.overflow == 0 and c.n_hits > 0
    finally:
        ds.close()


def test_kat_pairs_through_the_device():
    # SURVEY.md §8c KAT1-4 as a tiny world: exact expected values
    rows = [dict(kind=1, pos=(200.0, 100.0), radius=0.0, verts=worlds.rect_verts(32.0, 32.0), tags=1),
            dict(kind=0, pos=(200.0, 120.0), radius=8.0, tags=1),
            dict(kind=1, pos=(400.0, 200.0), radius=0.0, verts=worlds.rect_verts(10.0, 10.0), tags=1),
            dict(kind=1, pos=(408.0, 202.0), radius=0.0, verts=worlds.rect_verts(10.0, 10.0), tags=1),
            dict(kind=0, pos=(100.0, 300.0), radius=5.0, tags=1),
            dict(kind=0, pos=(106.0, 300.0), radius=3.0, tags=1)]
    w = worlds._from_rows(rows, (640, 360), (16, 16), "kat").finalize()
    ds = space_for_world(w)
    try:
        c = ds.frame(1)
        hits, con = ds.results()
        by = {(int(h["a"]), int(h["b"])): h for h in hits if h["count_rank"] & 0x7f}
        assert set(by) == {(0, 1), (1, 0), (2, 3), (3, 2), (4, 5), (5, 4)}
        assert by[(0, 1)]["mtv"].tolist() == [0.0, -4.0]
        assert by[(2, 3)]["mtv"].tolist() == [-2.0, 0.0]
        assert by[(4, 5)]["mtv"].tolist() == [-2.0, 0.0]
        f = int(by[(2, 3)]["first_contact"])
        assert np.abs(con["p"][f:f + 2] - np.array([[402.9, 205.0], [405.0, 196.9]])).max() < 1e-9
    finally:
        ds.close()


def test_tester_table_indexes_the_records_of_each_tester():
    w = worlds.make_mixed(6000, space=4000, cell=32, filtered=False)
    ds = space_for_world(w)
    try:
        c = ds.frame(1, _abi.FRAME_DOWNLOAD | _abi.FRAME_TESTER_TABLE, grow=True)
        hits, con = ds.results()
        first, count = ds.tester_table()
        assert int(count.sum()) == len(hits) == c.n_gated
        cnt = count.astype(np.int64)
        idx = np.repeat(first.astype(np.int64), cnt) + (np.arange(int(cnt.sum())) - np.repeat(np.cumsum(cnt) - cnt, cnt))
        assert np.array_equal(hits["a"][idx], np.repeat(np.arange(w.n), cnt))
        rk = (hits["count_rank"][idx] >> 8).astype(np.int64)
        same = hits["a"][idx][1:] == hits["a"][idx][:-1]
        assert np.all(rk[1:][same] > rk[:-1][same]), "records of a tester are not in generation order"
    finally:
        ds.close()


def test_two_frames_in_flight_give_the_same_results_as_sequential_frames():
    w = worlds.make_mixed(8000, space=4600, cell=32, filtered=False)
    p0 = w.pos.copy()
    worlds.advance(w, 1)
    p1 = w.pos.copy()
    ds = space_for_world(w)
    try:
        seq = []
        for p in (p0, p1):
            ds.set_positions(p)
            c = ds.frame(1, _abi.FRAME_DOWNLOAD, grow=True)
            h, k = ds.results()
            seq.append((c.as_dict(), h.copy(), k.copy()))
        pages = [ds.pos_page(0), ds.pos_page(1)]
        pages[0][:2 * w.n] = p0.reshape(-1)
        pages[1][:2 * w.n] = p1.reshape(-1)
        ds.commit_pos_page(0)
        ds.frame_launch(1, _abi.FRAME_DOWNLOAD)
        ds.commit_pos_page(1)
        ds.frame_launch(1, _abi.FRAME_DOWNLOAD)      # second frame queued before the first is finished
        for i in range(2):
            c = ds.frame_finish()
            h, k = ds.results()
            assert c.as_dict() == seq[i][0]
            o, so = np.lexsort((h["b"], h["a"])), np.lexsort((seq[i][1]["b"], seq[i][1]["a"]))
            assert np.array_equal(h["a"][o], seq[i][1]["a"][so]) and np.array_equal(h["b"][o], seq[i][1]["b"][so])
            assert np.array_equal(h["count_rank"][o], seq[i][1]["count_rank"][so])
            assert np.array_equal(h["mtv"][o].view(np.uint64), seq[i][1]["mtv"][so].view(np.uint64))
            n = h["count_rank"][o] & 0x7f
            dc = np.concatenate([k["p"][f:f + m] for f, m in zip(h["first_contact"][o], n)])
            sc = np.concatenate([seq[i][2]["p"][f:f + m] for f, m in zip(seq[i][1]["first_contact"][so], n)])
            assert np.array_equal(dc.view(np.uint64), sc.view(np.uint64))
        assert seq[0][0] != seq[1][0]
    finally:
        ds.close()


def test_replayed_frame_graph_matches_direct_launches(monkeypatch):
    """Steady-state frames replay a captured CUDA graph (from the third identical rsv_frame_launch on); the frames it
    produces must be the ones direct launches produce, frame after frame, while the shapes move."""
    w = worlds.make_mixed(6000, space=4000, cell=32, filtered=True)
    traj = [w.pos.copy()]
    for f in range(1, 7):
        worlds.advance(w, f)
        traj.append(w.pos.copy())

    def run(no_graph):
        monkeypatch.setenv("RSV_NO_GRAPH", "1" if no_graph else "0")
        ds = space_for_world(w)
        out = []
        try:
            ds.set_positions(traj[0])
            ds.frame(w.margin, _abi.FRAME_DOWNLOAD, grow=True)   # settle capacities first
            for p in traj:
                ds.set_positions(p)
                c = ds.frame(w.margin, _abi.FRAME_DOWNLOAD)
                h, k = ds.results()
                o = np.lexsort((h["b"], h["a"]))
                n = h["count_rank"][o] & 0x7f
                pts = [k["p"][f:f + m] for f, m in zip(h["first_contact"][o], n)]
                out.append((c.as_dict(), h["a"][o].copy(), h["b"][o].copy(), h["count_rank"][o].copy(),
                            h["mtv"][o].view(np.uint64).copy(),
                            np.concatenate(pts).view(np.uint64).copy() if pts else np.zeros(0, np.uint64)))
        finally:
            ds.close()
        return out

    direct, replay = run(True), run(False)
    assert len({tuple(sorted(f[0].items())) for f in direct}) > 1, "the world did not change between frames"
    for fd, fr in zip(direct, replay):
        assert fd[0] == fr[0]
        for x, y in zip(fd[1:], fr[1:]):
            assert np.array_equal(x, y)
    # and the replayed frames are the oracle's
    ob = OracleBatch(w)
    ds = space_for_world(w)
    try:
        for _ in range(4):
            compare_frame(ds, ob, w.margin, check_rank=True)
    finally:
        ds.close()
