"""The default frame's candidate path (resolv_b200/csrc/rsv_cellpairs.cuh) against the reference's literal walk and the
oracle: what is specific to it — records grouped through the tester table, the per-tester counters that must be zero
between frames (also after a frame that overflowed), cells left unordered until a consumer of the full grid follows,
cells crowded enough for the warp-cooperative path, and the flag that selects the walk.

(Every world of tests/test_parity_gpu.py and tests/test_fuzz_gpu.py already runs through all three candidate paths:
tests/util.py: compare_frame.)"""
import numpy as np
import pytest

from resolv_b200 import _abi, space_for_world, worlds
from tests.util import OracleBatch, bits, compare_frame, device_cells

pytestmark = pytest.mark.gpu


def _gated(ds):
    hits, con = ds.results()
    h = hits[(hits["count_rank"] & 0x80) == 0]
    o = np.lexsort((h["b"], h["a"]))
    h = h[o]
    n = (h["count_rank"] & 0x7f).astype(np.int64)
    pts = np.concatenate([np.concatenate([con["p"][f:f + k], con["n"][f:f + k]], -1) for f, k in zip(h["first_contact"], n)]) \
        if n.sum() else np.zeros((0, 4))
    return np.stack([h["a"], h["b"], h["count_rank"]], -1).copy(), bits(h["mtv"]).copy(), bits(pts).copy()


def test_the_walk_flag_counts_candidates_and_both_paths_return_the_same_records():
    w = worlds.make_mixed(12000, space=6000, cell=32)
    ds = space_for_world(w)
    try:
        cw = ds.frame(w.margin, _abi.FRAME_DOWNLOAD | _abi.FRAME_CANDIDATE_WALK, grow=True)
        walk = _gated(ds)
        cd = ds.frame(w.margin, _abi.FRAME_DOWNLOAD, grow=True)
        cell = _gated(ds)
        assert cw.n_candidates > 0 and cd.n_candidates == 0
        assert cd.n_gated == len(cell[0]) <= cw.n_gated          # (the walk also records float-gated pairs that fail the exact gate)
        assert (cd.n_testers, cd.n_entries, cd.n_hits, cd.n_contacts) == (cw.n_testers, cw.n_entries, cw.n_hits, cw.n_contacts)
        for a, b in zip(walk, cell):
            assert a.shape == b.shape and np.array_equal(a, b)
    finally:
        ds.close()


def test_a_default_frame_fills_the_tester_table_without_being_asked():
    w = worlds.make_mixed(6000, space=4000, cell=32, filtered=False)
    ds = space_for_world(w)
    try:
        c = ds.frame(1, _abi.FRAME_DOWNLOAD, grow=True)
        hits, _ = ds.results()
        first, count = ds.tester_table()
        assert int(count.sum()) == len(hits) == c.n_gated
        cnt = count.astype(np.int64)
        idx = np.repeat(first.astype(np.int64), cnt) + (np.arange(int(cnt.sum())) - np.repeat(np.cumsum(cnt) - cnt, cnt))
        assert np.array_equal(np.sort(idx), np.arange(len(hits))), "the testers' ranges do not tile the record buffer"
        assert np.array_equal(hits["a"][idx], np.repeat(np.arange(w.n), cnt))
        rk = (hits["count_rank"][idx] >> 8).astype(np.int64)
        same = hits["a"][idx][1:] == hits["a"][idx][:-1]
        assert np.all(rk[1:][same] > rk[:-1][same]), "records of a tester are not in generation order"
    finally:
        ds.close()


def test_an_overflowed_frame_leaves_nothing_behind():
    """The per-tester record counters are consumed (and zeroed) by the frame that filled them. A frame whose record buffer
    overflowed must leave them clean too: the retried frame and the frames after it equal a fresh space's."""
    w = worlds.make_mixed(5000, space=1200, cell=32, filtered=False)
    ds, fresh = space_for_world(w, max_pairs=128, max_contacts=1 << 16), space_for_world(w)
    try:
        fresh.frame(1, _abi.FRAME_DOWNLOAD, grow=True)
        want = _gated(fresh)
        for _ in range(2):                                   # twice: the second overflow starts from the first one's leftovers
            with pytest.raises(_abi.RsvError) as e:
                ds.frame(1, _abi.FRAME_DOWNLOAD)
            assert e.value.code == _abi.RSV_ERR_CAPACITY
        ds.reserve(pairs=4 * len(want[0]) + 1024, contacts=1 << 18)
        for _ in range(2):
            c = ds.frame(1, _abi.FRAME_DOWNLOAD)
            assert c.overflow == 0
            got = _gated(ds)
            for a, b in zip(want, got):
                assert a.shape == b.shape and np.array_equal(a, b)
    finally:
        ds.close()
        fresh.close()


def test_crowded_cells_take_the_cooperative_path():
    """Hundreds of shapes in a handful of cells: every listed cell is far above the four entries that are expanded into
    pair items, so all pairs come from the warp-per-cell loop; ranks reach the hundreds."""
    w = worlds.make_mixed(900, space=160, cell=32, filtered=True)
    ob = OracleBatch(w)
    ds = space_for_world(w, max_pairs=1 << 20, max_contacts=1 << 21)
    try:
        c, exact = compare_frame(ds, ob, 1)
        assert exact and c.n_candidates / c.n_testers > 100
        assert device_cells(ds, c.n_entries) == ob.cells()
    finally:
        ds.close()


def test_linetest_and_cell_readback_after_a_default_frame_see_the_full_grid():
    """A default frame leaves the cells unordered and without per-entry boxes; LineTest and the grid read-back complete
    the grid first (ensure_full_grid): ray sets equal to the oracle's and in generation order, cells in ascending slot order."""
    w = worlds.make_mixed(8000, space=3000, cell=32, filtered=False)
    ob = OracleBatch(w)
    ds = space_for_world(w, max_rays=512)
    try:
        c = ds.frame(1, _abi.FRAME_DOWNLOAD, grow=True)
        rays = worlds.make_rays(400, 77, w.space_w, w.space_h, 8, 600)
        ds.linetest(rays, flags=_abi.FRAME_DOWNLOAD)
        hits, con, first, count = ds.ray_results()
        fc, oh = ob.spaces[0].linetest(rays, mode="rect")
        assert len(hits) == len(oh.a) > 0
        o, dd = np.lexsort((oh.b, oh.a)), np.lexsort((hits["b"], hits["ray"]))
        assert np.array_equal(hits["ray"][dd], oh.a[o]) and np.array_equal(hits["b"][dd], oh.b[o]), "ray sets differ"
        assert np.array_equal(bits(hits["mtv"][dd]), bits(oh.mtv[o])) and np.array_equal(bits(hits["key"][dd]), bits(oh.dist[o]))
        # a ray's sets are contiguous on the device (ray_first / ray_count), in generation order (ascending rank)
        at = np.concatenate([np.arange(f, f + k) for f, k in zip(first, count) if k])
        assert np.array_equal(np.sort(at), np.arange(len(hits)))
        rk = (hits["count_rank"][at] >> 8).astype(np.int64)
        same = hits["ray"][at][1:] == hits["ray"][at][:-1]
        assert np.all(rk[1:][same] > rk[:-1][same])
        ds.frame(1, _abi.FRAME_DOWNLOAD)                      # and again: the read-back alone must complete it as well
        assert device_cells(ds, c.n_entries) == ob.cells()
    finally:
        ds.close()


def test_sets_of_more_than_twelve_contacts_are_sorted_by_the_host_with_gos_sort():
    """sort.Slice on more than 12 items is Go's pdqsort_func: unstable, its result depends on the input order. The device
    delivers such sets in generation order (the input the reference's sort sees) and the host layer sorts them with its
    port of Go's sort; compare_frame plays the host with the oracle's restatement of that sort — so the contact ORDER is
    compared exactly, not as a multiset. Octagons cut by a circle (16 points, all at distance r from the circle: every
    key ties) and by a rotated octagon (16 points at the same distance from the centre)."""
    def ngon(n, r, phase):
        t = phase + 2 * np.pi * np.arange(n) / n
        return np.stack([r * np.cos(t), r * np.sin(t)], -1)

    rows = [dict(kind=1, pos=(100.0, 100.0), radius=0.0, verts=ngon(8, 20.0, 0.0), tags=1),
            dict(kind=0, pos=(100.0, 100.0), radius=19.2, tags=1),
            dict(kind=1, pos=(300.0, 200.0), radius=0.0, verts=ngon(8, 30.0, 0.0), tags=1),
            dict(kind=1, pos=(300.0, 200.0), radius=0.0, verts=ngon(8, 30.0, np.pi / 8), tags=1),
            dict(kind=1, pos=(500.5, 300.25), radius=0.0, verts=ngon(7, 25.0, 0.3), tags=1),
            dict(kind=0, pos=(501.0, 300.0), radius=23.5, tags=1)]
    w = worlds._from_rows(rows, (640, 480), (32, 32), "many_contacts").finalize()
    ob = OracleBatch(w)
    ds = space_for_world(w)
    try:
        c, exact = compare_frame(ds, ob, 1)
        hits, _ = ds.results()
        assert exact and int((hits["count_rank"] & 0x7f).max()) > 12
    finally:
        ds.close()
