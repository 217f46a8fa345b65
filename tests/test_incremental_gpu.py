"""Incremental grid (RSV_META_STATIC + RSV_FRAME_INCREMENTAL): the registrations of the static
shapes are cached on the device and a frame re-registers only the others — what the reference does implicitly, since
only ShapeBase.update of a moved shape touches cells (shape.go:183-214,239-242).

Bar: an incremental frame is bit-identical to a full rebuild — same cells (content and in-cell order), same candidate
set with generation rank, same hits, MTVs and contacts — and therefore equal to the oracle like every other frame.
"""
import numpy as np
import pytest

from resolv_b200 import _abi, space_for_world, worlds
from tests.util import OracleBatch, compare_frame, device_cells

pytestmark = pytest.mark.gpu

INC = _abi.FRAME_INCREMENTAL


def _snapshot(ds, margin, flags):
    c = ds.frame(margin, flags | _abi.FRAME_DOWNLOAD | _abi.FRAME_ALL_CANDIDATES, grow=True)
    h, k = ds.results()
    o = np.lexsort((h["b"], h["a"]))
    n = h["count_rank"][o] & 0x7f
    pts = [np.concatenate([k["p"][f:f + m], k["n"][f:f + m]], -1) for f, m in zip(h["first_contact"][o], n)]
    con = np.concatenate(pts).view(np.uint64).copy() if pts else np.zeros(0, np.uint64)
    return (c.as_dict(), h["a"][o].copy(), h["b"][o].copy(), h["count_rank"][o].copy(), h["mtv"][o].view(np.uint64).copy(), con,
            device_cells(ds, c.n_entries))


def _assert_same(a, b):
    assert a[0] == b[0], (a[0], b[0])
    for x, y in zip(a[1:6], b[1:6]):
        assert np.array_equal(x, y)
    assert a[6] == b[6], "cells differ (content or in-cell order)"


def test_platformer_batch_tiles_static_bodies_dynamic():
    w = worlds.make_platformer_batch(16)
    static = w.tester == 0                       # the tiles; the 32 bodies per Space move every frame
    ob = OracleBatch(w)
    ds, full = space_for_world(w), space_for_world(w)
    try:
        ds.set_static(static)
        for f in range(5):
            if f:
                worlds.advance(w, f)
                ds.set_positions(w.pos)
                full.set_positions(w.pos)
            inc = _snapshot(ds, w.margin, INC | (0 if f % 2 else _abi.FRAME_TESTER_TABLE))
            _assert_same(inc, _snapshot(full, w.margin, 0))
            if f == 0:
                se, nd = ds.static_info()
                assert nd == int((~static).sum()) and se > 0 and se + 4 * nd >= inc[0]["n_entries"] > se
            c, exact = compare_frame(ds, OracleBatch(w), w.margin, check_rank=True, extra_flags=INC)
            assert exact and c.n_hits > 0
            assert device_cells(ds, c.n_entries) == OracleBatch(w).cells()
        ob.set_positions(w.pos)                   # the moved oracle (history-dependent in-cell order): sets only
        compare_frame(ds, ob, w.margin, check_rank=False, extra_flags=INC)
        # nothing moves any more: every shape static, no dynamic list at all
        ds.set_static(np.ones(w.n, dtype=bool))
        for _ in range(2):
            _assert_same(_snapshot(ds, w.margin, INC), _snapshot(full, w.margin, 0))
        assert ds.static_info()[1] == 0
    finally:
        ds.close()
        full.close()


@pytest.mark.parametrize("filtered", [False, True])
def test_mixed_world_with_a_random_static_subset(filtered):
    w = worlds.make_mixed(12000, space=6000, cell=32, filtered=filtered)
    rng = np.random.default_rng(5)
    static = rng.random(w.n) < 0.8
    base_flags = 0 if filtered else _abi.FRAME_NO_TAG_FILTERS
    ds, full = space_for_world(w), space_for_world(w)
    try:
        ds.set_static(static)
        p_static = w.pos.copy()
        for f in range(4):
            if f:
                worlds.advance(w, f)
                w.pos[static] = p_static[static]
                ds.set_positions(w.pos)
                full.set_positions(w.pos)
            _assert_same(_snapshot(ds, w.margin, INC | base_flags), _snapshot(full, w.margin, base_flags))
        c, exact = compare_frame(ds, OracleBatch(w), w.margin, check_rank=True, extra_flags=INC)
        assert exact and c.n_hits > 100
        # a static shape moves after all: the host says so and the cache is rebuilt
        mover = np.nonzero(static)[0][:50]
        w.pos[mover] += 40.0
        ds.set_positions(w.pos)
        full.set_positions(w.pos)
        ds.static_invalidate()
        _assert_same(_snapshot(ds, w.margin, INC | base_flags), _snapshot(full, w.margin, base_flags))
        # committing META (a shape leaves the Space) drops the cache by itself
        for d in (ds, full):
            d.buf[_abi.BUF_META][int(mover[0])] |= _abi.META_ABSENT
            d.commit(1 << _abi.BUF_META)
        _assert_same(_snapshot(ds, w.margin, INC | base_flags), _snapshot(full, w.margin, base_flags))
    finally:
        ds.close()
        full.close()


def test_static_shapes_outside_the_grid_and_empty_dynamic_set():
    # static testers beyond the border (orphan testers: margin reaches the grid), static shapes that touch no cell
    # at all, and a world in which nothing is dynamic
    w = worlds.make_mixed(3000, space=2000, cell=32, filtered=False)
    w.pos[:40, 0] = -40.0 - np.arange(40) % 3 * 16.0           # left of the grid, some within margin reach
    w.pos[40:60, 1] = 2000.0 + 500.0                           # far below: no cell, no reach
    static = np.ones(w.n, dtype=bool)
    ds, full = space_for_world(w), space_for_world(w)
    try:
        ds.set_static(static)
        for margin in (1, 2):
            _assert_same(_snapshot(ds, margin, INC), _snapshot(full, margin, 0))
        se, nd = ds.static_info()
        assert se > 0 and nd == 40 + 20    # the static testers that hold no cell are re-examined every frame
        compare_frame(ds, OracleBatch(w), 1, check_rank=True, extra_flags=INC)
        # mostly dynamic: the library falls back to the full rebuild on its own
        ds.set_static(np.arange(w.n) < 100)
        _assert_same(_snapshot(ds, 1, INC), _snapshot(full, 1, 0))
    finally:
        ds.close()
        full.close()


def test_incremental_frames_replay_as_a_graph_and_feed_linetest():
    w = worlds.make_platformer_batch(8)
    static = w.tester == 0
    rays, use_any, any_mask, sidx = worlds.make_platformer_probes(w)
    ds, full = space_for_world(w, max_rays=len(rays)), space_for_world(w, max_rays=len(rays))
    try:
        ds.set_static(static)
        ds.frame(w.margin, INC | _abi.FRAME_DOWNLOAD, grow=True)
        full.frame(w.margin, _abi.FRAME_DOWNLOAD, grow=True)
        for f in range(1, 7):                      # identical launch parameters: replayed from the third frame on
            worlds.advance(w, f)
            out = []
            for d, fl in ((ds, INC), (full, 0)):
                d.set_positions(w.pos)
                d.frame_launch(w.margin, fl | _abi.FRAME_DOWNLOAD)
                c = d.frame_finish()
                h, k = d.results()
                h = h[(h["count_rank"] & 0x80) == 0]      # (the incremental frame walks: float-gated records, exact gate per record;
                o = np.lexsort((h["b"], h["a"]))          #  the full one takes the cell-pair path: exactly the gated pairs, no candidate count)
                cd = {k_: v for k_, v in c.as_dict().items() if k_ not in ("n_candidates", "n_gated")}
                out.append((cd, h["a"][o].copy(), h["b"][o].copy(), h["count_rank"][o].copy(),
                            h["mtv"][o].view(np.uint64).copy()))
            assert out[0][0] == out[1][0]
            for x, y in zip(out[0][1:], out[1][1:]):
                assert np.array_equal(x, y)
        res = []
        for d in (ds, full):
            d.linetest(rays, use_any=use_any, any_mask=any_mask, space_idx=sidx)
            h, k, first, count = d.ray_results()
            o = np.lexsort((h["b"], h["ray"]))
            n = h["count_rank"][o] & 0xff
            pts = np.concatenate([np.concatenate([k["p"][f:f + m], k["n"][f:f + m]], -1) for f, m in zip(h["first_contact"][o], n)])
            res.append((h["ray"][o].copy(), h["b"][o].copy(), h["count_rank"][o].copy(), h["mtv"][o].view(np.uint64).copy(),
                        h["key"][o].view(np.uint64).copy(), pts.view(np.uint64).copy(), count.copy()))
        for x, y in zip(res[0], res[1]):
            assert np.array_equal(x, y)
        assert len(res[0][0]) > 0
    finally:
        ds.close()
        full.close()


def test_many_small_random_worlds_incremental_equals_full_and_oracle():
    """Randomised sweep: random static subsets (0 .. 100 %), shapes off the grid, non-testers, batched Spaces, several
    frames in which dynamic shapes move (sometimes across many cells) and static shapes are re-flagged."""
    rs = np.random.default_rng(11)
    hits = 0
    for trial in range(30):
        if trial % 5 == 4:
            w = worlds.make_platformer_batch(int(rs.integers(1, 5)), seed=int(rs.integers(1, 1 << 30)), n_tiles=int(rs.integers(140, 300)),
                                             n_bodies=int(rs.integers(1, 40)))
        else:
            n = int(rs.integers(2, 400))
            size = int(rs.choice([64, 257, 600]))
            cell = int(rs.choice([7, 16, 33]))
            w = worlds.make_mixed(n, seed=int(rs.integers(1, 1 << 30)), space=size, cell=cell, filtered=bool(trial % 2))
            w.pos = w.pos * 1.3 - 0.15 * size          # part of the shapes off-grid
            w.tester[:] = rs.integers(0, 2, n) | (rs.random(n) < 0.5)
            w.finalize()
        static = rs.random(w.n) < rs.choice([0.0, 0.5, 0.9, 1.0])
        margin = int(rs.integers(0, 4))
        ds, full = space_for_world(w), space_for_world(w)
        try:
            ds.set_static(static)
            for f in range(3):
                if f:
                    step = rs.normal(0.0, 1.0, (w.n, 2)) * float(rs.choice([2.0, 40.0]))
                    w.pos[~static] += step[~static]
                    if f == 2:                             # a few static shapes become dynamic and the other way round
                        flip = rs.random(w.n) < 0.1
                        static = static ^ flip
                        ds.set_static(static)
                    ds.set_positions(w.pos)
                    full.set_positions(w.pos)
                _assert_same(_snapshot(ds, margin, INC), _snapshot(full, margin, 0))
            c, exact = compare_frame(ds, OracleBatch(w), margin, check_rank=True, extra_flags=INC)
            assert exact
            hits += c.n_hits
        finally:
            ds.close()
            full.close()
    assert hits > 200
