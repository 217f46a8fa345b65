"""RSV_FRAME_OWN_RECT (unfiltered frames): a tester visits only the entries of its own cell rect; candidate counts and
generation ranks come in closed form from prefix counts of the entry flags (rsv_pairs.cuh, k_pairs_own). Replaces the
same reference code as the regular walk — SelectTouchingCells / CellSelection.ForEach / the gather loop of
IntersectionTest (shape.go:220-237,277-289, cell.go:86-121) — and must give the same answers: checked against the CPU
oracle and, record for record, against the regular walk."""
import numpy as np
import pytest

from resolv_b200 import _abi, space_for_world, worlds
from tests.util import OracleBatch, compare_frame

pytestmark = pytest.mark.gpu

OWN = _abi.FRAME_NO_TAG_FILTERS | _abi.FRAME_OWN_RECT


@pytest.mark.parametrize("name", ["rects", "bouncer"])
def test_own_rect_frames_match_the_oracle(name):
    w = worlds.make_rects(20000, space=9280, cell=32) if name == "rects" else worlds.make_bouncer()
    ds = space_for_world(w, max_pairs=64 * w.n, max_contacts=16 * w.n)
    try:
        ob = OracleBatch(w)
        for margin in (1, 2):
            c, exact = compare_frame(ds, ob, margin, check_candidates=False, extra_flags=OWN)   # n_candidates, hits, contacts, MTV
            assert exact and c.n_hits > 0
    finally:
        ds.close()


def _records(ds, margin, flags):
    c = ds.frame(margin, flags | _abi.FRAME_DOWNLOAD | _abi.FRAME_TESTER_TABLE, grow=True)
    hits, con = ds.results()
    hits, con = hits.copy(), con.copy()
    first, count = ds.tester_table()
    return c, hits, con, first.copy(), count.copy()


def _canon(hits, con):
    # records that only passed the conservative float32 pre-gate (RSV_HIT_NOT_GATED, empty sets) are not part of the
    # contract: the regular walk also stages such pairs from cells beyond the tester's own rect
    hits = hits[(hits["count_rank"] & 0x80) == 0]
    rank, n = hits["count_rank"] >> 8, hits["count_rank"] & 0x7f
    order = np.lexsort((rank, hits["a"]))
    h = hits[order]
    idx = np.concatenate([np.arange(f, f + k) for f, k in zip(h["first_contact"], n[order])] + [np.zeros(0, np.int64)]).astype(np.int64)
    return h["a"], h["b"], h["count_rank"], h["mtv"].copy().view(np.uint64), con[idx]


@pytest.mark.parametrize("world", ["dense_off_grid", "tall_wide", "crowded", "batch"])
def test_own_rect_records_equal_the_regular_walk(world):
    """Same records (tester, other, generation rank), same per-tester order, same tester table, same MTVs and contacts —
    for shapes straddling / outside the grid, rects of more than four cell rows, batches that overflow the record stage
    (fallback to the regular walk inside the kernel) and batched Spaces."""
    if world == "dense_off_grid":
        w = worlds.make_rects(3000, space=1024, cell=32)
        w.pos[:12] = np.array([[-5.0, 40.0], [3.0, -2.0], [1028.0, 100.0], [1021.0, 1022.0]] * 3)
        margins = (0, 1, 3)
    elif world == "tall_wide":
        w, margins = worlds.make_rects(300, space=1024, cell=16, lo=10.0, hi=150.0), (0, 1, 5)
    elif world == "crowded":
        w, margins = worlds.make_rects(500, space=512, cell=32), (1, 2)
    else:
        w, margins = worlds.make_platformer_batch(16), (1, 4)
    ds = space_for_world(w)
    try:
        for m in margins:
            c0, h0, k0, _, _ = _records(ds, m, _abi.FRAME_NO_TAG_FILTERS)
            c1, h1, k1, f1, n1 = _records(ds, m, OWN)
            for f in ("n_testers", "n_entries", "n_candidates", "n_hits", "n_contacts"):
                assert getattr(c0, f) == getattr(c1, f), (world, m, f)
            a0, a1 = _canon(h0, k0), _canon(h1, k1)
            for x, y in zip(a0[:4], a1[:4]):
                assert np.array_equal(x, y), (world, m)
            assert a0[4].tobytes() == a1[4].tobytes(), (world, m)
            # per tester contiguous, ascending generation rank, and indexed by the tester table
            a, r = h1["a"].astype(np.int64), (h1["count_rank"] >> 8).astype(np.int64)
            assert np.all((np.diff(r) > 0) | (np.diff(a) != 0))
            assert int(n1.sum()) == len(h1)
            for t in np.nonzero(n1)[0][:500]:
                assert np.all(h1["a"][f1[t]:f1[t] + n1[t]] == t)
    finally:
        ds.close()
