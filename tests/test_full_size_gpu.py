"""Parity at BASELINE size: the 1 M-shape frames the bench times, against the oracle on the SAME world.

The regular frame (what bench.py launches: the cell-pair path, records for the Bounds-gated pairs only) is compared — hit set, contact
counts, MTV and contacts, generation ranks of the hits, all counters — and, for config 3, the frame after one step of
motion as well (sets only: the oracle's in-cell order is history-dependent after moves).
The oracle frame runs single-threaded with record=1 (a few seconds per frame at 1 M shapes).
"""
import numpy as np
import pytest

from resolv_b200 import _abi, space_for_world, worlds
from tests.util import OracleBatch, compare_frame

pytestmark = pytest.mark.gpu


def _full(w, frames):
    ob = OracleBatch(w)
    ds = space_for_world(w)
    flags = 0 if (w.use_any.any() or w.none_mask.any()) else _abi.FRAME_NO_TAG_FILTERS
    try:
        for f in range(frames):
            if f:
                worlds.advance(w, f)
                ob.set_positions(w.pos)
                ds.set_positions(w.pos)
            # check_candidates=False: the regular frame (gated records only), exactly what the bench times
            c, exact = compare_frame(ds, ob, w.margin, check_candidates=False, check_rank=(f == 0), extra_flags=flags)
            assert exact, "MTV/contacts within 1e-9 but not bit-identical"
            if f == 0:
                # generation ranks of the gated records against the oracle's candidate list
                hits, _ = ds.results()
                _, cands, _ = ob.frame(w.margin)
                key = cands["a"].astype(np.int64) * w.n + cands["b"].astype(np.int64)
                order = np.argsort(key)
                hk = hits["a"].astype(np.int64) * w.n + hits["b"].astype(np.int64)
                at = np.searchsorted(key[order], hk)
                assert np.array_equal(key[order][at], hk), "a gated record is not a candidate of the oracle"
                assert np.array_equal(cands["rank"][order][at], hits["count_rank"] >> 8), "generation rank mismatch"
        return c
    finally:
        ds.close()


def test_cfg2_full_size():
    c = _full(worlds.make_rects(1_000_000), frames=1)
    assert c.n_shapes == 1_000_000 and c.n_hits > 300_000


def test_cfg3_full_size():
    c = _full(worlds.make_mixed(1_000_000), frames=2)
    assert c.n_shapes == 1_000_000 and c.n_hits > 100_000
