"""bench.py helpers that need no GPU: the host-side cell statistics behind the byte accounting of the cell-pair kernels
must agree with the oracle's cell membership."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from resolv_b200 import worlds  # noqa: E402
from tests.util import OracleBatch  # noqa: E402


def test_grid_stats_match_the_oracles_cells():
    for w in (worlds.make_mixed(3000, space=2000, cell=32), worlds.make_bouncer(120), worlds.make_rects(2500, space=1500, cell=16)):
        gs = bench.grid_stats(w, w.pos)
        counts = np.array([len(c) for c in OracleBatch(w).cells()])
        assert gs["multi_cells"] == int((counts >= 2).sum())
        assert gs["multi_entries"] == int(counts[counts >= 2].sum())
        assert gs["query_entries"] > 0


def test_cell_pair_accounting_is_smaller_than_the_walks():
    w = worlds.make_mixed(3000, space=2000, cell=32)

    class C:
        n_shapes, n_entries, n_gated, n_contacts = 3000, 7400, 520, 760
    n_cells = 63 * 63
    walk = bench.kernel_algorithmic_bytes(C, w, n_cells, filters=True)
    cp = bench.kernel_algorithmic_bytes(C, w, n_cells, filters=True, gs=bench.grid_stats(w, w.pos))
    assert set(walk) == set(cp) and cp["cellsort"] == 0 and cp["pairs"] < walk["pairs"] + walk["cellsort"]
    assert cp["narrow"] == walk["narrow"] and cp["bounds"] == walk["bounds"]
