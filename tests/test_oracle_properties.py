"""Independent checks of the oracle (SURVEY.md §8c "extra safety nets"): the closed-form candidate predicate of
SURVEY.md §8a/B6 and brute-force all-pairs narrowphase. Neither depends on how the restatement walks cells."""
import numpy as np
import pytest

from resolv_b200 import worlds
from tests.util import OracleBatch


def _rects(sp, n):
    return np.array([sp.cell_rect(i) for i in range(n)], dtype=np.int64)


@pytest.mark.parametrize("margin", [0, 1, 2])
def test_candidates_equal_closed_form_predicate(margin):
    # B in cand(A, m)  <=>  B != A and [Acx-m,Aex+m]x[Acy-m,Aey+m], [Bcx,Bex]x[Bcy,Bey], [0,W-1]x[0,H-1] share a cell
    w = worlds.make_mixed(700, space=900, cell=32, filtered=False)
    w.pos = w.pos * 1.2 - 80.0      # some shapes straddle / leave the grid
    w.pos0 = w.pos.copy()
    w.finalize()
    ob = OracleBatch(w)
    sp = ob.spaces[0]
    tot, cands, _ = ob.frame(margin)
    r = _rects(sp, w.n)
    W, H = sp.width_cells, sp.height_cells
    ax0 = np.maximum(r[:, 0] - margin, 0)[:, None]; ax1 = np.minimum(r[:, 2] + margin, W - 1)[:, None]
    ay0 = np.maximum(r[:, 1] - margin, 0)[:, None]; ay1 = np.minimum(r[:, 3] + margin, H - 1)[:, None]
    bx0 = np.maximum(r[:, 0], 0)[None, :]; bx1 = np.minimum(r[:, 2], W - 1)[None, :]
    by0 = np.maximum(r[:, 1], 0)[None, :]; by1 = np.minimum(r[:, 3], H - 1)[None, :]
    pred = (np.maximum(ax0, bx0) <= np.minimum(ax1, bx1)) & (np.maximum(ay0, by0) <= np.minimum(ay1, by1))
    np.fill_diagonal(pred, False)
    got = np.zeros_like(pred)
    got[cands["a"], cands["b"]] = True
    assert np.array_equal(got, pred)
    assert len(cands["a"]) == pred.sum()   # no duplicates: every shape is offered once per tester
    # first-occurrence order: rank follows (first shared cell row-major, then slot) for a freshly built Space
    fy = np.maximum(ay0, by0); fx = np.maximum(ax0, bx0)
    key = (fy * W + fx)[cands["a"], cands["b"]] * w.n + cands["b"]
    for a in np.unique(cands["a"])[:200]:
        s = cands["a"] == a
        assert np.array_equal(np.argsort(key[s], kind="stable"), np.argsort(cands["rank"][s], kind="stable"))


def test_grid_path_equals_brute_force_all_pairs():
    w = worlds.make_mixed(400, space=640, cell=32, filtered=False)   # everything inside the grid
    w.pos = 40.0 + w.pos * (560.0 / 640.0)
    w.pos0 = w.pos.copy()
    w.finalize()
    ob = OracleBatch(w)
    sp = ob.spaces[0]
    tot, cands, hits = ob.frame(1)
    grid = set(zip(hits["a"].tolist(), hits["b"].tolist()))
    brute = set()
    for a in range(w.n):
        for b in range(w.n):
            if a != b and len(sp.intersection(a, b)[1]):
                brute.add((a, b))
    assert grid == brute and len(grid) > 100


def test_world_generators_are_deterministic_and_match_oracle_bounds():
    for mk in (lambda: worlds.make_rects(500, space=1600), lambda: worlds.make_mixed(500, space=1600),
               lambda: worlds.make_bouncer(60), lambda: worlds.make_platformer_batch(2)):
        a, b = mk(), mk()
        assert np.array_equal(a.pos, b.pos) and np.array_equal(a.verts, b.verts) and np.array_equal(a.tags, b.tags)
        ob = OracleBatch(a)
        for sp, idx in zip(ob.spaces, ob.glob):
            lb = np.array([sp.local_bounds(i) for i in range(sp.num_shapes)])
            assert np.array_equal(lb.view(np.uint64), a.local_aabb[idx].view(np.uint64)), a.name
    assert worlds.rand_u01(1, 2, 4, start=3).tolist() == worlds.rand_u01(1, 2, 7)[3:].tolist()
