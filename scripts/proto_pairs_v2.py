"""CPU prototype (numpy + the oracle as judge) of the candidate-pair scheme planned for k_pairs in unfiltered frames:

  * a tester A only VISITS the entries of its own (un-grown, clamped) cell rect — every shape whose Bounds can pass the
    gate of utils.go:146-173 is registered there — instead of the margin-grown rect (config 2: 2.7 instead of 11 entries
    per tester);
  * the two things the grown walk of cell.go:86-121 also yields are recovered in closed form from three prefix counts
    over the cell-ordered entry array (entries with the first-column flag, with the first-row flag, with both):
      - the number of candidates of A (what `n_candidates` reports), and
      - the generation RANK of a surviving candidate B (its index in possibleIntersections before the distance sort).

The identities: an entry at cell (x, y) of the walk Q is B's first occurrence iff (FC or x == Q.x0) and (FR or y == Q.y0).
So a row y of Q (one contiguous run [s0, e) of the entry array, s1 = end of its first cell) holds
    y == Q.y0:  (s1 - s0) + FC(s1, e)          candidates,      else:  FR(s0, s1) + H(s1, e),
B's first occurrence is the entry of B in cell (max(B.x0, Q.x0), max(B.y0, Q.y0)), and
    rank(B) = sum of the rows above + the same expression cut at B's entry - [A's own home entry comes first].

Run: python scripts/proto_pairs_v2.py     (asserts against the oracle's candidate lists; prints the visit counts)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from resolv_b200 import worlds  # noqa: E402
from tests.util import OracleBatch  # noqa: E402


def check(w, margin):
    ob = OracleBatch(w)
    sp = ob.spaces[0]
    W, H = sp.width_cells, sp.height_cells
    counts, ent = sp.cells()
    counts = np.asarray(counts, np.int64)
    cell = np.concatenate([[0], np.cumsum(counts)])
    R = len(ent)
    rect = np.array([sp.cell_rect(i) for i in range(w.n)], np.int64)           # unclamped (x0, y0, x1, y1)
    cx0, cy0 = np.maximum(rect[:, 0], 0), np.maximum(rect[:, 1], 0)
    cx1, cy1 = np.minimum(rect[:, 2], W - 1), np.minimum(rect[:, 3], H - 1)
    in_grid = (cx0 <= cx1) & (cy0 <= cy1)
    # per entry: cell coordinates and first-column / first-row flags (k_scatter's flag bits)
    ecell = np.repeat(np.arange(W * H), counts)
    ex, ey = ecell % W, ecell // W
    for c in range(W * H):                                                        # in-cell order = ascending slot
        assert np.all(np.diff(ent[cell[c]:cell[c + 1]]) > 0)
    FC = (ex == cx0[ent]).astype(np.int64)
    FR = (ey == cy0[ent]).astype(np.int64)
    pfc = np.concatenate([[0], np.cumsum(FC)])
    pfr = np.concatenate([[0], np.cumsum(FR)])
    ph = np.concatenate([[0], np.cumsum(FC * FR)])

    def entry_of(b, x, y):                                                        # index of b's entry in cell (x, y)
        c = y * W + x
        k = cell[c] + int(np.searchsorted(ent[cell[c]:cell[c + 1]], b))
        assert ent[k] == b
        return k

    def row_count(y, qx0, qx1, qy0, cut=None):
        """Candidates (self included) of row y of the walk, optionally only those before entry index `cut`."""
        s0, s1, e = cell[y * W + qx0], cell[y * W + qx0 + 1], cell[y * W + qx1 + 1]
        if cut is not None:
            e = min(e, cut)
            s1 = min(s1, cut)
        if y == qy0:
            return (s1 - s0) + (pfc[e] - pfc[s1])
        return (pfr[s1] - pfr[s0]) + (ph[e] - ph[s1])

    _, cands, _ = ob.frame(margin)
    oracle = {}
    for a, b, r in zip(cands["a"], cands["b"], cands["rank"]):
        oracle.setdefault(int(a), {})[int(b)] = int(r)

    visited_own = visited_grown = n_checked = 0
    for a in range(w.n):
        if not w.tester[a]:
            continue
        qx0, qx1 = max(rect[a, 0] - margin, 0), min(rect[a, 2] + margin, W - 1)
        qy0, qy1 = max(rect[a, 1] - margin, 0), min(rect[a, 3] + margin, H - 1)
        mine = oracle.get(a, {})
        if not in_grid[a]:
            continue                       # orphan testers keep the full walk (k_pairs' orphan list)
        hk = entry_of(a, cx0[a], cy0[a])
        total = sum(row_count(y, qx0, qx1, qy0) for y in range(qy0, qy1 + 1)) - 1
        assert total == len(mine), (a, total, len(mine))
        visited_grown += sum(cell[y * W + qx1 + 1] - cell[y * W + qx0] for y in range(qy0, qy1 + 1))
        # own-rect walk with the owner-cell rule relative to the own rect
        for y in range(cy0[a], cy1[a] + 1):
            for x in range(cx0[a], cx1[a] + 1):
                for k in range(cell[y * W + x], cell[y * W + x + 1]):
                    visited_own += 1
                    b = int(ent[k])
                    if b == a or not ((FC[k] or x == cx0[a]) and (FR[k] or y == cy0[a])):
                        continue
                    # b's first occurrence in the grown walk
                    fx, fy = max(cx0[b], qx0), max(cy0[b], qy0)
                    kb = k if (fx == x and fy == y) else entry_of(b, fx, fy)
                    rank = sum(row_count(yy, qx0, qx1, qy0) for yy in range(qy0, fy)) + row_count(fy, qx0, qx1, qy0, cut=kb)
                    if cy0[a] < fy or (cy0[a] == fy and hk < kb):
                        rank -= 1
                    assert mine.get(b) == rank, (a, b, rank, mine.get(b))
                    n_checked += 1
        # every candidate whose Bounds touch A's is reachable from the own rect
        ba = sp.bounds(a)
        for b in mine:
            bb = sp.bounds(b)
            if not (bb[2] < ba[0] or bb[0] > ba[2] or bb[3] < ba[1] or bb[1] > ba[3]):
                assert cx0[b] <= cx1[a] and cx1[b] >= cx0[a] and cy0[b] <= cy1[a] and cy1[b] >= cy0[a], (a, b)
    return n_checked, visited_own, visited_grown


if __name__ == "__main__":
    tot = [0, 0, 0]
    for name, w in (("rects dense", worlds.make_rects(1500, space=1024, cell=32)),
                    ("rects big", worlds.make_rects(400, space=512, cell=16, lo=20.0, hi=90.0)),
                    ("bouncer", worlds.make_bouncer())):
        w.use_any[:] = 0
        w.none_mask[:] = 0
        # a few shapes straddling and outside the grid
        w.pos[:12] = np.array([[-5.0, 40.0], [3.0, -2.0], [w.space_w + 4.0, 100.0], [w.space_w - 3.0, w.space_h - 2.0]] * 3)
        for margin in (0, 1, 2, 3):
            r = check(w, margin)
            print(f"{name:12s} margin {margin}: {r[0]:6d} ranks checked, entries visited own-rect {r[1]:7d} vs grown walk {r[2]:7d}")
            for i in range(3):
                tot[i] += r[i]
    print("all ranks and candidate counts reproduced;", tot)
