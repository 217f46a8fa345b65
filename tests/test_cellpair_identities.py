"""The identities behind the default frame's candidate path (resolv_b200/csrc/rsv_cellpairs.cuh), checked on the CPU
against the oracle's literal walk — no GPU, no CUDA code: a plain-Python restatement of what k_cellpairs and k_rank
compute, fed with the oracle's cell membership.

  (1) Every candidate that passes Bounds().IsIntersecting (utils.go:146-173, with its IsEmpty typo) is found among the
      entries of ONE cell — the first cell the two shapes' clamped rects share — plus, for testers that hold no cell,
      a walk of their margin-grown rect.
  (2) Its index in possibleIntersections (the generation rank) equals the number of candidates in the cells of the
      tester's query rect in front of B's first shared cell (row-major), plus the candidates of that cell with a smaller
      slot — counted with the walk's own first-occurrence rule and tag filters.
"""
import numpy as np
import pytest

from resolv_b200 import worlds
from tests.util import OracleBatch


def _gate(a, o):
    """Bounds.IsIntersecting (utils.go:146-173): a, o = (minx, miny, maxx, maxy)."""
    if o[2] < a[0] or o[0] > a[2] or o[3] < a[1] or o[1] > a[3]:
        return False
    ovmaxx, ovminx, ovmaxy = min(o[2], a[2]), max(o[0], a[0]), min(o[3], a[3])
    return not ((ovmaxx - ovminx) == 0 and (ovmaxy - ovminx) == 0)     # IsEmpty's typo: Max.Y - Min.X


def _cell_pair_records(w, sp, cells, margin):
    W, H = sp.width_cells, sp.height_cells
    n = w.n
    rect = np.array([sp.cell_rect(i) for i in range(n)], dtype=np.int64)            # unclamped, inclusive
    box = np.asarray(w.local_aabb, dtype=np.float64).reshape(-1, 4) + np.concatenate([w.pos, w.pos], 1)
    x0c, y0c = np.maximum(rect[:, 0], 0), np.maximum(rect[:, 1], 0)
    x1c, y1c = np.minimum(rect[:, 2], W - 1), np.minimum(rect[:, 3], H - 1)
    in_grid = (x0c <= x1c) & (y0c <= y1c)

    def passes(a, b):       # ByTags / NotByTags of tester a against b (shapefilter.go:44-61, tags.go:29-31)
        if w.use_any[a] and (int(w.tags[b]) & int(w.any_mask[a])) == 0:
            return False
        return (int(w.tags[b]) & int(w.none_mask[a])) == 0

    def query(a):
        return (max(rect[a, 0] - margin, 0), max(rect[a, 1] - margin, 0), min(rect[a, 2] + margin, W - 1), min(rect[a, 3] + margin, H - 1))

    pairs = []
    # k_cellpairs: cells with two or more entries, every unordered pair, taken at the first cell the rects share
    for c, members in enumerate(cells):
        if len(members) < 2:
            continue
        x, y = c % W, c // W
        for i, a in enumerate(members):
            for b in members[i + 1:]:
                if x != max(x0c[a], x0c[b]) or y != max(y0c[a], y0c[b]) or not _gate(box[a], box[b]):
                    continue
                if w.tester[a] and passes(a, b):
                    pairs.append((a, b))
                if w.tester[b] and passes(b, a):
                    pairs.append((b, a))
    # ... and the testers that hold no cell but whose margin reaches the grid walk their query rect
    for a in np.nonzero(w.tester.astype(bool) & ~in_grid)[0]:
        qx0, qy0, qx1, qy1 = query(a)
        for y in range(qy0, qy1 + 1):
            for x in range(qx0, qx1 + 1):
                for b in cells[y * W + x]:
                    if b != a and x == max(qx0, x0c[b]) and y == max(qy0, y0c[b]) and _gate(box[a], box[b]) and passes(a, b):
                        pairs.append((a, b))

    # k_rank: candidates of the reference's walk in front of B
    def rank(a, b):
        qx0, qy0, qx1, qy1 = query(a)
        fx, fy = max(qx0, x0c[b]), max(qy0, y0c[b])
        r = 0
        for y in range(qy0, fy + 1):
            for x in range(qx0, (qx1 if y < fy else fx) + 1):
                for e in cells[y * W + x]:
                    if y == fy and x == fx and e >= b:
                        continue
                    if e != a and x == max(qx0, x0c[e]) and y == max(qy0, y0c[e]) and passes(a, e):
                        r += 1
        return r

    return sorted((a, b, rank(a, b)) for a, b in pairs), box


@pytest.mark.parametrize("margin", [0, 1, 3])
@pytest.mark.parametrize("world", ["mixed_filtered_offgrid", "rects_dense", "bouncer"])
def test_cell_pairs_and_ranks_equal_the_gated_candidates_of_the_walk(world, margin):
    if world == "mixed_filtered_offgrid":
        w = worlds.make_mixed(500, space=700, cell=32, filtered=True)
        w.pos = w.pos * 1.25 - 90.0          # some shapes straddle or leave the grid: orphan testers
        w.pos0 = w.pos.copy()
        w.tester[::7] = 0
        w.finalize()
    elif world == "rects_dense":
        w = worlds.make_rects(600, space=500, cell=16)
    else:
        w = worlds.make_bouncer(90)          # walls spanning dozens of cells
    ob = OracleBatch(w)
    sp = ob.spaces[0]
    got, box = _cell_pair_records(w, sp, ob.cells(), margin)
    _, cands, _ = ob.frame(margin)
    want = sorted((int(a), int(b), int(r)) for a, b, r in zip(cands["a"], cands["b"], cands["rank"]) if _gate(box[a], box[b]))
    assert len(want) > 20
    assert got == want
