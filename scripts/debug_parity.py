import sys
sys.path.insert(0, '.')
import numpy as np
from resolv_b200 import _abi, space_for_world, worlds
from tests.util import OracleBatch

w = worlds.make_bouncer(294)
ob = OracleBatch(w)
ds = space_for_world(w, max_pairs=64 * w.n, max_contacts=16 * w.n, max_entries=64 * w.n)
c = ds.frame(1, _abi.FRAME_DOWNLOAD | _abi.FRAME_ALL_CANDIDATES)
print(c.as_dict())
tot, cands, oh = ob.frame(1)
hits, con = ds.results()
rank = hits["count_rank"] >> 8
dev = set(zip(hits["a"].tolist(), hits["b"].tolist()))
ora = set(zip(cands["a"].tolist(), cands["b"].tolist()))
print("dev-ora", sorted(dev - ora)[:10], "ora-dev", sorted(ora - dev)[:10], len(dev), len(ora), len(hits))
dr = {(a, b): r for a, b, r in zip(hits["a"].tolist(), hits["b"].tolist(), rank.tolist())}
orr = {(a, b): r for a, b, r in zip(cands["a"].tolist(), cands["b"].tolist(), cands["rank"].tolist())}
bad = [(k, dr[k], orr[k]) for k in dr if k in orr and dr[k] != orr[k]]
print("rank mismatches", len(bad), bad[:10])
if bad:
    a = bad[0][0][0]
    print("tester", a, "rect", ds.debug_cell_rects()[a], "oracle rect", ob.spaces[0].cell_rect(a))
    print("dev order", sorted([(r, b) for (x, b), r in dr.items() if x == a]))
    print("ora order", sorted([(r, b) for (x, b), r in orr.items() if x == a]))
