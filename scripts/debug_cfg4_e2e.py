"""Where the time of config 4's end-to-end step goes (host clock around every ABI call)."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from resolv_b200 import _abi, space_for_world, worlds  # noqa: E402

w = worlds.make_platformer_batch(int(sys.argv[1]) if len(sys.argv) > 1 else 4096)
ds = space_for_world(w, max_rays=4 * int(w.tester.sum()))
ds.set_static(w.tester == 0)
L = ds.L
rays, use_any, any_mask, sidx = worlds.make_platformer_probes(w)
n_rays = len(rays)
ds.frame(w.margin, _abi.FRAME_INCREMENTAL | _abi.FRAME_DOWNLOAD, grow=True)
ds.linetest(rays, use_any=use_any, any_mask=any_mask, space_idx=sidx)
bodies = np.nonzero(w.tester != 0)[0].astype(np.uint32)
ls, lxy = ds.pos_list()
ls[:len(bodies)] = bodies
lxy[:2 * len(bodies)] = w.pos[bodies].reshape(-1)
acc = {}
def t(name, fn):
    t0 = time.perf_counter(); r = fn(); acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0; return r
K = 10
for i in range(K + 2):
    if i == 2:
        acc.clear()
    t("commit_pos_list", lambda: ds.commit_pos_list(len(bodies)))
    t("frame_launch", lambda: ds.frame_launch(w.margin, _abi.FRAME_INCREMENTAL | _abi.FRAME_DOWNLOAD))
    t("linetest_launch", lambda: ds._check(L.rsv_linetest_launch(ds.h, n_rays, _abi.FRAME_DOWNLOAD)))
    t("frame_finish", lambda: ds.frame_finish())
    rc = _abi.RayCounts()
    t("linetest_finish", lambda: ds._check(L.rsv_linetest_finish(ds.h, C.byref(rc))))
print({k: round(1e3 * v / K, 3) for k, v in acc.items()}, "ms per step;", int(rc.n_hits), "ray hits", int(rc.n_contacts), "ray contacts")
