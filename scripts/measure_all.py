"""Secondary measurements for DESIGN.md / profiles (not the bench.py contract): configs 3-5 building blocks on one GPU."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from resolv_b200 import _abi, space_for_world, worlds  # noqa: E402


def frame_time(ds, margin, flags, k=20):
    for _ in range(3):
        ds.frame(margin, flags, grow=True)
    ds.timer_start()
    for _ in range(k):
        ds.frame_launch(margin, flags)
    ms = ds.timer_stop()
    c = ds.frame_finish()
    ds.frame(margin, flags | _abi.FRAME_TIMING)
    return ms / k, c, ds.timing()


out = {}
# config 3: 1M mixed, tag-filtered
w = worlds.make_mixed(1_000_000)
ds = space_for_world(w)
ms, c, t = frame_time(ds, 1, 0)
out["cfg3_1M_mixed"] = dict(ms=ms, shapes_per_s=w.n / ms * 1e3, pair_tests_per_s=c.n_candidates / ms * 1e3, counts=c.as_dict(), kernels_us=t)
# rays over the same world (config 5's sweep, scaled to one GPU): 4M rays per call
ds.frame(1, _abi.FRAME_GRID_ONLY)
nr = 4_000_000
ds.close()
ds = _abi.DeviceSpace(w.space_w, w.space_h, w.cell_w, w.cell_h, max_shapes=w.n, max_verts=len(w.verts), max_rays=nr)
ds.upload_world(w)
ds.frame(1, _abi.FRAME_GRID_ONLY)
rays = worlds.make_rays(nr, 5, w.space_w, w.space_h)
for name, fl in (("rect", 0), ("dda", _abi.LINE_DDA)):
    rc = ds.linetest(rays, flags=fl)
    rc = ds.linetest(rays, flags=fl | _abi.LINE_RESIDENT)
    out[f"rays_{name}_4M"] = dict(kernel_us=rc.kernel_us, rays_per_s=nr / rc.kernel_us * 1e6, counts=rc.as_dict())
ds.close()
# config 5's ray sweep on one GPU: 10 M segments over the 1 M-shape mixed world
nr = 10_000_000
ds = _abi.DeviceSpace(w.space_w, w.space_h, w.cell_w, w.cell_h, max_shapes=w.n, max_verts=len(w.verts), max_rays=nr)
ds.upload_world(w)
ds.frame(1, _abi.FRAME_GRID_ONLY)
rays = worlds.make_rays(nr, 6, w.space_w, w.space_h)
rc = ds.linetest(rays, flags=_abi.LINE_DDA)
rc = ds.linetest(rays, flags=_abi.LINE_DDA | _abi.LINE_RESIDENT)
out["rays_dda_10M"] = dict(kernel_us=rc.kernel_us, rays_per_s=nr / rc.kernel_us * 1e6, counts=rc.as_dict())
ds.close()
# config 4: 512 platformer Spaces on one GPU (4096 / 8): the frame, then four LineTest probes per dynamic body
w = worlds.make_platformer_batch(512)
rays, use_any, any_mask, sidx = worlds.make_platformer_probes(w)
ds = space_for_world(w, max_rays=len(rays))
ms, c, t = frame_time(ds, w.margin, 0)
out["cfg4_512_spaces"] = dict(ms=ms, shapes=w.n, shapes_per_s=w.n / ms * 1e3, testers=int(c.n_testers), pair_tests_per_s=c.n_candidates / ms * 1e3,
                              counts=c.as_dict(), kernels_us=t)
for name, fl in (("rect", 0), ("dda", _abi.LINE_DDA)):
    rc = ds.linetest(rays, use_any=use_any, any_mask=any_mask, flags=fl, space_idx=sidx)
    rc = ds.linetest(rays, use_any=use_any, any_mask=any_mask, flags=fl | _abi.LINE_RESIDENT, space_idx=sidx)
    out[f"cfg4_512_spaces_probes_{name}"] = dict(rays=len(rays), kernel_us=rc.kernel_us, counts=rc.as_dict())
ds.close()
print(json.dumps(out, indent=1))
