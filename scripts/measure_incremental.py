"""Incremental grid (RSV_FRAME_INCREMENTAL) against the full rebuild, and the LineTest sweep, on one GPU.
Not the bench.py contract: secondary measurements for DESIGN.md / profiles."""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
from resolv_b200 import _abi, space_for_world, worlds  # noqa: E402


def frame_time(ds, margin, flags, k=20):
    for _ in range(6):
        ds.frame(margin, flags, grow=True)
    ds.timer_start()
    for _ in range(k):
        ds.frame_launch(margin, flags)
    ms = ds.timer_stop()
    c = ds.frame_finish()
    ds.frame(margin, flags | _abi.FRAME_TIMING)
    return ms / k, c, ds.timing()


out = {}
# config 4 per GPU: 512 platformer Spaces, tiles static, bodies dynamic
w = worlds.make_platformer_batch(512)
ds = space_for_world(w)
ds.set_static(w.tester == 0)
for name, fl in (("full", 0), ("incremental", _abi.FRAME_INCREMENTAL)):
    ms, c, t = frame_time(ds, w.margin, fl)
    out[f"cfg4_512_spaces_{name}"] = dict(ms=ms, counts=c.as_dict(), kernels_us=t, static_info=ds.static_info())
ds.close()
# 1 M rectangles (config 2's world), 90 % / 99 % of them static
w = worlds.make_rects(1_000_000)
ds = space_for_world(w)
for frac in (0.9, 0.99):
    static = worlds.rand_u01(77, 0, w.n) < frac
    ds.set_static(static)
    for name, fl in (("full", _abi.FRAME_NO_TAG_FILTERS), ("incremental", _abi.FRAME_NO_TAG_FILTERS | _abi.FRAME_INCREMENTAL)):
        ms, c, t = frame_time(ds, 1, fl)
        out[f"cfg2_1M_static{frac}_{name}"] = dict(ms=ms, counts=c.as_dict(), kernels_us=t, static_info=ds.static_info())
ds.close()
# LineTest sweep over the 1 M mixed world (config 5's rays on one GPU)
w = worlds.make_mixed(1_000_000)
nr = 4_000_000
ds = _abi.DeviceSpace(w.space_w, w.space_h, w.cell_w, w.cell_h, max_shapes=w.n, max_verts=len(w.verts), max_rays=nr)
ds.upload_world(w)
ds.frame(1, _abi.FRAME_GRID_ONLY)
rays = worlds.make_rays(nr, 5, w.space_w, w.space_h)
for name, fl in (("rect", 0), ("dda", _abi.LINE_DDA)):
    rc = ds.linetest(rays, flags=fl)
    rc = ds.linetest(rays, flags=fl | _abi.LINE_RESIDENT)
    out[f"rays_{name}_4M"] = dict(kernel_us=rc.kernel_us, rays_per_s=nr / rc.kernel_us * 1e6, counts=rc.as_dict())
ds.close()
print(json.dumps(out, indent=1))
