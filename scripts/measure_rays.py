"""LineTest sweep timing for one library build (RSV_LIB_PATH selects a tuning variant): 4 M rays over the 1 M mixed world."""
import json
import os
import sys

sys.path.insert(0, ".")
from resolv_b200 import _abi, worlds  # noqa: E402

w = worlds.make_mixed(1_000_000)
nr = 4_000_000
ds = _abi.DeviceSpace(w.space_w, w.space_h, w.cell_w, w.cell_h, max_shapes=w.n, max_verts=len(w.verts), max_rays=nr)
ds.upload_world(w)
ds.frame(1, _abi.FRAME_GRID_ONLY)
rays = worlds.make_rays(nr, 5, w.space_w, w.space_h)
out = {"lib": os.path.basename(_abi.LIB_PATH)}
for name, fl in (("rect", 0), ("dda", _abi.LINE_DDA)):
    rc = ds.linetest(rays, flags=fl)
    best = min(ds.linetest(rays, flags=fl | _abi.LINE_RESIDENT).kernel_us for _ in range(3))
    out[name] = dict(kernel_us=best, rays_per_s=nr / best * 1e6, hits=int(rc.n_hits), candidates=int(rc.n_candidates))
ds.close()
print(json.dumps(out))
