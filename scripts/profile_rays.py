"""Small driver for ncu: LineTest over the mixed bench world (no timing claims are taken from runs under ncu)."""
import sys
sys.path.insert(0, ".")
from resolv_b200 import _abi, worlds
n, nr = 1_000_000, int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
w = worlds.make_mixed(n)
ds = _abi.DeviceSpace(w.space_w, w.space_h, w.cell_w, w.cell_h, max_shapes=w.n, max_verts=len(w.verts), max_rays=nr)
ds.upload_world(w)
ds.frame(1, _abi.FRAME_GRID_ONLY)
rays = worlds.make_rays(nr, 5, w.space_w, w.space_h)
for name, fl in (("rect", 0), ("dda", _abi.LINE_DDA)):
    rc = ds.linetest(rays, flags=fl)
    print(name, round(rc.kernel_us), "us", rc.as_dict())
ds.close()
