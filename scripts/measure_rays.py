import sys
sys.path.insert(0, ".")
from resolv_b200 import _abi, worlds
w = worlds.make_mixed(1_000_000)
nr = 4_000_000
ds = _abi.DeviceSpace(w.space_w, w.space_h, w.cell_w, w.cell_h, max_shapes=w.n, max_verts=len(w.verts), max_rays=nr)
ds.upload_world(w)
ds.frame(1, _abi.FRAME_GRID_ONLY)
rays = worlds.make_rays(nr, 5, w.space_w, w.space_h)
for name, fl in (("rect", 0), ("dda", _abi.LINE_DDA)):
    rc = ds.linetest(rays, flags=fl)
    rc = ds.linetest(rays, flags=fl | _abi.LINE_RESIDENT)
    print(name, round(rc.kernel_us), "us", round(nr / rc.kernel_us, 1), "M rays/s", rc.as_dict())
ds.close()
