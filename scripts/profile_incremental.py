"""Small driver for ncu: incremental frames (1 M rectangles, 99 % static) — no timing claims are taken from runs under ncu."""
import sys

sys.path.insert(0, ".")
from resolv_b200 import _abi, space_for_world, worlds  # noqa: E402

w = worlds.make_rects(1_000_000)
ds = space_for_world(w)
static = worlds.rand_u01(77, 0, w.n) < 0.99
ds.set_static(static)
flags = _abi.FRAME_NO_TAG_FILTERS | _abi.FRAME_INCREMENTAL
ds.frame(1, flags | _abi.FRAME_DOWNLOAD, grow=True)
p0 = w.pos.copy()
for f in range(3):
    worlds.advance(w, f + 1)
    w.pos[static] = p0[static]
    ds.set_positions(w.pos)
    c = ds.frame(1, flags | _abi.FRAME_DOWNLOAD)
print(c.as_dict(), ds.static_info())
ds.close()
