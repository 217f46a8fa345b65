"""Small driver for ncu: config 4's per-GPU block (512 platformer Spaces in one rsv_space)."""
import sys
sys.path.insert(0, ".")
from resolv_b200 import _abi, space_for_world, worlds
w = worlds.make_platformer_batch(512)
ds = space_for_world(w)
for f in range(4):
    c = ds.frame(w.margin, _abi.FRAME_DOWNLOAD, grow=True)
print(c.as_dict())
ds.close()
