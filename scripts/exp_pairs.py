import sys
sys.path.insert(0, ".")
from resolv_b200 import _abi, space_for_world, worlds
w = worlds.make_rects(1_000_000)
ds = space_for_world(w)
base = _abi.FRAME_NO_TAG_FILTERS
for name, fl in (("full", 0), ("walk only (no gate/alloc/write)", 0x20000), ("no items (chunk scan + preamble)", 0x10000 | 0x20000)):
    for _ in range(3):
        ds.frame(1, base | fl, grow=True)
    acc = 0.0
    for _ in range(10):
        ds.frame(1, base | fl | _abi.FRAME_TIMING)
        acc += ds.timing()["pairs"] / 10
    print(f"{name:40s} pairs = {acc:7.1f} us")
ds.close()
