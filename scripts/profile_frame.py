"""Small driver for ncu: a few frames of the bench workload (no timing claims are taken from runs under ncu)."""
import argparse
import sys

sys.path.insert(0, ".")
from resolv_b200 import _abi, space_for_world, worlds  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg3")
ap.add_argument("--shapes", type=int, default=1_000_000)
ap.add_argument("--frames", type=int, default=3)
a = ap.parse_args()
w = worlds.make_rects(a.shapes) if a.workload == "cfg2" else worlds.make_mixed(a.shapes)
ds = space_for_world(w)
flags = 0 if (w.use_any.any() or w.none_mask.any()) else _abi.FRAME_NO_TAG_FILTERS   # as bench.py
ds.frame(w.margin, _abi.FRAME_DOWNLOAD | flags, grow=True)                            # sizes the buffers
for f in range(a.frames):
    worlds.advance(w, f + 1)
    ds.set_positions(w.pos)
    c = ds.frame(w.margin, _abi.FRAME_DOWNLOAD | flags)
print(c.as_dict())
ds.close()
