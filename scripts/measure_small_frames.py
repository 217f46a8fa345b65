"""Frame / k_pairs timing on small and medium worlds for one library build (RSV_LIB_PATH selects a tuning variant)."""
import json
import os
import sys

sys.path.insert(0, ".")
from resolv_b200 import _abi, space_for_world, worlds  # noqa: E402

out = {"lib": os.path.basename(_abi.LIB_PATH)}
for name, w, fl in (("cfg4_512", worlds.make_platformer_batch(512), 0), ("cfg4_64", worlds.make_platformer_batch(64), 0),
                    ("bouncer", worlds.make_bouncer(294), _abi.FRAME_NO_TAG_FILTERS),
                    ("cfg2_100k", worlds.make_rects(100_000, space=20736), _abi.FRAME_NO_TAG_FILTERS),
                    ("cfg2_300k", worlds.make_rects(300_000, space=35904), _abi.FRAME_NO_TAG_FILTERS),
                    ("cfg2_400k", worlds.make_rects(400_000, space=41440), _abi.FRAME_NO_TAG_FILTERS),
                    ("cfg2_500k", worlds.make_rects(500_000, space=46336), _abi.FRAME_NO_TAG_FILTERS),
                    ("cfg2_650k", worlds.make_rects(650_000, space=52832), _abi.FRAME_NO_TAG_FILTERS),
                    ("cfg2_1M", worlds.make_rects(1_000_000), _abi.FRAME_NO_TAG_FILTERS)):
    ds = space_for_world(w)
    for _ in range(6):
        ds.frame(w.margin, fl, grow=True)
    ds.timer_start()
    for _ in range(20):
        ds.frame_launch(w.margin, fl)
    ms = ds.timer_stop() / 20
    ds.frame_finish()
    acc = 0.0
    for _ in range(5):
        ds.frame(w.margin, fl | _abi.FRAME_TIMING)
        acc += ds.timing()["pairs"] / 5
    out[name] = dict(us=round(1e3 * ms, 1), pairs=round(acc, 1))
    ds.close()
print(json.dumps(out))
