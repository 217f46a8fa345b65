"""Frame timing for one library build (RSV_LIB_PATH selects a tuning variant): configs 2 and 3 at 1 M shapes."""
import json
import os
import sys

sys.path.insert(0, ".")
from resolv_b200 import _abi, space_for_world, worlds  # noqa: E402

out = {"lib": os.path.basename(_abi.LIB_PATH)}
for name, w, fl in (("cfg2", worlds.make_rects(1_000_000), _abi.FRAME_NO_TAG_FILTERS), ("cfg3", worlds.make_mixed(1_000_000), 0)):
    ds = space_for_world(w)
    for _ in range(4):
        ds.frame(1, fl, grow=True)
    ds.timer_start()
    for _ in range(20):
        ds.frame_launch(1, fl)
    ms = ds.timer_stop() / 20
    c = ds.frame_finish()
    acc = {}
    for _ in range(5):
        ds.frame(1, fl | _abi.FRAME_TIMING)
        for k, v in ds.timing().items():
            acc[k] = acc.get(k, 0.0) + v / 5
    out[name] = dict(ms=round(ms, 4), pairs=round(acc["pairs"], 1), narrow=round(acc["narrow"], 1), hits=int(c.n_hits), cands=int(c.n_candidates))
    ds.close()
print(json.dumps(out))
