"""How much of a frame is random access by slot: config 3 with the slots in the bench's order (uniformly random in
space) against the same world with the shapes renumbered in row-major order of their home cells (what a host that
inserts a level row by row produces)."""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
from resolv_b200 import _abi, space_for_world, worlds  # noqa: E402


def timed(w, fl):
    ds = space_for_world(w)
    for _ in range(4):
        c = ds.frame(1, fl, grow=True)
    ds.timer_start()
    for _ in range(20):
        ds.frame_launch(1, fl)
    ms = ds.timer_stop() / 20
    ds.drain()
    acc = {}
    for _ in range(5):
        ds.frame(1, fl | _abi.FRAME_TIMING)
        for k, v in ds.timing().items():
            acc[k] = round(acc.get(k, 0.0) + v / 5, 1)
    ds.close()
    return dict(ms=round(ms, 4), kernels_us=acc, hits=int(c.n_hits), gated=int(c.n_gated))


n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
out = {}
for name, mk, fl in (("cfg3", worlds.make_mixed, 0), ("cfg2", worlds.make_rects, _abi.FRAME_NO_TAG_FILTERS)):
    w = mk(n)
    out[name] = {"random_slots": timed(w, fl)}
    b = np.asarray(w.local_aabb).reshape(-1, 4)[:, :2] + w.pos
    cx = np.clip(np.floor(b[:, 0] / w.cell_w), 0, None).astype(np.int64)
    cy = np.clip(np.floor(b[:, 1] / w.cell_h), 0, None).astype(np.int64)
    p = np.argsort(cy * (1 << 20) + cx, kind="stable")
    for k, v in list(w.__dict__.items()):
        if isinstance(v, np.ndarray) and len(v) == n and k != "verts":
            setattr(w, k, np.ascontiguousarray(v[p]))
    out[name]["slots_in_cell_order"] = timed(w, fl)
print(json.dumps(out))
