"""Cell-pair path against the literal walk on the 1 M-shape configs: frame and per-kernel times of both, and a
record-for-record comparison of their results (pairs, ranks, contact counts, MTV and contact bits)."""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
from resolv_b200 import _abi, space_for_world, worlds  # noqa: E402


def canon(ds):
    hits, con = ds.results()
    hits = hits[(hits["count_rank"] & 0x80) == 0]
    o = np.lexsort((hits["b"], hits["a"]))
    h = hits[o]
    n = (h["count_rank"] & 0x7f).astype(np.int64)
    idx = np.concatenate([np.arange(f, f + k) for f, k in zip(h["first_contact"], n) if k]) if n.sum() else np.zeros(0, np.int64)
    return (np.stack([h["a"], h["b"], h["count_rank"]], -1).copy(), h["mtv"].copy().view(np.uint64),
            np.concatenate([con["p"][idx], con["n"][idx]], -1).copy().view(np.uint64))


out = {}
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
for name, w, fl in (("cfg2", worlds.make_rects(n), _abi.FRAME_NO_TAG_FILTERS), ("cfg3", worlds.make_mixed(n), 0)):
    ds = space_for_world(w)
    res = {}
    for path, pf in (("walk", _abi.FRAME_CANDIDATE_WALK), ("cellpairs", 0)):
        for _ in range(4):
            c = ds.frame(1, fl | pf | _abi.FRAME_DOWNLOAD, grow=True)
        ref = canon(ds)
        if path == "walk":
            keep = ref
        else:
            res["identical"] = bool(all(a.shape == b.shape and np.array_equal(a, b) for a, b in zip(keep, ref)))
        ds.timer_start()
        for _ in range(20):
            ds.frame_launch(1, fl | pf)
        ms = ds.timer_stop() / 20
        ds.drain()
        acc = {}
        for _ in range(5):
            ds.frame(1, fl | pf | _abi.FRAME_TIMING)
            for k, v in ds.timing().items():
                acc[k] = round(acc.get(k, 0.0) + v / 5, 1)
        res[path] = dict(ms=round(ms, 4), kernels_us=acc, counts=c.as_dict())
    out[name] = res
    ds.close()
print(json.dumps(out))
