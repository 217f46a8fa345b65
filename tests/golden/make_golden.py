"""Generates the golden fixtures under tests/golden/ from the CPU oracle (oracle/resolv_oracle.cpp).

The reference (pure Go) cannot be run in this image, so these vectors pin the ORACLE's current behaviour (which is
itself pinned by the hand-derived KATs of tests/test_oracle_kat.py); they guard oracle and CUDA path against drift
and travel to the GPU box, where /root/reference and Go do not exist either.
Usage: python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from resolv_b200 import worlds  # noqa: E402
from tests.util import OracleBatch  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def cases():
    yield "cfg1_bouncer", worlds.make_bouncer(294), 2
    yield "cfg2_rects", worlds.make_rects(1500, space=2560, cell=32), 2
    yield "cfg3_mixed", worlds.make_mixed(1500, space=2560, cell=32), 2
    yield "cfg4_platformer", worlds.make_platformer_batch(2), 2


def frame_record(ob, w):
    tot, cands, hits = ob.frame(w.margin)
    o = np.lexsort((hits["b"], hits["a"]))
    con = np.concatenate([hits["contacts"][f:f + k] for f, k in zip(hits["first"][o], hits["count"][o])]) if len(o) else np.zeros((0, 4))
    co = np.lexsort((cands["b"], cands["a"]))
    return dict(cand_a=cands["a"][co].astype(np.uint32), cand_b=cands["b"][co].astype(np.uint32),
                hit_a=hits["a"][o].astype(np.uint32), hit_b=hits["b"][o].astype(np.uint32), hit_count=hits["count"][o].astype(np.uint32),
                hit_order=hits["order"][o].astype(np.uint32), mtv=hits["mtv"][o], contacts=con)


def main():
    for name, w, frames in cases():
        ob = OracleBatch(w)
        out = {}
        for f in range(frames):
            if f:
                worlds.advance(w, f)
                ob.set_positions(w.pos)
            for k, v in frame_record(ob, w).items():
                out[f"f{f}_{k}"] = v
            out[f"f{f}_pos"] = w.pos.copy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, {k: v.shape for k, v in out.items() if k.startswith("f0")})
    # line tests on the mixed world
    w = worlds.make_mixed(1500, space=2560, cell=32)
    ob = OracleBatch(w)
    rays = worlds.make_rays(400, 11, w.space_w, w.space_h, 16, 300)
    fc, h = ob.spaces[0].linetest(rays, mode="rect")
    o = np.lexsort((h.b, h.a))
    con = np.concatenate([h.contacts[f:f + k] for f, k in zip(h.first[o], h.count[o])])
    np.savez_compressed(os.path.join(HERE, "lines_rect.npz"), rays=rays, ray=h.a[o], b=h.b[o], count=h.count[o], order=h.order[o],
                        mtv=h.mtv[o], center=h.center[o], key=h.dist[o], contacts=con)
    print("lines_rect", len(o))


if __name__ == "__main__":
    main()
