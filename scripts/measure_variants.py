"""Times the candidate-pair and narrowphase stages under the tuning switches of rsv_abi.cu (read at rsv_create; build
variants come in through RSV_LIB_PATH). Same world per workload; counts must agree across variants."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from resolv_b200 import _abi, space_for_world, worlds  # noqa: E402

VARIANTS = [
    ("auto", {}),
    ("coop_never", {"RSV_COOP_MAX_PAIRS": "0"}),
    ("coop_always", {"RSV_COOP_MAX_PAIRS": "4000000000"}),
]
KEYS = ["RSV_COOP_MAX_PAIRS"]


def run(w, name, env, steps=12):
    for k in KEYS:
        os.environ.pop(k, None)
    os.environ.update(env)
    ds = space_for_world(w)
    fl = 0 if (w.use_any.any() or w.none_mask.any()) else _abi.FRAME_NO_TAG_FILTERS
    c = ds.frame(w.margin, fl, grow=True)
    for _ in range(4):
        c = ds.frame(w.margin, fl, grow=True)
    acc = {}
    for _ in range(steps):
        ds.frame(w.margin, fl | _abi.FRAME_TIMING)
        t = ds.timing()
        for k, v in t.items():
            acc[k] = acc.get(k, 0.0) + v / steps
    ds.close()
    return {"variant": name, "pairs_us": round(acc["pairs"], 1), "narrow_us": round(acc["narrow"], 1), "frame_us": round(acc["frame"], 1),
            "counts": [int(c.n_candidates), int(c.n_gated), int(c.n_hits), int(c.n_contacts)]}


def main():
    only = sys.argv[1:]
    out = {}
    for wl, mk in (("cfg2", lambda: worlds.make_rects(1_000_000)), ("cfg3", lambda: worlds.make_mixed(1_000_000)),
                   ("cfg4_512", lambda: worlds.make_platformer_batch(512))):
        if only and wl not in only:
            continue
        w = mk()
        rows = [run(w, n, e) for n, e in VARIANTS]
        ref = rows[0]["counts"]
        for r in rows:
            r["counts_match"] = r["counts"] == ref
            print(wl, json.dumps(r), flush=True)
        out[wl] = rows
    json.dump(out, open(os.path.join("gpurun_out", "measure_variants.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
