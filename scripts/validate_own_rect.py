"""RSV_FRAME_OWN_RECT against the regular candidate walk, on the device: same counts, same records (tester, other,
generation rank), same per-tester order, same MTVs and contacts, bit for bit — then the timing of both on configs 2
and 3 (unfiltered). The regular walk is itself checked against the CPU oracle by tests/test_parity_gpu.py."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from resolv_b200 import _abi, space_for_world, worlds  # noqa: E402

F0 = _abi.FRAME_NO_TAG_FILTERS | _abi.FRAME_DOWNLOAD | _abi.FRAME_TESTER_TABLE


def snapshot(ds, margin, flags):
    c = ds.frame(margin, flags, grow=True)
    hits, con = ds.results()
    hits, con = hits.copy(), con.copy()
    first, count = ds.tester_table()
    return c, hits, con, first.copy(), count.copy()


def canon(hits, con):
    """Records that passed the exact Bounds gate, ordered by (tester, rank). Records that only passed the conservative
    float32 pre-gate (flagged RSV_HIT_NOT_GATED by k_bin, empty sets) are not part of the contract: the regular walk
    also stages such pairs from cells beyond the tester's own rect (boxes within a float32 ulp of each other)."""
    hits = hits[(hits["count_rank"] & 0x80) == 0]
    rank = hits["count_rank"] >> 8
    n = hits["count_rank"] & 0x7f
    order = np.lexsort((rank, hits["a"]))
    h = hits[order]
    idx = np.concatenate([np.arange(f, f + k) for f, k in zip(h["first_contact"], n[order])] + [np.zeros(0, np.int64)]).astype(np.int64)
    return h["a"].copy(), h["b"].copy(), h["count_rank"].copy(), h["mtv"].copy().view(np.uint64), con[idx].copy()


def compare(name, w, margins=(1,), extra=0):
    ds = space_for_world(w)
    if extra & _abi.FRAME_INCREMENTAL:
        ds.set_static(w.tester == 0)
    ok = True
    for m in margins:
        c0, h0, k0, f0, n0 = snapshot(ds, m, F0 | extra)
        c1, h1, k1, f1, n1 = snapshot(ds, m, F0 | extra | _abi.FRAME_OWN_RECT)
        same_counts = all(getattr(c0, f) == getattr(c1, f) for f in ("n_testers", "n_entries", "n_candidates", "n_hits", "n_contacts"))
        a0, a1 = canon(h0, k0), canon(h1, k1)
        same_rec = all(np.array_equal(x, y) for x, y in zip(a0[:4], a1[:4])) and a0[4].tobytes() == a1[4].tobytes()
        # per tester: contiguous, in ascending rank, and the tester table points at them
        r1 = h1["count_rank"] >> 8
        brk = np.nonzero(np.diff(h1["a"].astype(np.int64)) != 0)[0] + 1
        starts = np.concatenate([[0], brk]) if len(h1) else np.zeros(0, np.int64)
        grouped = len(np.unique(h1["a"][starts])) == len(starts) if len(h1) else True
        asc = bool(np.all((np.diff(r1.astype(np.int64)) > 0) | (np.diff(h1["a"].astype(np.int64)) != 0))) if len(h1) > 1 else True
        table = int(n1.sum()) == len(h1) and all(bool(np.all(h1["a"][f1[a]:f1[a] + n1[a]] == a)) for a in np.nonzero(n1)[0][:3000])
        good = same_counts and same_rec and grouped and asc and table
        ok &= good
        print(f"{name:28s} margin {m}: N={c0.n_shapes} P={c0.n_candidates}/{c1.n_candidates} G={c0.n_gated}/{c1.n_gated} "
              f"H={c0.n_hits}/{c1.n_hits} counts={same_counts} records={same_rec} grouped={grouped} ascending={asc} table={table}",
              flush=True)
    ds.close()
    return ok


def timing(name, w, flags, diag=False):
    out = {}
    ds = space_for_world(w)
    O = flags | _abi.FRAME_OWN_RECT
    variants = [("walk", flags), ("own", O)]
    if diag:   # experiment switches of rsv_pairs.cuh (results meaningless, timing only)
        variants += [("prefix+walk", O | 0x80000), ("own-rank", O | 0x20000), ("own-walk-rank", O | 0x40000),
                     ("own-walk-rank-rows", O | 0x40000 | 0x100000)]
    for label, fl in variants:
        for _ in range(4):
            ds.frame(1, fl, grow=True)
        ds.timer_start()
        for _ in range(20):
            ds.frame_launch(1, fl)
        ms = ds.timer_stop() / 20
        ds.frame_finish()
        acc = {}
        for _ in range(5):
            ds.frame(1, fl | _abi.FRAME_TIMING)
            for k, v in ds.timing().items():
                acc[k] = acc.get(k, 0.0) + v / 5
        out[label] = dict(ms=round(ms, 4), pairs_us=round(acc["pairs"], 1), narrow_us=round(acc["narrow"], 1))
    ds.close()
    print(name, json.dumps(out), flush=True)
    return out


if __name__ == "__main__":
    t0 = time.time()
    ok = True
    w = worlds.make_rects(3000, space=1024, cell=32)
    w.pos[:12] = np.array([[-5.0, 40.0], [3.0, -2.0], [1028.0, 100.0], [1021.0, 1022.0]] * 3)   # straddling / outside the grid
    ok &= compare("rects dense + off-grid", w, (0, 1, 2, 3))
    ok &= compare("rects big", worlds.make_rects(600, space=512, cell=16, lo=20.0, hi=90.0), (0, 1, 2))
    ok &= compare("rects crowded (stage overflow)", worlds.make_rects(4000, space=256, cell=32, lo=8.0, hi=40.0), (1,))
    ok &= compare("rects semi-crowded", worlds.make_rects(500, space=512, cell=32), (1, 2))
    ok &= compare("bouncer", worlds.make_bouncer(), (1, 2))
    ok &= compare("mixed unfiltered", worlds.make_mixed(20000, space=4096, cell=32, filtered=False), (1,))
    ok &= compare("platformer batch 16", worlds.make_platformer_batch(16), (1, 4))
    ok &= compare("platformer incremental", worlds.make_platformer_batch(8), (4,), extra=_abi.FRAME_INCREMENTAL)
    ok &= compare("rects 200k", worlds.make_rects(200_000, space=29312, cell=32), (1,))
    ok &= compare("rects tall/wide (> 4 rows)", worlds.make_rects(300, space=1024, cell=16, lo=10.0, hi=150.0), (0, 1, 5))
    ok &= compare("first form: rects dense", w, (1,), extra=0x10000)
    print("PARITY", "OK" if ok else "MISMATCH", f"({time.time() - t0:.1f} s)", flush=True)
    res = {"parity": ok}
    res["cfg2"] = timing("cfg2", worlds.make_rects(1_000_000), _abi.FRAME_NO_TAG_FILTERS, diag=True)
    res["cfg3_unfiltered"] = timing("cfg3u", worlds.make_mixed(1_000_000, filtered=False), _abi.FRAME_NO_TAG_FILTERS)
    print(json.dumps(res), flush=True)
    print(f"total {time.time() - t0:.1f} s")
    sys.exit(0 if ok else 1)
