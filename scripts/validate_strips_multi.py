"""Multi-process, multi-GPU correctness of the strip partition (torchrun, one rank per GPU):
the union of all ranks' results must equal the single-Space result on the same world — same summed counters, same hit
set with the same generation ranks and contact counts, bit-identical MTVs and contacts — over several frames of motion,
through the real transport (CUDA IPC mailboxes over NVLink, exchange fused into the frame).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        scripts/validate_strips_multi.py [--workload cfg3] [--shapes 1000000] [--frames 3]

Rank 0 holds the whole world as one Space on its own GPU (it fits) and does the comparison."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def table(hits, con, to_global):
    live = (hits["count_rank"] & 0x7f) > 0
    h = hits[live]
    n = (h["count_rank"] & 0x7f).astype(np.int64)
    cons = np.concatenate([np.concatenate([con["p"][f:f + k], con["n"][f:f + k]], -1) for f, k in zip(h["first_contact"], n)]) \
        if len(h) else np.zeros((0, 4))
    return dict(a=to_global(h["a"]), b=to_global(h["b"]), n=n, rank=(h["count_rank"] >> 8).astype(np.int64),
                mtv=h["mtv"].copy().view(np.uint64), cons=cons.view(np.uint64))


def main():
    import torch
    import torch.distributed as dist
    from resolv_b200 import _abi, space_for_world, strips
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg3")
    ap.add_argument("--shapes", type=int, default=1_000_000, help="per GPU (cfg2/cfg3) or in total (cfg5)")
    ap.add_argument("--frames", type=int, default=3)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = a.shapes // world if a.workload == "cfg5" else a.shapes
    sr = strips.StripRank(a.workload, n, rank, world, local, margin=1, span_rows=1)
    ok_ipc = sr.connect_ipc(dist)
    ref = ref_world = None
    if rank == 0:
        ref_world = strips.whole_world(a.workload, n, world)
        ref = space_for_world(ref_world, device=local)
    own0 = sr.w.pos[n:2 * n].copy()
    vel = sr.w.vel[n:2 * n].copy()
    report = []
    for f in range(a.frames):
        if f:
            sr.ds.buf[_abi.BUF_POS][2 * n:4 * n] = (own0 + f * vel).reshape(-1)
            sr.ds.commit(1 << _abi.BUF_POS, n, 2 * n)
            if rank == 0:
                ref.set_positions(ref_world.pos + f * ref_world.vel)
        for attempt in range(4):
            sr.step_launch(_abi.FRAME_DOWNLOAD)
            try:
                c = sr.ds.frame_finish()
                over = 0
            except _abi.RsvError as e:
                if e.code != _abi.RSV_ERR_CAPACITY:
                    raise
                lc = sr.ds.last_counts()
                sr.ds.reserve(int(lc.n_entries * 1.3) + 1024, int(lc.n_gated * 1.3) + 1024, int(lc.n_contacts * 1.3) + 1024)
                over = 1
            t = torch.tensor([over], device=torch.device("cuda", local))
            dist.all_reduce(t)
            if int(t.item()) == 0:
                break
        hits, con = sr.ds.results()
        mine = table(hits, con, sr.global_slot)
        mine["counts"] = [int(c.n_testers), int(c.n_candidates), int(c.n_hits), int(c.n_contacts)]
        parts = [None] * world if rank == 0 else None
        dist.gather_object(mine, parts, dst=0)
        if rank == 0:
            cr = ref.frame(1, sr.flags | _abi.FRAME_DOWNLOAD, grow=True)
            rt = table(*ref.results(), lambda x: np.asarray(x, dtype=np.int64))
            summed = np.sum([p["counts"] for p in parts], axis=0).tolist()
            want = [int(cr.n_testers), int(cr.n_candidates), int(cr.n_hits), int(cr.n_contacts)]
            cat = lambda k: np.concatenate([p[k] for p in parts])
            A, B = cat("a"), cat("b")
            o, ro = np.lexsort((B, A)), np.lexsort((rt["b"], rt["a"]))
            same_set = len(A) == len(rt["a"]) and np.array_equal(A[o], rt["a"][ro]) and np.array_equal(B[o], rt["b"][ro])
            res = {"frame": f, "counts_equal": summed == want, "summed": summed, "single_space": want, "hit_set_equal": bool(same_set)}
            if same_set:
                res["ranks_and_contact_counts_equal"] = bool(np.array_equal(cat("n")[o], rt["n"][ro]) and np.array_equal(cat("rank")[o], rt["rank"][ro]))
                res["mtv_bit_identical"] = bool(np.array_equal(cat("mtv")[o], rt["mtv"][ro]))
                # contacts: per hit in canonical order
                def by_hit(tab, order):
                    first = np.cumsum(tab["n"]) - tab["n"]
                    idx = np.concatenate([np.arange(first[i], first[i] + tab["n"][i]) for i in order]) if len(order) else np.zeros(0, np.int64)
                    return tab["cons"][idx]
                allt = {k: cat(k) for k in ("n", "cons")}
                res["contacts_bit_identical"] = bool(np.array_equal(by_hit(allt, o), by_hit(rt, ro)))
            res["ok"] = all(v for k, v in res.items() if isinstance(v, bool))
            report.append(res)
    dist.barrier()
    sr.close()
    if rank == 0:
        ref.close()
        out = {"workload": a.workload, "n_gpus": world, "shapes_total": n * world, "transport": "peer (CUDA IPC), fused" if ok_ipc and sr.fused else ("peer" if ok_ipc else "nccl"),
               "frames": report, "ok": all(r["ok"] for r in report)}
        print(json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
