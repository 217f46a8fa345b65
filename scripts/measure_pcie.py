"""What the box's host <-> device path can do with every rank copying at once (torchrun, one rank per GPU): plain
cudaMemcpyAsync between PINNED host memory and the device — H2D alone, D2H alone, both directions together — timed with
CUDA events after a barrier. The e2e leg of bench.py moves 16 B/shape up and ~13-36 B/shape down per step and rank;
this is its ceiling (`e2e.pcie_frac` in the bench line is computed against the committed copy of this output).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29531 scripts/measure_pcie.py
"""
import json
import os

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    nbytes = 64 << 20
    h_up, h_dn = torch.empty(nbytes, dtype=torch.uint8).pin_memory(), torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d_up, d_dn = torch.empty(nbytes, dtype=torch.uint8, device=dev), torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s_up, s_dn = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    reps = 20

    def run(up, dn):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s_up.wait_stream(torch.cuda.current_stream()); s_dn.wait_stream(torch.cuda.current_stream())
        for _ in range(reps):
            if up:
                with torch.cuda.stream(s_up):
                    d_up.copy_(h_up, non_blocking=True)
            if dn:
                with torch.cuda.stream(s_dn):
                    h_dn.copy_(d_dn, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s_up); torch.cuda.current_stream().wait_stream(s_dn)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return reps * nbytes / (float(ms.item()) * 1e-3) / 1e9      # GB/s per rank and direction, slowest rank

    for _ in range(2):
        run(True, True)
    out = {"n_gpus": world, "bytes_per_copy": nbytes, "h2d_alone_gbs_per_rank": run(True, False), "d2h_alone_gbs_per_rank": run(False, True),
           "both_gbs_per_rank_per_direction": run(True, True)}
    out["aggregate_both_directions_gbs"] = 2 * world * out["both_gbs_per_rank_per_direction"]
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
