"""torchrun --nproc-per-node 2 scripts/measure_strips.py : where a strip step's time goes (exchange vs frame)."""
import os, sys
sys.path.insert(0, ".")
import torch, torch.distributed as dist
from resolv_b200 import _abi
from resolv_b200.strips import StripRank

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = 1_000_000
sr = StripRank("cfg2", n, rank, world, lr, margin=1, span_rows=1)
ds = sr.ds
for i in range(5):
    sr.pack(); sr.exchange(); sr.apply()
    ds.frame(sr.margin, sr.flags | _abi.FRAME_DOWNLOAD, grow=True)
K = 30
def timed(fn, label):
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(sr.stream)
    for i in range(K):
        fn()
    e1.record(sr.stream)
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / K], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{label:32s} {ms.item() * 1e3:8.1f} us/step", flush=True)
timed(lambda: sr.pack(), "pack only")
timed(lambda: (sr.pack(), sr.exchange()), "pack + exchange")
timed(lambda: (sr.pack(), sr.exchange(), sr.apply()), "pack + exchange + apply")
timed(lambda: ds.frame_launch(sr.margin, sr.flags), "frame only")
ds.frame_finish()
timed(lambda: sr.step_launch(), "full step (NCCL)")
ds.frame_finish()
if sr.connect_ipc(dist):
    for i in range(3):
        sr.halo_step()
        ds.frame(sr.margin, sr.flags | _abi.FRAME_DOWNLOAD, grow=True)
    timed(lambda: sr.exchange_peer(), "peer exchange only")
    timed(lambda: sr.step_launch(), "full step (peer)")
    c = ds.frame_finish()
    if rank == 0:
        print(c.as_dict(), flush=True)
elif rank == 0:
    print("CUDA IPC not available", flush=True)
dist.barrier()
sr.close()
dist.destroy_process_group()
