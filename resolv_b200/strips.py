"""Strip partitioning of one large Space across GPUs (BASELINE.json configs[4]).

One process per GPU. Rank g stores the cell rows of its strip plus a halo, owns the slot range of "its" shapes and,
once per frame, exchanges (slot, position) records with its two neighbours (NCCL send/recv through
torch.distributed — plumbing only; the records are produced / consumed by the library's rsv_halo_pack /
rsv_halo_apply kernels on the same CUDA stream, so a frame needs no host synchronisation).

The reference has no distributed layer; what is preserved is its candidate relation (shape.go:220-237 +
cell.go:86-121): a tester A sees B iff B != A and the cell rects of A grown by `margin` and of B share an in-grid
cell. Rank g therefore needs every shape whose rows reach  [home_lo - margin, home_hi - 1 + span + margin]  where
span is the largest number of extra rows a shape covers.
"""
from __future__ import annotations

import ctypes as C
import os
import time
from dataclasses import dataclass

import numpy as np

HALO_REC_BYTES = 24
NO_UP = -(2 ** 31)      # up_row_max that selects nothing
NO_DN = 2 ** 31 - 1     # dn_row_min that selects nothing


@dataclass
class StripPlan:
    rank: int
    world: int
    H: int                 # grid rows of the whole Space
    home_lo: int           # rows whose shapes are testers on this rank
    home_hi: int
    row_lo: int            # stored rows (home rows + halo)
    row_hi: int
    up_row_max: int        # owned shapes whose first row <= this go to rank-1 ...
    dn_row_min: int        # ... whose last row >= this go to rank+1


def make_plan(rank: int, world: int, H: int, margin: int, span_rows: int) -> StripPlan:
    """Even split of H rows; span_rows = max over shapes of (last row - first row)."""
    base, rem = divmod(H, world)
    los = [r * base + min(r, rem) for r in range(world + 1)]
    lo, hi = los[rank], los[rank + 1]
    halo = margin + span_rows
    if world > 1 and min(los[r + 1] - los[r] for r in range(world)) <= 2 * halo:
        raise ValueError("strips must be thicker than twice the halo")

    def window(r):
        return max(0, los[r] - margin), min(H, los[r + 1] + span_rows + margin)

    row_lo, row_hi = window(rank)
    up_row_max = window(rank - 1)[1] - 1 if rank > 0 else NO_UP
    dn_row_min = window(rank + 1)[0] if rank + 1 < world else NO_DN
    return StripPlan(rank, world, H, lo, hi, row_lo, row_hi, up_row_max, dn_row_min)


def pack_reference(pos, local_aabb, cell_h, own, up_row_max, dn_row_min):
    """numpy restatement of k_halo_pack for the owned slots `own` (used by the CPU/gloo tests)."""
    miny = local_aabb[own, 1] + pos[own, 1]
    maxy = local_aabb[own, 3] + pos[own, 1]
    cy = np.floor(miny / float(cell_h)).astype(np.int64)
    ey = np.floor(maxy / float(cell_h)).astype(np.int64)
    return own[cy <= up_row_max], own[ey >= dn_row_min]


def exchange(send_up, send_dn, recv_up, recv_dn, rank, world):
    """Neighbour exchange of fixed-size halo buffers (torch tensors on any device / backend)."""
    import torch.distributed as dist
    ops = []
    if rank > 0:
        ops += [dist.P2POp(dist.isend, send_up, rank - 1), dist.P2POp(dist.irecv, recv_up, rank - 1)]
    if rank + 1 < world:
        ops += [dist.P2POp(dist.isend, send_dn, rank + 1), dist.P2POp(dist.irecv, recv_dn, rank + 1)]
    if ops:
        for r in dist.batch_isend_irecv(ops):
            r.wait()


def tile_world(kind: str, n: int, tile: int, seed: int = 20265, size: int = 65536):
    """Tile `tile` of the weak-scaling world: the single-GPU workload generator shifted down by tile * size."""
    from . import worlds
    w = (worlds.make_rects(n, seed=seed + 101 * tile, space=size) if kind == "cfg2"
         else worlds.make_mixed(n, seed=seed + 101 * tile, space=size))
    w.pos[:, 1] += tile * float(size)
    w.pos0 = w.pos.copy()
    return w.finalize()


def concat_worlds(ws, space_h):
    """Concatenate tile worlds into one World (slot = tile * n + k)."""
    from . import worlds
    w0 = ws[0]
    cat = np.concatenate
    voff, acc = [], 0
    for w in ws:
        voff.append(w.vert_off + acc)
        acc += len(w.verts)
    out = worlds.World(space_w=w0.space_w, space_h=space_h, cell_w=w0.cell_w, cell_h=w0.cell_h,
                       kind=cat([w.kind for w in ws]), pos=cat([w.pos for w in ws]), pos0=cat([w.pos0 for w in ws]),
                       radius=cat([w.radius for w in ws]), tags=cat([w.tags for w in ws]), nv=cat([w.nv for w in ws]),
                       vert_off=cat(voff).astype(np.uint32), verts=cat([w.verts for w in ws]),
                       closed=cat([w.closed for w in ws]), tester=cat([w.tester for w in ws]),
                       use_any=cat([w.use_any for w in ws]), any_mask=cat([w.any_mask for w in ws]),
                       none_mask=cat([w.none_mask for w in ws]), space_idx=cat([w.space_idx for w in ws]),
                       n_spaces=1, margin=w0.margin, name="tiles",
                       vel=cat([w.vel for w in ws]) if w0.vel is not None else None)
    out.local_aabb = cat([w.local_aabb for w in ws])
    return out


class StripRank:
    """One rank of a strip-partitioned Space: the local DeviceSpace holds tiles rank-1, rank, rank+1 (3n slots; the
    owned range is [n, 2n)), stores rows [row_lo,row_hi) and runs frames on torch's current CUDA stream."""

    def __init__(self, kind, n, rank, world, device, margin=1, span_rows=1, cap_halo=65536, seed=20265, size=65536, max_rays=0):
        import torch
        from . import _abi, worlds
        self.n, self.rank, self.world = n, rank, world
        self.torch = torch
        tiles = []
        for t in (rank - 1, rank, rank + 1):
            w = tile_world(kind, n, min(max(t, 0), world - 1), seed, size)
            if t < 0 or t >= world:
                w.tester[:] = 0
                absent = True
            else:
                absent = False
            tiles.append((w, absent))
        space_h = size * world
        self.w = concat_worlds([t[0] for t in tiles], space_h)
        self.absent = np.concatenate([np.full(n, a) for _, a in tiles])
        cell_h = self.w.cell_h
        H = -(-space_h // cell_h)
        self.plan = make_plan(rank, world, H, margin, span_rows)
        self.ds = _abi.DeviceSpace(self.w.space_w, space_h, self.w.cell_w, cell_h, max_shapes=3 * n,
                                   max_verts=max(len(self.w.verts), 1), device=device, row_lo=self.plan.row_lo,
                                   row_hi=self.plan.row_hi, max_rays=max_rays)
        self.ds.upload_world(self.w)
        meta = self.ds.buf[_abi.BUF_META][:3 * n]
        meta[self.absent] |= _abi.META_ABSENT
        self.ds.commit(1 << _abi.BUF_META)
        self.ds.sync()
        L = self.ds.L
        L.rsv_strip_setup.restype = C.c_int32
        L.rsv_strip_setup.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int32, C.c_int32, C.c_uint32]
        L.rsv_set_stream.restype = C.c_int32
        L.rsv_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        L.rsv_halo_pack.restype = C.c_int32
        L.rsv_halo_pack.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_uint32]
        L.rsv_halo_apply.restype = C.c_int32
        L.rsv_halo_apply.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_uint32]
        L.rsv_halo_mailbox.restype = C.c_int32
        L.rsv_halo_mailbox.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.rsv_ipc_export.restype = C.c_int32
        L.rsv_ipc_export.argtypes = [C.c_void_p, C.c_char_p]
        L.rsv_ipc_open.restype = C.c_int32
        L.rsv_ipc_open.argtypes = [C.c_char_p, C.c_int32, C.POINTER(C.c_void_p)]
        L.rsv_ipc_close.restype = C.c_int32
        L.rsv_ipc_close.argtypes = [C.c_void_p]
        L.rsv_halo_connect.restype = C.c_int32
        L.rsv_halo_connect.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.rsv_halo_exchange.restype = C.c_int32
        L.rsv_halo_exchange.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_uint32]
        self.peer_mode = False
        self._opened = []
        self.device = device
        self.ds._check(L.rsv_strip_setup(self.ds.h, n, 2 * n, self.plan.home_lo, self.plan.home_hi, 2 * cap_halo))
        self.cap = cap_halo
        dev = torch.device("cuda", device)
        words = (cap_halo + 1) * HALO_REC_BYTES // 8
        self.send_up, self.send_dn, self.recv_up, self.recv_dn = (torch.zeros(words, dtype=torch.int64, device=dev) for _ in range(4))
        # a dedicated (non-default) torch stream shared by the library's kernels and the NCCL calls
        self.stream = torch.cuda.Stream(device=dev)
        self.ds._check(L.rsv_set_stream(self.ds.h, C.c_void_p(self.stream.cuda_stream)))
        self.margin = margin
        self.flags = 0 if (self.w.use_any.any() or self.w.none_mask.any()) else _abi.FRAME_NO_TAG_FILTERS

    # --- one frame = pack -> exchange -> apply -> frame
    def pack(self):
        p = self.plan
        up = self.send_up.data_ptr() if self.rank > 0 else None
        dn = self.send_dn.data_ptr() if self.rank + 1 < self.world else None
        self.ds._check(self.ds.L.rsv_halo_pack(self.ds.h, p.up_row_max, p.dn_row_min, up, dn, self.cap))

    def apply(self):
        up = self.recv_up.data_ptr() if self.rank > 0 else None
        dn = self.recv_dn.data_ptr() if self.rank + 1 < self.world else None
        self.ds._check(self.ds.L.rsv_halo_apply(self.ds.h, up, -self.n, dn, self.n, self.cap))

    def exchange(self):
        with self.torch.cuda.stream(self.stream):
            exchange(self.send_up, self.send_dn, self.recv_up, self.recv_dn, self.rank, self.world)

    # --- peer-memory transport (NVLink P2P through CUDA IPC mailboxes): no NCCL call per frame
    def mailbox(self):
        """Allocates this rank's mailbox; returns (device pointer, 64-byte CUDA IPC handle)."""
        ptr, nbytes = C.c_void_p(), C.c_size_t()
        self.ds._check(self.ds.L.rsv_halo_mailbox(self.ds.h, self.cap, C.byref(ptr), C.byref(nbytes)))
        handle = C.create_string_buffer(64)
        rc = self.ds.L.rsv_ipc_export(ptr, handle)
        return ptr.value, (bytes(handle.raw) if rc == 0 else None)

    def connect(self, up_ptr, dn_ptr):
        """Neighbour mailboxes as device pointers valid in this process (None: no neighbour on that side)."""
        self.ds._check(self.ds.L.rsv_halo_connect(self.ds.h, C.c_void_p(up_ptr), C.c_void_p(dn_ptr)))
        self.peer_mode = True

    def connect_ipc(self, dist):
        """All ranks: exchange IPC handles over torch.distributed (host plumbing, once) and map both neighbours'
        mailboxes. Returns False (and stays on the NCCL transport) if CUDA IPC is not available."""
        _, handle = self.mailbox()
        handles = [None] * self.world
        dist.all_gather_object(handles, handle)
        if any(h is None for h in handles):
            return False
        ptrs = {}
        ok = True
        for r in (self.rank - 1, self.rank + 1):
            if 0 <= r < self.world:
                p = C.c_void_p()
                if self.ds.L.rsv_ipc_open(handles[r], self.device, C.byref(p)) != 0:
                    ok = False
                    break
                ptrs[r] = p.value
                self._opened.append(p.value)
        flags = [None] * self.world
        dist.all_gather_object(flags, ok)
        if not all(flags):
            return False
        self.connect(ptrs.get(self.rank - 1), ptrs.get(self.rank + 1))
        dist.barrier()
        return True

    def exchange_peer(self, phases=3):
        """phases: 1 = send half, 2 = receive half, 3 = both (see rsv_halo_exchange)."""
        p = self.plan
        self.ds._check(self.ds.L.rsv_halo_exchange(self.ds.h, p.up_row_max, p.dn_row_min, -self.n, self.n, phases))

    def halo_step(self):
        """pack -> neighbour exchange -> apply, on the rank's stream, by whichever transport is connected."""
        if self.peer_mode:
            self.exchange_peer()
        else:
            self.pack()
            self.exchange()
            self.apply()

    def step_launch(self, extra_flags=0):
        self.halo_step()
        self.ds.frame_launch(self.margin, self.flags | extra_flags)

    def linetest(self, rays, flags=0, **kw):
        """LineTest of a batch of rays against the shapes HOMED on this rank (RSV_LINE_HOME_ONLY): every rank is given
        the same rays, each (ray, shape) set is produced by exactly one rank, and the union over ranks is the result on
        the whole Space. Returns (hits with GLOBAL shape slots, contacts, ray_first, ray_count) of this rank; callers
        gather them at the ray's owner (e.g. dist.all_gather_object) and order each ray's sets by `key`."""
        from . import _abi
        self.ds.linetest(rays, flags=flags | _abi.FRAME_DOWNLOAD | _abi.LINE_HOME_ONLY, **kw)
        hits, con, first, count = self.ds.ray_results()
        hits = hits.copy()
        hits["b"] = self.global_slot(hits["b"]).astype(np.uint32)
        return hits, con.copy(), first.copy(), count.copy()

    def global_slot(self, local):
        return np.asarray(local, dtype=np.int64) + (self.rank - 1) * self.n

    def close(self):
        self.ds.sync()
        for p in self._opened:
            self.ds.L.rsv_ipc_close(C.c_void_p(p))
        self._opened = []
        self.ds.close()


def run_bench(args, rank, world, local_rank):
    """bench.py leg for N > 1 GPUs: weak scaling, args.shapes shapes per GPU."""
    import torch
    import torch.distributed as dist
    from . import _abi, worlds
    from bench import ClockSampler, METRIC, UNIT, workload_name, measured_peaks

    K, W = args.steps, max(args.warmup, 3)
    n = args.shapes
    sr = StripRank(args.workload, n, rank, world, local_rank, margin=1, span_rows=1)
    ds = sr.ds
    dev = torch.device("cuda", local_rank)
    transport = "NCCL send/recv"
    if os.environ.get("RSV_HALO_TRANSPORT", "peer") == "peer" and sr.connect_ipc(dist):
        transport = "peer-memory mailboxes over NVLink (CUDA IPC, stream-ordered waits; no NCCL call per step)"
    # warm-up (also sizes the output buffers: grow on capacity errors)
    for i in range(W):
        sr.halo_step()
        c = ds.frame(sr.margin, sr.flags | _abi.FRAME_DOWNLOAD, grow=True)
    for i in range(max(W, 6)):      # the timed loop's exact call pattern (its flags select its own CUDA graphs)
        sr.step_launch()
    ds.frame_finish()
    sampler = ClockSampler(local_rank)
    sampler.start()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(sr.stream)
    for i in range(K):
        sr.step_launch()
    e1.record(sr.stream)
    torch.cuda.synchronize()
    c = ds.frame_finish()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dist.barrier()
    # e2e, pipelined like the N = 1 leg: every step each rank uploads the positions of its owned range from a pinned
    # page, exchanges halos, runs the frame and downloads its results; the download of step i-1 (copy stream)
    # overlaps the upload + exchange + kernels of step i. Two position sets alternate (pages 0 / 1).
    own_pos = sr.w.pos[n:2 * n].copy()
    vel = sr.w.vel[n:2 * n].copy() if sr.w.vel is not None else np.zeros_like(own_pos)
    pages = [ds.pos_page(0), ds.pos_page(1)]
    pages[0][2 * n:4 * n] = own_pos.reshape(-1)
    pages[1][2 * n:4 * n] = (own_pos + vel).reshape(-1)
    # one sequential warm pass sizes the buffers for both position sets
    for pg in (0, 1):
        ds.commit_pos_page(pg, n, 2 * n)
        sr.halo_step()
        ds.frame(sr.margin, sr.flags | _abi.FRAME_DOWNLOAD, grow=True)
    h2d = d2h = 0

    def launch(i):
        ds.commit_pos_page(i % 2, n, 2 * n)
        sr.halo_step()
        ds.frame_launch(sr.margin, sr.flags | _abi.FRAME_DOWNLOAD)

    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    launch(0)
    for i in range(1, K):
        launch(i)
        ce = ds.frame_finish()
        d2h += 32 * int(ce.n_gated) + 32 * int(ce.n_contacts) + 128
    ce = ds.frame_finish()
    d2h += 32 * int(ce.n_gated) + 32 * int(ce.n_contacts) + 128
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    h2d = 16 * n * K
    te = torch.tensor([t_e2e], device=dev, dtype=torch.float64)
    dist.all_reduce(te, op=dist.ReduceOp.MAX)
    tot = torch.tensor([int(c.n_candidates), int(c.n_hits), int(c.n_contacts), int(c.n_entries), h2d // K, d2h // K], device=dev, dtype=torch.int64)
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    clocks = sampler.stop()
    sr_peer = sr.peer_mode
    dist.barrier()      # (a rank must not unmap a neighbour's mailbox while that neighbour is still running)
    sr.close()
    if rank != 0:
        return None
    ms_per_step = float(ms.item()) / K
    N = n * world
    peak, peak_src = measured_peaks()
    return {
        "metric": METRIC, "value": N / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(args, world) + f"; Space grown to 65536 x {65536 * world}, strip-partitioned by cell rows, "
                   "halo exchange of (slot, position) records with both neighbours every step: " + transport,
                   "l2": "inputs + per-frame intermediates (~0.3 GB per GPU) exceed the 126 MB L2", "margin": 1},
        "pair_tests_per_sec": int(tot[0].item()) / (ms_per_step * 1e-3),
        "counts": {"n_shapes": N, "n_candidates": int(tot[0].item()), "n_hits": int(tot[1].item()), "n_contacts": int(tot[2].item()),
                   "n_entries": int(tot[3].item())},
        "clocks": clocks,
        "e2e": {"value": N * K / float(te.item()), "unit": UNIT, "h2d_bytes_per_step": int(tot[4].item()),
                "d2h_bytes_per_step": int(tot[5].item()), "ms_per_step": 1e3 * float(te.item()) / K},
        # per rank and step: 11 frame kernels + the halo exchange (4 kernels on the peer transport; pack + 2 applies on NCCL)
        "gpu_launches": K * (11 + (4 if sr_peer else 3)) * world,
        "roofline": {"bound": "hbm", "kernel": "frame", "achieved": None, "peak": peak, "unit": "GB/s", "frac": None,
                     "traffic": None, "peak_source": peak_src, "note": "per-kernel roofline is reported by the N=1 run"},
    }


HALO_DTYPE = np.dtype([("slot", np.uint32), ("pad", np.uint32), ("x", np.float64), ("y", np.float64)])


def encode_records(slots, pos, cap):
    """Host-side builder of a halo buffer in the device layout (record 0 = count); used by the CPU tests."""
    buf = np.zeros(cap + 1, dtype=HALO_DTYPE)
    k = min(len(slots), cap)
    buf["slot"][0] = len(slots)
    buf["slot"][1:1 + k] = slots[:k]
    buf["x"][1:1 + k] = pos[slots[:k], 0]
    buf["y"][1:1 + k] = pos[slots[:k], 1]
    return buf.view(np.int64)


def decode_records(words):
    rec = np.asarray(words).view(HALO_DTYPE)
    k = int(rec["slot"][0])
    return rec["slot"][1:1 + k].astype(np.int64), np.stack([rec["x"][1:1 + k], rec["y"][1:1 + k]], -1)
