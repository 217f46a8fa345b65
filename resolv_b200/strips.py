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


def tile_world(kind: str, n: int, tile: int, seed: int = 20265, size: int = 65536, world: int = 1):
    """Tile `tile` (= the shapes owned by rank `tile`) of a strip-partitioned world.
    cfg2 / cfg3 (weak scaling): the single-GPU workload generator shifted down by tile * size.
    cfg5 (strong scaling): bands [tile * 8 / world, (tile + 1) * 8 / world) of the ONE 16M-shape config-5 world
    (worlds.make_cfg5_band), n = shapes per rank."""
    from . import worlds
    if kind == "cfg5":
        per = worlds.CFG5_BANDS // world
        ws = [worlds.make_cfg5_band(b, n * world, seed) for b in range(tile * per, (tile + 1) * per)]
        return ws[0] if len(ws) == 1 else worlds.concat(ws, name=f"cfg5_tile{tile}")
    w = (worlds.make_rects(n, seed=seed + 101 * tile, space=size) if kind == "cfg2"
         else worlds.make_mixed(n, seed=seed + 101 * tile, space=size))
    w.pos[:, 1] += tile * float(size)
    w.pos0 = w.pos.copy()
    return w.finalize()


def whole_world(kind: str, n: int, world: int, seed: int = 20265, size: int = 65536):
    """The same world as ONE Space (slot = tile * n + k): what the union of all ranks must reproduce."""
    from . import worlds
    ws = [tile_world(kind, n, t, seed, size, world) for t in range(world)]
    return worlds.concat(ws, space_h=space_height(kind, world, size), name=f"{kind}_whole")


def space_height(kind: str, world: int, size: int = 65536) -> int:
    from . import worlds
    return worlds.CFG5_SPACE if kind == "cfg5" else size * world


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
            w = tile_world(kind, n, min(max(t, 0), world - 1), seed, size, world)
            if t < 0 or t >= world:
                w.tester[:] = 0
                absent = True
            else:
                absent = False
            tiles.append((w, absent))
        space_h = space_height(kind, world, size)
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
        self.fused = os.environ.get("RSV_HALO_FUSED", "1") != "0"   # (0: the separate rsv_halo_exchange chain of round 1)
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

    def step_launch(self, extra_flags=0, phases=3):
        """One frame of this rank. Peer transport: the fused form (rsv_strip_frame_launch — pack inside k_bounds, waits
        on the device, one CUDA graph per frame); phases 1 / 2 split it for processes that drive several ranks on one
        device. NCCL transport: pack -> send/recv -> apply -> frame."""
        if self.peer_mode and self.fused:
            p = self.plan
            self.ds._check(self.ds.L.rsv_strip_frame_launch(self.ds.h, self.margin, self.flags | extra_flags, p.up_row_max,
                                                            p.dn_row_min, -self.n, self.n, phases))
            return
        if phases & 1:
            self.halo_step()
        if phases & 2:
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


def _summed_counts(c, dev, dist):
    import torch
    t = torch.tensor([int(c.n_testers), int(c.n_candidates), int(c.n_hits), int(c.n_contacts), int(c.n_gated)], device=dev, dtype=torch.int64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return dict(zip(("n_testers", "n_candidates", "n_hits", "n_contacts", "n_gated"), (int(v) for v in t.tolist())))


def run_bench(args, rank, world, local_rank):
    """bench.py leg for N > 1 GPUs. cfg2 / cfg3: weak scaling, args.shapes shapes per GPU in a Space that grows with N.
    cfg5: strong scaling, ONE world of args.shapes shapes in a 262144^2 Space cut into N strips (+ the ray sweep).
    Rank 0 additionally runs the SAME world as one Space on its own GPU and the summed counters of all ranks must equal
    that frame's (`strip_counts_ok`)."""
    import torch
    import torch.distributed as dist
    from . import _abi, worlds, space_for_world
    from bench import ClockSampler, METRIC, UNIT, workload_name, measured_peaks, N_POS_PAGES, pcie_fraction

    K, W = args.steps, max(args.warmup, 3)
    kind = args.workload
    strong = kind == "cfg5"
    n = args.shapes // world if strong else args.shapes
    n_rays = args.rays if strong else 0
    sr = StripRank(kind, n, rank, world, local_rank, margin=1, span_rows=1, max_rays=min(n_rays, 2_000_000))
    ds = sr.ds
    dev = torch.device("cuda", local_rank)
    transport = "NCCL send/recv"
    if os.environ.get("RSV_HALO_TRANSPORT", "peer") == "peer" and sr.connect_ipc(dist):
        transport = ("peer-memory mailboxes over NVLink (CUDA IPC), exchange folded into the frame: pack inside k_bounds, device-side "
                     "flag waits, one CUDA graph per frame" if sr.fused else
                     "peer-memory mailboxes over NVLink (CUDA IPC, stream-ordered waits; no NCCL call per step)")
    # position sets of the owned shapes resident in HBM: the timed frames cycle through them (per-frame motion)
    own0 = sr.w.pos[n:2 * n].copy()
    vel = sr.w.vel[n:2 * n].copy() if sr.w.vel is not None else np.zeros_like(own0)
    sets = [sr.w.pos.copy() for _ in range(N_POS_PAGES)]
    for k in range(N_POS_PAGES):
        sets[k][n:2 * n] = own0 + k * vel
    # warm-up (also sizes the output buffers: grow on capacity errors)
    for i in range(W):
        if sr.peer_mode and sr.fused:
            sr.step_launch(_abi.FRAME_DOWNLOAD)
            try:
                c = ds.frame_finish()
            except _abi.RsvError as e:      # (every rank has done its exchange: epochs stay in step across ranks)
                if e.code != _abi.RSV_ERR_CAPACITY:
                    raise
                lc = ds.last_counts()
                ds.reserve(int(lc.n_entries * 1.3) + 1024, int(lc.n_gated * 1.3) + 1024, int(lc.n_contacts * 1.3) + 1024)
        else:
            sr.halo_step()
            c = ds.frame(sr.margin, sr.flags | _abi.FRAME_DOWNLOAD, grow=True)
    ds.device_pos_pages(sets)
    # ---- the same world as ONE Space on rank 0's GPU: the summed counters of the ranks must equal it (frame of set 0)
    ds.select_pos_page(0)
    sr.step_launch(_abi.FRAME_DOWNLOAD)
    c0 = ds.frame_finish()
    got = _summed_counts(c0, dev, dist)
    check = None
    if not getattr(args, "no_parity", False):
        ok = torch.zeros(1, device=dev, dtype=torch.int64)
        if rank == 0:
            whole = whole_world(kind, n, world)
            one = space_for_world(whole, device=local_rank)
            cw = one.frame(1, sr.flags, grow=True)
            one.close()
            # (default frames take the cell-pair path on both sides: n_candidates is not counted there, n_gated — the
            # Bounds-gated ordered pairs, found exactly once across the strips — is)
            want = {k: int(getattr(cw, k)) for k in ("n_testers", "n_candidates", "n_gated", "n_hits", "n_contacts")}
            check = {"ok": all(got[k] == want[k] for k in want), "ranks_summed": got, "single_space": want}
            ok[0] = int(check["ok"])
            del whole
        dist.broadcast(ok, 0)
        dist.barrier()
    for i in range(3 * N_POS_PAGES):      # the timed loop's exact call pattern (its flags and pages select its own CUDA graphs)
        ds.select_pos_page(i % N_POS_PAGES)
        sr.step_launch()
    ds.drain()
    sampler = ClockSampler(local_rank)
    sampler.start()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(sr.stream)
    for i in range(K):
        ds.select_pos_page(i % N_POS_PAGES)
        sr.step_launch()
    e1.record(sr.stream)
    torch.cuda.synchronize()
    c = ds.frame_finish()
    ds.drain()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dist.barrier()
    # per-kernel durations on this rank over another K frames (events between the kernels; the exchange is part of `bounds`)
    acc = {k: 0.0 for k in _abi.KERNEL_NAMES + ["frame"]}
    for i in range(K):
        ds.select_pos_page(i % N_POS_PAGES)
        sr.step_launch(_abi.FRAME_TIMING)
        ds.frame_finish()
        t = ds.timing()
        for k in acc:
            acc[k] += t[k] / K
    ds.select_pos_page(0)
    # e2e, pipelined like the N = 1 leg: every step each rank uploads the positions of its owned range from a pinned
    # page, exchanges halos, runs the frame and downloads its results; the upload of step i+1 (upload stream) and the
    # download of step i-1 (copy stream) overlap the exchange + kernels of step i. Two position sets alternate.
    pages = [ds.pos_page(0), ds.pos_page(1)]
    pages[0][2 * n:4 * n] = own0.reshape(-1)
    pages[1][2 * n:4 * n] = (own0 + vel).reshape(-1)
    h2d = d2h = 0

    def upload(i):
        ds.pos_page_upload(i % 3, i % 2, n, 2 * n)      # (upload stream: overlaps the kernels of the frame in flight)

    def launch(i):
        ds.select_pos_page(i % 3)
        sr.step_launch(_abi.FRAME_DOWNLOAD)

    for i in range(6):    # warm this call pattern (graphs of the DOWNLOAD flag set)
        upload(i)
        launch(i)
        ds.frame_finish()
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    upload(0)
    launch(0)
    if K > 1:
        upload(1)
    for i in range(1, K):
        launch(i)
        if i + 1 < K:
            upload(i + 1)      # three stages at once: upload of step i+1, kernels of step i, download of step i-1
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
    # ---- cfg5: the ray sweep across strips. Every rank tests all rays against the shapes HOMED on it (RSV_LINE_HOME_ONLY):
    # the union over ranks is the sweep on the whole Space; time = max over ranks of the summed kernel times.
    rays = None
    if n_rays:
        batch = 2_000_000
        ds.select_pos_page(0)
        sr.step_launch()
        ds.frame_finish()
        rays = {"n_rays": n_rays}
        for name, fl in (("rect", 0), ("dda", _abi.LINE_DDA)):
            t_us, hits, done = 0.0, 0, 0
            while done < n_rays:
                m = min(batch, n_rays - done)
                seg = worlds.make_rays(m, 20265 + done, sr.w.space_w, sr.w.space_h)
                rc = ds.linetest(seg, flags=fl | _abi.LINE_HOME_ONLY)
                t_us += rc.kernel_us
                hits += int(rc.n_hits)
                done += m
            rt = torch.tensor([t_us], device=dev, dtype=torch.float64)
            dist.all_reduce(rt, op=dist.ReduceOp.MAX)
            rh = torch.tensor([hits], device=dev, dtype=torch.int64)
            dist.all_reduce(rh, op=dist.ReduceOp.SUM)
            rays[name] = {"kernel_ms": float(rt.item()) / 1e3, "rays_per_sec": n_rays / (float(rt.item()) * 1e-6), "n_hits": int(rh.item())}
        rays["iterator"] = ("RECT == space.FilterCells(bounds of the cast line).FilterShapes(); every rank tests all rays against the shapes "
                            "homed on it (RSV_LINE_HOME_ONLY): the union over ranks is the sweep on the whole Space")
        rays["kernel_ms"], rays["rays_per_sec"], rays["n_hits"] = rays["rect"]["kernel_ms"], rays["rect"]["rays_per_sec"], rays["rect"]["n_hits"]
    sr_peer, sr_fused = sr.peer_mode, sr.fused
    dist.barrier()      # (a rank must not unmap a neighbour's mailbox while that neighbour is still running)
    sr.close()
    if rank != 0:
        return None
    ms_per_step = float(ms.item()) / K
    N = n * world
    peak, peak_src = measured_peaks()
    out = {
        "metric": METRIC, "value": N / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(args, world) + ("" if strong else f"; Space grown to 65536 x {65536 * world}") +
                   ", strip-partitioned by cell rows, halo exchange of (slot, position) records with both neighbours every step: " + transport,
                   "l2": "inputs + per-frame intermediates (~0.3 GB per GPU) exceed the 126 MB L2", "margin": 1,
                   "motion": f"timed frames cycle through {N_POS_PAGES} position sets resident in HBM"},
        "pair_tests_per_sec": int(tot[0].item()) / (ms_per_step * 1e-3),
        "counts": {"n_shapes": N, "n_candidates": int(tot[0].item()), "n_hits": int(tot[1].item()), "n_contacts": int(tot[2].item()),
                   "n_entries": int(tot[3].item())},
        "strip_counts_ok": None if check is None else check["ok"], "strip_check": check,
        "clocks": clocks,
        "kernels_us_rank0": {k: round(v, 2) for k, v in acc.items()},
        "e2e": dict({"value": N * K / float(te.item()), "unit": UNIT, "h2d_bytes_per_step": int(tot[4].item()),
                     "d2h_bytes_per_step": int(tot[5].item()), "ms_per_step": 1e3 * float(te.item()) / K},
                    **pcie_fraction(world, int(tot[4].item()), int(tot[5].item()), float(te.item()) / K)),
        # per rank and step: 11 frame kernels + the exchange (fused: 1 extra kernel; separate peer chain: 4; NCCL: pack + 2 applies)
        "gpu_launches": K * (11 + (1 if (sr_peer and sr_fused) else (4 if sr_peer else 3))) * world,
        "roofline": {"bound": "hbm", "kernel": "frame", "achieved": None, "peak": peak, "unit": "GB/s", "frac": None,
                     "traffic": None, "peak_source": peak_src, "note": "per-kernel roofline is reported by the N=1 run"},
    }
    if rays:
        out["rays"] = rays
    return out


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
