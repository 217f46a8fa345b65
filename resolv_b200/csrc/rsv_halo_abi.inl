// rsv_halo_abi.inl — C ABI of strip partitioning (included at the end of rsv_abi.cu).

extern "C" {

int32_t rsv_set_stream(rsv_space* s, void* cuda_stream) {
    if (!s) return RSV_ERR_BAD_ARG;
    CU(s, cudaSetDevice(s->device));
    CU(s, cudaStreamSynchronize(s->stream));
    if (cuda_stream) {
        s->stream = (cudaStream_t)cuda_stream;
    } else {
        s->stream = s->own_stream;
    }
    return RSV_OK;
}

int32_t rsv_strip_setup(rsv_space* s, uint32_t own_lo, uint32_t own_hi, int32_t home_lo, int32_t home_hi, uint32_t max_ghosts) {
    if (!s) return RSV_ERR_BAD_ARG;
    if (own_lo > own_hi || own_hi > s->cfg.max_shapes) return fail(s, RSV_ERR_BAD_ARG, "rsv_strip_setup: bad owned range");
    if (home_lo < 0 || home_hi > s->d.H || home_lo >= home_hi) return fail(s, RSV_ERR_BAD_ARG, "rsv_strip_setup: bad home rows");
    CU(s, cudaSetDevice(s->device));
    CU(s, cudaStreamSynchronize(s->stream));
    DevView& d = s->d;
    if (max_ghosts > d.cap_ghosts) {
        if (d.ghosts) cudaFree(d.ghosts);
        d.ghosts = nullptr; d.cap_ghosts = 0;
        CU(s, cudaMalloc(&d.ghosts, (size_t)max_ghosts * 4));
        d.cap_ghosts = max_ghosts;
    }
    if (!d.n_ghosts) {
        CU(s, cudaMalloc(&d.n_ghosts, 8));   // [0] ghosts of the current frame, [1] sticky "a halo batch was truncated" word
        CU(s, cudaMemset(d.n_ghosts, 0, 8));
    }
    d.own_lo = own_lo; d.own_hi = own_hi;
    d.home_lo = home_lo; d.home_hi = home_hi;
    s->strip_mode = true;
    return RSV_OK;
}

int32_t rsv_halo_pack(rsv_space* s, int32_t up_row_max, int32_t dn_row_min, void* dev_send_up, void* dev_send_dn, uint32_t cap_records) {
    if (!s) return RSV_ERR_BAD_ARG;
    CU(s, cudaSetDevice(s->device));
    cudaStream_t st = s->stream;
    if (dev_send_up) CU(s, cudaMemsetAsync(dev_send_up, 0, sizeof(HaloRec), st));
    if (dev_send_dn) CU(s, cudaMemsetAsync(dev_send_dn, 0, sizeof(HaloRec), st));
    const uint32_t n = s->d.own_hi - s->d.own_lo;
    if (n && (dev_send_up || dev_send_dn))
        k_halo_pack<<<blocks_for(n, kGridBlock, s->sm_count * 8), kGridBlock, 0, st>>>(s->d, up_row_max, dn_row_min, (HaloRec*)dev_send_up,
                                                                                      (HaloRec*)dev_send_dn, cap_records,
                                                                                      dev_send_up ? &((HaloRec*)dev_send_up)->slot : nullptr,
                                                                                      dev_send_dn ? &((HaloRec*)dev_send_dn)->slot : nullptr);
    CU(s, cudaGetLastError());
    return RSV_OK;
}

int32_t rsv_halo_apply(rsv_space* s, const void* dev_recv_up, int32_t slot_off_up, const void* dev_recv_dn, int32_t slot_off_dn,
                       uint32_t cap_records) {
    if (!s) return RSV_ERR_BAD_ARG;
    if (!s->strip_mode) return fail(s, RSV_ERR_BAD_ARG, "rsv_halo_apply: call rsv_strip_setup first");
    CU(s, cudaSetDevice(s->device));
    cudaStream_t st = s->stream;
    CU(s, cudaMemsetAsync(s->d.n_ghosts, 0, 4, st));
    const int blocks = blocks_for(cap_records, kGridBlock, s->sm_count * 2);
    if (dev_recv_up) k_halo_apply<<<blocks, kGridBlock, 0, st>>>(s->d, (const HaloRec*)dev_recv_up, slot_off_up, cap_records);
    if (dev_recv_dn) k_halo_apply<<<blocks, kGridBlock, 0, st>>>(s->d, (const HaloRec*)dev_recv_dn, slot_off_dn, cap_records);
    CU(s, cudaGetLastError());
    return RSV_OK;
}

// ---- peer-memory transport: mailbox + IPC plumbing + one stream-ordered exchange call
typedef CUresult (*rsv_stream_wait32_fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
static rsv_stream_wait32_fn stream_wait32() {
    static rsv_stream_wait32_fn fn = []() -> rsv_stream_wait32_fn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        return (rsv_stream_wait32_fn)p;
    }();
    return fn;
}

int32_t rsv_halo_mailbox(rsv_space* s, uint32_t cap_records, void** dev_mailbox, size_t* bytes) {
    if (!s || !cap_records) return RSV_ERR_BAD_ARG;
    if (!s->strip_mode) return fail(s, RSV_ERR_BAD_ARG, "rsv_halo_mailbox: call rsv_strip_setup first");
    CU(s, cudaSetDevice(s->device));
    if (!stream_wait32()) return fail(s, RSV_ERR_CUDA, "rsv_halo_mailbox: the driver has no cuStreamWaitValue32");
    if (s->mailbox && s->mailbox_cap != cap_records) return fail(s, RSV_ERR_BAD_ARG, "rsv_halo_mailbox: capacity is fixed once the mailbox exists");
    if (!s->mailbox) {
        const size_t b = halo_mailbox_bytes(cap_records);
        CU(s, cudaMalloc(&s->mailbox, b));   // (a plain allocation: pool memory cannot be exported through CUDA IPC)
        CU(s, cudaMemset(s->mailbox, 0, b));
        HaloMailboxHeader h{};
        h.cap = cap_records;
        CU(s, cudaMemcpy(s->mailbox, &h, sizeof h, cudaMemcpyHostToDevice));
        s->mailbox_cap = cap_records;
        s->halo_epoch = 0;
        CU(s, cudaDeviceSynchronize());
    }
    if (dev_mailbox) *dev_mailbox = s->mailbox;
    if (bytes) *bytes = halo_mailbox_bytes(cap_records);
    return RSV_OK;
}

int32_t rsv_ipc_export(void* dev_ptr, unsigned char handle[RSV_IPC_HANDLE_BYTES]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == RSV_IPC_HANDLE_BYTES, "IPC handle size");
    if (!dev_ptr || !handle) return RSV_ERR_BAD_ARG;
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, dev_ptr) != cudaSuccess) {
        g_create_error = std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(cudaGetLastError());
        return RSV_ERR_CUDA;
    }
    memcpy(handle, &h, sizeof h);
    return RSV_OK;
}

int32_t rsv_ipc_open(const unsigned char handle[RSV_IPC_HANDLE_BYTES], int32_t device, void** dev_ptr) {
    if (!handle || !dev_ptr) return RSV_ERR_BAD_ARG;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    if (cudaSetDevice(device) != cudaSuccess || cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        g_create_error = std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(cudaGetLastError());
        return RSV_ERR_CUDA;
    }
    return RSV_OK;
}

int32_t rsv_ipc_close(void* dev_ptr) {
    if (!dev_ptr) return RSV_OK;
    return cudaIpcCloseMemHandle(dev_ptr) == cudaSuccess ? RSV_OK : RSV_ERR_CUDA;
}

int32_t rsv_halo_connect(rsv_space* s, void* up_mailbox, void* dn_mailbox) {
    if (!s) return RSV_ERR_BAD_ARG;
    if (!s->mailbox) return fail(s, RSV_ERR_BAD_ARG, "rsv_halo_connect: call rsv_halo_mailbox first");
    CU(s, cudaSetDevice(s->device));
    CU(s, cudaStreamSynchronize(s->stream));
    void* peers[2] = {up_mailbox, dn_mailbox};
    for (int i = 0; i < 2; i++) {
        if (peers[i]) {
            HaloMailboxHeader h{};
            CU(s, cudaMemcpy(&h, peers[i], sizeof h, cudaMemcpyDefault));
            if (h.cap != s->mailbox_cap) return fail(s, RSV_ERR_BAD_ARG, "rsv_halo_connect: neighbour mailbox holds %u records, ours %u", h.cap, s->mailbox_cap);
        }
        s->peer_mailbox[i] = peers[i];
    }
    // epochs restart at 1: clear our flag words (callers synchronise all devices between connect and the first exchange)
    CU(s, cudaMemset(s->mailbox, 0, offsetof(HaloMailboxHeader, cap)));
    CU(s, cudaMemset((char*)s->mailbox + offsetof(HaloMailboxHeader, cnt), 0, sizeof(uint32_t) * 6));   // cnt, epoch, block counters
    s->halo_epoch = 0;
    return RSV_OK;
}

// pack straight into both neighbours' mailboxes -> publish counts + `ready` -> wait for both neighbours' records ->
// apply -> tell them their buffers are free. Four small kernels and a few stream memory waits, all enqueued on the
// space's stream; the host never blocks. Every device of the strip chain has to call this the same number of times
// (epochs), like a collective.
int32_t rsv_halo_exchange(rsv_space* s, int32_t up_row_max, int32_t dn_row_min, int32_t slot_off_up, int32_t slot_off_dn,
                          uint32_t phases) {
    if (!s) return RSV_ERR_BAD_ARG;
    if (!s->mailbox) return fail(s, RSV_ERR_BAD_ARG, "rsv_halo_exchange: call rsv_halo_mailbox / rsv_halo_connect first");
    if (!(phases & (RSV_HALO_SEND | RSV_HALO_RECV))) return fail(s, RSV_ERR_BAD_ARG, "rsv_halo_exchange: no phase selected");
    CU(s, cudaSetDevice(s->device));
    cudaStream_t st = s->stream;
    rsv_stream_wait32_fn wait32 = stream_wait32();
    if (phases & RSV_HALO_SEND) ++s->halo_epoch;
    const uint32_t cap = s->mailbox_cap, e = s->halo_epoch;
    const int par = (int)(e & 1u);
    HaloMailboxHeader* mine = (HaloMailboxHeader*)s->mailbox;
    HaloMailboxHeader* peer[2] = {(HaloMailboxHeader*)s->peer_mailbox[0], (HaloMailboxHeader*)s->peer_mailbox[1]};
    HaloRec* dst[2] = {nullptr, nullptr};   // where our records go: the neighbour's buffer for data from ITS other side
    if (phases & RSV_HALO_SEND) {
    for (int side = 0; side < 2; side++) {  // side: where the neighbour sits (0 upper, 1 lower)
        if (!peer[side]) continue;
        dst[side] = halo_mailbox_recv(peer[side], cap, 1 - side, par);
        // the neighbour must have applied what we stored into this parity two epochs ago
        if (e > 2 && wait32(st, (CUdeviceptr)&mine->done[side][par], e - 2, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
            return fail(s, RSV_ERR_CUDA, "cuStreamWaitValue32 failed");
    }
    const uint32_t n = s->d.own_hi - s->d.own_lo;
    if (n && (dst[0] || dst[1]))
        k_halo_pack<<<blocks_for(n, kGridBlock, s->sm_count * 8), kGridBlock, 0, st>>>(s->d, up_row_max, dn_row_min, dst[0], dst[1], cap,
                                                                                      &mine->cnt[0], &mine->cnt[1]);
    k_halo_publish<<<1, 1, 0, st>>>(mine->cnt, dst[0], dst[1], peer[0] ? &peer[0]->ready[1][par] : nullptr,
                                    peer[1] ? &peer[1]->ready[0][par] : nullptr, e, s->d.n_ghosts);
    }
    if (!(phases & RSV_HALO_RECV)) {
        CU(s, cudaGetLastError());
        return RSV_OK;
    }
    for (int side = 0; side < 2; side++)
        if (peer[side] && wait32(st, (CUdeviceptr)&mine->ready[side][par], e, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
            return fail(s, RSV_ERR_CUDA, "cuStreamWaitValue32 failed");
    k_halo_apply2<<<dim3(blocks_for(cap, kGridBlock, s->sm_count), 2), kGridBlock, 0, st>>>(
        s->d, peer[0] ? halo_mailbox_recv(mine, cap, 0, par) : nullptr, slot_off_up,
        peer[1] ? halo_mailbox_recv(mine, cap, 1, par) : nullptr, slot_off_dn, cap, s->d.n_ghosts + 1);
    k_halo_signal<<<1, 1, 0, st>>>(peer[0] ? &peer[0]->done[1][par] : nullptr, peer[1] ? &peer[1]->done[0][par] : nullptr, e);
    CU(s, cudaGetLastError());
    return RSV_OK;
}

// The frame of a strip-partitioned space with the halo exchange folded into it (peer-memory transport; call
// rsv_halo_mailbox / rsv_halo_connect first). Replaces rsv_halo_exchange + rsv_frame_launch:
//   k_clear -> k_bounds over the owned shapes, which also packs the boundary shapes straight into the neighbours'
//   mailboxes; its last block publishes counts + `ready` flags                                   [RSV_HALO_SEND]
//   k_halo_apply_bounds waits ON THE DEVICE for both neighbours' batches, applies them and registers the ghosts ->
//   scan, scatter, ... narrowphase as in rsv_frame_launch                                           [RSV_HALO_RECV]
// Epochs are kept on the device, so the sequence is the same every frame and is replayed as one CUDA graph. While a
// device waits for its neighbours it has already done its own registration work; the exchange costs one small kernel.
// A space must use either this call or rsv_halo_exchange after rsv_halo_connect, not both.
int32_t rsv_strip_frame_launch(rsv_space* s, int32_t margin, uint32_t flags, int32_t up_row_max, int32_t dn_row_min,
                               int32_t slot_off_up, int32_t slot_off_dn, uint32_t phases) {
    if (!s) return RSV_ERR_BAD_ARG;
    if (!s->mailbox || !s->strip_mode) return fail(s, RSV_ERR_BAD_ARG, "rsv_strip_frame_launch: call rsv_strip_setup / rsv_halo_mailbox / rsv_halo_connect first");
    if (!(phases & (RSV_HALO_SEND | RSV_HALO_RECV))) return fail(s, RSV_ERR_BAD_ARG, "rsv_strip_frame_launch: no phase selected");
    if (margin < 0 || margin > (1 << 20)) return fail(s, RSV_ERR_BAD_ARG, "rsv_frame: margin %d out of range", margin);
    CU(s, cudaSetDevice(s->device));
    const uint32_t cap = s->mailbox_cap;
    HaloArgs h{};
    h.mine = (HaloMailboxHeader*)s->mailbox;
    for (int side = 0; side < 2; side++) {
        HaloMailboxHeader* peer = (HaloMailboxHeader*)s->peer_mailbox[side];
        for (int par = 0; par < 2; par++) {
            h.recv[side][par] = halo_mailbox_recv(s->mailbox, cap, side, par);
            if (!peer) continue;
            h.dst[side][par] = halo_mailbox_recv(peer, cap, 1 - side, par);
            h.peer_ready[side][par] = &peer->ready[1 - side][par];
            h.peer_done[side][par] = &peer->done[1 - side][par];
        }
    }
    h.up_row_max = up_row_max; h.dn_row_min = dn_row_min;
    h.slot_off[0] = slot_off_up; h.slot_off[1] = slot_off_dn;
    h.cap = cap;
    s->d.subset = kSubsetAll;
    DevView& d = s->d;
    cudaStream_t st = s->stream;
    const bool timing = (flags & RSV_FRAME_TIMING) != 0;
    if (phases & RSV_HALO_SEND) {
        rsv_space::FrameSlot& sl0 = s->slot[s->launch_slot];
        if (sl0.pending) {  // both slots busy: the oldest frame was never finished -> its results are dropped
            sl0.pending = false;
            if (s->finish_slot == s->launch_slot) s->finish_slot ^= 1;
        }
    }
    rsv_space::FrameSlot& sl = s->slot[s->launch_slot];
    d.hits = sl.hits; d.contacts = sl.contacts; d.kind_queue = sl.kind_queue;
    int32_t submitted = 0;
    if (s->graphs && !timing && phases == (RSV_HALO_SEND | RSV_HALO_RECV)) {
        submitted = frame_via_graph(s, d, margin, flags | 0x80000000u, sl, s->graph[s->launch_slot][s->cur_pos_page], st, &h);
        if (submitted < 0) return submitted;
    }
    if (!submitted) {
        const int32_t rc = enqueue_frame(s, d, margin, flags, sl, st, timing, &h, phases);
        if (rc != RSV_OK) return rc;
    }
    if (!(phases & RSV_HALO_RECV)) return RSV_OK;
    CU(s, cudaEventRecord(sl.ev_done, st));
    if (s->upload_stream) {
        CU(s, cudaEventRecord(s->ev_page_rd[s->cur_pos_page], st));
        s->page_rd_valid[s->cur_pos_page] = true;
    }
    sl.flags = flags;
    sl.pending = true;
    sl.cellpairs = cell_pair_frame(flags, false);
    sl.pair_list = false;
    s->grid_complete = !sl.cellpairs;
    s->launch_slot ^= 1;
    return RSV_OK;
}

// Per-tester (first record, record count) of the last RSV_FRAME_TESTER_TABLE frame, copied to host arrays.
int32_t rsv_tester_table(rsv_space* s, uint32_t* first, uint32_t* count, uint64_t n_shapes) {
    if (!s || !first || !count) return RSV_ERR_BAD_ARG;
    if (n_shapes != s->d.n_shapes) return fail(s, RSV_ERR_BAD_ARG, "rsv_tester_table: expected %u shapes", s->d.n_shapes);
    if (!(s->launched_flags & RSV_FRAME_TESTER_TABLE) && !s->launched_cellpairs) return fail(s, RSV_ERR_BAD_ARG, "rsv_tester_table: last frame ran without RSV_FRAME_TESTER_TABLE");
    CU(s, cudaSetDevice(s->device));
    CU(s, cudaStreamSynchronize(s->stream));
    if (n_shapes) {
        CU(s, cudaMemcpy(first, s->d.tester_start, (size_t)n_shapes * 4, cudaMemcpyDeviceToHost));
        CU(s, cudaMemcpy(count, s->d.tester_count, (size_t)n_shapes * 4, cudaMemcpyDeviceToHost));
    }
    return RSV_OK;
}

}  // extern "C"
