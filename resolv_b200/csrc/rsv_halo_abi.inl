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
        CU(s, cudaMalloc(&d.n_ghosts, 4));
        CU(s, cudaMemset(d.n_ghosts, 0, 4));
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
                                                                                      (HaloRec*)dev_send_dn, cap_records);
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

// Per-tester (first record, record count) of the last RSV_FRAME_TESTER_TABLE frame, copied to host arrays.
int32_t rsv_tester_table(rsv_space* s, uint32_t* first, uint32_t* count, uint64_t n_shapes) {
    if (!s || !first || !count) return RSV_ERR_BAD_ARG;
    if (n_shapes != s->d.n_shapes) return fail(s, RSV_ERR_BAD_ARG, "rsv_tester_table: expected %u shapes", s->d.n_shapes);
    if (!(s->launched_flags & RSV_FRAME_TESTER_TABLE)) return fail(s, RSV_ERR_BAD_ARG, "rsv_tester_table: last frame ran without RSV_FRAME_TESTER_TABLE");
    CU(s, cudaSetDevice(s->device));
    CU(s, cudaStreamSynchronize(s->stream));
    if (n_shapes) {
        CU(s, cudaMemcpy(first, s->d.tester_start, (size_t)n_shapes * 4, cudaMemcpyDeviceToHost));
        CU(s, cudaMemcpy(count, s->d.tester_count, (size_t)n_shapes * 4, cudaMemcpyDeviceToHost));
    }
    return RSV_OK;
}

}  // extern "C"
