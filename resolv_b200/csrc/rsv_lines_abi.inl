// rsv_lines_abi.inl — C ABI of the batched LineTest (included at the end of rsv_abi.cu).

static int32_t lines_ensure(rsv_space* s) {
    LineState& L = s->lines;
    if (L.max_rays) return RSV_OK;
    uint32_t n = s->cfg.max_rays ? s->cfg.max_rays : std::max<uint32_t>(65536u, s->cfg.max_shapes);
    CU(s, cudaHostAlloc(&L.h_seg, (size_t)n * 32, cudaHostAllocDefault));
    CU(s, cudaHostAlloc(&L.h_mask, (size_t)n * 24, cudaHostAllocDefault));
    memset(L.h_mask, 0, (size_t)n * 24);
    CU(s, cudaMalloc(&L.seg, (size_t)n * 32));
    CU(s, cudaMalloc(&L.mask, (size_t)n * 24));
    CU(s, cudaMalloc(&L.ray_first, (size_t)n * 4));
    CU(s, cudaMalloc(&L.ray_count, (size_t)n * 4));
    L.cap_hits = (uint32_t)std::min<unsigned long long>(2ull * n + 4096, 0xfffffff0ull);
    L.cap_contacts = (uint32_t)std::min<unsigned long long>(4ull * n + 4096, 0xfffffff0ull);
    CU(s, cudaMalloc(&L.hits, (size_t)L.cap_hits * sizeof(rsv_ray_hit)));
    CU(s, cudaMalloc(&L.contacts, (size_t)L.cap_contacts * sizeof(rsv_contact)));
    for (auto& e : L.ev) CU(s, cudaEventCreate(&e));
    L.max_rays = n;
    return RSV_OK;
}

extern "C" {

void* rsv_ray_staging(rsv_space* s, int32_t which, size_t* bytes) {
    if (!s || which < 0 || which > 1) return nullptr;
    if (cudaSetDevice(s->device) != cudaSuccess) return nullptr;
    if (lines_ensure(s) != RSV_OK) return nullptr;
    if (bytes) *bytes = (size_t)s->lines.max_rays * (which == 0 ? 32 : 24);
    return which == 0 ? (void*)s->lines.h_seg : (void*)s->lines.h_mask;
}

int32_t rsv_ray_reserve(rsv_space* s, uint32_t max_hits, uint32_t max_contacts) {
    if (!s) return RSV_ERR_BAD_ARG;
    CU(s, cudaSetDevice(s->device));
    int32_t rc = lines_ensure(s);
    if (rc != RSV_OK) return rc;
    CU(s, cudaStreamSynchronize(s->stream));
    LineState& L = s->lines;
    if (max_hits > L.cap_hits) {
        cudaFree(L.hits); L.hits = nullptr; L.cap_hits = 0;
        CU(s, cudaMalloc(&L.hits, (size_t)max_hits * sizeof(rsv_ray_hit)));
        L.cap_hits = max_hits;
    }
    if (max_contacts > L.cap_contacts) {
        cudaFree(L.contacts); L.contacts = nullptr; L.cap_contacts = 0;
        CU(s, cudaMalloc(&L.contacts, (size_t)max_contacts * sizeof(rsv_contact)));
        L.cap_contacts = max_contacts;
    }
    return RSV_OK;
}

int32_t rsv_linetest_launch(rsv_space* s, uint32_t n_rays, uint32_t flags) {
    if (!s) return RSV_ERR_BAD_ARG;
    CU(s, cudaSetDevice(s->device));
    int32_t rc = lines_ensure(s);
    if (rc != RSV_OK) return rc;
    rc = ensure_full_grid(s);   // (the last frame may have taken the cell-pair path: unordered cells, no per-entry boxes)
    if (rc != RSV_OK) return rc;
    LineState& L = s->lines;
    if (n_rays > L.max_rays) return fail(s, RSV_ERR_BAD_ARG, "rsv_linetest: %u rays > max_rays %u", n_rays, L.max_rays);
    cudaStream_t st = s->stream;
    if (!(flags & RSV_LINE_RESIDENT) && n_rays) {
        CU(s, cudaMemcpyAsync(L.seg, L.h_seg, (size_t)n_rays * 32, cudaMemcpyHostToDevice, st));
        CU(s, cudaMemcpyAsync(L.mask, L.h_mask, (size_t)n_rays * 24, cudaMemcpyHostToDevice, st));
    }
    CU(s, cudaMemsetAsync(&s->d.ctr->ray_hit_cursor, 0, sizeof(Counters) - offsetof(Counters, ray_hit_cursor), st));
    CU(s, cudaMemsetAsync(&s->d.ctr->overflow, 0, sizeof(unsigned int), st));
    LinesView v{L.seg, L.mask, L.ray_first, L.ray_count, L.hits, L.contacts, n_rays, L.cap_hits, L.cap_contacts};
    CU(s, cudaEventRecord(L.ev[0], st));
    if (n_rays) {
        int blocks = blocks_for(n_rays, kLinesBlock, s->sm_count * 16);
        if (flags & RSV_LINE_DDA) k_linetest<true><<<blocks, kLinesBlock, 0, st>>>(s->d, v, flags);
        else k_linetest<false><<<blocks, kLinesBlock, 0, st>>>(s->d, v, flags);
    }
    CU(s, cudaEventRecord(L.ev[1], st));
    CU(s, cudaGetLastError());
    CU(s, cudaMemcpyAsync(s->h_ctr, s->d.ctr, sizeof(Counters), cudaMemcpyDeviceToHost, st));
    L.launched_flags = flags;
    L.launched_rays = n_rays;
    L.pending = true;
    return RSV_OK;
}

int32_t rsv_linetest_finish(rsv_space* s, rsv_ray_counts* out) {
    if (!s) return RSV_ERR_BAD_ARG;
    LineState& L = s->lines;
    if (!L.pending) return fail(s, RSV_ERR_BAD_ARG, "rsv_linetest_finish: nothing launched");
    CU(s, cudaSetDevice(s->device));
    CU(s, cudaStreamSynchronize(s->stream));
    L.pending = false;
    const Counters& c = *s->h_ctr;
    rsv_ray_counts rc{};
    rc.n_rays = L.launched_rays;
    rc.n_candidates = c.n_ray_candidates;
    rc.n_hits = c.ray_hit_cursor;
    rc.n_contacts = c.ray_contact_cursor;
    rc.overflow = c.overflow & (RSV_OVF_RAY_HITS | RSV_OVF_RAY_CONTACTS);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, L.ev[0], L.ev[1]);
    rc.kernel_us = ms * 1000.f;
    if (out) *out = rc;
    L.last_hits = L.last_contacts = 0;
    L.last_rays = 0;
    if (rc.overflow)
        return fail(s, RSV_ERR_CAPACITY, "rsv_linetest: capacity exceeded (need hits=%llu contacts=%llu; have %u/%u): rsv_ray_reserve + retry",
                    (unsigned long long)rc.n_hits, (unsigned long long)rc.n_contacts, L.cap_hits, L.cap_contacts);
    if (L.launched_flags & RSV_FRAME_DOWNLOAD) {
        if (L.h_cap_hits < L.cap_hits) {
            if (L.h_hits) cudaFreeHost(L.h_hits);
            L.h_hits = nullptr; L.h_cap_hits = 0;
            CU(s, cudaHostAlloc(&L.h_hits, (size_t)L.cap_hits * sizeof(rsv_ray_hit), cudaHostAllocDefault));
            L.h_cap_hits = L.cap_hits;
        }
        if (L.h_cap_contacts < L.cap_contacts) {
            if (L.h_contacts) cudaFreeHost(L.h_contacts);
            L.h_contacts = nullptr; L.h_cap_contacts = 0;
            CU(s, cudaHostAlloc(&L.h_contacts, (size_t)L.cap_contacts * sizeof(rsv_contact), cudaHostAllocDefault));
            L.h_cap_contacts = L.cap_contacts;
        }
        if (!L.h_ray_first) {
            CU(s, cudaHostAlloc(&L.h_ray_first, (size_t)L.max_rays * 4, cudaHostAllocDefault));
            CU(s, cudaHostAlloc(&L.h_ray_count, (size_t)L.max_rays * 4, cudaHostAllocDefault));
        }
        cudaStream_t st = s->stream;
        if (rc.n_hits) CU(s, cudaMemcpyAsync(L.h_hits, L.hits, (size_t)rc.n_hits * sizeof(rsv_ray_hit), cudaMemcpyDeviceToHost, st));
        if (rc.n_contacts) CU(s, cudaMemcpyAsync(L.h_contacts, L.contacts, (size_t)rc.n_contacts * sizeof(rsv_contact), cudaMemcpyDeviceToHost, st));
        if (rc.n_rays) {
            CU(s, cudaMemcpyAsync(L.h_ray_first, L.ray_first, (size_t)rc.n_rays * 4, cudaMemcpyDeviceToHost, st));
            CU(s, cudaMemcpyAsync(L.h_ray_count, L.ray_count, (size_t)rc.n_rays * 4, cudaMemcpyDeviceToHost, st));
        }
        CU(s, cudaStreamSynchronize(st));
        L.last_hits = rc.n_hits;
        L.last_contacts = rc.n_contacts;
        L.last_rays = (uint32_t)rc.n_rays;
    }
    return RSV_OK;
}

int32_t rsv_linetest(rsv_space* s, uint32_t n_rays, uint32_t flags, rsv_ray_counts* out) {
    int32_t rc = rsv_linetest_launch(s, n_rays, flags);
    if (rc != RSV_OK) return rc;
    return rsv_linetest_finish(s, out);
}

int32_t rsv_ray_results_view(rsv_space* s, rsv_ray_results* out) {
    if (!s || !out) return RSV_ERR_BAD_ARG;
    LineState& L = s->lines;
    out->hits = L.h_hits;
    out->n_hits = L.last_hits;
    out->contacts = L.h_contacts;
    out->n_contacts = L.last_contacts;
    out->ray_first = L.h_ray_first;
    out->ray_count = L.h_ray_count;
    out->n_rays = L.last_rays;
    return RSV_OK;
}

}  // extern "C"
