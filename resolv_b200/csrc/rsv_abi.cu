// rsv_abi.cu — the C ABI of libresolv_b200.so (include/resolv_b200.h): memory ownership, staging, frame
// orchestration on one CUDA stream per space, result mirrors, timing and parity dumps.
//
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false
//             -Xcompiler -fPIC -shared rsv_abi.cu -o libresolv_b200.so
// -fmad=false is part of the numerical contract (Go/amd64 does not fuse multiply-add), see rsv_math.cuh.
#include <cuda.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/resolv_b200.h"
#include "rsv_device.cuh"
#include "rsv_grid.cuh"
#include "rsv_halo.cuh"
#include "rsv_lines.cuh"
#include "rsv_narrow.cuh"
#include "rsv_pairs.cuh"
#include "rsv_cellpairs.cuh"

using namespace rsv;

static thread_local std::string g_create_error;

struct rsv_space {
    rsv_config cfg{};
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;      // stream every kernel of this space is launched on
    cudaStream_t own_stream = nullptr;  // the library-created stream (stream == own_stream unless rsv_set_stream)
    bool strip_mode = false;
    DevView d{};
    // device allocations behind the const views
    void* dev_buf[RSV_BUF_COUNT] = {};
    void* host_buf[RSV_BUF_COUNT] = {};
    size_t buf_bytes[RSV_BUF_COUNT] = {};
    size_t elem_bytes[RSV_BUF_COUNT] = {};
    uint32_t n_shapes = 0, n_verts = 0;
    uint32_t n_tiles = 0;          // scan tiles of the cell array
    bool grid_complete = true;     // false after a cell-pair frame: cells are unordered and carry no per-entry boxes / tags / home records
    // Two frames can be in flight: frame f's results are copied to the host on `copy_stream` while frame f+1's
    // upload and kernels run on `stream`, so each slot has its own result buffers, pinned mirrors and counters.
    struct FrameSlot {
        rsv_hit* hits = nullptr;
        rsv_contact* contacts = nullptr;
        uint32_t* kind_queue = nullptr;
        Counters* h_ctr = nullptr;  // pinned
        rsv_hit* h_hits = nullptr;  // pinned mirrors, allocated on first download
        rsv_contact* h_contacts = nullptr;
        uint32_t h_cap_pairs = 0, h_cap_contacts = 0;
        cudaEvent_t ev_done = nullptr;
        uint32_t flags = 0;
        bool pending = false;
        bool cellpairs = false;     // the frame took the cell-pair path (tester table always filled, n_candidates not counted)
        bool pair_list = false;     // rsv_test_pairs, not a frame
    } slot[2];
    int launch_slot = 0, finish_slot = 0, last_slot = 0;
    // captured launch sequence of a frame, one per result slot (frame_via_graph)
    struct FrameGraph {
        struct Key {
            int32_t margin; uint32_t flags; cudaStream_t stream; uint32_t n_tiles; int blocks[3]; HaloArgs halo;
        } key{};
        DevView view{};
        cudaGraphExec_t exec = nullptr;
        int seen = 0;
    } graph[2][4];   // per result slot and device position page
    void* pos_page[4] = {};   // device position pages (rsv_pos_pages); page 0 is dev_buf[RSV_BUF_POS]
    int n_pos_pages = 1, cur_pos_page = 0;
    // uploads into a position page run on their own stream so that they overlap the kernels of the frame in flight:
    // ev_page_up[p] = last upload into page p, ev_page_rd[p] = last frame launched on page p
    cudaStream_t upload_stream = nullptr;
    cudaEvent_t ev_page_up[4] = {}, ev_page_rd[4] = {};
    bool page_up_pending[4] = {}, page_rd_valid[4] = {};
    bool graphs = true;  // RSV_NO_GRAPH=1 in the environment turns graph replay off
    cudaStream_t copy_stream = nullptr;
    cudaStream_t side_stream = nullptr;  // second narrowphase kernel
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    Counters* h_ctr = nullptr;       // pinned counters of the line tests
    void* host_pos_page1 = nullptr;  // second pinned staging page for RSV_BUF_POS (rsv_staging_page)
    // sparse position updates (rsv_commit_pos_list): pinned (slot, x, y) lists and their device copies
    uint32_t* h_pair_list = nullptr; uint32_t* d_pair_list = nullptr; uint32_t cap_pair_list = 0;   // rsv_test_pairs
    uint32_t* h_list_slots = nullptr; double* h_list_xy = nullptr;
    uint32_t* d_list_slots = nullptr; double2* d_list_xy = nullptr;
    uint32_t launched_flags = 0;     // flags of the last finished frame
    bool launched_cellpairs = false; // ... and whether it took the cell-pair path
    rsv_counts last_counts{};
    uint64_t last_records = 0, last_contacts = 0;
    cudaEvent_t ev[RSV_K_COUNT + 1] = {};
    cudaEvent_t ev_timer[2] = {};
    rsv_timing_info timing{};
    int narrow_blocks = 148 * 4;
    int narrow_blocks1 = 148 * 4;
    int narrow_order = 1;           // 1: submit k_narrow<0> before k_narrow<1> (RSV_NARROW_ORDER=0: the other way round)
    int narrow_blocks0_fixed = 0;   // RSV_NARROW_BLOCKS0 given: no adaptive grid for k_narrow<0>
    int narrow_per_sm = 0;          // blocks per SM of k_narrow<0> chosen from the last finished frames (0: not known yet)
    int pairs_blocks[2] = {148 * 8, 148 * 8};  // k_pairs<false> (no tag filters), k_pairs<true>
    // peer-memory halo transport (rsv_halo_abi.inl)
    void* mailbox = nullptr;            // this device's mailbox
    uint32_t mailbox_cap = 0;
    void* peer_mailbox[2] = {};         // neighbours' mailboxes as mapped in this process: 0 upper, 1 lower
    uint32_t halo_epoch = 0;
    // incremental grid (RSV_FRAME_INCREMENTAL): cached registrations of the RSV_META_STATIC shapes
    StaticGrid sg{};
    bool static_valid = false;
    // line tests
    LineState lines{};
    std::string error;
};

static int32_t fail(rsv_space* s, int32_t code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (s) s->error = buf;
    else g_create_error = buf;
    return code;
}

#define CU(s, call)                                                                                   \
    do {                                                                                              \
        cudaError_t e__ = (call);                                                                     \
        if (e__ != cudaSuccess)                                                                       \
            return fail((s), e__ == cudaErrorMemoryAllocation ? RSV_ERR_OOM : RSV_ERR_CUDA, "%s: %s", #call, \
                        cudaGetErrorString(e__));                                                     \
    } while (0)

static inline int blocks_for(uint32_t n, int block, int cap) {
    long long b = ((long long)n + block - 1) / block;
    if (b < 1) b = 1;
    if (b > cap) b = cap;
    return (int)b;
}

static void drop_graph(rsv_space::FrameGraph& g) {
    if (g.exec) cudaGraphExecDestroy(g.exec);
    g.exec = nullptr;
    g.seen = 0;
}

static size_t buf_elem_bytes(int which) {
    switch (which) {
        case RSV_BUF_POS: return 16;
        case RSV_BUF_LOCAL_AABB: return 32;
        case RSV_BUF_RADIUS: return 8;
        case RSV_BUF_TAGS: return 8;
        case RSV_BUF_META: return 4;
        case RSV_BUF_VERT_OFF: return 4;
        case RSV_BUF_VERTS: return 16;
        case RSV_BUF_QMASK: return 16;
        case RSV_BUF_SPACE_IDX: return 4;
    }
    return 0;
}

static void bind_views(rsv_space* s) {
    DevView& d = s->d;
    d.pos = (const double2*)s->dev_buf[RSV_BUF_POS];
    d.laabb = (const Box*)s->dev_buf[RSV_BUF_LOCAL_AABB];
    d.radius = (const double*)s->dev_buf[RSV_BUF_RADIUS];
    d.tags = (const unsigned long long*)s->dev_buf[RSV_BUF_TAGS];
    d.meta = (const uint32_t*)s->dev_buf[RSV_BUF_META];
    d.vert_off = (const uint32_t*)s->dev_buf[RSV_BUF_VERT_OFF];
    d.verts = (const double2*)s->dev_buf[RSV_BUF_VERTS];
    d.qmask = (const ulonglong2*)s->dev_buf[RSV_BUF_QMASK];
    d.space_idx = (const uint32_t*)s->dev_buf[RSV_BUF_SPACE_IDX];
}

extern "C" {

int32_t rsv_abi_version(void) { return RSV_ABI_VERSION; }

int32_t rsv_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char* rsv_last_error(const rsv_space* s) { return s ? s->error.c_str() : g_create_error.c_str(); }

static int32_t alloc_outputs(rsv_space* s, uint32_t cap_entries, uint32_t cap_pairs, uint32_t cap_contacts) {
    DevView& d = s->d;
    if (cap_entries > d.cap_entries) {
        if (d.entries) cudaFree(d.entries);
        if (d.ent_box) cudaFree(d.ent_box);
        if (d.ent_home) cudaFree(d.ent_home);
        if (d.ent_tags) cudaFree(d.ent_tags);
        d.entries = nullptr; d.ent_box = nullptr; d.ent_home = nullptr; d.ent_tags = nullptr; d.cap_entries = 0;
        CU(s, cudaMalloc(&d.ent_tags, ((size_t)cap_entries + 8) * 8));
        if (d.multi) cudaFree(d.multi);
        d.multi = nullptr;
        d.cap_multi = cap_entries + 1;  // (incremental frames list every cell that received a dynamic registration)
        CU(s, cudaMalloc(&d.multi, (size_t)d.cap_multi * sizeof(uint2)));
        CU(s, cudaMalloc(&d.entries, ((size_t)cap_entries + 8) * 4));
        CU(s, cudaMalloc(&d.ent_box, ((size_t)cap_entries + 8) * sizeof(float4)));
        CU(s, cudaMalloc(&d.ent_home, ((size_t)cap_entries + 8) * sizeof(uint2)));
        d.cap_entries = cap_entries;
    }
    if (cap_pairs > d.cap_pairs) {
        d.cap_pairs = 0;
        for (auto& sl : s->slot) {
            if (sl.hits) cudaFree(sl.hits);
            if (sl.kind_queue) cudaFree(sl.kind_queue);
            sl.hits = nullptr; sl.kind_queue = nullptr;
            CU(s, cudaMalloc(&sl.hits, (size_t)cap_pairs * sizeof(rsv_hit)));
            CU(s, cudaMalloc(&sl.kind_queue, (size_t)cap_pairs * sizeof(uint32_t)));
        }
        if (d.ptmp) cudaFree(d.ptmp);
        if (d.pgrp) cudaFree(d.pgrp);
        d.ptmp = nullptr; d.pgrp = nullptr;
        CU(s, cudaMalloc(&d.ptmp, (size_t)cap_pairs * sizeof(uint4)));
        CU(s, cudaMalloc(&d.pgrp, (size_t)cap_pairs * sizeof(uint4)));
        d.cap_pairs = cap_pairs;
    }
    if (cap_contacts > d.cap_contacts) {
        d.cap_contacts = 0;
        for (auto& sl : s->slot) {
            if (sl.contacts) cudaFree(sl.contacts);
            sl.contacts = nullptr;
            CU(s, cudaMalloc(&sl.contacts, (size_t)cap_contacts * sizeof(rsv_contact)));
        }
        d.cap_contacts = cap_contacts;
    }
    d.hits = s->slot[0].hits; d.kind_queue = s->slot[0].kind_queue; d.contacts = s->slot[0].contacts;
    return RSV_OK;
}

int32_t rsv_create(const rsv_config* cfg, rsv_space** out) {
    if (!cfg || !out) return fail(nullptr, RSV_ERR_BAD_ARG, "rsv_create: null argument");
    *out = nullptr;
    if (cfg->space_w <= 0 || cfg->space_h <= 0 || cfg->cell_w <= 0 || cfg->cell_h <= 0)
        return fail(nullptr, RSV_ERR_BAD_ARG, "rsv_create: space and cell sizes must be positive");
    if (cfg->max_shapes == 0 || cfg->max_shapes >= (1u << 26))
        return fail(nullptr, RSV_ERR_BAD_ARG, "rsv_create: max_shapes must be in [1, 2^26)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(nullptr, RSV_ERR_NO_DEVICE, "rsv_create: no CUDA device (there is no CPU fallback)");
    }
    if (cfg->device < 0 || cfg->device >= ndev)
        return fail(nullptr, RSV_ERR_BAD_ARG, "rsv_create: device %d out of range (have %d)", cfg->device, ndev);
    rsv_space* s = new (std::nothrow) rsv_space;
    if (!s) return fail(nullptr, RSV_ERR_OOM, "rsv_create: out of host memory");
    s->cfg = *cfg;
    s->device = cfg->device;
    // Grid dimensions: NewSpace -> Resize(ceil(w/cw), ceil(h/ch)) (space.go:23)
    int W = (int)std::ceil((double)cfg->space_w / (double)cfg->cell_w);
    int H = (int)std::ceil((double)cfg->space_h / (double)cfg->cell_h);
    int row_lo = 0, row_hi = H;
    if (!(cfg->row_lo == 0 && cfg->row_hi == 0)) {
        row_lo = cfg->row_lo; row_hi = cfg->row_hi;
        if (row_lo < 0 || row_hi > H || row_lo >= row_hi) {
            delete s;
            return fail(nullptr, RSV_ERR_BAD_ARG, "rsv_create: strip [%d,%d) not inside [0,%d)", row_lo, row_hi, H);
        }
    }
    // (home records pack a shape's clamped column / row span into 16 bits each: rsv_grid.cuh k_entboxes)
    if (W > 65536 || row_hi - row_lo > 65536) {
        delete s;
        return fail(nullptr, RSV_ERR_BAD_ARG, "rsv_create: %d x %d cells: at most 65536 columns and 65536 stored rows are supported", W, row_hi - row_lo);
    }
    uint32_t n_spaces = cfg->n_spaces ? cfg->n_spaces : 1;
    unsigned long long cps = (unsigned long long)W * (unsigned long long)(row_hi - row_lo);
    unsigned long long n_cells = cps * n_spaces;
    if (n_cells + 1 >= (1ull << 32)) {
        delete s;
        return fail(nullptr, RSV_ERR_BAD_ARG, "rsv_create: %llu cells exceed the 32-bit cell index", n_cells);
    }
    DevView& d = s->d;
    d.W = W; d.H = H; d.row_lo = row_lo; d.row_hi = row_hi;
    d.cell_w = (double)cfg->cell_w; d.cell_h = (double)cfg->cell_h;
    d.cells_per_space = (uint32_t)cps; d.n_cells = (uint32_t)n_cells; d.n_spaces = n_spaces;
    d.n_shapes = 0;

    auto bail = [&](int32_t code) {
        std::string msg = s->error;
        rsv_destroy(s);
        g_create_error = msg;
        return code;
    };
#define CUC(call)                                                                                                  \
    do {                                                                                                           \
        cudaError_t e__ = (call);                                                                                  \
        if (e__ != cudaSuccess) {                                                                                  \
            fail(s, e__ == cudaErrorMemoryAllocation ? RSV_ERR_OOM : RSV_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e__)); \
            return bail(e__ == cudaErrorMemoryAllocation ? RSV_ERR_OOM : RSV_ERR_CUDA);                            \
        }                                                                                                          \
    } while (0)

    CUC(cudaSetDevice(s->device));
    cudaDeviceProp prop;
    CUC(cudaGetDeviceProperties(&prop, s->device));
    s->sm_count = prop.multiProcessorCount;
    CUC(cudaStreamCreateWithFlags(&s->own_stream, cudaStreamNonBlocking));
    s->stream = s->own_stream;
    const uint32_t N = cfg->max_shapes;
    const uint32_t NV = cfg->max_verts ? cfg->max_verts : 1;
    for (int b = 0; b < RSV_BUF_COUNT; b++) {
        s->elem_bytes[b] = buf_elem_bytes(b);
        size_t n = (b == RSV_BUF_VERTS) ? NV : N;
        s->buf_bytes[b] = n * s->elem_bytes[b];
        CUC(cudaMalloc(&s->dev_buf[b], s->buf_bytes[b]));
        CUC(cudaMemsetAsync(s->dev_buf[b], 0, s->buf_bytes[b], s->stream));
        CUC(cudaHostAlloc(&s->host_buf[b], s->buf_bytes[b], cudaHostAllocDefault));
        memset(s->host_buf[b], 0, s->buf_bytes[b]);
    }
    bind_views(s);
    CUC(cudaMalloc(&d.aabb, (size_t)N * sizeof(Box)));
    CUC(cudaMalloc(&d.crect, (size_t)N * sizeof(int4)));
    s->n_tiles = (uint32_t)((n_cells + 1 + kScanTile - 1) / kScanTile);
    CUC(cudaMalloc(&d.cell, ((size_t)s->n_tiles * kScanTile + 16) * 4));   // (+16: 16-byte-aligned row windows of the tile kernel)
    CUC(cudaMemsetAsync(d.cell, 0, ((size_t)s->n_tiles * kScanTile) * 4, s->stream));
    CUC(cudaMalloc(&d.cnt, ((size_t)s->n_tiles * kScanTile + 16) * 4));
    CUC(cudaMemsetAsync(d.cnt, 0, ((size_t)s->n_tiles * kScanTile + 16) * 4, s->stream));
    CUC(cudaMalloc(&d.tile_state, (size_t)s->n_tiles * 8));
    CUC(cudaMalloc(&d.tester_start, (size_t)N * 4));
    CUC(cudaMalloc(&d.tester_count, (size_t)N * 4));
    CUC(cudaMalloc(&d.grp_count, (size_t)N * 4));
    CUC(cudaMemsetAsync(d.grp_count, 0, (size_t)N * 4, s->stream));
    CUC(cudaMalloc(&d.orphans, (size_t)N * 4));
    CUC(cudaMalloc(&d.ctr, sizeof(Counters)));
    CUC(cudaHostAlloc(&s->h_ctr, sizeof(Counters), cudaHostAllocDefault));
    CUC(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
    CUC(cudaStreamCreateWithFlags(&s->side_stream, cudaStreamNonBlocking));
    if (const char* e = getenv("RSV_NO_GRAPH")) s->graphs = !(e[0] && e[0] != '0');
    CUC(cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming));
    CUC(cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming));
    for (auto& sl : s->slot) {
        CUC(cudaHostAlloc(&sl.h_ctr, sizeof(Counters), cudaHostAllocDefault));
        CUC(cudaEventCreateWithFlags(&sl.ev_done, cudaEventDisableTiming));
    }
    uint32_t ce = cfg->max_entries ? cfg->max_entries : (uint32_t)std::min<unsigned long long>(4ull * N + 4096, 0xfffffff0ull);
    uint32_t cp = cfg->max_pairs ? cfg->max_pairs : (uint32_t)std::min<unsigned long long>(2ull * N + 4096, 0xfffffff0ull);
    uint32_t cc = cfg->max_contacts ? cfg->max_contacts : (uint32_t)std::min<unsigned long long>(2ull * N + 4096, 0xfffffff0ull);
    if (alloc_outputs(s, ce, cp, cc) != RSV_OK) return bail(RSV_ERR_OOM);
    for (auto& e : s->ev) CUC(cudaEventCreate(&e));
    for (auto& e : s->ev_timer) CUC(cudaEventCreate(&e));
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_narrow<0>, kNarrowBlock, 0) == cudaSuccess && occ > 0)
        s->narrow_blocks = s->sm_count * occ;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_narrow<1>, kNarrowBlock, 0) == cudaSuccess && occ > 0)
        s->narrow_blocks1 = s->sm_count * occ;

    if (const char* e = getenv("RSV_NARROW_ORDER")) s->narrow_order = atoi(e);
    // (tuning: blocks per SM of the two narrowphase kernels, which run side by side)
    if (const char* e = getenv("RSV_NARROW_BLOCKS0")) { const int v = atoi(e); if (v >= 1 && v <= 16) { s->narrow_blocks = s->sm_count * v; s->narrow_blocks0_fixed = 1; } }
    if (const char* e = getenv("RSV_NARROW_BLOCKS1")) { const int v = atoi(e); if (v >= 1 && v <= 16) s->narrow_blocks1 = s->sm_count * v; }
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_pairs<false>, kPairsBlock, 0) == cudaSuccess && occ > 0)
        s->pairs_blocks[0] = s->sm_count * occ;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_pairs<true>, kPairsBlock, 0) == cudaSuccess && occ > 0)
        s->pairs_blocks[1] = s->sm_count * occ;
    CUC(cudaStreamSynchronize(s->stream));
#undef CUC
    *out = s;
    return RSV_OK;
}

int32_t rsv_destroy(rsv_space* s) {
    if (!s) return RSV_OK;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    for (int b = 0; b < RSV_BUF_COUNT; b++) {
        if (s->dev_buf[b]) cudaFree(s->dev_buf[b]);
        if (s->host_buf[b]) cudaFreeHost(s->host_buf[b]);
    }
    DevView& d = s->d;
    void* dev[] = {d.aabb, d.crect, d.cell, d.cnt, d.entries, d.ent_box, d.ent_home, d.ent_tags, d.multi, d.tile_state, d.tester_start, d.tester_count, d.grp_count, d.ptmp, d.pgrp, d.orphans, d.ghosts, d.n_ghosts, d.ctr};
    for (void* p : dev) if (p) cudaFree(p);
    if (s->h_ctr) cudaFreeHost(s->h_ctr);
    if (s->host_pos_page1) cudaFreeHost(s->host_pos_page1);
    if (s->h_pair_list) cudaFreeHost(s->h_pair_list);
    if (s->d_pair_list) cudaFree(s->d_pair_list);
    if (s->h_list_slots) cudaFreeHost(s->h_list_slots);
    if (s->h_list_xy) cudaFreeHost(s->h_list_xy);
    if (s->d_list_slots) cudaFree(s->d_list_slots);
    if (s->d_list_xy) cudaFree(s->d_list_xy);
    for (auto& sl : s->slot) {
        void* dv[] = {sl.hits, sl.contacts, sl.kind_queue};
        for (void* p : dv) if (p) cudaFree(p);
        if (sl.h_ctr) cudaFreeHost(sl.h_ctr);
        if (sl.h_hits) cudaFreeHost(sl.h_hits);
        if (sl.h_contacts) cudaFreeHost(sl.h_contacts);
        if (sl.ev_done) cudaEventDestroy(sl.ev_done);
    }
    for (auto& gs : s->graph) for (auto& g : gs) drop_graph(g);
    {
        void* sv[] = {s->sg.cell, s->sg.count, s->sg.entries, s->sg.ent_box, s->sg.ent_home, s->sg.ent_tags, s->sg.cell_of, d.dyn_list, d.n_dyn_dev};
        for (void* p : sv) if (p) cudaFree(p);
    }
    for (int p = 1; p < 4; p++) if (s->pos_page[p]) cudaFree(s->pos_page[p]);
    if (s->upload_stream) { cudaStreamSynchronize(s->upload_stream); cudaStreamDestroy(s->upload_stream); }
    for (int p = 0; p < 4; p++) {
        if (s->ev_page_up[p]) cudaEventDestroy(s->ev_page_up[p]);
        if (s->ev_page_rd[p]) cudaEventDestroy(s->ev_page_rd[p]);
    }
    if (s->mailbox) cudaFree(s->mailbox);
    if (s->copy_stream) cudaStreamDestroy(s->copy_stream);
    if (s->side_stream) cudaStreamDestroy(s->side_stream);
    if (s->ev_fork) cudaEventDestroy(s->ev_fork);
    if (s->ev_join) cudaEventDestroy(s->ev_join);
    lines_free(s->lines);
    for (auto& e : s->ev) if (e) cudaEventDestroy(e);
    for (auto& e : s->ev_timer) if (e) cudaEventDestroy(e);
    if (s->own_stream) cudaStreamDestroy(s->own_stream);
    cudaGetLastError();
    delete s;
    return RSV_OK;
}

int32_t rsv_grid_dims(const rsv_space* s, int32_t* w, int32_t* h) {
    if (!s) return RSV_ERR_BAD_ARG;
    if (w) *w = s->d.W;
    if (h) *h = s->d.H;
    return RSV_OK;
}

int32_t rsv_reserve(rsv_space* s, uint32_t max_entries, uint32_t max_pairs, uint32_t max_contacts) {
    if (!s) return RSV_ERR_BAD_ARG;
    CU(s, cudaSetDevice(s->device));
    CU(s, cudaStreamSynchronize(s->stream));
    CU(s, cudaStreamSynchronize(s->copy_stream));
    for (auto& sl : s->slot) sl.pending = false;  // frames in flight are dropped: their buffers are about to move
    s->launch_slot = s->finish_slot = 0;
    return alloc_outputs(s, max_entries, max_pairs, max_contacts);
}

void* rsv_staging(rsv_space* s, int32_t which, size_t* bytes) {
    if (!s || which < 0 || which >= RSV_BUF_COUNT) return nullptr;
    if (bytes) *bytes = s->buf_bytes[which];
    return s->host_buf[which];
}

// Second pinned page for the positions: the host writes frame f+1 into one page while frame f's upload reads the
// other. page 0 is rsv_staging(RSV_BUF_POS).
void* rsv_staging_page(rsv_space* s, int32_t which, int32_t page, size_t* bytes) {
    if (!s || which != RSV_BUF_POS || page < 0 || page > 1) return nullptr;
    if (bytes) *bytes = s->buf_bytes[which];
    if (page == 0) return s->host_buf[which];
    if (!s->host_pos_page1) {
        if (cudaSetDevice(s->device) != cudaSuccess) return nullptr;
        if (cudaHostAlloc(&s->host_pos_page1, s->buf_bytes[which], cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        memcpy(s->host_pos_page1, s->host_buf[which], s->buf_bytes[which]);
    }
    return s->host_pos_page1;
}

int32_t rsv_commit_pos_page(rsv_space* s, int32_t page, uint32_t lo, uint32_t hi) {
    if (!s || page < 0 || page > 1) return RSV_ERR_BAD_ARG;
    if (page == 1 && !s->host_pos_page1) return fail(s, RSV_ERR_BAD_ARG, "rsv_commit_pos_page: page 1 was never requested");
    if (hi > s->n_shapes) hi = s->n_shapes;
    if (lo >= hi) return RSV_OK;
    CU(s, cudaSetDevice(s->device));
    const char* src = (const char*)(page ? s->host_pos_page1 : s->host_buf[RSV_BUF_POS]);
    CU(s, cudaMemcpyAsync((char*)s->dev_buf[RSV_BUF_POS] + (size_t)lo * 16, src + (size_t)lo * 16, (size_t)(hi - lo) * 16,
                          cudaMemcpyHostToDevice, s->stream));
    return RSV_OK;
}

// Device-resident position pages: a host (or a device-side integrator) that keeps several position sets in HBM selects
// which one the next frames read, without any copy. Page 0 is RSV_BUF_POS's device buffer.
int32_t rsv_pos_pages(rsv_space* s, int32_t n_pages) {
    if (!s || n_pages < 1 || n_pages > 4) return RSV_ERR_BAD_ARG;
    CU(s, cudaSetDevice(s->device));
    s->pos_page[0] = s->dev_buf[RSV_BUF_POS];
    if (!s->upload_stream) {
        CU(s, cudaStreamCreateWithFlags(&s->upload_stream, cudaStreamNonBlocking));
        for (int p = 0; p < 4; p++) {
            CU(s, cudaEventCreateWithFlags(&s->ev_page_up[p], cudaEventDisableTiming));
            CU(s, cudaEventCreateWithFlags(&s->ev_page_rd[p], cudaEventDisableTiming));
        }
    }
    for (int p = 1; p < n_pages; p++)
        if (!s->pos_page[p]) {
            CU(s, cudaMalloc(&s->pos_page[p], s->buf_bytes[RSV_BUF_POS]));
            CU(s, cudaMemcpyAsync(s->pos_page[p], s->dev_buf[RSV_BUF_POS], s->buf_bytes[RSV_BUF_POS], cudaMemcpyDeviceToDevice, s->stream));
        }
    CU(s, cudaStreamSynchronize(s->stream));
    if (n_pages > s->n_pos_pages) s->n_pos_pages = n_pages;
    return RSV_OK;
}

int32_t rsv_pos_page_upload(rsv_space* s, int32_t device_page, int32_t staging_page, uint32_t lo, uint32_t hi) {
    if (!s || device_page < 0 || device_page >= s->n_pos_pages || staging_page < 0 || staging_page > 1) return RSV_ERR_BAD_ARG;
    if (!s->upload_stream) return fail(s, RSV_ERR_BAD_ARG, "rsv_pos_page_upload: call rsv_pos_pages first");
    if (staging_page == 1 && !s->host_pos_page1) return fail(s, RSV_ERR_BAD_ARG, "rsv_pos_page_upload: staging page 1 was never requested");
    if (hi > s->n_shapes) hi = s->n_shapes;
    if (lo >= hi) return RSV_OK;
    CU(s, cudaSetDevice(s->device));
    void* dst = device_page ? s->pos_page[device_page] : s->dev_buf[RSV_BUF_POS];
    const char* src = (const char*)(staging_page ? s->host_pos_page1 : s->host_buf[RSV_BUF_POS]);
    // the frames already launched on this page must have read it before it is overwritten
    if (s->page_rd_valid[device_page]) CU(s, cudaStreamWaitEvent(s->upload_stream, s->ev_page_rd[device_page], 0));
    CU(s, cudaMemcpyAsync((char*)dst + (size_t)lo * 16, src + (size_t)lo * 16, (size_t)(hi - lo) * 16, cudaMemcpyHostToDevice, s->upload_stream));
    CU(s, cudaEventRecord(s->ev_page_up[device_page], s->upload_stream));
    s->page_up_pending[device_page] = true;
    return RSV_OK;
}

int32_t rsv_pos_page_select(rsv_space* s, int32_t device_page) {
    if (!s || device_page < 0 || device_page >= s->n_pos_pages) return RSV_ERR_BAD_ARG;
    CU(s, cudaSetDevice(s->device));
    s->cur_pos_page = device_page;
    s->d.pos = (const double2*)(device_page ? s->pos_page[device_page] : s->dev_buf[RSV_BUF_POS]);
    if (s->page_up_pending[device_page]) {   // frames on this page start after its upload has landed
        CU(s, cudaStreamWaitEvent(s->stream, s->ev_page_up[device_page], 0));
        s->page_up_pending[device_page] = false;
    }
    return RSV_OK;
}

// Sparse position updates: the host lists the slots that moved and their new positions (two pinned arrays owned by the
// library) and only those 20 bytes per moved shape cross the bus; a small kernel scatters them into the position page
// the next frame reads. The write side of ShapeBase.update (shape.go:239-242) for worlds in which few shapes move
// (config 4: 32 bodies among 500 tiles per Space).
static __global__ void k_pos_scatter(double2* __restrict__ pos, const uint32_t* __restrict__ slots, const double2* __restrict__ xy,
                                     uint32_t n, uint32_t n_shapes) {
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const uint32_t i = slots[k];
        if (i < n_shapes) pos[i] = xy[k];
    }
}

void* rsv_pos_list_staging(rsv_space* s, int32_t which, size_t* bytes) {
    if (!s || which < 0 || which > 1) return nullptr;
    if (cudaSetDevice(s->device) != cudaSuccess) return nullptr;
    const size_t N = s->cfg.max_shapes;
    if (!s->h_list_slots) {
        if (cudaHostAlloc(&s->h_list_slots, N * 4, cudaHostAllocDefault) != cudaSuccess ||
            cudaHostAlloc(&s->h_list_xy, N * 16, cudaHostAllocDefault) != cudaSuccess ||
            cudaMalloc(&s->d_list_slots, N * 4) != cudaSuccess || cudaMalloc(&s->d_list_xy, N * 16) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
    }
    if (bytes) *bytes = which ? N * 16 : N * 4;
    return which ? (void*)s->h_list_xy : (void*)s->h_list_slots;
}

int32_t rsv_commit_pos_list(rsv_space* s, uint32_t count) {
    if (!s) return RSV_ERR_BAD_ARG;
    if (!s->h_list_slots) return fail(s, RSV_ERR_BAD_ARG, "rsv_commit_pos_list: call rsv_pos_list_staging first");
    if (count > s->cfg.max_shapes) return fail(s, RSV_ERR_BAD_ARG, "rsv_commit_pos_list: %u entries > capacity %u", count, s->cfg.max_shapes);
    if (!count) return RSV_OK;
    CU(s, cudaSetDevice(s->device));
    CU(s, cudaMemcpyAsync(s->d_list_slots, s->h_list_slots, (size_t)count * 4, cudaMemcpyHostToDevice, s->stream));
    CU(s, cudaMemcpyAsync(s->d_list_xy, s->h_list_xy, (size_t)count * 16, cudaMemcpyHostToDevice, s->stream));
    k_pos_scatter<<<blocks_for(count, 256, s->sm_count * 4), 256, 0, s->stream>>>(const_cast<double2*>(s->d.pos), s->d_list_slots, s->d_list_xy,
                                                                                 count, s->n_shapes);
    CU(s, cudaGetLastError());
    return RSV_OK;
}

void* rsv_device_buffer(rsv_space* s, int32_t which) {
    if (!s || which < 0 || which >= RSV_BUF_COUNT) return nullptr;
    return s->dev_buf[which];
}

int32_t rsv_set_counts(rsv_space* s, uint32_t n_shapes, uint32_t n_verts) {
    if (!s) return RSV_ERR_BAD_ARG;
    if (n_shapes > s->cfg.max_shapes) return fail(s, RSV_ERR_BAD_ARG, "rsv_set_counts: %u shapes > capacity %u", n_shapes, s->cfg.max_shapes);
    if (n_verts > (s->cfg.max_verts ? s->cfg.max_verts : 1)) return fail(s, RSV_ERR_BAD_ARG, "rsv_set_counts: %u vertices > capacity %u", n_verts, s->cfg.max_verts);
    if (n_shapes != s->n_shapes) s->static_valid = false;
    s->n_shapes = n_shapes;
    s->n_verts = n_verts;
    s->d.n_shapes = n_shapes;
    if (!s->strip_mode) { s->d.own_lo = 0; s->d.own_hi = n_shapes; s->d.home_lo = 0; s->d.home_hi = s->d.H; }
    return RSV_OK;
}

int32_t rsv_commit(rsv_space* s, uint32_t which_mask, uint32_t lo, uint32_t hi) {
    if (!s) return RSV_ERR_BAD_ARG;
    if (hi > s->n_shapes) hi = s->n_shapes;
    CU(s, cudaSetDevice(s->device));
    // anything that can change a static shape's registrations drops the cached static grid (positions are the
    // host's promise: RSV_META_STATIC means "not moved since the flag was committed")
    if (which_mask & ((1u << RSV_BUF_META) | (1u << RSV_BUF_LOCAL_AABB) | (1u << RSV_BUF_TAGS) | (1u << RSV_BUF_SPACE_IDX)))
        s->static_valid = false;
    for (int b = 0; b < RSV_BUF_COUNT; b++) {
        if (!(which_mask & (1u << b))) continue;
        size_t off, len;
        if (b == RSV_BUF_VERTS) { off = 0; len = (size_t)s->n_verts * s->elem_bytes[b]; }
        else {
            if (lo >= hi) continue;
            off = (size_t)lo * s->elem_bytes[b]; len = (size_t)(hi - lo) * s->elem_bytes[b];
        }
        if (len == 0) continue;
        CU(s, cudaMemcpyAsync((char*)s->dev_buf[b] + off, (char*)s->host_buf[b] + off, len, cudaMemcpyHostToDevice, s->stream));
    }
    return RSV_OK;
}

int32_t rsv_sync(rsv_space* s) {
    if (!s) return RSV_ERR_BAD_ARG;
    CU(s, cudaSetDevice(s->device));
    if (s->upload_stream) CU(s, cudaStreamSynchronize(s->upload_stream));
    CU(s, cudaStreamSynchronize(s->stream));
    return RSV_OK;
}


// Grid of k_narrow<0> (the staged convex-convex bins): its occupancy limit when the last finished frame had more than two
// chunks per warp of four blocks per SM (a world of rectangles: several waves), else four or three blocks per SM, so that
// k_narrow<1> finds free registers next to it (measured on one box, blocks per SM of k_narrow<0> = 2 / 3 / 4 / 5:
// config 3 0.288 / 0.286 / 0.287 / 0.296 ms, config 2 0.384 / 0.335 / 0.308 / 0.298 ms).
static inline int narrow_grid0(const rsv_space* s) {
    if (s->narrow_blocks0_fixed || s->narrow_per_sm == 0) return s->narrow_blocks;
    return s->sm_count * std::min(s->narrow_per_sm, s->narrow_blocks / s->sm_count);
}
// Called with the bin counts of a finished frame. Up to two chunks per resident warp still count as "one wave of long
// chains" (config 3 has ~4 k chunks of convex-convex pairs). A count within 12 % of a threshold keeps the current choice:
// the grid is part of the frame graph's key, and a world sitting on a threshold must not re-capture it every frame.
static inline void narrow_grid0_update(rsv_space* s, uint32_t cc_chunks) {
    const double w4 = 2.0 * s->sm_count * 4 * (kNarrowBlock / 32), w3 = 2.0 * s->sm_count * 3 * (kNarrowBlock / 32);
    const int occ = std::max(s->narrow_blocks / s->sm_count, 1);
    auto cls = [&](double f) { return cc_chunks > f * w4 ? occ : (cc_chunks > f * w3 ? 4 : 3); };
    const int lo = cls(1.12), hi = cls(0.88);   // the class with the thresholds moved up / down
    if (lo == hi || s->narrow_per_sm == 0 || (s->narrow_per_sm != lo && s->narrow_per_sm != hi)) s->narrow_per_sm = lo == hi ? lo : cls(1.0);
}

// Which candidate path a frame takes (see rsv_cellpairs.cuh).
static inline bool cell_pair_frame(uint32_t flags, bool incremental) {
    return !incremental && !(flags & (RSV_FRAME_ALL_CANDIDATES | RSV_FRAME_CANDIDATE_WALK | RSV_FRAME_GRID_ONLY));
}

// After a cell-pair frame the cells are unordered and the per-entry boxes / tags / home records are not filled: the
// consumers of the full grid (LineTest, the grid read-back) complete it first. The frame's counters (n_multi,
// n_entries) are still on the device: they are only cleared by the next frame or rsv_test_pairs.
static int32_t ensure_full_grid(rsv_space* s) {
    if (s->grid_complete) return RSV_OK;
    const DevView& d = s->d;
    const int gcap = s->sm_count * 8;
    k_cellsort<<<blocks_for(std::min<uint32_t>(d.n_cells, d.cap_multi), kGridBlock, gcap), kGridBlock, 0, s->stream>>>(d);
    k_entboxes<<<gcap, kGridBlock, 0, s->stream>>>(d, 0u);
    CU(s, cudaGetLastError());
    s->grid_complete = true;
    return RSV_OK;
}

// The kernels of one frame plus the counter read-back, enqueued on `st` (directly, or into a stream capture).
// halo != nullptr: the frame of a strip-partitioned space with the halo exchange folded in (rsv_strip_frame_launch);
// phases selects the part before the wait on the neighbours (RSV_HALO_SEND) and / or the part after it (RSV_HALO_RECV).
static int32_t enqueue_frame(rsv_space* s, const DevView& d, int32_t margin, uint32_t flags, rsv_space::FrameSlot& sl,
                             cudaStream_t st, bool timing, const HaloArgs* halo = nullptr, uint32_t phases = 3u) {
    auto mark = [&](int k) { if (timing) cudaEventRecord(s->ev[k], st); };
    const int gcap = s->sm_count * 8;
    const bool inc = d.subset == kSubsetDynamic;   // incremental frame: only dyn_list is re-registered
    const uint32_t max_items = inc ? d.n_dyn : (d.own_hi - d.own_lo) + ((d.n_ghosts && !halo) ? d.cap_ghosts : 0u);
    if (phases & RSV_HALO_SEND) {
        mark(RSV_K_CLEAR);
        k_clear<<<4, kGridBlock, 0, st>>>(d.tile_state, s->n_tiles, reinterpret_cast<uint32_t*>(d.ctr),
                                         (uint32_t)(offsetof(Counters, ray_hit_cursor) / 4));
        mark(RSV_K_BOUNDS);
        // (an incremental frame launches k_bounds even without dynamic shapes: it carries the static counts over)
        if (halo) {
            DevView v = d;
            v.part = kPartOwned;
            k_bounds<true><<<blocks_for(std::max<uint32_t>(max_items, 1u), kGridBlock, gcap), kGridBlock, 0, st>>>(v, margin, *halo);
        } else if (d.n_shapes) {
            k_bounds<false><<<blocks_for(max_items, kGridBlock, gcap), kGridBlock, 0, st>>>(d, margin, HaloArgs{});
        }
    }
    if (!(phases & RSV_HALO_RECV)) {
        CU(s, cudaGetLastError());
        return RSV_OK;
    }
    if (halo) k_halo_apply_bounds<<<dim3(blocks_for(halo->cap, kGridBlock, s->sm_count), 2), kGridBlock, 0, st>>>(d, margin, *halo);
    mark(RSV_K_SCAN);
    const int scan_blocks = (d.n_cells + 1 + kScanTile - 1) / kScanTile;
    if (inc) k_scan<true><<<scan_blocks, kScanBlock, 0, st>>>(d.cnt, d.cell, d.n_cells + 1, d.tile_state, &d.ctr->scan_ticket, d.multi,
                                                              &d.ctr->n_multi, d.cap_multi, s->sg.count);
    else k_scan<false><<<scan_blocks, kScanBlock, 0, st>>>(d.cnt, d.cell, d.n_cells + 1, d.tile_state, &d.ctr->scan_ticket, d.multi,
                                                           &d.ctr->n_multi, d.cap_multi, nullptr);
    mark(RSV_K_SCATTER);
    if (inc && d.static_entries) k_static_place<<<blocks_for(d.static_entries, kGridBlock, gcap), kGridBlock, 0, st>>>(d, s->sg, flags);
    const uint32_t scatter_items = max_items + (halo ? d.cap_ghosts : 0u);   // (k_scatter walks own + ghosts)
    const bool grid_only = (flags & RSV_FRAME_GRID_ONLY) != 0;
    if (d.n_shapes && scatter_items) k_scatter<<<blocks_for(scatter_items, kGridBlock, gcap), kGridBlock, 0, st>>>(d);
    mark(RSV_K_CELLSORT);
    // Cell-pair path (rsv_cellpairs.cuh): frames that only need the Bounds-gated pairs. The reference's literal walk
    // (k_pairs) serves the frames that dump every candidate or count them, incremental frames (their k_scan lists
    // other cells) and grid-only frames (a LineTest follows: it needs ordered cells and per-entry boxes).
    const bool cellpairs = cell_pair_frame(flags, inc);
    if (inc) {
        k_cellfix<<<blocks_for(std::min<uint32_t>(d.n_cells, 4u * std::max<uint32_t>(d.n_dyn, 1u)), kGridBlock, gcap), kGridBlock, 0, st>>>(d, flags);
    } else if (!cellpairs) {
        k_cellsort<<<blocks_for(std::min<uint32_t>(d.n_cells, d.cap_multi), kGridBlock, gcap), kGridBlock, 0, st>>>(d);
        k_entboxes<<<gcap, kGridBlock, 0, st>>>(d, flags);
    }
    mark(RSV_K_PAIRS);
    if (d.n_shapes && !grid_only) {
        if (cellpairs) {
            const bool nf = (flags & RSV_FRAME_NO_TAG_FILTERS) != 0;
            if (nf) k_cellpairs<false><<<s->sm_count * 6, kCpBlock, 0, st>>>(d, margin);
            else k_cellpairs<true><<<s->sm_count * 6, kCpBlock, 0, st>>>(d, margin);
            k_alloc<<<s->sm_count * 8, 256, 0, st>>>(d);
            if (nf) k_rank<false><<<s->sm_count * 8, 256, 0, st>>>(d, margin);
            else k_rank<true><<<s->sm_count * 8, 256, 0, st>>>(d, margin);
            k_place_bin<<<s->sm_count * 8, 256, 0, st>>>(d);   // (place + bin in one pass over the grouped records)
        } else if (flags & RSV_FRAME_NO_TAG_FILTERS) k_pairs<false><<<s->pairs_blocks[0], kPairsBlock, 0, st>>>(d, margin, flags);
        else k_pairs<true><<<s->pairs_blocks[1], kPairsBlock, 0, st>>>(d, margin, flags);
    }
    mark(RSV_K_NARROW);
    if (!grid_only) {
        if (!cellpairs) k_bin<<<s->sm_count * 8, 256, 0, st>>>(d, false);
        k_bin_scatter<<<s->sm_count * 4, 256, 0, st>>>(d);
        // the two narrowphase kernels touch disjoint records: run them side by side (their tails overlap)
        CU(s, cudaEventRecord(s->ev_fork, st));
        CU(s, cudaStreamWaitEvent(s->side_stream, s->ev_fork, 0));
        // Submission order and grid sizes matter: whichever kernel is submitted first fills the SMs' registers and the
        // other one only gets in as its blocks retire. The convex-convex kernel (the longer chain) goes first, on a grid
        // just large enough for its work (narrow_grid0: from the bin counts of the last finished frame), so that the
        // mixed kernel starts beside it at once: config 3 0.321 -> 0.286 ms, config 2 0.312 -> 0.298 ms.
        if (s->narrow_order) {
            k_narrow<0><<<narrow_grid0(s), kNarrowBlock, 0, st>>>(d);
            k_narrow<1><<<s->narrow_blocks1, kNarrowBlock, 0, s->side_stream>>>(d);
            CU(s, cudaEventRecord(s->ev_join, s->side_stream));
        } else {
            k_narrow<1><<<s->narrow_blocks1, kNarrowBlock, 0, s->side_stream>>>(d);
            CU(s, cudaEventRecord(s->ev_join, s->side_stream));
            k_narrow<0><<<s->narrow_blocks, kNarrowBlock, 0, st>>>(d);
        }
        CU(s, cudaStreamWaitEvent(st, s->ev_join, 0));
    }
    mark(RSV_K_COUNT);
    CU(s, cudaGetLastError());
    CU(s, cudaMemcpyAsync(sl.h_ctr, d.ctr, sizeof(Counters), cudaMemcpyDeviceToHost, st));
    return RSV_OK;
}

// A frame is a fixed chain of a dozen dependent launches whose arguments only change when capacities, counts or the
// strip layout change, so steady-state frames replay one captured CUDA graph per result slot: small worlds (the
// reference's own examples are a few hundred shapes) are bound by launch overhead, not by the kernels.
// Returns 1 if the frame was submitted as a graph, 0 if the caller has to launch directly, < 0 on error.
static int32_t frame_via_graph(rsv_space* s, const DevView& d, int32_t margin, uint32_t flags, rsv_space::FrameSlot& sl,
                               rsv_space::FrameGraph& g, cudaStream_t st, const HaloArgs* halo = nullptr) {
    rsv_space::FrameGraph::Key key{};
    key.margin = margin; key.flags = flags; key.stream = st; key.n_tiles = s->n_tiles;
    if (halo) key.halo = *halo;   // (flags carries bit 31 for fused strip frames)
    key.blocks[0] = s->pairs_blocks[0] ^ (s->pairs_blocks[1] << 16);
    key.blocks[1] = narrow_grid0(s) ^ (s->narrow_blocks1 << 16);
    const bool same = g.seen && memcmp(&key, &g.key, sizeof key) == 0 && memcmp(&d, &g.view, sizeof d) == 0;
    if (!same) {
        drop_graph(g);
        g.key = key;
        memcpy(&g.view, &d, sizeof d);
        g.seen = 1;
        return 0;  // first frame with these parameters: launch directly, capture if they come again
    }
    if (!g.exec) {
        cudaGraph_t graph = nullptr;
        if (cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
            cudaGetLastError();
            s->graphs = false;
            return 0;
        }
        const int32_t rc = enqueue_frame(s, d, margin, flags & 0x7fffffffu, sl, st, false, halo, 3u);
        const cudaError_t e = cudaStreamEndCapture(st, &graph);
        if (rc != RSV_OK || e != cudaSuccess || !graph || cudaGraphInstantiate(&g.exec, graph, 0) != cudaSuccess) {
            cudaGetLastError();
            if (graph) cudaGraphDestroy(graph);
            g.exec = nullptr;
            s->graphs = false;  // capture is not possible on this stream: stay with direct launches
            return 0;
        }
        cudaGraphDestroy(graph);
    }
    CU(s, cudaGraphLaunch(g.exec, st));
    return 1;
}

// One-off build of the cached static grid: the regular grid kernels over the RSV_META_STATIC shapes that hold
// cells, written into s->sg; every other live shape is appended to dyn_list. Synchronous (two small read-backs).
static int32_t build_static(rsv_space* s) {
    DevView& d = s->d;
    cudaStream_t st = s->stream;
    StaticGrid& sg = s->sg;
    const size_t cell_bytes = (size_t)s->n_tiles * kScanTile * 4;
    if (!sg.cell) CU(s, cudaMalloc(&sg.cell, cell_bytes));
    if (!sg.count) CU(s, cudaMalloc(&sg.count, cell_bytes));
    CU(s, cudaMemsetAsync(sg.count, 0, cell_bytes, st));
    if (!d.dyn_list) CU(s, cudaMalloc(&d.dyn_list, (size_t)s->cfg.max_shapes * 4));
    if (!d.n_dyn_dev) CU(s, cudaMalloc(&d.n_dyn_dev, 4));
    for (auto& gs : s->graph) for (auto& g : gs) drop_graph(g);
    DevView v = d;
    v.subset = kSubsetStaticBuild;
    v.cell = sg.cell;
    const int gcap = s->sm_count * 8;
    k_clear<<<4, kGridBlock, 0, st>>>(d.tile_state, s->n_tiles, reinterpret_cast<uint32_t*>(d.ctr),
                                     (uint32_t)(offsetof(Counters, ray_hit_cursor) / 4));
    CU(s, cudaMemsetAsync(d.n_dyn_dev, 0, 4, st));
    const uint32_t items = d.own_hi - d.own_lo;
    if (items) k_bounds<false><<<blocks_for(items, kGridBlock, gcap), kGridBlock, 0, st>>>(v, 0, HaloArgs{});
    Counters hc_local{};   // (pageable on purpose: the pinned mirrors may belong to a frame / ray batch in flight)
    Counters* hc = &hc_local;
    uint32_t n_dyn = 0;
    CU(s, cudaMemcpyAsync(hc, d.ctr, sizeof(Counters), cudaMemcpyDeviceToHost, st));
    CU(s, cudaMemcpyAsync(&n_dyn, d.n_dyn_dev, 4, cudaMemcpyDeviceToHost, st));
    CU(s, cudaStreamSynchronize(st));
    if (hc->n_entries >= 0xfffffff0ull) return fail(s, RSV_ERR_CAPACITY, "static grid: %llu registrations", (unsigned long long)hc->n_entries);
    const uint32_t Rs = (uint32_t)hc->n_entries;
    if (Rs > sg.cap) {
        void* old[] = {sg.entries, sg.ent_box, sg.ent_home, sg.ent_tags, sg.cell_of};
        for (void* p : old) if (p) cudaFree(p);
        sg.entries = nullptr; sg.ent_box = nullptr; sg.ent_home = nullptr; sg.ent_tags = nullptr; sg.cell_of = nullptr; sg.cap = 0;
        const uint32_t cap = (uint32_t)std::min<unsigned long long>((unsigned long long)Rs + Rs / 8 + 1024, 0xfffffff0ull);
        CU(s, cudaMalloc(&sg.entries, (size_t)cap * 4));
        CU(s, cudaMalloc(&sg.ent_box, (size_t)cap * sizeof(float4)));
        CU(s, cudaMalloc(&sg.ent_home, (size_t)cap * sizeof(uint2)));
        CU(s, cudaMalloc(&sg.ent_tags, (size_t)cap * 8));
        CU(s, cudaMalloc(&sg.cell_of, (size_t)cap * 4));
        sg.cap = cap;
    }
    v.entries = sg.entries; v.ent_box = sg.ent_box; v.ent_home = sg.ent_home; v.ent_tags = sg.ent_tags; v.cap_entries = sg.cap;
    k_scan<false><<<(d.n_cells + 1 + kScanTile - 1) / kScanTile, kScanBlock, 0, st>>>(d.cnt, sg.cell, d.n_cells + 1, d.tile_state, &d.ctr->scan_ticket,
                                                                                     d.multi, &d.ctr->n_multi, d.cap_multi, nullptr);
    if (items) k_scatter<<<blocks_for(items, kGridBlock, gcap), kGridBlock, 0, st>>>(v);
    k_cellsort<<<blocks_for(std::min<uint32_t>(d.n_cells, d.cap_multi), kGridBlock, gcap), kGridBlock, 0, st>>>(v);
    k_entboxes<<<gcap, kGridBlock, 0, st>>>(v, 0u);   // (tags always: later frames may or may not use filters)
    k_static_cells<<<blocks_for(d.n_cells, kGridBlock, gcap), kGridBlock, 0, st>>>(sg.cell, d.n_cells, sg.cell_of, sg.count, sg.cap);
    CU(s, cudaGetLastError());
    CU(s, cudaStreamSynchronize(st));
    d.n_dyn = n_dyn;
    d.static_entries = Rs;
    d.static_testers = (uint32_t)hc->n_testers;
    s->static_valid = true;
    return RSV_OK;
}

int32_t rsv_static_invalidate(rsv_space* s) {
    if (!s) return RSV_ERR_BAD_ARG;
    s->static_valid = false;
    return RSV_OK;
}

int32_t rsv_static_info(rsv_space* s, uint32_t* static_entries, uint32_t* n_dynamic) {
    if (!s) return RSV_ERR_BAD_ARG;
    if (static_entries) *static_entries = s->static_valid ? s->d.static_entries : 0u;
    if (n_dynamic) *n_dynamic = s->static_valid ? s->d.n_dyn : s->n_shapes;
    return RSV_OK;
}

int32_t rsv_frame_launch(rsv_space* s, int32_t margin, uint32_t flags) {
    if (!s) return RSV_ERR_BAD_ARG;
    if (margin < 0 || margin > (1 << 20)) return fail(s, RSV_ERR_BAD_ARG, "rsv_frame: margin %d out of range", margin);
    CU(s, cudaSetDevice(s->device));
    // Incremental frames: single-device Spaces only (strip frames rebuild fully: their ghost set changes every frame).
    bool inc = (flags & RSV_FRAME_INCREMENTAL) && !s->strip_mode && !s->d.n_ghosts && s->d.n_shapes;
    if (inc && !s->static_valid) {
        const int32_t rc = build_static(s);
        if (rc != RSV_OK) return rc;
    }
    // mostly dynamic worlds are cheaper to rebuild in one go
    if (inc && (unsigned long long)s->d.n_dyn * 2ull > s->d.n_shapes) inc = false;
    s->d.subset = inc ? kSubsetDynamic : kSubsetAll;
    DevView& d = s->d;
    cudaStream_t st = s->stream;
    rsv_space::FrameSlot& sl = s->slot[s->launch_slot];
    if (sl.pending) {  // both slots busy: the oldest frame was never finished -> its results are dropped
        sl.pending = false;
        if (s->finish_slot == s->launch_slot) s->finish_slot ^= 1;
    }
    d.hits = sl.hits; d.contacts = sl.contacts; d.kind_queue = sl.kind_queue;
    const bool timing = (flags & RSV_FRAME_TIMING) != 0;
    int32_t submitted = 0;
    // (timed frames record events between the kernels; the ghost count of strip frames is read on the device)
    if (s->graphs && !timing) {
        submitted = frame_via_graph(s, d, margin, flags, sl, s->graph[s->launch_slot][s->cur_pos_page], st);
        if (submitted < 0) return submitted;
    }
    if (!submitted) {
        const int32_t rc = enqueue_frame(s, d, margin, flags, sl, st, timing);
        if (rc != RSV_OK) return rc;
    }
    CU(s, cudaEventRecord(sl.ev_done, st));
    if (s->upload_stream) {
        CU(s, cudaEventRecord(s->ev_page_rd[s->cur_pos_page], st));
        s->page_rd_valid[s->cur_pos_page] = true;
    }
    sl.flags = flags;
    sl.pending = true;
    s->launch_slot ^= 1;
    sl.cellpairs = cell_pair_frame(flags, inc);
    sl.pair_list = false;
    s->grid_complete = !sl.cellpairs;
    return RSV_OK;
}

static int32_t ensure_mirrors(rsv_space* s, rsv_space::FrameSlot& sl) {
    if (sl.h_cap_pairs < s->d.cap_pairs) {
        if (sl.h_hits) cudaFreeHost(sl.h_hits);
        sl.h_hits = nullptr; sl.h_cap_pairs = 0;
        CU(s, cudaHostAlloc(&sl.h_hits, (size_t)s->d.cap_pairs * sizeof(rsv_hit), cudaHostAllocDefault));
        sl.h_cap_pairs = s->d.cap_pairs;
    }
    if (sl.h_cap_contacts < s->d.cap_contacts) {
        if (sl.h_contacts) cudaFreeHost(sl.h_contacts);
        sl.h_contacts = nullptr; sl.h_cap_contacts = 0;
        CU(s, cudaHostAlloc(&sl.h_contacts, (size_t)s->d.cap_contacts * sizeof(rsv_contact), cudaHostAllocDefault));
        sl.h_cap_contacts = s->d.cap_contacts;
    }
    return RSV_OK;
}

int32_t rsv_frame_finish(rsv_space* s, rsv_counts* out) {
    if (!s) return RSV_ERR_BAD_ARG;
    rsv_space::FrameSlot& sl = s->slot[s->finish_slot];
    if (!sl.pending) return fail(s, RSV_ERR_BAD_ARG, "rsv_frame_finish: no frame launched");
    CU(s, cudaSetDevice(s->device));
    CU(s, cudaEventSynchronize(sl.ev_done));  // this frame's kernels and counters are done; a newer frame may still run
    sl.pending = false;
    s->last_slot = s->finish_slot;
    s->finish_slot ^= 1;
    s->launched_flags = sl.flags;
    s->launched_cellpairs = sl.cellpairs;
    const Counters& c = *sl.h_ctr;
    {
        uint32_t chunks = 0;
        for (int b = 17; b < 81; b++) chunks += (c.kind_count[b] + 31u) / 32u;
        if (!sl.pair_list) narrow_grid0_update(s, chunks);   // (a pair list says nothing about the next frame's narrowphase)
    }
    rsv_counts rc{};
    rc.n_shapes = s->d.n_shapes;
    rc.n_testers = c.n_testers;
    rc.n_entries = c.n_entries;
    rc.n_candidates = c.n_candidates;
    rc.n_gated = c.pair_cursor;
    rc.n_hits = c.n_hits;
    rc.n_contacts = c.n_contacts_needed;
    rc.overflow = c.overflow;
    if (c.n_entries > s->d.cap_entries) rc.overflow |= RSV_OVF_ENTRIES;
    if (c.pair_cursor > s->d.cap_pairs) rc.overflow |= RSV_OVF_PAIRS;
    if (c.n_contacts_needed > s->d.cap_contacts) rc.overflow |= RSV_OVF_CONTACTS;
    s->last_counts = rc;
    if (out) *out = rc;
    if (s->launched_flags & RSV_FRAME_TIMING) {
        for (int k = 0; k < RSV_K_COUNT; k++) cudaEventElapsedTime(&s->timing.kernel_us[k], s->ev[k], s->ev[k + 1]);
        for (int k = 0; k < RSV_K_COUNT; k++) s->timing.kernel_us[k] *= 1000.f;
        cudaEventElapsedTime(&s->timing.frame_us, s->ev[0], s->ev[RSV_K_COUNT]);
        s->timing.frame_us *= 1000.f;
    }
    s->last_records = 0; s->last_contacts = 0;
    if (rc.overflow & RSV_OVF_PAIR_CONTACTS)
        return fail(s, RSV_ERR_CAPACITY, "rsv_frame: a pair produced more than 127 contacts (the per-pair count is 7 bits wide); "
                    "rsv_reserve cannot fix this: split polygons with very many vertices");
    if (rc.overflow & RSV_OVF_GHOSTS)
        return fail(s, RSV_ERR_CAPACITY, "rsv_frame: more halo records than the ghost list / mailbox holds (results incomplete): "
                    "grow max_ghosts (rsv_strip_setup) and the mailbox capacity (rsv_halo_mailbox)");
    if (rc.overflow)
        return fail(s, RSV_ERR_CAPACITY, "rsv_frame: capacity exceeded (need entries=%llu pairs=%llu contacts=%llu; have %u/%u/%u)",
                    (unsigned long long)rc.n_entries, (unsigned long long)rc.n_gated, (unsigned long long)rc.n_contacts,
                    s->d.cap_entries, s->d.cap_pairs, s->d.cap_contacts);
    if (s->launched_flags & RSV_FRAME_DOWNLOAD) {
        // (both result slots at once: pinning a few hundred MB takes tens of milliseconds and must not land in the middle
        // of a host's steady state when the second slot is first used)
        for (auto& each : s->slot) {
            int32_t rcode = ensure_mirrors(s, each);
            if (rcode != RSV_OK) return rcode;
        }
        // on the copy stream: overlaps the upload and kernels of a newer frame already queued on `stream`
        if (rc.n_gated) CU(s, cudaMemcpyAsync(sl.h_hits, sl.hits, (size_t)rc.n_gated * sizeof(rsv_hit), cudaMemcpyDeviceToHost, s->copy_stream));
        if (rc.n_contacts) CU(s, cudaMemcpyAsync(sl.h_contacts, sl.contacts, (size_t)rc.n_contacts * sizeof(rsv_contact), cudaMemcpyDeviceToHost, s->copy_stream));
        CU(s, cudaStreamSynchronize(s->copy_stream));
        s->last_records = rc.n_gated;
        s->last_contacts = rc.n_contacts;
    }
    return RSV_OK;
}

int32_t rsv_frame(rsv_space* s, int32_t margin, uint32_t flags, rsv_counts* out) {
    if (!s) return RSV_ERR_BAD_ARG;
    // launch + finish of THIS frame: frames launched earlier and never finished are dropped
    for (auto& sl : s->slot) sl.pending = false;
    s->finish_slot = s->launch_slot;
    int32_t rc = rsv_frame_launch(s, margin, flags);
    if (rc != RSV_OK) return rc;
    return rsv_frame_finish(s, out);
}

// ---- explicit pair tests: Shape.Intersection(other) for a host-supplied list of ordered pairs ----------------------
// One record per listed pair, in list order: Bounds() of both shapes from the committed positions, the exact Bounds
// gate, the narrowphase. Used by the host layer when a callback moved a shape in the middle of an IntersectionTest
// (shape.go:296-312 evaluates s.owner.Intersection(p.Shape) live, candidate by candidate) and for single
// Shape.Intersection / IsIntersecting calls (circle.go:69-82, convexPolygon.go:288-300). The grid of the last frame is
// left alone.
static __global__ void k_pairs_from_list(DevView d, const uint32_t* __restrict__ list, uint32_t n) {
    if (blockIdx.x == 0 && threadIdx.x == 0) d.ctr->pair_cursor = n;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const uint32_t A = list[2 * k], B = list[2 * k + 1];
        for (int w = 0; w < 2; w++) {
            const uint32_t i = w ? B : A;
            const double2 p = __ldg(d.pos + i);
            const Box lb = load_box(d.laabb + i);
            store_box(d.aabb + i, Box{rsv::dadd(lb.minx, p.x), rsv::dadd(lb.miny, p.y), rsv::dadd(lb.maxx, p.x), rsv::dadd(lb.maxy, p.y)});   // (same value from every writer)
        }
        *reinterpret_cast<uint4*>(d.hits + k) = make_uint4(A, B, 0u, k << 8);
    }
}

void* rsv_pair_staging(rsv_space* s, uint32_t max_pairs, size_t* bytes) {
    if (!s || !max_pairs) return nullptr;
    if (cudaSetDevice(s->device) != cudaSuccess) return nullptr;
    if (max_pairs > s->cap_pair_list) {
        if (s->h_pair_list) cudaFreeHost(s->h_pair_list);
        if (s->d_pair_list) cudaFree(s->d_pair_list);
        s->h_pair_list = nullptr; s->d_pair_list = nullptr; s->cap_pair_list = 0;
        if (cudaHostAlloc(&s->h_pair_list, (size_t)max_pairs * 8, cudaHostAllocDefault) != cudaSuccess ||
            cudaMalloc(&s->d_pair_list, (size_t)max_pairs * 8) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        s->cap_pair_list = max_pairs;
    }
    if (bytes) *bytes = (size_t)s->cap_pair_list * 8;
    return s->h_pair_list;
}

int32_t rsv_test_pairs(rsv_space* s, uint32_t n_pairs, rsv_counts* out) {
    if (!s) return RSV_ERR_BAD_ARG;
    if (n_pairs > s->cap_pair_list) return fail(s, RSV_ERR_BAD_ARG, "rsv_test_pairs: %u pairs > staged capacity %u", n_pairs, s->cap_pair_list);
    if (n_pairs > s->d.cap_pairs) return fail(s, RSV_ERR_CAPACITY, "rsv_test_pairs: %u pairs > pair capacity %u (rsv_reserve)", n_pairs, s->d.cap_pairs);
    for (uint32_t k = 0; k < 2 * n_pairs; k++)
        if (s->h_pair_list[k] >= s->n_shapes) return fail(s, RSV_ERR_BAD_ARG, "rsv_test_pairs: slot %u out of range", s->h_pair_list[k]);
    CU(s, cudaSetDevice(s->device));
    for (auto& sl : s->slot) sl.pending = false;   // frames launched earlier and never finished are dropped
    s->finish_slot = s->launch_slot;
    rsv_space::FrameSlot& sl = s->slot[s->launch_slot];
    DevView& d = s->d;
    d.hits = sl.hits; d.contacts = sl.contacts; d.kind_queue = sl.kind_queue;
    cudaStream_t st = s->stream;
    {
        const int32_t rcg = ensure_full_grid(s);   // (its counters are about to be cleared)
        if (rcg != RSV_OK) return rcg;
    }
    k_clear<<<1, kGridBlock, 0, st>>>(d.tile_state, 0u, reinterpret_cast<uint32_t*>(d.ctr), (uint32_t)(offsetof(Counters, ray_hit_cursor) / 4));
    if (n_pairs) {
        CU(s, cudaMemcpyAsync(s->d_pair_list, s->h_pair_list, (size_t)n_pairs * 8, cudaMemcpyHostToDevice, st));
        k_pairs_from_list<<<blocks_for(n_pairs, 256, s->sm_count * 4), 256, 0, st>>>(d, s->d_pair_list, n_pairs);
        k_bin<<<blocks_for(n_pairs, 256, s->sm_count * 8), 256, 0, st>>>(d, false);
        k_bin_scatter<<<blocks_for(n_pairs, 256, s->sm_count * 4), 256, 0, st>>>(d);
        k_narrow<1><<<s->narrow_blocks1, kNarrowBlock, 0, st>>>(d);
        k_narrow<0><<<s->narrow_blocks, kNarrowBlock, 0, st>>>(d);
    }
    CU(s, cudaGetLastError());
    CU(s, cudaMemcpyAsync(sl.h_ctr, d.ctr, sizeof(Counters), cudaMemcpyDeviceToHost, st));
    CU(s, cudaEventRecord(sl.ev_done, st));
    sl.flags = RSV_FRAME_DOWNLOAD;
    sl.pending = true;
    sl.cellpairs = false;
    sl.pair_list = true;
    s->launch_slot ^= 1;
    const int32_t rc = rsv_frame_finish(s, out);
    if (out) { out->n_testers = 0; out->n_entries = 0; out->n_candidates = n_pairs; }
    return rc;
}

int32_t rsv_results_view(rsv_space* s, rsv_results* out) {
    if (!s || !out) return RSV_ERR_BAD_ARG;
    out->hits = s->slot[s->last_slot].h_hits;
    out->n_hit_records = s->last_records;
    out->contacts = s->slot[s->last_slot].h_contacts;
    out->n_contacts = s->last_contacts;
    return RSV_OK;
}

int32_t rsv_timing(rsv_space* s, rsv_timing_info* out) {
    if (!s || !out) return RSV_ERR_BAD_ARG;
    *out = s->timing;
    return RSV_OK;
}

int32_t rsv_timer_start(rsv_space* s) {
    if (!s) return RSV_ERR_BAD_ARG;
    CU(s, cudaSetDevice(s->device));
    CU(s, cudaEventRecord(s->ev_timer[0], s->stream));
    return RSV_OK;
}

int32_t rsv_timer_stop(rsv_space* s, float* ms) {
    if (!s || !ms) return RSV_ERR_BAD_ARG;
    CU(s, cudaSetDevice(s->device));
    CU(s, cudaEventRecord(s->ev_timer[1], s->stream));
    CU(s, cudaEventSynchronize(s->ev_timer[1]));
    CU(s, cudaEventElapsedTime(ms, s->ev_timer[0], s->ev_timer[1]));
    return RSV_OK;
}

int32_t rsv_debug_cells(rsv_space* s, uint32_t* cell_end, uint64_t n_cells, uint32_t* entries, uint64_t cap_entries) {
    if (!s) return RSV_ERR_BAD_ARG;
    CU(s, cudaSetDevice(s->device));
    {
        const int32_t rcg = ensure_full_grid(s);
        if (rcg != RSV_OK) return rcg;
    }
    CU(s, cudaStreamSynchronize(s->stream));
    if (cell_end) {
        if (n_cells != s->d.n_cells) return fail(s, RSV_ERR_BAD_ARG, "rsv_debug_cells: expected %u cells", s->d.n_cells);
        CU(s, cudaMemcpy(cell_end, s->d.cell + 1, (size_t)n_cells * 4, cudaMemcpyDeviceToHost));
    }
    if (entries) {
        uint64_t R = s->last_counts.n_entries;
        if (cap_entries < R) return fail(s, RSV_ERR_BAD_ARG, "rsv_debug_cells: need room for %llu entries", (unsigned long long)R);
        if (R > s->d.cap_entries) return fail(s, RSV_ERR_CAPACITY, "rsv_debug_cells: last frame overflowed the entry buffer");
        std::vector<uint32_t> tmp(R);
        if (R) CU(s, cudaMemcpy(tmp.data(), s->d.entries, (size_t)R * 4, cudaMemcpyDeviceToHost));
        for (uint64_t k = 0; k < R; k++) entries[k] = tmp[k] >> kEntShift;
    }
    return RSV_OK;
}

int32_t rsv_debug_cell_rects(rsv_space* s, int32_t* rects4, uint64_t n_shapes) {
    if (!s || !rects4) return RSV_ERR_BAD_ARG;
    if (n_shapes != s->d.n_shapes) return fail(s, RSV_ERR_BAD_ARG, "rsv_debug_cell_rects: expected %u shapes", s->d.n_shapes);
    CU(s, cudaSetDevice(s->device));
    CU(s, cudaStreamSynchronize(s->stream));
    if (n_shapes) CU(s, cudaMemcpy(rects4, s->d.crect, (size_t)n_shapes * 16, cudaMemcpyDeviceToHost));
    return RSV_OK;
}

}  // extern "C"

#include "rsv_lines_abi.inl"
#include "rsv_halo_abi.inl"
