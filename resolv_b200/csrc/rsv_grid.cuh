// rsv_grid.cuh — per-frame bulk rebuild of the cell grid.
//
// Replaces the reference's incremental, pointer-chasing maintenance
//   ShapeBase.update -> removeFromTouchingCells / addToTouchingCells   (shape.go:183-214,239-242)
//   Cell.register / unregister / Contains                              (cell.go:19-48)
//   Space.Cell (nil outside the grid)                                  (space.go:135-142)
// with a counting sort of (cell, shape) registrations: k_bounds counts per cell (RED atomics), k_scan turns
// counts into offsets (single-pass chained scan), k_scatter places shape slots, k_cellsort orders each cell by
// slot so that the in-cell order equals the reference's for a freshly built Space (append order, cell.go:21).
// The resulting membership relation "shape in cell(x,y)" is the reference's, bit for bit.
#pragma once
#include "rsv_device.cuh"

namespace rsv {

constexpr int kGridBlock = 256;

// In-grid part of a shape's inclusive cell rect; false if the shape touches no stored cell.
__device__ __forceinline__ bool clamp_rect(const DevView& d, int4 r, int& x0, int& y0, int& x1, int& y1) {
    x0 = max(r.x, 0);
    y0 = max(r.y, d.row_lo);
    x1 = min(r.z, d.W - 1);
    y1 = min(r.w, d.row_hi - 1);
    return x0 <= x1 && y0 <= y1;
}

// The shapes a frame processes: the owned slot range plus the halo ("ghost") slots received this frame. For a
// single-GPU Space the owned range is [0, n_shapes) and there are no ghosts.
// In the per-frame part of an incremental frame (kSubsetDynamic) they are the slots of dyn_list instead.
__device__ __forceinline__ uint32_t n_items(const DevView& d) {
    if (d.subset == kSubsetDynamic) return d.n_dyn;
    if (d.part == kPartOwned) return d.own_hi - d.own_lo;   // (the ghosts are registered by k_halo_apply_bounds)
    uint32_t ng = d.n_ghosts ? min(*d.n_ghosts, d.cap_ghosts) : 0u;
    return (d.own_hi - d.own_lo) + ng;
}
__device__ __forceinline__ uint32_t item_slot(const DevView& d, uint32_t j) {
    if (d.subset == kSubsetDynamic) return d.dyn_list[j];
    const uint32_t n_own = d.own_hi - d.own_lo;
    return j < n_own ? d.own_lo + j : d.ghosts[j - n_own];
}
// Strip partitioning: a shape runs its IntersectionTest on the device whose home rows contain its first cell row
// (clamped into the grid), so every ordered pair is tested exactly once across devices.
__device__ __forceinline__ bool is_home(const DevView& d, int4 r) {
    const int hy = min(max(r.y, 0), d.H - 1);
    return hy >= d.home_lo && hy < d.home_hi;
}

__device__ __forceinline__ uint32_t cell_base(const DevView& d, uint32_t i) {
    return d.n_spaces > 1 ? __ldg(d.space_idx + i) * d.cells_per_space : 0u;
}

// Shapes covering more than kBigCells cells (walls, floors) are not walked by their own thread: the warp takes them
// one at a time and strides its 32 lanes over the cells (a 640x16 wall on 16 px cells is 80 dependent atomics
// otherwise — 75 us of a 300-shape frame were measured in k_scatter alone).
constexpr int kBigCells = 16;

// Calls f(x, y) for every cell of the rect [x0,x1]x[y0,y1] of every lane that has `big` set, cooperatively.
template <class F>
__device__ __forceinline__ void warp_big_rects(bool big, int x0, int y0, int x1, int y1, uint32_t payload, uint32_t base, F f) {
    const int lane = threadIdx.x & 31;
    for (uint32_t m = __ballot_sync(0xffffffffu, big); m; m &= m - 1) {
        const int src = __ffs(m) - 1;
        const int bx0 = __shfl_sync(0xffffffffu, x0, src), by0 = __shfl_sync(0xffffffffu, y0, src);
        const int bx1 = __shfl_sync(0xffffffffu, x1, src), by1 = __shfl_sync(0xffffffffu, y1, src);
        const uint32_t pl = __shfl_sync(0xffffffffu, payload, src), bb = __shfl_sync(0xffffffffu, base, src);
        const int w = bx1 - bx0 + 1;
        const long long n = (long long)w * (long long)(by1 - by0 + 1);
        for (long long c = lane; c < n; c += 32) f(bx0 + (int)(c % w), by0 + (int)(c / w), bx0, by0, bx1, by1, pl, bb);
    }
}

// Halo exchange folded into the grid kernels (strip partitioning over peer memory, rsv_halo.cuh): k_bounds<true> packs
// the owned shapes that reach a neighbour's stored rows straight into that neighbour's mailbox while it registers
// them, its last block publishes the counts and raises the neighbours' `ready` flags; k_halo_apply_bounds then waits
// for this device's own flags, applies the received positions and registers the ghosts. Epochs live on the device, so
// a strip frame is one fixed launch sequence (one CUDA graph). (Tried: a separate pack + publish pass IN FRONT of
// k_clear / k_bounds so that the neighbours' batches travel while the owned shapes are registered — the extra pass
// over the owned shapes cost 28 us and the frame got slower, 0.501 against 0.485 ms at N = 2.)
struct HaloRec;
struct HaloMailboxHeader;
struct HaloArgs {
    HaloMailboxHeader* mine;
    HaloRec* dst[2][2];          // [side: 0 upper, 1 lower neighbour][epoch parity]: where our records go (peer memory)
    uint32_t* peer_ready[2][2];  // the neighbours' `ready` words for data from our side
    uint32_t* peer_done[2][2];   // the neighbours' `done` words for our side
    const HaloRec* recv[2][2];   // our own receive buffers
    int32_t up_row_max, dn_row_min;
    int32_t slot_off[2];
    uint32_t cap;
};

// Bounds() -> toCellSpace() per shape, and the per-cell registration counts.
//   Circle.Bounds                circle.go:33-39      (pos +- r  ==  pos + (-r, +r), exact in IEEE)
//   ConvexPolygon.Bounds         convexPolygon.go:158-161 (cached local bounds + position)
//   Bounds.toCellSpace           utils.go:90-98
//   addToTouchingCells           shape.go:191-214     (cells outside the grid are skipped)
// One shape per lane; the whole warp must call this together (shapes covering many cells are registered cooperatively).
// Returns the shape's cell rect in `r_out` (for the halo pack).
__device__ __forceinline__ void bounds_item(const DevView& d, int margin, bool valid, uint32_t i, double2 p, unsigned long long& my_entries,
                                            unsigned long long& my_testers, int4& r_out, uint32_t& meta_out) {
    bool big = false, listed = false;
    int x0 = 0, y0 = 0, x1 = -1, y1 = -1;
    uint32_t base = 0;
    if (valid) {
        uint32_t meta = __ldg(d.meta + i);
        meta_out = meta;
        Box lb = load_box(d.laabb + i);
        Box b{dadd(lb.minx, p.x), dadd(lb.miny, p.y), dadd(lb.maxx, p.x), dadd(lb.maxy, p.y)};
        int4 r;
        r.x = cell_of(b.minx, d.cell_w);
        r.y = cell_of(b.miny, d.cell_h);
        r.z = cell_of(b.maxx, d.cell_w);
        r.w = cell_of(b.maxy, d.cell_h);
        if (meta & RSV_META_ABSENT) r = make_int4(kClampCell, kClampCell, -kClampCell, -kClampCell);
        r_out = r;
        store_box(d.aabb + i, b);
        d.crect[i] = r;
        d.tester_count[i] = 0;
        const bool tester = (meta & RSV_META_TESTER) && !(meta & RSV_META_ABSENT) && is_home(d, r);
        if (tester) my_testers++;
        bool in_grid = clamp_rect(d, r, x0, y0, x1, y1);
        if (d.subset == kSubsetStaticBuild && !((meta & RSV_META_STATIC) && in_grid)) {
            // Static build: only RSV_META_STATIC shapes that hold cells go into the cached grid. Non-static shapes are
            // appended to dyn_list and handled by the per-frame part; a static tester outside the grid is listed too
            // (it may be an orphan tester, whose list is rebuilt every frame, and it has no registrations to repeat).
            listed = !(meta & RSV_META_ABSENT) && (!(meta & RSV_META_STATIC) || tester);
            in_grid = false;      // not counted here
            if (tester) my_testers--;   // (a tester that is not cached is always listed: counted per frame)
        }
        if (tester && !in_grid && d.subset != kSubsetStaticBuild) {
            // SelectTouchingCells(margin) can still reach stored cells (shape.go:220-237): k_pairs finds testers
            // through their home registration, so these go to a side list.
            int qx0 = max(r.x - margin, 0), qx1 = min(r.z + margin, d.W - 1);
            int qy0 = max(r.y - margin, d.row_lo), qy1 = min(r.w + margin, d.row_hi - 1);
            if (qx0 <= qx1 && qy0 <= qy1) d.orphans[atomicAdd(&d.ctr->n_orphans, 1u)] = i;
        }
        if (in_grid) {
            base = cell_base(d, i);
            const long long ncell = (long long)(x1 - x0 + 1) * (long long)(y1 - y0 + 1);
            my_entries += (unsigned long long)ncell;
            big = ncell > kBigCells;
            if (!big)
                for (int y = y0; y <= y1; y++) {
                    uint32_t row = base + (uint32_t)(y - d.row_lo) * (uint32_t)d.W;
                    for (int x = x0; x <= x1; x++) atomicAdd(d.cnt + row + x + 1, 1u);
                }
        }
    }
    warp_big_rects(big, x0, y0, x1, y1, 0u, base, [&](int x, int y, int, int, int, int, uint32_t, uint32_t bb) {
        atomicAdd(d.cnt + bb + (uint32_t)(y - d.row_lo) * (uint32_t)d.W + x + 1, 1u);
    });
    if (d.subset == kSubsetStaticBuild) {   // warp-aggregated append keeps dyn_list nearly in slot order
        const uint32_t lm = __ballot_sync(0xffffffffu, listed);
        uint32_t b0 = 0;
        if ((threadIdx.x & 31) == 0 && lm) b0 = atomicAdd(d.n_dyn_dev, (uint32_t)__popc(lm));
        b0 = __shfl_sync(0xffffffffu, b0, 0);
        if (listed) d.dyn_list[b0 + __popc(lm & ((1u << (threadIdx.x & 31)) - 1u))] = i;
    }
}

// Block reduction of the two statistics -> one atomic each per block.
__device__ __forceinline__ void bounds_stats(const DevView& d, unsigned long long my_entries, unsigned long long my_testers, bool add_static) {
    for (int o = 16; o > 0; o >>= 1) {
        my_entries += __shfl_down_sync(0xffffffffu, my_entries, o);
        my_testers += __shfl_down_sync(0xffffffffu, my_testers, o);
    }
    __shared__ unsigned long long s_e[kGridBlock / 32], s_t[kGridBlock / 32];
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { s_e[warp] = my_entries; s_t[warp] = my_testers; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long e = 0, t = 0;
        for (int w = 0; w < kGridBlock / 32; w++) { e += s_e[w]; t += s_t[w]; }
        if (add_static) { e += d.static_entries; t += d.static_testers; }
        if (e) atomicAdd(&d.ctr->n_entries, e);
        if (t) atomicAdd(&d.ctr->n_testers, t);
    }
}

__device__ __forceinline__ uint32_t ld_vol_u32(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }
__device__ __forceinline__ void st_vol_u32(uint32_t* p, uint32_t v) { *reinterpret_cast<volatile uint32_t*>(p) = v; }

template <bool HALO>
__global__ void __launch_bounds__(kGridBlock) k_bounds(DevView d, int margin, HaloArgs h);

// Zeroes what a frame accumulates into besides the per-cell counters (those are zeroed by k_scan as it consumes them):
// scan tile states and frame counters. One small launch.
__global__ void __launch_bounds__(kGridBlock) k_clear(unsigned long long* __restrict__ state, uint32_t n_tiles,
                                                      uint32_t* __restrict__ ctr_words, uint32_t n_ctr_words) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    for (uint32_t i = tid; i < n_tiles; i += stride) state[i] = 0ull;
    for (uint32_t i = tid; i < n_ctr_words; i += stride) ctr_words[i] = 0u;
}

// ------------------------------------------------------------------------------------------------------------
// Single-pass chained exclusive scan of uint32: one read of the counters, one write of the offsets — and one write of
// zeros over the counters, so that the next frame finds them cleared (a separate 4 (C + 1)-byte clear kernel cost
// 10 us per frame; this kernel is one latency-bound wave and the extra stores ride along). Tiles take their id from an
// atomic ticket so a tile only ever waits on tiles that already started.
constexpr int kScanBlock = 256;
#ifndef RSV_SCAN_ITEMS
#define RSV_SCAN_ITEMS 32
#endif
constexpr int kScanItems = RSV_SCAN_ITEMS;  // per thread, as 8 lane-striped uint4 (every load instruction reads 512 contiguous bytes); 64 -> 32: 25 -> 20 us (513 tiles, 64 registers)
constexpr int kScanTile = kScanBlock * kScanItems;
// Tiles are large on purpose: all tiles of a frame are resident at once, so a tile's look-back finds only
// aggregates (no finished prefixes) behind it and needs ~tile_index/32 serial rounds; 16 K-item tiles keep that short.

__device__ __forceinline__ unsigned long long ld_state(const unsigned long long* p) {
    return *reinterpret_cast<const volatile unsigned long long*>(p);
}
__device__ __forceinline__ void st_state(unsigned long long* p, unsigned long long v) {
    *reinterpret_cast<volatile unsigned long long*>(p) = v;
}

// Exclusive scan of the per-cell counters (data[c + 1] = count of cell c) into out[c + 1] = first entry of cell c
// (out[0] = 0); data is left zeroed. Also appends (first entry, entry count) of every cell that holds two or more
// entries to `multi` — the cells whose entries k_cellsort has to order and k_cellpairs has to pair up; they get the
// cell's run of the entry array from one coalesced load and never touch the offset array.
//
// MERGE (incremental frames): `data` holds the counts of the DYNAMIC registrations only and `scount` the static
// ones of the cached static grid, in the same layout. The scan runs over dynamic + static counts, and
// data[c + 1] is left at "first entry of cell c + its static count": the static entries are copied to the front of
// the cell by k_static_place and the cursor atomics of k_scatter append the dynamic ones behind them. `multi` then
// lists the cells that received at least one dynamic registration (k_cellfix).
template <bool MERGE>
__global__ void __launch_bounds__(kScanBlock) k_scan(uint32_t* __restrict__ data, uint32_t* __restrict__ out, uint32_t n,
                                                     unsigned long long* __restrict__ state, unsigned int* ticket,
                                                     uint2* __restrict__ multi, unsigned int* n_multi, uint32_t cap_multi,
                                                     const uint32_t* __restrict__ scount) {
    __shared__ uint32_t s_tile, s_prefix, s_warp[kScanBlock / 32], s_mw[kScanBlock / 32], s_mbase;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int kGroups = kScanItems / 4;                  // uint4 groups per lane
    constexpr uint32_t kWarpItems = 32 * kScanItems;         // items per warp
    const uint32_t wbase = tile * kScanTile + warp * kWarpItems;
    // group g of lane l covers items wbase + (g * 32 + l) * 4 .. + 3 (the array is allocated in whole tiles)
    uint4 q[kGroups];
    uint4* src = reinterpret_cast<uint4*>(data + wbase) + lane;
#pragma unroll
    for (int g = 0; g < kGroups; g++) {
        const uint32_t e0 = wbase + (g * 32 + lane) * 4;
        q[g] = e0 < n ? src[g * 32] : make_uint4(0u, 0u, 0u, 0u);
        if (e0 < n) src[g * 32] = make_uint4(0u, 0u, 0u, 0u);   // consumed: cleared for the next frame
        if (e0 < n && e0 + 4 > n) {  // last, partial group
            if (e0 + 1 >= n) q[g].y = 0u;
            if (e0 + 2 >= n) q[g].z = 0u;
            q[g].w = 0u;
        }
    }
    // cells with >= 2 entries (array index e holds the count of cell e - 1); MERGE: cells with a dynamic entry.
    // Their places in the list are reserved now; the records are written at the end, when the offsets are known.
    constexpr uint32_t kListed = MERGE ? 1u : 2u;
    uint32_t m = 0;
#pragma unroll
    for (int g = 0; g < kGroups; g++) m += (q[g].x >= kListed) + (q[g].y >= kListed) + (q[g].z >= kListed) + (q[g].w >= kListed);
    uint32_t minc = m;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, minc, o);
        if (lane >= o) minc += t;
    }
    // (one allocation per BLOCK, issued by thread 0 behind the first barrier and picked up at the output: a returning
    // atomic per warp on the one list counter — 4 k of them, each waited for on the spot — was 12 % of this kernel's stalls)
    if (lane == 31) s_mw[warp] = minc;
    // MERGE: static counts of the same cells, in the same layout (index e: static entries of cell e - 1)
    // (they are read again for the output instead of being kept: 32 registers less, and the second read hits L1/L2)
    const uint4* ssrc = reinterpret_cast<const uint4*>(scount + wbase) + lane;
    if (MERGE) {
#pragma unroll
        for (int g = 0; g < kGroups; g++) {
            const uint32_t e0 = wbase + (g * 32 + lane) * 4;
            const uint4 c4 = e0 < n ? __ldg(ssrc + g * 32) : make_uint4(0u, 0u, 0u, 0u);   // (zero past n_cells)
            q[g].x += c4.x; q[g].y += c4.y; q[g].z += c4.z; q[g].w += c4.w;   // totals from here on
        }
    }
    // warp-level exclusive scan in memory order: groups are g-major, lanes within a group
    uint32_t excl[kGroups];
    uint32_t carry = 0;
#pragma unroll
    for (int g = 0; g < kGroups; g++) {
        const uint32_t s4 = q[g].x + q[g].y + q[g].z + q[g].w;
        uint32_t inc = s4;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        excl[g] = carry + inc - s4;
        carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) s_warp[warp] = carry;  // warp total
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t mt = 0;
#pragma unroll
        for (int w = 0; w < kScanBlock / 32; w++) mt += s_mw[w];
        s_mbase = mt ? atomicAdd(n_multi, mt) : 0u;
    }
    uint32_t warp_excl = 0, block_sum = 0;
#pragma unroll
    for (int w = 0; w < kScanBlock / 32; w++) {
        uint32_t t = s_warp[w];
        if (w < warp) warp_excl += t;
        block_sum += t;
    }
    // Prefix of this tile = sum of the aggregates of ALL earlier tiles, gathered by the whole block at once (thread t
    // takes tiles t, t + 256, ...). Every tile publishes its aggregate as soon as it is reduced and tiles take their
    // ids from the ticket, so a tile only ever waits on tiles that started before it. (The usual decoupled look-back
    // walks back 32 tiles per step until it meets a finished prefix; with every tile of a frame resident at once there
    // are none to meet, and the last of 513 tiles paid 16 dependent L2 round trips: 21 us for 34 MB.)
    if (threadIdx.x == 0) st_state(state + tile, (1ull << 32) | block_sum);
    uint32_t mine = 0;
    for (uint32_t j = threadIdx.x; j < tile; j += kScanBlock) {
        unsigned long long sv;
        do { sv = ld_state(state + j); } while ((sv >> 32) == 0);
        mine += (uint32_t)sv;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    __syncthreads();                 // (everyone has read s_warp[])
    if (lane == 0) s_warp[warp] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t ex = 0;
#pragma unroll
        for (int w = 0; w < kScanBlock / 32; w++) ex += s_warp[w];
        s_prefix = ex;
    }
    __syncthreads();
    const uint32_t pre = s_prefix + warp_excl;
    uint32_t mbase = s_mbase + (minc - m);
#pragma unroll
    for (int w = 0; w < kScanBlock / 32; w++)
        if (w < warp) mbase += s_mw[w];
    uint4* dst = reinterpret_cast<uint4*>(out + wbase) + lane;
#pragma unroll
    for (int g = 0; g < kGroups; g++) {
        const uint32_t e0 = wbase + (g * 32 + lane) * 4;
        if (e0 < n) {  // (elements past n inside the last group belong to the allocation's padding)
            uint32_t run = pre + excl[g];
            uint4 o;
            o.x = run; run += q[g].x;
            o.y = run; run += q[g].y;
            o.z = run; run += q[g].z;
            o.w = run;
            uint4 c4 = make_uint4(0u, 0u, 0u, 0u);
            if (MERGE) c4 = __ldg(ssrc + g * 32);
            if (m) {   // (first entry, entries) of the listed cells; listed by their DYNAMIC count in merged frames
                if (q[g].x - c4.x >= kListed) { if (mbase < cap_multi) multi[mbase] = make_uint2(o.x, q[g].x); mbase++; }
                if (q[g].y - c4.y >= kListed) { if (mbase < cap_multi) multi[mbase] = make_uint2(o.y, q[g].y); mbase++; }
                if (q[g].z - c4.z >= kListed) { if (mbase < cap_multi) multi[mbase] = make_uint2(o.z, q[g].z); mbase++; }
                if (q[g].w - c4.w >= kListed) { if (mbase < cap_multi) multi[mbase] = make_uint2(o.w, q[g].w); mbase++; }
            }
            if (MERGE) { o.x += c4.x; o.y += c4.y; o.z += c4.z; o.w += c4.w; }
            dst[g * 32] = o;
        }
    }
}

// Entry encoding: slot << 6 | useAny << 5 | lastCol << 4 | lastRow << 3 | tester << 2 | firstCol << 1 | firstRow. The first/last flags say that this cell is
// the first in-grid column / row of the shape's own cell rect; with them the candidate walk can apply the
// owner-cell rule (emit a shape only from the first cell it shares with the query rect, i.e. the first
// occurrence in the reference's row-major walk, cell.go:86-121) without loading the other shape's rect.
constexpr uint32_t kEntFirstRow = 1u, kEntFirstCol = 2u, kEntTester = 4u, kEntLastRow = 8u, kEntLastCol = 16u, kEntUseAny = 32u;
constexpr int kEntShift = 6;

__global__ void __launch_bounds__(kGridBlock) k_scatter(DevView d) {
    const uint32_t n = n_items(d);
    const uint32_t stride = gridDim.x * blockDim.x;
    auto place = [&](int x, int y, int x0, int y0, int x1, int y1, uint32_t tag, uint32_t base) {
        const uint32_t slot = atomicAdd(d.cell + base + (uint32_t)(y - d.row_lo) * (uint32_t)d.W + x + 1, 1u);
        if (slot < d.cap_entries)
            d.entries[slot] = tag | (x == x0 ? kEntFirstCol : 0u) | (y == y0 ? kEntFirstRow : 0u) |
                              (x == x1 ? kEntLastCol : 0u) | (y == y1 ? kEntLastRow : 0u);
    };
    for (uint32_t j0 = blockIdx.x * blockDim.x; j0 < n; j0 += stride) {  // (warp-uniform trip count)
        const uint32_t j = j0 + threadIdx.x;
        bool big = false;
        int x0 = 0, y0 = 0, x1 = -1, y1 = -1;
        uint32_t base = 0, tag = 0;
        if (j < n) {
            const uint32_t i = item_slot(d, j);
            const int4 r = d.crect[i];
            if (clamp_rect(d, r, x0, y0, x1, y1)) {
                const uint32_t meta = __ldg(d.meta + i);
                base = cell_base(d, i);
                tag = (i << kEntShift) | (((meta & RSV_META_TESTER) && is_home(d, r)) ? kEntTester : 0u) |
                      ((meta & RSV_META_USE_ANY) ? kEntUseAny : 0u);
                big = (long long)(x1 - x0 + 1) * (long long)(y1 - y0 + 1) > kBigCells;
                // (static build: only the RSV_META_STATIC shapes are placed)
                const bool skip = d.subset == kSubsetStaticBuild && !(meta & RSV_META_STATIC);
                if (skip) big = false;
                if (!big && !skip)
                    for (int y = y0; y <= y1; y++)
                        for (int x = x0; x <= x1; x++) place(x, y, x0, y0, x1, y1, tag, base);
            }
        }
        warp_big_rects(big, x0, y0, x1, y1, tag, base, place);
    }
}

// Orders the entries of every cell that holds two or more by shape slot (the atomics above place them in arrival
// order). The list of such cells comes from k_scan.
__global__ void __launch_bounds__(kGridBlock) k_cellsort(DevView d) {
    const uint32_t nm = min(d.ctr->n_multi, d.cap_multi);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nm; i += gridDim.x * blockDim.x) {
        const uint2 sn = d.multi[i];
        const uint32_t s = sn.x, n = sn.y;
        if (n < 2 || (unsigned long long)s + n > d.cap_entries) continue;
        uint32_t* a = d.entries + s;
        for (uint32_t i = 1; i < n; i++) {
            uint32_t key = a[i];
            uint32_t j = i;
            while (j > 0 && a[j - 1] > key) { a[j] = a[j - 1]; j--; }
            a[j] = key;
        }
    }
}

// Fills the per-entry float32 boxes in final (sorted) entry order: one gather of the shape's AABB per entry,
// coalesced 16 B stores. (Writing them from k_scatter instead would be 16 B stores to random slots.) For a shape's
// "home" entry (first row and first column of its cell rect) it also records where that cell is and how many
// columns / rows the clamped rect spans, so that k_pairs can build the tester's query rect from coalesced loads
// instead of gathering crect[] / space_idx[] by shape slot inside its latency-critical prologue.
__global__ void __launch_bounds__(kGridBlock) k_entboxes(DevView d, uint32_t flags) {
    const bool filters = !(flags & RSV_FRAME_NO_TAG_FILTERS);
    const uint32_t R = (uint32_t)min((unsigned long long)d.cap_entries, d.ctr->n_entries);
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < R; k += gridDim.x * blockDim.x) {
        const uint32_t ent = d.entries[k];
        const uint32_t i = ent >> kEntShift;
        d.ent_box[k] = float_box(load_box_nc(d.aabb + i));
        if (filters) d.ent_tags[k] = __ldg(d.tags + i);  // so that k_pairs filters on a coalesced stream
        if ((ent & (kEntFirstRow | kEntFirstCol)) == (kEntFirstRow | kEntFirstCol)) {
            int x0, y0, x1, y1;
            clamp_rect(d, d.crect[i], x0, y0, x1, y1);
            const uint32_t c = cell_base(d, i) + (uint32_t)(y0 - d.row_lo) * (uint32_t)d.W + (uint32_t)x0;
            d.ent_home[k] = make_uint2(c, (uint32_t)(x1 - x0) | ((uint32_t)(y1 - y0) << 16));
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Incremental grid (RSV_FRAME_INCREMENTAL): the registrations of RSV_META_STATIC shapes are kept from a one-off
// static build (same kernels, kSubsetStaticBuild, into the s* arrays below); a frame only re-registers the shapes
// of dyn_list. Mirrors what the reference does for a mostly static world: ShapeBase.update (shape.go:239-242)
// unregisters / re-registers the shape that moved and leaves every other cell membership alone. The merged grid
// is bit-identical to a full rebuild (same cell contents, ascending slot within a cell).
struct StaticGrid {
    uint32_t* cell = nullptr;      // [n_cells + 1] (whole scan tiles): offsets into the arrays below
    uint32_t* entries = nullptr;   // static registrations, cell-major, ascending slot within a cell
    float4* ent_box = nullptr;
    uint2* ent_home = nullptr;
    unsigned long long* ent_tags = nullptr;
    uint32_t* cell_of = nullptr;   // per static registration: its cell
    uint32_t* count = nullptr;     // [n_cells + 1] (whole scan tiles): count[c + 1] = static entries of cell c (the scan's input layout)
    uint32_t cap = 0;              // capacity of the per-registration arrays
};

// Cell of every static registration and the per-cell static counts (static build only; count[] was zeroed).
__global__ void __launch_bounds__(kGridBlock) k_static_cells(const uint32_t* __restrict__ scell, uint32_t n_cells,
                                                             uint32_t* __restrict__ cell_of, uint32_t* __restrict__ count, uint32_t cap) {
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < n_cells; c += gridDim.x * blockDim.x) {
        const uint32_t s = scell[c], e1 = scell[c + 1], e = min(e1, cap);
        count[c + 1] = e1 - s;
        for (uint32_t k = s; k < e; k++) cell_of[k] = c;
    }
}

// Copies the static registrations (entry word, float box, home record, tags) to the front of their cells in this
// frame's grid. After k_scan<true>, d.cell[c + 1] = first entry of cell c + static count of c, so static entry ks of
// cell c lands at ks + (d.cell[c + 1] - scell[c + 1]): fully coalesced reads, piecewise-contiguous writes.
__global__ void __launch_bounds__(kGridBlock) k_static_place(DevView d, StaticGrid sg, uint32_t flags) {
    const bool filters = !(flags & RSV_FRAME_NO_TAG_FILTERS);
    const uint32_t Rs = min(d.static_entries, sg.cap);
    for (uint32_t ks = blockIdx.x * blockDim.x + threadIdx.x; ks < Rs; ks += gridDim.x * blockDim.x) {
        const uint32_t c = __ldg(sg.cell_of + ks);
        const uint32_t dst = ks + (d.cell[c + 1] - __ldg(sg.cell + c + 1));
        if (dst >= d.cap_entries) continue;
        const uint32_t ent = __ldg(sg.entries + ks);
        d.entries[dst] = ent;
        d.ent_box[dst] = __ldg(sg.ent_box + ks);
        if ((ent & (kEntFirstRow | kEntFirstCol)) == (kEntFirstRow | kEntFirstCol)) d.ent_home[dst] = __ldg(sg.ent_home + ks);
        if (filters) d.ent_tags[dst] = __ldg(sg.ent_tags + ks);
    }
}

// Cells that received dynamic registrations this frame (listed by k_scan<true>): order the cell by slot (its static
// prefix is already in order) and refill the per-entry records of the whole cell, as k_cellsort + k_entboxes do for
// every cell in a full rebuild.
__global__ void __launch_bounds__(kGridBlock) k_cellfix(DevView d, uint32_t flags) {
    const bool filters = !(flags & RSV_FRAME_NO_TAG_FILTERS);
    const uint32_t nm = min(d.ctr->n_multi, d.cap_multi);
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < nm; t += gridDim.x * blockDim.x) {
        const uint2 sn = d.multi[t];
        const uint32_t s = sn.x, n = sn.y, e = s + n;
        if (n == 0 || (unsigned long long)s + n > d.cap_entries) continue;
        uint32_t* a = d.entries + s;
        for (uint32_t i = 1; i < n; i++) {
            const uint32_t key = a[i];
            uint32_t j = i;
            while (j > 0 && a[j - 1] > key) { a[j] = a[j - 1]; j--; }
            a[j] = key;
        }
        // refill in batches of four: all gathers of a batch are issued before its stores (the loop is a chain of
        // dependent round trips otherwise: ~15 us for ten thousand touched cells)
        for (uint32_t k0 = s; k0 < e; k0 += 4) {
            uint32_t ent[4];
            Box bx[4];
            unsigned long long tg[4];
            int4 cr[4];
#pragma unroll
            for (int u = 0; u < 4; u++) ent[u] = k0 + u < e ? a[k0 - s + u] : 0u;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t i = ent[u] >> kEntShift;
                const bool live = k0 + u < e;
                const bool home = (ent[u] & (kEntFirstRow | kEntFirstCol)) == (kEntFirstRow | kEntFirstCol);
                bx[u] = live ? load_box_nc(d.aabb + i) : Box{0.0, 0.0, 0.0, 0.0};
                tg[u] = (live && filters) ? __ldg(d.tags + i) : 0ull;
                cr[u] = (live && home) ? d.crect[i] : make_int4(0, 0, 0, 0);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t k = k0 + u;
                if (k >= e) break;
                const uint32_t i = ent[u] >> kEntShift;
                d.ent_box[k] = float_box(bx[u]);
                if (filters) d.ent_tags[k] = tg[u];
                if ((ent[u] & (kEntFirstRow | kEntFirstCol)) == (kEntFirstRow | kEntFirstCol)) {
                    int x0, y0, x1, y1;
                    clamp_rect(d, cr[u], x0, y0, x1, y1);
                    const uint32_t hc = cell_base(d, i) + (uint32_t)(y0 - d.row_lo) * (uint32_t)d.W + (uint32_t)x0;
                    d.ent_home[k] = make_uint2(hc, (uint32_t)(x1 - x0) | ((uint32_t)(y1 - y0) << 16));
                }
            }
        }
    }
}

}  // namespace rsv
