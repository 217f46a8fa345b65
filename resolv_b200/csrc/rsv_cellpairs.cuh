// rsv_cellpairs.cuh — the cell-pair path: Bounds-gated pairs found cell by cell, ranks derived afterwards.
//
// What a frame has to deliver for tester A is the reference's
//   for B in A.SelectTouchingCells(margin).FilterShapes()[.ByTags][.NotByTags]:  A.Intersection(B)      shape.go:220-237,273-316
// i.e. one IntersectionSet per candidate B that passes Bounds().IsIntersecting (the first statement of every pair
// test, shape.go:322,360,403,442; utils.go:146-173), plus each such B's position in possibleIntersections (the tie
// break of the distance sort, shape.go:291-293). The walk kernel (rsv_pairs.cuh) reproduces the reference's loop
// literally: every tester visits every entry of its margin-grown cell rect (11 per tester in the 1 M-shape configs)
// although 92 % of the candidates it finds fail the Bounds gate at once. This path turns the loop inside out:
//
//  1. k_cellpairs — two shapes whose Bounds intersect share a point, and the cell that holds the point holds both
//     registrations (both rects are clamped to the same grid; three pairwise-intersecting intervals have a common
//     point, so clamping cannot separate them). Gated pairs are therefore found among the entries of ONE cell: every
//     cell with two or more entries (listed by k_scan) is expanded into its unordered entry pairs, a pair is taken at
//     the first cell the two rects share ((A.firstCol | B.firstCol) & (A.firstRow | B.firstRow), the owner-cell rule
//     of cell.go:105 restricted to the tester's own rect), the exact float64 gate runs on the two AABBs, the tag
//     filters (shapefilter.go:44-61, tags.go:29-31) on the survivors, and (A, B) / (B, A) records are appended for
//     the sides that are testers. 1.5 M unordered pairs instead of 11 M visited entries in config 3. Testers that hold
//     no cell but whose margin reaches the grid (the orphan list of k_bounds) walk their query rect directly.
//  2. k_alloc — every tester with records reserves its range of the hit buffer (records are grouped by tester, as the
//     walk kernel's are; the tester table comes for free).
//  3. k_rank — per record, the candidate's index in possibleIntersections: the number of candidates the reference's
//     row-major walk over A's margin-grown rect (cell.go:86-121) offers before B — whole cell rows in front of B's
//     first shared row, whole cells in front of its first shared cell, and the entries of that cell with a smaller
//     slot (in-cell order of a freshly built Space is ascending slot, cell.go:21). Only the 0.17 records per tester
//     pay for a walk, and no in-cell ordering of the entry array is needed (k_cellsort / k_entboxes are skipped; they
//     run lazily if a LineTest or a grid read-back follows).
//  4. k_place_bin (rsv_narrow.cuh) — orders a tester's records by rank (generation order) and bins them for the
//     narrowphase in the same pass.
//
// The frame's outputs (hit records incl. rank, MTV, contacts, tester table) are identical to the walk kernel's;
// what is not produced is the statistic n_candidates (it needs the full walk: RSV_FRAME_CANDIDATE_WALK) and the
// RSV_FRAME_ALL_CANDIDATES dump.
#pragma once
#include "rsv_device.cuh"
#include "rsv_grid.cuh"

namespace rsv {

constexpr int kCpBlock = 128;
constexpr int kCpWarps = kCpBlock / 32;
constexpr int kCpSmall = 6;                                          // cells up to this size are expanded into pair items
constexpr int kCpQueue = 32 * (kCpSmall * (kCpSmall - 1) / 2) + 64;  // 32 cells' items + a partial double round carried over
constexpr int kCpStage = 288;                                        // staged records per warp (flushed from 224 on; an emit adds <= 64)

struct CpWarpSmem {
    uint32_t qs[kCpQueue];    // pair item: first entry of its cell
    uint16_t qij[kCpQueue];   //            i | j << 8 (entry indices within the cell, i < j)
    uint32_t stA[kCpStage];   // staged records
    uint32_t stB[kCpStage];
};

// ByTags / NotByTags (shapefilter.go:44-61; Has == any bit, tags.go:29-31) of a tester with masks qm against tags tg.
__device__ __forceinline__ bool cp_tags_ok(bool useAny, ulonglong2 qm, unsigned long long tg) {
    if (useAny && (tg & qm.x) == 0) return false;
    return (tg & qm.y) == 0;
}

__device__ __forceinline__ bool cp_gate(const Box& a, const Box& o) {
    return bounds_gate(a.minx, a.miny, a.maxx, a.maxy, o.minx, o.miny, o.maxx, o.maxy);
}

// Appends the staged records of the warp to the unordered record list; each record takes its index within its
// tester's records from the tester's counter.
__device__ __forceinline__ void cp_flush(const DevView& d, CpWarpSmem& sm, uint32_t& staged) {
    const int lane = threadIdx.x & 31;
    uint32_t base = 0;
    if (lane == 0 && staged) {
        base = atomicAdd(&d.ctr->pair_cursor, staged);
        if (base + staged > d.cap_pairs || base + staged < base) atomicOr(&d.ctr->overflow, RSV_OVF_PAIRS);
    }
    base = __shfl_sync(0xffffffffu, base, 0);
    for (uint32_t r = lane; r < staged; r += 32) {
        const unsigned long long pos = (unsigned long long)base + r;
        if (pos < d.cap_pairs) {
            const uint32_t A = sm.stA[r];
            const uint32_t idx = atomicAdd(d.grp_count + A, 1u);
            d.ptmp[pos] = make_uint4(A, sm.stB[r], idx, 0u);
        }
    }
    __syncwarp();
    staged = 0;
}

// Warp-wide: stages (A0, A1) for lanes with e01 and (A1, A0) for lanes with e10.
__device__ __forceinline__ void cp_emit(const DevView& d, CpWarpSmem& sm, uint32_t& staged, bool e01, bool e10, uint32_t A0, uint32_t A1) {
    const uint32_t lt = (1u << (threadIdx.x & 31)) - 1u;
    const uint32_t m0 = __ballot_sync(0xffffffffu, e01), m1 = __ballot_sync(0xffffffffu, e10);
    if (!(m0 | m1)) return;
    if (e01) { const uint32_t p = staged + __popc(m0 & lt); sm.stA[p] = A0; sm.stB[p] = A1; }
    if (e10) { const uint32_t p = staged + __popc(m0) + __popc(m1 & lt); sm.stA[p] = A1; sm.stB[p] = A0; }
    staged += __popc(m0) + __popc(m1);
    __syncwarp();
    if (staged >= kCpStage - 64) cp_flush(d, sm, staged);
}

// Two pair items per lane (queue slots r and r + 32). The kernel is bound by the latency of its gathers, not by bytes
// or instructions, so every phase issues the loads of both items before it looks at any of them: four entry words,
// then four AABBs, then the filter masks.
template <bool FILTERS>
__device__ __forceinline__ void cp_items(const DevView& d, CpWarpSmem& sm, uint32_t& staged, uint32_t r, uint32_t qn) {
    uint32_t e[2][2] = {{0u, 0u}, {0u, 0u}};
    bool pass[2];
#pragma unroll
    for (int u = 0; u < 2; u++) {
        pass[u] = r + 32u * u < qn;
        if (pass[u]) {
            const uint32_t s = sm.qs[r + 32u * u], ij = sm.qij[r + 32u * u];
            e[u][0] = d.entries[s + (ij & 0xffu)];
            e[u][1] = d.entries[s + (ij >> 8)];
        }
    }
    Box b[2][2];
#pragma unroll
    for (int u = 0; u < 2; u++) {
        const uint32_t w = e[u][0] | e[u][1];
        // some side is a tester, and this is the first cell the two rects share
        pass[u] = pass[u] && (w & kEntTester) && (w & kEntFirstCol) && (w & kEntFirstRow);
        if (pass[u]) {
            b[u][0] = load_box_nc(d.aabb + (e[u][0] >> kEntShift));
            b[u][1] = load_box_nc(d.aabb + (e[u][1] >> kEntShift));
        }
    }
    ulonglong2 qm[2][2];
    unsigned long long tg[2][2];
#pragma unroll
    for (int u = 0; u < 2; u++) {
        pass[u] = pass[u] && cp_gate(b[u][0], b[u][1]);   // (symmetric predicate)
        if (FILTERS && pass[u]) {
#pragma unroll
            for (int v = 0; v < 2; v++) {
                qm[u][v] = __ldg(d.qmask + (e[u][v] >> kEntShift));
                tg[u][v] = __ldg(d.tags + (e[u][v] >> kEntShift));
            }
        }
    }
#pragma unroll
    for (int u = 0; u < 2; u++) {
        bool e01 = pass[u] && (e[u][0] & kEntTester), e10 = pass[u] && (e[u][1] & kEntTester);
        if (FILTERS && pass[u]) {
            e01 = e01 && cp_tags_ok((e[u][0] & kEntUseAny) != 0, qm[u][0], tg[u][1]);
            e10 = e10 && cp_tags_ok((e[u][1] & kEntUseAny) != 0, qm[u][1], tg[u][0]);
        }
        cp_emit(d, sm, staged, e01, e10, e[u][0] >> kEntShift, e[u][1] >> kEntShift);
    }
}

template <bool FILTERS>
__global__ void __launch_bounds__(kCpBlock, 6) k_cellpairs(DevView d, int margin) {
    __shared__ CpWarpSmem s_warp[kCpWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    CpWarpSmem& sm = s_warp[warp];
    const uint32_t nm = min(d.ctr->n_multi, d.cap_multi);
    const uint32_t n_warps = gridDim.x * kCpWarps, wid = blockIdx.x * kCpWarps + warp;
    uint32_t qn = 0, staged = 0;
    uint2 nxt = make_uint2(0u, 0u);
    if (wid * 32u + lane < nm) nxt = d.multi[wid * 32u + lane];
    for (uint32_t base = wid * 32u; base < nm; base += n_warps * 32u) {
        const uint2 sn = nxt;                                // (first entry, entries) of a cell, straight from k_scan
        nxt = make_uint2(0u, 0u);
        if (base + n_warps * 32u + lane < nm) nxt = d.multi[base + n_warps * 32u + lane];   // next round's cells: in flight during this one
        const uint32_t s = sn.x;
        const uint32_t n = (unsigned long long)s + sn.y <= d.cap_entries ? sn.y : 0u;   // (an overflowed entry array fails the frame anyway)
        const bool small = n >= 2 && n <= (uint32_t)kCpSmall;
        const uint32_t np = small ? n * (n - 1) / 2 : 0u;
        uint32_t inc = np;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (small) {
            uint32_t at = qn + inc - np;
            for (uint32_t i = 0; i + 1 < n; i++)
                for (uint32_t j = i + 1; j < n; j++) { sm.qs[at] = s; sm.qij[at] = (uint16_t)(i | (j << 8)); at++; }
        }
        qn += __shfl_sync(0xffffffffu, inc, 31);
        __syncwarp();
        uint32_t r = 0;
        for (; r + 64 <= qn; r += 64) cp_items<FILTERS>(d, sm, staged, r + lane, qn);
        if (r) {   // carry the partial double round to the front
            const uint32_t rest = qn - r;
            uint32_t ts[2] = {0u, 0u}; uint16_t tij[2] = {0, 0};
#pragma unroll
            for (int u = 0; u < 2; u++)
                if (lane + 32u * u < rest) { ts[u] = sm.qs[r + lane + 32u * u]; tij[u] = sm.qij[r + lane + 32u * u]; }
            __syncwarp();
#pragma unroll
            for (int u = 0; u < 2; u++)
                if (lane + 32u * u < rest) { sm.qs[lane + 32u * u] = ts[u]; sm.qij[lane + 32u * u] = tij[u]; }
            qn = rest;
            __syncwarp();
        }
        // crowded cells: the warp takes them one at a time, lane = tester entry, all lanes stream the cell's entries
        for (uint32_t m = __ballot_sync(0xffffffffu, n > (uint32_t)kCpSmall); m; m &= m - 1) {
            const int src = __ffs(m) - 1;
            const uint32_t bs = __shfl_sync(0xffffffffu, s, src), bn = __shfl_sync(0xffffffffu, n, src);
            for (uint32_t i0 = 0; i0 < bn; i0 += 32) {
                const uint32_t ia = i0 + lane;
                const bool va = ia < bn;
                const uint32_t eA = va ? d.entries[bs + ia] : 0u;
                const uint32_t A = eA >> kEntShift;
                const bool ta = va && (eA & kEntTester);
                Box ba{0.0, 0.0, 0.0, 0.0};
                ulonglong2 qmA = make_ulonglong2(0ull, 0ull);
                if (ta) ba = load_box_nc(d.aabb + A);
                if (FILTERS && ta) qmA = __ldg(d.qmask + A);
                if (!__any_sync(0xffffffffu, ta)) continue;
                for (uint32_t j = 0; j < bn; j++) {
                    const uint32_t eB = d.entries[bs + j], B = eB >> kEntShift;   // (uniform address: one broadcast load)
                    const uint32_t u = eA | eB;
                    bool ok = ta && j != ia && (u & kEntFirstCol) && (u & kEntFirstRow);
                    if (ok) ok = cp_gate(ba, load_box_nc(d.aabb + B));
                    if (FILTERS && ok) ok = cp_tags_ok((eA & kEntUseAny) != 0, qmA, __ldg(d.tags + B));
                    cp_emit(d, sm, staged, ok, false, A, B);
                }
            }
        }
    }
    for (uint32_t r = 0; r < qn; r += 64) cp_items<FILTERS>(d, sm, staged, r + lane, qn);
    cp_flush(d, sm, staged);
    // testers that hold no stored cell but whose margin reaches the grid: the reference's walk, one thread each
    const uint32_t n_orph = min(d.ctr->n_orphans, d.n_shapes);
    for (uint32_t o = blockIdx.x * kCpBlock + threadIdx.x; o < n_orph; o += gridDim.x * kCpBlock) {
        const uint32_t A = d.orphans[o];
        const int4 r = d.crect[A];
        const int qx0 = max(r.x - margin, 0), qx1 = min(r.z + margin, d.W - 1);
        const int qy0 = max(r.y - margin, d.row_lo), qy1 = min(r.w + margin, d.row_hi - 1);
        if (qx0 > qx1 || qy0 > qy1) continue;
        const bool useAny = FILTERS && (__ldg(d.meta + A) & RSV_META_USE_ANY);
        ulonglong2 qmA = make_ulonglong2(0ull, 0ull);
        if (FILTERS) qmA = __ldg(d.qmask + A);
        const Box ba = load_box_nc(d.aabb + A);
        const uint32_t* cellp = d.cell + cell_base(d, A);
        for (int y = qy0; y <= qy1; y++) {
            const uint32_t* row = cellp + (size_t)(y - d.row_lo) * (size_t)d.W;
            const uint32_t s0 = min(row[qx0], d.cap_entries), s1 = min(row[qx0 + 1], d.cap_entries), e = min(row[qx1 + 1], d.cap_entries);
            for (uint32_t k = s0; k < e; k++) {
                const uint32_t eB = d.entries[k], B = eB >> kEntShift;
                if (B == A || !((eB & kEntFirstCol) || k < s1) || !((eB & kEntFirstRow) || y == qy0)) continue;
                if (!cp_gate(ba, load_box_nc(d.aabb + B))) continue;
                if (FILTERS && !cp_tags_ok(useAny, qmA, __ldg(d.tags + B))) continue;
                const uint32_t pos = atomicAdd(&d.ctr->pair_cursor, 1u);
                if (pos < d.cap_pairs) d.ptmp[pos] = make_uint4(A, B, atomicAdd(d.grp_count + A, 1u), 0u);
                else atomicOr(&d.ctr->overflow, RSV_OVF_PAIRS);
            }
        }
    }
}

// Per-tester record groups: the record that took index 0 of its tester reserves the tester's range of the hit buffer
// (one atomic per warp), publishes the tester table and re-arms the tester's counter for the next frame.
__global__ void __launch_bounds__(256) k_alloc(DevView d) {
    const uint32_t n_rec = min(d.ctr->pair_cursor, d.cap_pairs);
    const int lane = threadIdx.x & 31;
    for (uint32_t t0 = blockIdx.x * blockDim.x; t0 < n_rec; t0 += gridDim.x * blockDim.x) {   // (warp-uniform trip count)
        const uint32_t t = t0 + threadIdx.x;
        uint32_t A = 0, cnt = 0;
        if (t < n_rec) {
            const uint4 rec = d.ptmp[t];
            if (rec.z == 0) { A = rec.x; cnt = d.grp_count[A]; }
        }
        uint32_t inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        uint32_t base = 0;
        if (lane == 31 && inc) base = atomicAdd(&d.ctr->grp_cursor, inc);
        base = __shfl_sync(0xffffffffu, base, 31);
        if (cnt) {
            d.tester_start[A] = base + inc - cnt;
            d.tester_count[A] = cnt;
            d.grp_count[A] = 0u;
        }
    }
}

// Rank of every record = number of candidates the reference's walk offers before B (see the header of this file).
// Four lanes per record: lane r counts the query-rect rows qy0 + r, qy0 + r + 4, ... up to B's first shared row; a row's
// entries are taken four at a time (entry words, then their tags, all in flight together).
template <bool FILTERS>
__global__ void __launch_bounds__(256) k_rank(DevView d, int margin) {
    if (d.ctr->overflow & RSV_OVF_PAIRS) return;
    const uint32_t n_rec = min(d.ctr->pair_cursor, d.cap_pairs);
    const int sub = threadIdx.x & 3;
    const uint32_t n_groups = gridDim.x * (blockDim.x >> 2);
    for (uint32_t t0 = blockIdx.x * (blockDim.x >> 2); t0 < n_rec; t0 += n_groups) {   // (warp-uniform trip count)
        const uint32_t t = t0 + (threadIdx.x >> 2);
        uint32_t rank = 0, A = 0, B = 0, idx = 0;
        if (t < n_rec) {
            const uint4 rec = d.ptmp[t];
            A = rec.x; B = rec.y;
            idx = rec.z + d.tester_start[A];   // (requested now, needed at the very end: was 9 % of this kernel's stall samples there)
            const int4 ra = d.crect[A], rb = d.crect[B];
            bool filtered = false, useAny = false;
            ulonglong2 qm = make_ulonglong2(0ull, 0ull);
            if (FILTERS) {
                useAny = (__ldg(d.meta + A) & RSV_META_USE_ANY) != 0;
                qm = __ldg(d.qmask + A);
                filtered = useAny || qm.y != 0;
            }
            const int qx0 = max(ra.x - margin, 0), qx1 = min(ra.z + margin, d.W - 1);
            const int qy0 = max(ra.y - margin, d.row_lo), qy1 = min(ra.w + margin, d.row_hi - 1);
            // first cell of the walk that holds B: where B's (clamped) rect starts inside the query rect
            const int fx = min(max(qx0, max(rb.x, 0)), qx1), fy = min(max(qy0, max(rb.y, d.row_lo)), qy1);
            const uint32_t* cellp = d.cell + cell_base(d, A);
            for (int y = qy0 + sub; y <= fy; y += 4) {
                const uint32_t* row = cellp + (size_t)(y - d.row_lo) * (size_t)d.W;
                const int xe = y < fy ? qx1 + 1 : fx;                // whole cells [qx0, xe), then (last row) B's cell
                const uint32_t s0 = min(row[qx0], d.cap_entries), s1 = min(row[qx0 + 1], d.cap_entries);
                const uint32_t e = xe > qx0 ? min(row[xe], d.cap_entries) : s0;
                const uint32_t pe = y == fy ? min(row[fx + 1], d.cap_entries) : e;
                const bool first_row = y == qy0;
                for (uint32_t k0 = s0; k0 < pe; k0 += 4) {
                    uint32_t E[4];
                    bool c[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) E[u] = k0 + u < pe ? d.entries[k0 + u] : 0u;
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const uint32_t k = k0 + u, S = E[u] >> kEntShift;
                        // whole cells: every entry; B's cell: the entries in front of B (in-cell order = ascending slot)
                        const bool first_col_cell = k < e ? k < s1 : fx == qx0;
                        c[u] = k < pe && (k < e || S < B) && S != A &&                                   // cell.go:101
                               ((E[u] & kEntFirstCol) || first_col_cell) && ((E[u] & kEntFirstRow) || first_row);   // cell.go:105
                    }
                    if (FILTERS && filtered) {
                        unsigned long long tg[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) tg[u] = c[u] ? __ldg(d.tags + (E[u] >> kEntShift)) : 0ull;
#pragma unroll
                        for (int u = 0; u < 4; u++) c[u] = c[u] && cp_tags_ok(useAny, qm, tg[u]);
                    }
                    rank += (uint32_t)c[0] + (uint32_t)c[1] + (uint32_t)c[2] + (uint32_t)c[3];
                }
            }
        }
        rank += __shfl_xor_sync(0xffffffffu, rank, 1);
        rank += __shfl_xor_sync(0xffffffffu, rank, 2);
        if (t < n_rec && sub == 0) {
            if (idx < d.cap_pairs) d.pgrp[idx] = make_uint4(A, B, rank, 0u);
        }
    }
}

// (Generation order within a tester — a record's place is the number of its tester's records with a smaller rank — is
// established by k_place_bin, rsv_narrow.cuh, together with the narrowphase binning.)

}  // namespace rsv
