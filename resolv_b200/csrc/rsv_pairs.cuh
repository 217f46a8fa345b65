// rsv_pairs.cuh — candidate-pair emitter.
//
// Replaces, for every tester A at once,
//   A.SelectTouchingCells(margin)            shape.go:220-237   (cell rect grown by margin, not clamped)
//   CellSelection.ForEach                    cell.go:86-121     (row-major walk, skip self, dedupe by ID)
//   ShapeFilter.ForEach / ByTags / NotByTags shapefilter.go:26-61, tags.go:29-31
//   the gather loop of IntersectionTest      shape.go:277-289   (skip self)
//   Bounds.IsIntersecting (first gate of every pair test) utils.go:146-173, shape.go:322,360,403,442
//
// Dedupe uses the owner-cell rule instead of the reference's seen-ID list: a shape B registered in several
// cells of A's query rect is emitted only from the first cell (row-major) the two rects share, which is exactly
// where the reference's linear-scan dedupe lets it through. Because both rects are clamped to the same grid, a
// cell (x,y) of the walk is that first shared cell iff (x is B's first column or x is the walk's first column)
// and (y is B's first row or y is the walk's first row) — two flag bits stored in the entry, no gather of B's rect.
// Emission order per tester therefore equals the reference's candidate generation order (for a Space whose
// cells are in append order), and `rank` is the candidate's index in possibleIntersections before the sort.
#pragma once
#include "rsv_device.cuh"
#include "rsv_grid.cuh"

namespace rsv {

constexpr int kPairsBlock = 128;
constexpr int kStash = 8;  // pair records a thread can hold in shared memory before it needs a second walk

// Walks tester A's query rect and calls emit(b, rank, gated) for every candidate in generation order.
// Returns the number of candidates (P_A).
template <class Emit>
__device__ __forceinline__ uint32_t walk_candidates(const DevView& d, uint32_t A, uint32_t metaA, int margin,
                                                    Emit emit) {
    int4 r = d.crect[A];
    int qx0 = max(r.x - margin, 0), qx1 = min(r.z + margin, d.W - 1);
    int qy0 = max(r.y - margin, d.row_lo), qy1 = min(r.w + margin, d.row_hi - 1);
    if (qx0 > qx1 || qy0 > qy1) return 0;
    const Box a = load_box_nc(d.aabb + A);
    const bool useAny = (metaA & RSV_META_USE_ANY) != 0;
    ulonglong2 qm = __ldg(d.qmask + A);
    const bool filtered = useAny || qm.y != 0;
    const uint32_t base = cell_base(d, A);
    uint32_t rank = 0;
    for (int y = qy0; y <= qy1; y++) {
        const uint32_t* row = d.cell + base + (uint32_t)(y - d.row_lo) * (uint32_t)d.W;
        uint32_t e = row[qx0];
        for (int x = qx0; x <= qx1; x++) {
            uint32_t s = e;
            e = row[x + 1];
            uint32_t lim = min(e, d.cap_entries);
            for (uint32_t k = s; k < lim; k++) {
                uint32_t ent = d.entries[k];
                uint32_t B = ent >> 3;
                if (B == A) continue;                                         // cell.go:101, shape.go:279
                if (!((ent & kEntFirstCol) || x == qx0)) continue;            // owner-cell rule == cell.go:105
                if (!((ent & kEntFirstRow) || y == qy0)) continue;
                if (filtered) {
                    unsigned long long t = __ldg(d.tags + B);
                    if (useAny && (t & qm.x) == 0) continue;                  // ByTags: Has == any bit, tags.go:29-31
                    if ((t & qm.y) != 0) continue;                            // NotByTags
                }
                Box o = load_box_nc(d.aabb + B);
                bool gated = bounds_gate(a.minx, a.miny, a.maxx, a.maxy, o.minx, o.miny, o.maxx, o.maxy);
                emit(B, rank, gated);
                rank++;
            }
        }
    }
    return rank;
}

// One thread per tester. Each thread keeps its first kStash records in shared memory, the block allocates one
// contiguous chunk of the record buffer (one atomic per block) and writes it tester by tester, so all records of
// a tester are contiguous and in generation order. Threads with more than kStash records walk a second time.
__global__ void __launch_bounds__(kPairsBlock) k_pairs(DevView d, int margin, uint32_t flags) {
    __shared__ uint2 s_stash[kStash][kPairsBlock];
    __shared__ uint32_t s_warp[kPairsBlock / 32];
    __shared__ uint32_t s_chunk;
    const bool all = (flags & RSV_FRAME_ALL_CANDIDATES) != 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long my_cands = 0;
    const uint32_t rounds = (d.n_shapes + gridDim.x * kPairsBlock - 1) / (gridDim.x * kPairsBlock);
    for (uint32_t it = 0; it < rounds; it++) {
        const uint32_t A = (it * gridDim.x + blockIdx.x) * kPairsBlock + threadIdx.x;
        uint32_t metaA = 0;
        bool active = false;
        if (A < d.n_shapes) {
            metaA = __ldg(d.meta + A);
            active = (metaA & RSV_META_TESTER) && !(metaA & RSV_META_ABSENT);
        }
        uint32_t cnt = 0;
        if (active) {
            uint32_t P = walk_candidates(d, A, metaA, margin, [&](uint32_t B, uint32_t rank, bool gated) {
                if (gated || all) {
                    if (cnt < kStash) s_stash[cnt][threadIdx.x] = make_uint2(B, (rank << 8) | (gated ? 0u : RSV_HIT_NOT_GATED));
                    cnt++;
                }
            });
            my_cands += P;
        }
        // block-wide exclusive scan of cnt
        uint32_t inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        uint32_t off = inc - cnt, total = 0;
#pragma unroll
        for (int w = 0; w < kPairsBlock / 32; w++) {
            uint32_t t = s_warp[w];
            if (w < warp) off += t;
            total += t;
        }
        if (threadIdx.x == 0) {
            uint32_t chunk = 0;
            if (total) {
                chunk = atomicAdd(&d.ctr->pair_cursor, total);
                atomicAdd(&d.ctr->n_pairs_needed, (unsigned long long)total);
                if (chunk + total > d.cap_pairs || chunk + total < chunk) atomicOr(&d.ctr->overflow, RSV_OVF_PAIRS);
            }
            s_chunk = chunk;
        }
        __syncthreads();
        const uint32_t start = s_chunk + off;
        if (A < d.n_shapes) {
            d.tester_start[A] = start;
            d.tester_count[A] = cnt;
        }
        if (cnt && (unsigned long long)start + cnt <= d.cap_pairs) {
            uint32_t n0 = min(cnt, (uint32_t)kStash);
            for (uint32_t k = 0; k < n0; k++) {
                uint2 sb = s_stash[k][threadIdx.x];
                *reinterpret_cast<uint4*>(d.hits + start + k) = make_uint4(A, sb.x, 0u, sb.y);
            }
            if (cnt > kStash) {
                uint32_t k = 0;
                walk_candidates(d, A, metaA, margin, [&](uint32_t B, uint32_t rank, bool gated) {
                    if (gated || all) {
                        if (k >= kStash)
                            *reinterpret_cast<uint4*>(d.hits + start + k) =
                                make_uint4(A, B, 0u, (rank << 8) | (gated ? 0u : RSV_HIT_NOT_GATED));
                        k++;
                    }
                });
            }
        }
        __syncthreads();  // s_stash / s_warp are reused by the next round
    }
    for (int o = 16; o > 0; o >>= 1) my_cands += __shfl_down_sync(0xffffffffu, my_cands, o);
    if (lane == 0 && my_cands) atomicAdd(&d.ctr->n_candidates, my_cands);
}

}  // namespace rsv
