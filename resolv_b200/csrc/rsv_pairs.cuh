// rsv_pairs.cuh — warp-cooperative candidate-pair emitter.
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
//
// Work distribution
//  * Testers are taken in CELL ORDER straight from the sorted entry array: the entry carrying both first-row and
//    first-column flags is the shape's "home" registration, so ballot-compacting those entries yields testers
//    sorted by home cell for free; a warp's 32 testers walk overlapping cell rows. Warps grab chunks of the entry
//    array from an atomic cursor. Testers that touch no stored cell but whose margin reaches the grid come from a
//    small orphan list built by k_bounds.
//  * A cell row of a query rect is ONE contiguous run of the entry array (cells are consecutive in a row), so a
//    tester's walk is rows x (3 offsets + a run). Rows are taken four at a time: the warp prefix-sums the run
//    lengths, flattens the (tester, entry) items and moves them 32 per step with cp.async into shared memory — all
//    loads of a round are in flight together (the kernel is latency-bound otherwise) — then tests them from shared
//    memory, every lane one item per step no matter how entries are distributed over testers.
//  * Each entry carries an outward-rounded float32 copy of its shape's AABB. Disjoint float boxes imply disjoint
//    float64 boxes, so ~90 % of candidates are rejected without touching the other shape's data; survivors become
//    records, and the exact float64 Bounds gate (utils.go:146-173) is applied per record by k_bin, where the two
//    AABB gathers are fully parallel instead of sitting in this kernel's serial chain.
//  * Records are written with one allocation per batch, 16 B per lane, grouped by tester in generation order. A
//    batch that overflows the stage falls back to an inline count pass + direct-write pass.
#pragma once
#include <cuda_pipeline.h>

#include "rsv_device.cuh"
#include "rsv_grid.cuh"

namespace rsv {

#ifndef RSV_PAIRS_BLOCK
#define RSV_PAIRS_BLOCK 128
#endif
#ifndef RSV_ITEM_ROUND
#define RSV_ITEM_ROUND 160
#endif
constexpr int kPairsBlock = RSV_PAIRS_BLOCK;
constexpr int kPairsWarps = kPairsBlock / 32;
#ifndef RSV_PAIRS_CHUNK
#define RSV_PAIRS_CHUNK 256
#endif
#ifndef RSV_PAIRS_ONE_CHUNK_FACTOR
#define RSV_PAIRS_ONE_CHUNK_FACTOR 2u
#endif
constexpr int kPairsChunk = RSV_PAIRS_CHUNK;  // entries per grabbed chunk
constexpr int kPairStage = 96;    // staged records per warp batch
constexpr int kRowGroup = 4;      // query-rect rows flattened together
constexpr int kItemRound = RSV_ITEM_ROUND;   // items staged through shared memory per round
constexpr int kCandWords = 60;    // super-round = kCandWords * 32 items = 12 rounds
constexpr uint32_t kSuper = kCandWords * 32;
static_assert(kSuper % kItemRound == 0, "super-rounds are whole rounds");
static_assert(kItemRound % 32 == 0, "a round is whole warp steps (one candidate-bit word per step)");
constexpr uint32_t kNoEntry = 0xffffffffu;

// Number of set bits of the candidate words in the item range [lo, hi) (indices relative to the super-round).
__device__ __forceinline__ uint32_t popc_range(const uint32_t* cw, uint32_t lo, uint32_t hi) {
    if (hi <= lo) return 0;
    const uint32_t wl = lo >> 5, wh = (hi - 1) >> 5;
    uint32_t n = 0;
    for (uint32_t w = wl; w <= wh; w++) {
        uint32_t m = cw[w];
        if (w == wl) m &= 0xffffffffu << (lo & 31u);
        if (w == wh) m &= 0xffffffffu >> (31u - ((hi - 1) & 31u));
        n += __popc(m);
    }
    return n;
}

template <bool FILTERS>
struct PairsWarpSmem {
    float4 itBox[kItemRound];   // cp.async landing zone: entry float boxes
    float4 fbox[32];            // outward-rounded float AABB of each tester of the batch
    uint4 cum[32];              // per tester: cumulative run lengths of the current row group
    uint2 run[kRowGroup][32];   // per (row, tester): (start of run, end of the run's first cell)
    unsigned long long itTags[FILTERS ? kItemRound : 1];  // cp.async landing zone: entry tags (frames with tag filters only)
    uint32_t itEnt[kItemRound]; // cp.async landing zone: entry words
    uint16_t itMeta[kItemRound];  // tester lane | first-column-cell << 5 | first-row << 6
    uint32_t queue[64];         // home entries waiting for a full batch
    uint32_t A[32];             // tester slot
    uint32_t flags[32];         // bit0 useAny, bit1 filtered
    uint32_t cand[32];          // candidates so far (== next rank)
    uint32_t gated[32];         // records so far
    uint32_t excl[32];          // exclusive prefix of gated over the batch
    uint32_t stB[kPairStage];   // staged records: other slot
    uint32_t stR[kPairStage];   // rank << 8 | flags
    uint32_t stT[kPairStage];   // tester lane << 24 | index within the tester's records
    uint32_t cw[kCandWords];    // one candidate-bit word per 32 items of the current super-round (lazy ranks)
    uint2 rng[32];              // per tester: its item range [first, last) in the current row group
};

enum PairsMode { kStage = 0, kCount = 1, kWrite = 2 };

__device__ __forceinline__ bool exact_gate(const DevView& d, uint32_t A, uint32_t B) {
    const Box a = load_box_nc(d.aabb + A), o = load_box_nc(d.aabb + B);
    return bounds_gate(a.minx, a.miny, a.maxx, a.maxy, o.minx, o.miny, o.maxx, o.maxy);
}

// One pass over all (tester, entry) items of the batch.
//  kStage: count candidates; stage every candidate that survives the float gate (or every candidate in `all` mode).
//  kCount: count candidates and records with the exact gate inline.   kWrite: same walk, records written in place.
// Returns the number of staged (kStage) or produced (kCount/kWrite) records; uniform across the warp.
template <int MODE, bool FILTERS>
__device__ __forceinline__ uint32_t pairs_pass(const DevView& d, bool all, unsigned long long myAny,
                                               unsigned long long myNone, PairsWarpSmem<FILTERS>& sm, bool have, int rows,
                                               const uint32_t* row0, int ncols, uint32_t chunk) {
    constexpr bool filters = FILTERS;
    const int lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t produced = 0;
    for (int g = 0; __any_sync(0xffffffffu, have && g < rows); g += kRowGroup) {
        uint32_t s0[kRowGroup], s1[kRowGroup], e[kRowGroup];
#pragma unroll
        for (int r = 0; r < kRowGroup; r++) {
            s0[r] = s1[r] = e[r] = 0;
            if (have && g + r < rows) {
                const uint32_t* row = row0 + (size_t)(g + r) * (size_t)d.W;
                s0[r] = row[0];
                s1[r] = row[1];
                e[r] = ncols > 1 ? row[ncols] : s1[r];
            }
        }
        uint32_t c = 0;
        uint32_t cum[kRowGroup];
#pragma unroll
        for (int r = 0; r < kRowGroup; r++) {
            s0[r] = min(s0[r], d.cap_entries); s1[r] = min(s1[r], d.cap_entries); e[r] = min(e[r], d.cap_entries);
            c += e[r] - s0[r];
            cum[r] = c;
            sm.run[r][lane] = make_uint2(s0[r], s1[r]);
        }
        sm.cum[lane] = make_uint4(cum[0], cum[1], cum[2], cum[3]);
        uint32_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        sm.rng[lane] = make_uint2(incl - c, incl);
        __syncwarp();
        for (uint32_t sbeg = 0; sbeg < total; sbeg += kSuper) {
        const uint32_t send = min(sbeg + kSuper, total);
        const uint32_t stage0 = produced;  // records staged before this super-round already carry their rank
        for (uint32_t rbeg = sbeg; rbeg < send; rbeg += kItemRound) {
            const uint32_t rend = min(rbeg + (uint32_t)kItemRound, send);
            // ---- issue: locate every item of the round and start its copy into shared memory
            for (uint32_t it = rbeg; it < rend; it += 32) {
                const uint32_t i = min(it + lane, total - 1);  // (clamped lanes search too: shuffles need the full warp)
                // owner tester t: smallest lane with incl[t] > i (5-step search over the warp's prefix sums)
                int t = 0;
#pragma unroll
                for (int step = 16; step > 0; step >>= 1) {
                    uint32_t v = __shfl_sync(0xffffffffu, incl, t + step - 1);
                    if (v <= i) t += step;
                }
                t &= 31;
                const uint32_t t_incl = __shfl_sync(0xffffffffu, incl, t);
                if (it + lane < rend) {
                    const uint4 tc = sm.cum[t];
                    const uint32_t off = i - (t_incl - tc.w);
                    const int r = (off >= tc.x) + (off >= tc.y) + (off >= tc.z);
                    const uint32_t rb = r == 0 ? 0u : (r == 1 ? tc.x : (r == 2 ? tc.y : tc.z));
                    const uint2 rn = sm.run[r][t];
                    const uint32_t k = rn.x + (off - rb);
                    const uint32_t slot = i - rbeg;
                    __pipeline_memcpy_async(&sm.itEnt[slot], d.entries + k, 4);
                    __pipeline_memcpy_async(&sm.itBox[slot], d.ent_box + k, 16);
                    if (filters) __pipeline_memcpy_async(&sm.itTags[slot], d.ent_tags + k, 8);
                    sm.itMeta[slot] = (uint16_t)(t | (k < rn.y ? 32 : 0) | ((g + r) == 0 ? 64 : 0));
                }
            }
            __pipeline_commit();
            __pipeline_wait_prior(0);
            __syncwarp();
            // ---- test: one item per lane per step, from shared memory
            for (uint32_t it = rbeg; it < rend; it += 32) {
                const uint32_t i = it + lane;
                const bool valid = i < rend;
                bool cand = false, maybe = false, gate = false;
                uint32_t B = 0, t = 0, fl = 0;
                unsigned long long tg = 0;
                float4 fb = make_float4(0.f, 0.f, 0.f, 0.f), fa = fb;
                if (valid) {
                    const uint32_t slot = i - rbeg;
                    const uint32_t ent = sm.itEnt[slot];
                    const uint32_t m = sm.itMeta[slot];
                    t = m & 31u;
                    B = ent >> kEntShift;
                    cand = B != sm.A[t]                                     // cell.go:101, shape.go:279
                           && ((ent & kEntFirstCol) || (m & 32u))           // owner-cell rule == cell.go:105
                           && ((ent & kEntFirstRow) || (m & 64u));
                    fl = sm.flags[t];
                    if (filters) tg = sm.itTags[slot];
                    fb = sm.itBox[slot];
                    fa = sm.fbox[t];
                }
                if (filters) {
                    // the tester's masks live in its lane's registers (ByTags: Has == any bit, tags.go:29-31; NotByTags)
                    const unsigned long long tAny = __shfl_sync(0xffffffffu, myAny, t), tNone = __shfl_sync(0xffffffffu, myNone, t);
                    if (cand && (fl & 2u)) {
                        if ((fl & 1u) && (tg & tAny) == 0) cand = false;
                        if ((tg & tNone) != 0) cand = false;
                    }
                }
                // float boxes are rounded outward, so "disjoint in float" implies the exact reject of utils.go:156
                maybe = cand && !(fb.z < fa.x || fb.x > fa.z || fb.w < fa.y || fb.y > fa.w);
                gate = maybe;  // the exact float64 gate runs per record in k_bin
                if (MODE == kStage) {
                    // slim path: no per-item rank. One candidate-bit word per step; ranks are derived from the words at
                    // the end of the super-round, and only for the few staged survivors.
                    const uint32_t cmask = __ballot_sync(0xffffffffu, cand);
                    const bool rec = maybe || (all && cand);
                    const uint32_t rmask = __ballot_sync(0xffffffffu, rec);
                    if (lane == 0) sm.cw[(it - sbeg) >> 5] = cmask;
                    if (rec) {
                        const uint32_t slot = produced + __popc(rmask & lt);
                        if (slot < kPairStage) {
                            sm.stB[slot] = B;
                            sm.stR[slot] = ((i - sbeg) << 8) | (maybe ? 0u : RSV_HIT_NOT_GATED);  // item index for now
                            sm.stT[slot] = t << 24;
                        }
                    }
                    produced += __popc(rmask);
                    continue;
                }
                const uint32_t peers = __match_any_sync(0xffffffffu, valid ? t : 32 + lane);
                const uint32_t cmask = __ballot_sync(0xffffffffu, cand);
                const uint32_t rank = sm.cand[t] + __popc(cmask & peers & lt);
                const bool rec = gate || (all && cand);
                const uint32_t rmask = __ballot_sync(0xffffffffu, rec);
                const uint32_t gidx = sm.gated[t] + __popc(rmask & peers & lt);
                __syncwarp();
                if (valid && (peers & lt) == 0) {  // first lane of each tester's run updates its counters
                    sm.cand[t] += __popc(cmask & peers);
                    sm.gated[t] += __popc(rmask & peers);
                }
                if (rec && MODE == kWrite) {
                    const unsigned long long pos = (unsigned long long)chunk + sm.excl[t] + gidx;
                    if (pos < d.cap_pairs)
                        *reinterpret_cast<uint4*>(d.hits + pos) =
                            make_uint4(sm.A[t], B, 0u, (rank << 8) | (gate ? 0u : RSV_HIT_NOT_GATED));
                }
                produced += __popc(rmask);
                __syncwarp();
            }
        }
        if (MODE == kStage) {
            // ---- end of the super-round: ranks of the survivors staged in it, then fold its candidates into the bases
            __syncwarp();
            const uint32_t stage1 = min(produced, (uint32_t)kPairStage);
            for (uint32_t r = stage0 + lane; r < stage1; r += 32) {
                const uint32_t t = sm.stT[r] >> 24, w = sm.stR[r];
                const uint2 rg = sm.rng[t];
                const uint32_t lo = max(rg.x, sbeg) - sbeg;
                sm.stR[r] = ((sm.cand[t] + popc_range(sm.cw, lo, w >> 8)) << 8) | (w & 0xffu);
            }
            __syncwarp();
            const uint2 rg = sm.rng[lane];
            const uint32_t lo = max(rg.x, sbeg), hi = min(rg.y, send);
            if (hi > lo) sm.cand[lane] += popc_range(sm.cw, lo - sbeg, hi - sbeg);
            __syncwarp();
        }
        }
    }
    return produced;
}

// Assigns every staged record its index within its tester's records (stage order is generation order within a
// tester), one record per lane.
template <bool FILTERS>
__device__ __forceinline__ uint32_t pairs_assign(PairsWarpSmem<FILTERS>& sm, uint32_t staged) {
    const int lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    for (uint32_t base = 0; base < staged; base += 32) {
        const uint32_t r = base + lane;
        const bool valid = r < staged;
        const uint32_t t = valid ? (sm.stT[r] >> 24) : 0u;
        const uint32_t peers = __match_any_sync(0xffffffffu, valid ? t : 32 + lane);
        const uint32_t gidx = sm.gated[t] + __popc(peers & lt);
        __syncwarp();
        if (valid) {
            if ((peers & lt) == 0) sm.gated[t] += __popc(peers);
            sm.stT[r] = (t << 24) | (gidx & 0xffffffu);
        }
        __syncwarp();
    }
    return staged;
}

// One batch: lane `lane` owns a tester, given by its home entry index `hk` (or, for orphans, directly by slot A).
template <bool FILTERS>
__device__ __forceinline__ void pairs_batch(const DevView& d, int margin, uint32_t flags, bool have, uint32_t hk,
                                            uint32_t A, PairsWarpSmem<FILTERS>& sm, unsigned long long& my_cands) {
    const int lane = threadIdx.x & 31;
    const bool all = (flags & RSV_FRAME_ALL_CANDIDATES) != 0;
    int rows = 0, ncols = 0;
    bool useAnyBit = false;
    unsigned long long myAny = 0, myNone = 0;
    const uint32_t* row0 = d.cell;
    if (have) {
        constexpr bool filters0 = FILTERS;
        int qx0, qx1, qy0, qy1;
        uint32_t base_cell;
        if (hk != kNoEntry) {
            // everything about the tester comes from its home entry: three coalesced loads, no gather by slot
            const uint32_t ent = d.entries[hk];
            A = ent >> kEntShift;
            useAnyBit = (ent & kEntUseAny) != 0;
            sm.fbox[lane] = d.ent_box[hk];
            const uint2 hm = d.ent_home[hk];
            const uint32_t local = hm.x % d.cells_per_space;
            base_cell = hm.x - local;
            const int x0 = (int)(local % (uint32_t)d.W), y0 = d.row_lo + (int)(local / (uint32_t)d.W);
            // max(cx - m, 0) == max(x0 - m, 0) and min(ex + m, W-1) == min(x1 + m, W-1) for the clamped x0, x1
            qx0 = max(x0 - margin, 0); qx1 = min(x0 + (int)(hm.y & 0xffffu) + margin, d.W - 1);
            qy0 = max(y0 - margin, d.row_lo); qy1 = min(y0 + (int)(hm.y >> 16) + margin, d.row_hi - 1);
        } else {
            sm.fbox[lane] = float_box(load_box_nc(d.aabb + A));
            useAnyBit = filters0 && (__ldg(d.meta + A) & RSV_META_USE_ANY) != 0;
            const int4 r = d.crect[A];
            qx0 = max(r.x - margin, 0); qx1 = min(r.z + margin, d.W - 1);
            qy0 = max(r.y - margin, d.row_lo); qy1 = min(r.w + margin, d.row_hi - 1);
            base_cell = cell_base(d, A);
        }
        if (qx0 <= qx1 && qy0 <= qy1) {
            rows = qy1 - qy0 + 1;
            ncols = qx1 - qx0 + 1;
            row0 = d.cell + base_cell + (size_t)(qy0 - d.row_lo) * (size_t)d.W + qx0;
        }
        constexpr bool filters = FILTERS;
        const bool useAny = filters && useAnyBit;
        ulonglong2 qm = make_ulonglong2(0ull, 0ull);
        if (filters) qm = __ldg(d.qmask + A);
        myAny = qm.x; myNone = qm.y;
        sm.flags[lane] = (useAny ? 1u : 0u) | ((useAny || qm.y != 0) ? 2u : 0u);
    }
    sm.A[lane] = A;
    sm.cand[lane] = 0;
    sm.gated[lane] = 0;
    __syncwarp();
    const uint32_t staged = pairs_pass<kStage, FILTERS>(d, all, myAny, myNone, sm, have, rows, row0, ncols, 0u);
    const bool fits = staged <= kPairStage;
    uint32_t total;
    if (fits) {
        total = pairs_assign(sm, staged);
    } else {
        sm.cand[lane] = 0;
        __syncwarp();
        total = pairs_pass<kCount, FILTERS>(d, all, myAny, myNone, sm, have, rows, row0, ncols, 0u);
    }
    const uint32_t cnt = sm.gated[lane];
    my_cands += sm.cand[lane];
    uint32_t inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    uint32_t chunk = 0;
    if (lane == 0 && total) {
        chunk = atomicAdd(&d.ctr->pair_cursor, total);
        if (chunk + total > d.cap_pairs || chunk + total < chunk) atomicOr(&d.ctr->overflow, RSV_OVF_PAIRS);
    }
    chunk = __shfl_sync(0xffffffffu, chunk, 0);
    sm.excl[lane] = inc - cnt;
    if (have && (flags & RSV_FRAME_TESTER_TABLE)) {
        d.tester_start[A] = chunk + (inc - cnt);
        d.tester_count[A] = cnt;
    }
    __syncwarp();
    if (total == 0) return;
    if (fits) {
        for (uint32_t r = lane; r < staged; r += 32) {
            const uint32_t tt = sm.stT[r];
            if (tt == 0xffffffffu) continue;
            const uint32_t t = tt >> 24;
            const unsigned long long pos = (unsigned long long)chunk + sm.excl[t] + (tt & 0xffffffu);
            if (pos < d.cap_pairs) *reinterpret_cast<uint4*>(d.hits + pos) = make_uint4(sm.A[t], sm.stB[r], 0u, sm.stR[r]);
        }
    } else {
        sm.cand[lane] = 0;
        sm.gated[lane] = 0;
        __syncwarp();
        pairs_pass<kWrite, FILTERS>(d, all, myAny, myNone, sm, have, rows, row0, ncols, chunk);
    }
    __syncwarp();
}

// FILTERS = false: frames launched with RSV_FRAME_NO_TAG_FILTERS (no tags staged: less shared memory per warp, so
// more warps per SM for this latency-bound kernel).
template <bool FILTERS>
__global__ void __launch_bounds__(kPairsBlock) k_pairs(DevView d, int margin, uint32_t flags) {
    __shared__ PairsWarpSmem<FILTERS> s_warp[kPairsWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    PairsWarpSmem<FILTERS>& sm = s_warp[warp];
    uint32_t* q = sm.queue;
    unsigned long long my_cands = 0;
    const uint32_t R = (uint32_t)min((unsigned long long)d.cap_entries, d.ctr->n_entries);
    // chunk size: kPairsChunk entries for large frames (dynamic balancing over many chunks); frames with at most two
    // such chunks per warp are cut into exactly one chunk per warp, so that every warp gets work and none gets a
    // second helping (config 4 per GPU: 298 entries per warp; 256-entry chunks left 17 % of the warps with two: 80 -> 71 us)
    const uint32_t n_warps = gridDim.x * kPairsWarps;
    const uint32_t per_warp = max(32u, ((R + n_warps - 1) / n_warps + 31u) & ~31u);
    const uint32_t chunk_sz = per_warp <= RSV_PAIRS_ONE_CHUNK_FACTOR * (uint32_t)kPairsChunk ? per_warp : (uint32_t)kPairsChunk;
    const uint32_t n_chunks = (R + chunk_sz - 1) / chunk_sz;
    uint32_t qn = 0;
    for (;;) {
        uint32_t c = 0;
        if (lane == 0) c = atomicAdd(&d.ctr->chunk_cursor, 1u);
        c = __shfl_sync(0xffffffffu, c, 0);
        if (c >= n_chunks) break;
        const uint32_t lo = c * chunk_sz, hi = min(lo + chunk_sz, R);
        for (uint32_t base = lo; base < hi; base += 32) {
            uint32_t k = base + lane;
            uint32_t ent = k < hi ? d.entries[k] : 0u;
            const uint32_t want = kEntTester | kEntFirstCol | kEntFirstRow;   // the tester's home registration
            bool home = (ent & want) == want;
            uint32_t m = __ballot_sync(0xffffffffu, home);
            if (home) q[qn + __popc(m & ((1u << lane) - 1u))] = k;
            qn += __popc(m);
            __syncwarp();
            if (qn >= 32) {
                uint32_t hk = q[lane];
                __syncwarp();
                pairs_batch<FILTERS>(d, margin, flags, true, hk, 0u, sm, my_cands);
                if (lane + 32 < qn) { uint32_t t = q[lane + 32]; q[lane] = t; }
                qn -= 32;
                __syncwarp();
            }
        }
    }
    if (qn) {
        uint32_t hk = lane < qn ? q[lane] : kNoEntry;
        pairs_batch<FILTERS>(d, margin, flags, lane < qn, hk, 0u, sm, my_cands);
    }
    // testers with no stored cell whose margin still reaches the grid
    const uint32_t n_orph = min(d.ctr->n_orphans, d.n_shapes);
    for (;;) {
        uint32_t o = 0;
        if (lane == 0) o = atomicAdd(&d.ctr->orphan_cursor, 32u);
        o = __shfl_sync(0xffffffffu, o, 0);
        if (o >= n_orph) break;
        bool have = o + lane < n_orph;
        uint32_t A = have ? d.orphans[o + lane] : 0u;
        pairs_batch<FILTERS>(d, margin, flags, have, kNoEntry, A, sm, my_cands);
    }
    for (int o = 16; o > 0; o >>= 1) my_cands += __shfl_down_sync(0xffffffffu, my_cands, o);
    if (lane == 0 && my_cands) atomicAdd(&d.ctr->n_candidates, my_cands);
}

}  // namespace rsv
