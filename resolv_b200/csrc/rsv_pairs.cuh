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

// ------------------------------------------------------------------------------------------------------------
// RSV_FRAME_OWN_RECT (unfiltered frames): the tester visits only the entries of its OWN clamped cell rect — every shape
// whose Bounds can pass the gate of utils.go:146-173 is registered there — instead of the margin-grown rect of
// SelectTouchingCells (config 2: 2.7 instead of 11 entries per tester). The two things the grown walk of
// cell.go:86-121 also yields, the number of candidates and the generation rank of a survivor, come in closed form from
// three prefix counts over the cell-ordered entry array (ent_pref[k] = entries j < k with the first-column flag /
// the first-row flag / both): an entry at cell (x, y) of the walk Q is B's first occurrence iff
// (firstCol or x == Q.x0) and (firstRow or y == Q.y0), so a row y of Q — one contiguous run [s0, e) of the entry array
// whose first cell ends at s1 — holds
//     y == Q.y0:  (s1 - s0) + FC(s1, e)  candidates,      else:  FR(s0, s1) + H(s1, e);
// B's first occurrence is its entry in cell (max(B.x0, Q.x0), max(B.y0, Q.y0)), and
//     rank(B) = rows above + the same expression cut at that entry - [the tester's own home entry comes first].
// scripts/proto_pairs_v2.py checks these identities against the oracle's candidate lists.

constexpr int kPrefBlock = 256;
constexpr int kPrefItems = 8;
constexpr int kPrefTile = kPrefBlock * kPrefItems;

__device__ __forceinline__ unsigned long long pref_pack(uint32_t ent) {  // fc | fr << 21 | both << 42
    const unsigned long long fc = (ent >> 1) & 1u, fr = ent & 1u;
    return fc | (fr << 21) | ((fc & fr) << 42);
}

// Pass 1: per tile of kPrefTile entries, the three flag counts.
__global__ void __launch_bounds__(kPrefBlock) k_pref_tiles(DevView d, uint4* __restrict__ tiles) {
    __shared__ unsigned long long s_w[kPrefBlock / 32];
    const uint32_t R = (uint32_t)min((unsigned long long)d.cap_entries, d.ctr->n_entries);
    const uint32_t k0 = blockIdx.x * kPrefTile + threadIdx.x * kPrefItems;
    unsigned long long acc = 0;
    if (blockIdx.x * (uint32_t)kPrefTile <= R) {
#pragma unroll
        for (int u = 0; u < kPrefItems; u++)
            if (k0 + u < R) acc += pref_pack(d.entries[k0 + u]);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < kPrefBlock / 32; w++) t += s_w[w];
        tiles[blockIdx.x] = make_uint4((uint32_t)(t & 0x1fffffu), (uint32_t)((t >> 21) & 0x1fffffu), (uint32_t)(t >> 42), 0u);
    }
}

// Pass 2 (one block): exclusive scan of the tile counts, in place.
__global__ void __launch_bounds__(1024) k_pref_scan_tiles(uint4* __restrict__ tiles, uint32_t n_tiles) {
    __shared__ uint32_t s_w[3][32];
    __shared__ uint32_t s_carry[3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < 3) s_carry[threadIdx.x] = 0u;
    __syncthreads();
    for (uint32_t b = 0; b < n_tiles; b += 1024) {
        const uint32_t i = b + threadIdx.x;
        const uint4 v = i < n_tiles ? tiles[i] : make_uint4(0u, 0u, 0u, 0u);
        uint32_t inc[3] = {v.x, v.y, v.z};
#pragma unroll
        for (int c = 0; c < 3; c++) {
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, inc[c], o);
                if (lane >= o) inc[c] += t;
            }
            if (lane == 31) s_w[c][warp] = inc[c];
        }
        __syncthreads();
        uint32_t pre[3], tot[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            pre[c] = s_carry[c];
            tot[c] = 0u;
            for (int w = 0; w < 32; w++) {
                const uint32_t t = s_w[c][w];
                if (w < warp) pre[c] += t;
                tot[c] += t;
            }
        }
        if (i < n_tiles) tiles[i] = make_uint4(pre[0] + inc[0] - v.x, pre[1] + inc[1] - v.y, pre[2] + inc[2] - v.z, 0u);
        __syncthreads();
        if (threadIdx.x < 3) s_carry[threadIdx.x] += tot[threadIdx.x];
        __syncthreads();
    }
}

// Pass 3: ent_pref[k] for k in [0, R] (R + 1 values: ent_pref[R] is the total).
__global__ void __launch_bounds__(kPrefBlock) k_pref_write(DevView d, const uint4* __restrict__ tiles, uint4* __restrict__ pref) {
    __shared__ unsigned long long s_w[kPrefBlock / 32];
    const uint32_t R = (uint32_t)min((unsigned long long)d.cap_entries, d.ctr->n_entries);
    if (blockIdx.x * (uint32_t)kPrefTile > R) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t k0 = blockIdx.x * kPrefTile + threadIdx.x * kPrefItems;
    unsigned long long f[kPrefItems];
    unsigned long long mine = 0;
#pragma unroll
    for (int u = 0; u < kPrefItems; u++) {
        f[u] = k0 + u < R ? pref_pack(d.entries[k0 + u]) : 0ull;
        mine += f[u];
    }
    unsigned long long inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    unsigned long long run = inc - mine;
    for (int w = 0; w < warp; w++) run += s_w[w];
    const uint4 tp = tiles[blockIdx.x];
    uint32_t fc = tp.x + (uint32_t)(run & 0x1fffffu), fr = tp.y + (uint32_t)((run >> 21) & 0x1fffffu), hh = tp.z + (uint32_t)(run >> 42);
#pragma unroll
    for (int u = 0; u < kPrefItems; u++) {
        if (k0 + u <= R) pref[k0 + u] = make_uint4(fc, fr, hh, 0u);
        fc += (uint32_t)(f[u] & 1ull);
        fr += (uint32_t)((f[u] >> 21) & 1ull);
        hh += (uint32_t)(f[u] >> 42);
    }
}

// Per-warp scratch of the own-rect path; it lives in PairsWarpSmem::itBox, which that path does not use.
struct OwnWarpSmem {
    int4 q[32];                // walk rect Q of each tester: (x0, y0, x1, y1), clamped
    uint32_t hk[32];           // its home entry
    int32_t x0[32], y0[32];    // its home cell
    uint32_t base[32];         // first cell of its Space
    uint32_t stK[kPairStage];  // staged records: entry index where the other shape was met
    int32_t stY[kPairStage];   // ... and the row of that cell
};
static_assert(sizeof(OwnWarpSmem) <= sizeof(float4) * kItemRound, "OwnWarpSmem must fit PairsWarpSmem::itBox");

// Candidates of row y of the walk Q (the tester itself included) whose entry index is below `cut`.
__device__ __forceinline__ uint32_t own_row_count(const DevView& d, const uint4* __restrict__ P, uint32_t base, int y, int qx0,
                                                  int qx1, int qy0, uint32_t cut) {
    const uint32_t* row = d.cell + (size_t)base + (size_t)(y - d.row_lo) * (size_t)d.W + (size_t)qx0;
    uint32_t s0 = row[0], s1 = row[1], e = qx1 > qx0 ? row[qx1 - qx0 + 1] : s1;
    const uint32_t lim = min(d.cap_entries, cut);
    s0 = min(s0, lim); s1 = min(s1, lim); e = min(e, lim);
    if (y == qy0) return (s1 - s0) + (P[e].x - P[s1].x);
    return (P[s1].y - P[s0].y) + (P[e].z - P[s1].z);
}

// One batch of 32 testers (lane = tester, given by its home entry hk). Returns false — before anything global was
// touched — if the batch produced more records than the stage holds: the caller then runs the full walk on it.
__device__ __forceinline__ bool pairs_batch_own(const DevView& d, const uint4* __restrict__ P, int margin, uint32_t flags, bool have,
                                                uint32_t hk, PairsWarpSmem<false>& sm, unsigned long long& my_cands) {
    OwnWarpSmem& os = *reinterpret_cast<OwnWarpSmem*>(sm.itBox);
    const int lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t A = 0, base = 0, ncand = 0;
    float4 fa = make_float4(0.f, 0.f, 0.f, 0.f);
    int x0 = 0, y0 = 0, sx = 0, sy = 0, qx0 = 0, qx1 = -1, qy0 = 0, qy1 = -1;
    if (have) {
        const uint32_t ent = d.entries[hk];
        A = ent >> kEntShift;
        fa = d.ent_box[hk];
        const uint2 hm = d.ent_home[hk];
        const uint32_t local = hm.x % d.cells_per_space;
        base = hm.x - local;
        x0 = (int)(local % (uint32_t)d.W); y0 = d.row_lo + (int)(local / (uint32_t)d.W);
        sx = (int)(hm.y & 0xffffu); sy = (int)(hm.y >> 16);
        qx0 = max(x0 - margin, 0); qx1 = min(x0 + sx + margin, d.W - 1);
        qy0 = max(y0 - margin, d.row_lo); qy1 = min(y0 + sy + margin, d.row_hi - 1);
        for (int y = qy0; y <= qy1; y++) ncand += own_row_count(d, P, base, y, qx0, qx1, qy0, 0xffffffffu);
        ncand -= 1u;  // the tester's own home entry (cell.go:101)
    }
    sm.A[lane] = A;
    sm.gated[lane] = 0;
    os.q[lane] = make_int4(qx0, qy0, qx1, qy1);
    os.hk[lane] = hk; os.x0[lane] = x0; os.y0[lane] = y0; os.base[lane] = base;
    __syncwarp();
    // ---- walk the own rect: every lane steps through the entries of its tester's rows, one entry per warp step
    int y = y0;
    uint32_t k = 0, kend = 0, s1 = 0;
    bool active = have;
    auto load_row = [&]() {
        const uint32_t* row = d.cell + (size_t)base + (size_t)(y - d.row_lo) * (size_t)d.W + (size_t)x0;
        const uint32_t a = row[0], b = row[1], c = sx > 0 ? row[sx + 1] : b;
        k = min(a, d.cap_entries); s1 = min(b, d.cap_entries); kend = min(c, d.cap_entries);
    };
    if (active) load_row();
    uint32_t produced = 0;
    for (;;) {
        while (active && k >= kend) {
            if (++y > y0 + sy) active = false;
            else load_row();
        }
        if (!__any_sync(0xffffffffu, active)) break;
        bool rec = false;
        uint32_t B = 0;
        if (active) {
            const uint32_t ent = d.entries[k];
            const float4 fb = d.ent_box[k];
            B = ent >> kEntShift;
            const bool cand = B != A && ((ent & kEntFirstCol) || k < s1) && ((ent & kEntFirstRow) || y == y0);
            rec = cand && !(fb.z < fa.x || fb.x > fa.z || fb.w < fa.y || fb.y > fa.w);
        }
        const uint32_t rmask = __ballot_sync(0xffffffffu, rec);
        if (rec) {
            const uint32_t slot = produced + __popc(rmask & lt);
            if (slot < kPairStage) {
                sm.stB[slot] = B;
                sm.stT[slot] = (uint32_t)lane << 24;
                os.stK[slot] = k;
                os.stY[slot] = y;
            }
        }
        produced += __popc(rmask);
        if (active) k++;
    }
    if (produced > kPairStage) return false;
    __syncwarp();
    // ---- generation rank of every staged record, one record per lane
    for (uint32_t r = lane; r < produced; r += 32) {
        const uint32_t t = sm.stT[r] >> 24, B = sm.stB[r], kk = os.stK[r];
        const int yk = os.stY[r];
        const int4 q = os.q[t];
        const uint32_t bs = os.base[t];
        int bx0, by0, bx1, by1;
        clamp_rect(d, d.crect[B], bx0, by0, bx1, by1);
        const int fx = max(bx0, q.x), fy = max(by0, q.y);   // B's first occurrence in the walk
        const int xk = (d.entries[kk] & kEntFirstCol) ? bx0 : os.x0[t];
        uint32_t kb = kk;
        if (fx != xk || fy != yk) {
            const uint32_t* cp = d.cell + (size_t)bs + (size_t)(fy - d.row_lo) * (size_t)d.W + (size_t)fx;
            uint32_t lo = min(cp[0], d.cap_entries), hi = min(cp[1], d.cap_entries);
            while (lo < hi) {   // entry words ascend with the slot inside a cell
                const uint32_t mid = (lo + hi) >> 1;
                if ((d.entries[mid] >> kEntShift) < B) lo = mid + 1;
                else hi = mid;
            }
            kb = lo;
        }
        uint32_t rank = 0;
        for (int yy = q.y; yy < fy; yy++) rank += own_row_count(d, P, bs, yy, q.x, q.z, q.y, 0xffffffffu);
        rank += own_row_count(d, P, bs, fy, q.x, q.z, q.y, kb);
        const int y0t = os.y0[t];
        if (y0t < fy || (y0t == fy && os.hk[t] < kb)) rank -= 1u;
        sm.stR[r] = rank << 8;
    }
    __syncwarp();
    // ---- order each tester's records by rank (generation order); ranks of one tester are distinct
    for (uint32_t r = lane; r < produced; r += 32) {
        const uint32_t t = sm.stT[r] >> 24, rk = sm.stR[r];
        uint32_t pos = 0, cnt = 0;
        for (uint32_t r2 = 0; r2 < produced; r2++)
            if ((sm.stT[r2] >> 24) == t) {
                cnt++;
                if (sm.stR[r2] < rk) pos++;
            }
        sm.stT[r] = (t << 24) | pos;   // (the top byte, which other lanes are reading, does not change)
        if (pos == 0) sm.gated[t] = cnt;
    }
    __syncwarp();
    const uint32_t cnt = sm.gated[lane];
    if (have) my_cands += ncand;
    uint32_t inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
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
    for (uint32_t r = lane; r < produced; r += 32) {
        const uint32_t tt = sm.stT[r], t = tt >> 24;
        const unsigned long long pos = (unsigned long long)chunk + sm.excl[t] + (tt & 0xffffffu);
        if (pos < d.cap_pairs) *reinterpret_cast<uint4*>(d.hits + pos) = make_uint4(sm.A[t], sm.stB[r], 0u, sm.stR[r]);
    }
    __syncwarp();
    return true;
}

// Second form of the own-rect batch: the same results with shorter chains of dependent loads. The offsets of the
// first kOwnRows rows of the walk and of the own rows are requested together, then all prefix counts together; the
// per-row counts stay in shared memory for the rank phase (no second pass over the rows); the walk takes two entries
// of a lane's row per step; the entry of B in a small cell is found with one round of loads.
constexpr int kOwnRows = 4;
struct OwnWarpSmem2 {
    int4 q[32];
    uint32_t hk[32];
    int32_t x0[32], y0[32];
    uint32_t base[32];
    uint32_t stK[kPairStage];
    int32_t stY[kPairStage];
    uint2 rs[kOwnRows][32];     // per (row of the walk, tester): (s0, s1), clamped to the entry capacity
    uint32_t rc[kOwnRows][32];  // candidates of the whole row
    uint32_t ru[kOwnRows][32];  // first row of the walk: FC(s1); other rows: FR(s0)
    uint32_t rv[kOwnRows][32];  // candidates of the row's first cell
    uint32_t rw[kOwnRows][32];  // H(s1)
};
static_assert(sizeof(OwnWarpSmem2) <= offsetof(PairsWarpSmem<false>, queue), "OwnWarpSmem2 must end before PairsWarpSmem::queue");

__device__ __forceinline__ bool pairs_batch_own2(const DevView& d, const uint4* __restrict__ P, int margin, uint32_t flags, bool have,
                                                 uint32_t hk, PairsWarpSmem<false>& sm, unsigned long long& my_cands) {
    OwnWarpSmem2& os = *reinterpret_cast<OwnWarpSmem2*>(&sm);
    const int lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    const uint32_t cap = d.cap_entries;
    uint32_t A = 0, base = 0, ncand = 0;
    float4 fa = make_float4(0.f, 0.f, 0.f, 0.f);
    int x0 = 0, y0 = 0, sx = 0, sy = -1, qx0 = 0, qx1 = -1, qy0 = 0, qy1 = -1;
    if (have) {
        const uint32_t ent = d.entries[hk];
        A = ent >> kEntShift;
        fa = d.ent_box[hk];
        const uint2 hm = d.ent_home[hk];
        const uint32_t local = hm.x % d.cells_per_space;
        base = hm.x - local;
        x0 = (int)(local % (uint32_t)d.W); y0 = d.row_lo + (int)(local / (uint32_t)d.W);
        sx = (int)(hm.y & 0xffffu); sy = (int)(hm.y >> 16);
        qx0 = max(x0 - margin, 0); qx1 = min(x0 + sx + margin, d.W - 1);
        qy0 = max(y0 - margin, d.row_lo); qy1 = min(y0 + sy + margin, d.row_hi - 1);
    }
    // (timing experiments, results meaningless: 0x20000 no rank phase, 0x40000 no walk, 0x100000 no row counts)
    const int rows = (have && !(flags & 0x100000u)) ? qy1 - qy0 + 1 : 0;
    const uint32_t* cell0 = d.cell + (size_t)base;
    // ---- one round of offset loads: the first kOwnRows rows of the walk and of the own rect
    uint32_t s0[kOwnRows], s1[kOwnRows], e[kOwnRows], oa[kOwnRows], ob[kOwnRows], oc[kOwnRows];
#pragma unroll
    for (int r = 0; r < kOwnRows; r++) {
        s0[r] = s1[r] = e[r] = oa[r] = ob[r] = oc[r] = 0u;
        if (r < rows) {
            const uint32_t* row = cell0 + (size_t)(qy0 + r - d.row_lo) * (size_t)d.W + (size_t)qx0;
            s0[r] = row[0]; s1[r] = row[1]; e[r] = qx1 > qx0 ? row[qx1 - qx0 + 1] : s1[r];
        }
        if (have && r <= sy) {
            const uint32_t* row = cell0 + (size_t)(y0 + r - d.row_lo) * (size_t)d.W + (size_t)x0;
            oa[r] = row[0]; ob[r] = row[1]; oc[r] = sx > 0 ? row[sx + 1] : ob[r];
        }
    }
#pragma unroll
    for (int r = 0; r < kOwnRows; r++) {
        s0[r] = min(s0[r], cap); s1[r] = min(s1[r], cap); e[r] = min(e[r], cap);
        oa[r] = min(oa[r], cap); ob[r] = min(ob[r], cap); oc[r] = min(oc[r], cap);
    }
    // ---- one round of prefix-count loads
    uint4 p0[kOwnRows], p1[kOwnRows], pe[kOwnRows];
#pragma unroll
    for (int r = 0; r < kOwnRows; r++) {
        p0[r] = p1[r] = pe[r] = make_uint4(0u, 0u, 0u, 0u);
        if (r < rows) { p0[r] = P[s0[r]]; p1[r] = P[s1[r]]; pe[r] = P[e[r]]; }
    }
#pragma unroll
    for (int r = 0; r < kOwnRows; r++) {
        uint32_t u = 0, v = 0, cnt = 0;
        if (r < rows) {
            if (r == 0) { v = s1[r] - s0[r]; cnt = v + (pe[r].x - p1[r].x); u = p1[r].x; }
            else { v = p1[r].y - p0[r].y; cnt = v + (pe[r].z - p1[r].z); u = p0[r].y; }
            ncand += cnt;
        }
        os.rs[r][lane] = make_uint2(s0[r], s1[r]);
        os.rc[r][lane] = cnt; os.ru[r][lane] = u; os.rv[r][lane] = v; os.rw[r][lane] = p1[r].z;
    }
    if (have && !(flags & 0x100000u)) {
        for (int y = qy0 + kOwnRows; y <= qy1; y++) ncand += own_row_count(d, P, base, y, qx0, qx1, qy0, 0xffffffffu);
        ncand -= 1u;  // the tester's own home entry (cell.go:101)
    }
    sm.A[lane] = A;
    sm.gated[lane] = 0;
    os.q[lane] = make_int4(qx0, qy0, qx1, qy1);
    os.hk[lane] = hk; os.x0[lane] = x0; os.y0[lane] = y0; os.base[lane] = base;
    __syncwarp();
    // ---- walk the own rect, two entries of the lane's current row per warp step
    int ry = 0;
    uint32_t k = 0, kend = 0, fs1 = 0;
    bool active = have && !(flags & 0x40000u);
    auto set_row = [&]() {
        uint32_t a, b, c;
        if (ry < kOwnRows) {
            a = ry == 0 ? oa[0] : (ry == 1 ? oa[1] : (ry == 2 ? oa[2] : oa[3]));
            b = ry == 0 ? ob[0] : (ry == 1 ? ob[1] : (ry == 2 ? ob[2] : ob[3]));
            c = ry == 0 ? oc[0] : (ry == 1 ? oc[1] : (ry == 2 ? oc[2] : oc[3]));
        } else {
            const uint32_t* row = cell0 + (size_t)(y0 + ry - d.row_lo) * (size_t)d.W + (size_t)x0;
            a = min(row[0], cap); b = min(row[1], cap); c = sx > 0 ? min(row[sx + 1], cap) : b;
        }
        k = a; fs1 = b; kend = c;
    };
    if (active) set_row();
    uint32_t produced = 0;
    for (;;) {
        while (active && k >= kend) {
            if (++ry > sy) active = false;
            else set_row();
        }
        if (!__any_sync(0xffffffffu, active)) break;
        bool rec0 = false, rec1 = false;
        uint32_t B0 = 0, B1 = 0;
        const bool two = active && k + 1 < kend;
        if (active) {
            const uint32_t ent0 = d.entries[k];
            const float4 fb0 = d.ent_box[k];
            uint32_t ent1 = 0;
            float4 fb1 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (two) { ent1 = d.entries[k + 1]; fb1 = d.ent_box[k + 1]; }
            B0 = ent0 >> kEntShift;
            B1 = ent1 >> kEntShift;
            const bool c0 = B0 != A && ((ent0 & kEntFirstCol) || k < fs1) && ((ent0 & kEntFirstRow) || ry == 0);
            const bool c1 = two && B1 != A && ((ent1 & kEntFirstCol) || k + 1 < fs1) && ((ent1 & kEntFirstRow) || ry == 0);
            rec0 = c0 && !(fb0.z < fa.x || fb0.x > fa.z || fb0.w < fa.y || fb0.y > fa.w);
            rec1 = c1 && !(fb1.z < fa.x || fb1.x > fa.z || fb1.w < fa.y || fb1.y > fa.w);
        }
        const uint32_t m0 = __ballot_sync(0xffffffffu, rec0), m1 = __ballot_sync(0xffffffffu, rec1);
        if (rec0) {
            const uint32_t slot = produced + __popc(m0 & lt);
            if (slot < kPairStage) { sm.stB[slot] = B0; sm.stT[slot] = (uint32_t)lane << 24; os.stK[slot] = k; os.stY[slot] = y0 + ry; }
        }
        if (rec1) {
            const uint32_t slot = produced + __popc(m0) + __popc(m1 & lt);
            if (slot < kPairStage) { sm.stB[slot] = B1; sm.stT[slot] = (uint32_t)lane << 24; os.stK[slot] = k + 1; os.stY[slot] = y0 + ry; }
        }
        produced += __popc(m0) + __popc(m1);
        if (active) k += two ? 2u : 1u;
    }
    if (produced > kPairStage) return false;
    __syncwarp();
    // ---- generation rank of every staged record, one record per lane
    for (uint32_t r = lane; r < produced; r += 32) {
        const uint32_t t = sm.stT[r] >> 24, B = sm.stB[r], kk = os.stK[r];
        const int yk = os.stY[r];
        const int4 q = os.q[t];
        const uint32_t bs = os.base[t];
        if (flags & 0x20000u) { sm.stR[r] = r << 8; continue; }
        const int4 cr = d.crect[B];
        const uint32_t entk = d.entries[kk];
        int bx0, by0, bx1, by1;
        clamp_rect(d, cr, bx0, by0, bx1, by1);
        const int fx = max(bx0, q.x), fy = max(by0, q.y);   // B's first occurrence in the walk
        const int xk = (entk & kEntFirstCol) ? bx0 : os.x0[t];
        uint32_t kb = kk;
        if (fx != xk || fy != yk) {
            const uint32_t* cp = d.cell + (size_t)bs + (size_t)(fy - d.row_lo) * (size_t)d.W + (size_t)fx;
            uint32_t lo = min(cp[0], cap), hi = min(cp[1], cap);
            if (hi - lo <= 4u) {   // entry words ascend with the slot inside a cell: count the ones below B
                const uint32_t e0 = lo < hi ? d.entries[lo] : 0xffffffffu, e1 = lo + 1 < hi ? d.entries[lo + 1] : 0xffffffffu;
                const uint32_t e2 = lo + 2 < hi ? d.entries[lo + 2] : 0xffffffffu, e3 = lo + 3 < hi ? d.entries[lo + 3] : 0xffffffffu;
                kb = lo + ((e0 >> kEntShift) < B) + ((e1 >> kEntShift) < B) + ((e2 >> kEntShift) < B) + ((e3 >> kEntShift) < B);
            } else {
                while (lo < hi) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if ((d.entries[mid] >> kEntShift) < B) lo = mid + 1;
                    else hi = mid;
                }
                kb = lo;
            }
        }
        const int ri = fy - q.y;
        uint32_t rank = 0;
        if (ri < kOwnRows) {
#pragma unroll
            for (int rr = 0; rr < kOwnRows - 1; rr++)
                if (rr < ri) rank += os.rc[rr][t];
            const uint2 ss = os.rs[ri][t];
            const uint32_t u = os.ru[ri][t], v = os.rv[ri][t], w = os.rw[ri][t];
            const uint4 pk = P[kb];
            if (ri == 0) rank += kb < ss.y ? kb - ss.x : v + (pk.x - u);
            else rank += kb < ss.y ? pk.y - u : v + (pk.z - w);
        } else {
            for (int yy = q.y; yy < fy; yy++) rank += own_row_count(d, P, bs, yy, q.x, q.z, q.y, 0xffffffffu);
            rank += own_row_count(d, P, bs, fy, q.x, q.z, q.y, kb);
        }
        const int y0t = os.y0[t];
        if (y0t < fy || (y0t == fy && os.hk[t] < kb)) rank -= 1u;
        sm.stR[r] = rank << 8;
    }
    __syncwarp();
    // ---- order each tester's records by rank (generation order); ranks of one tester are distinct
    for (uint32_t r = lane; r < produced; r += 32) {
        const uint32_t t = sm.stT[r] >> 24, rk = sm.stR[r];
        uint32_t pos = 0, cnt = 0;
        for (uint32_t r2 = 0; r2 < produced; r2++)
            if ((sm.stT[r2] >> 24) == t) {
                cnt++;
                if (sm.stR[r2] < rk) pos++;
            }
        sm.stT[r] = (t << 24) | pos;   // (the top byte, which other lanes are reading, does not change)
        if (pos == 0) sm.gated[t] = cnt;
    }
    __syncwarp();
    const uint32_t cnt = sm.gated[lane];
    if (have) my_cands += ncand;
    uint32_t inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
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
    for (uint32_t r = lane; r < produced; r += 32) {
        const uint32_t tt = sm.stT[r], t = tt >> 24;
        const unsigned long long pos = (unsigned long long)chunk + sm.excl[t] + (tt & 0xffffffu);
        if (pos < d.cap_pairs) *reinterpret_cast<uint4*>(d.hits + pos) = make_uint4(sm.A[t], sm.stB[r], 0u, sm.stR[r]);
    }
    __syncwarp();
    return true;
}

template <int FORM>
__global__ void __launch_bounds__(kPairsBlock) k_pairs_own(DevView d, const uint4* __restrict__ pref, int margin, uint32_t flags) {
    __shared__ PairsWarpSmem<false> s_warp[kPairsWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    PairsWarpSmem<false>& sm = s_warp[warp];
    uint32_t* q = sm.queue;
    unsigned long long my_cands = 0;
    const uint32_t R = (uint32_t)min((unsigned long long)d.cap_entries, d.ctr->n_entries);
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
            const uint32_t k = base + lane;
            const uint32_t ent = k < hi ? d.entries[k] : 0u;
            const uint32_t want = kEntTester | kEntFirstCol | kEntFirstRow;
            const bool home = (ent & want) == want;
            const uint32_t m = __ballot_sync(0xffffffffu, home);
            if (home) q[qn + __popc(m & ((1u << lane) - 1u))] = k;
            qn += __popc(m);
            __syncwarp();
            if (qn >= 32) {
                const uint32_t hk = q[lane];
                __syncwarp();
                const bool done = FORM ? pairs_batch_own2(d, pref, margin, flags, true, hk, sm, my_cands)
                                       : pairs_batch_own(d, pref, margin, flags, true, hk, sm, my_cands);
                if (!done) pairs_batch<false>(d, margin, flags, true, hk, 0u, sm, my_cands);
                if (lane + 32 < qn) { const uint32_t t = q[lane + 32]; q[lane] = t; }
                qn -= 32;
                __syncwarp();
            }
        }
    }
    if (qn) {
        const uint32_t hk = lane < qn ? q[lane] : kNoEntry;
        const bool done = FORM ? pairs_batch_own2(d, pref, margin, flags, lane < qn, hk, sm, my_cands)
                               : pairs_batch_own(d, pref, margin, flags, lane < qn, hk, sm, my_cands);
        if (!done) pairs_batch<false>(d, margin, flags, lane < qn, hk, 0u, sm, my_cands);
    }
    // testers with no stored cell whose margin still reaches the grid: the full walk (their own rect is empty, yet a
    // shape straddling the border can still pass the Bounds gate)
    const uint32_t n_orph = min(d.ctr->n_orphans, d.n_shapes);
    for (;;) {
        uint32_t o = 0;
        if (lane == 0) o = atomicAdd(&d.ctr->orphan_cursor, 32u);
        o = __shfl_sync(0xffffffffu, o, 0);
        if (o >= n_orph) break;
        const bool have = o + lane < n_orph;
        const uint32_t A = have ? d.orphans[o + lane] : 0u;
        pairs_batch<false>(d, margin, flags, have, kNoEntry, A, sm, my_cands);
    }
    for (int o = 16; o > 0; o >>= 1) my_cands += __shfl_down_sync(0xffffffffu, my_cands, o);
    if (lane == 0 && my_cands) atomicAdd(&d.ctr->n_candidates, my_cands);
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
