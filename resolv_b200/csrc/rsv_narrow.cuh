// rsv_narrow.cuh — SAT / analytic narrowphase over the gated pair records.
//
// Replaces Shape.Intersection(other) and everything below it:
//   Circle.Intersection / ConvexPolygon.Intersection   circle.go:69-82, convexPolygon.go:288-300
//   convexConvexTest / convexCircleTest / circleConvexTest / circleCircleTest   shape.go:318-479
//   ConvexPolygon.Transformed / Lines / Project / SATAxes / Center / calculateMTV
//                                                       convexPolygon.go:117-154,200-243,418-514
//   Circle.Project                                      circle.go:41-54
//   collidingLine.Normal / IntersectionPointsLine / IntersectionPointsCircle   convexPolygon.go:705-789
//
// Every float64 operation is performed in the reference's order with IEEE rounding and no FMA, so results are
// bit-identical to the Go code on amd64 (where the reference re-derives Transformed() / Unit() several times it
// gets the same bits each time, so deriving them once per pair is exact). One thread owns one ordered pair;
// the transformed vertices of both polygons are staged once in shared memory ([vertex][thread] layout, conflict
// free 128-bit accesses). Contacts are produced in two passes (count, then write) around one warp-aggregated
// allocation in the contact buffer.
#pragma once
#include <cfloat>

#include "rsv_device.cuh"

namespace rsv {

constexpr int kNarrowBlock = 128;
constexpr int kStageVerts = 8;  // polygons with more vertices are read through L1 instead of shared memory
#ifndef RSV_NARROW_MIN_BLOCKS
#define RSV_NARROW_MIN_BLOCKS 4
#endif

// STAGED: the transformed vertices of this thread's polygon sit in shared memory (n <= kStageVerts); otherwise they
// are re-derived from the local vertices in global memory on every access (same bits: one add per coordinate).
template <bool STAGED>
struct PolyT {
    const double2* sm;  // staged transformed vertices of this thread (stride kNarrowBlock)
    const double2* gl;  // local vertices in global memory
    V2 pos;
    int n;   // vertices
    int nl;  // lines: n if Closed and n > 2, else n-1 (convexPolygon.go:117-141)
    __device__ __forceinline__ V2 get(int i) const {
        if (STAGED) {
            double2 v = sm[i * kNarrowBlock];
            return V2{v.x, v.y};
        }
        double2 v = __ldg(gl + i);
        return V2{dadd(v.x, pos.x), dadd(v.y, pos.y)};  // Transformed(): p + position (convexPolygon.go:151)
    }
    __device__ __forceinline__ V2 line_end(int i) const { return get(i + 1 == n ? 0 : i + 1); }
};

struct Proj {
    double mn, mx;
};

// ConvexPolygon.Project after axis.Unit() (convexPolygon.go:218-232); u is already normalised.
template <class Poly>
__device__ __forceinline__ Proj project_poly(const Poly& p, V2 u) {
    double mn = vdot(u, p.get(0)), mx = mn;
    for (int i = 1; i < p.n; i++) {
        double q = vdot(u, p.get(i));
        if (q < mn) mn = q;
        else if (q > mx) mx = q;
    }
    return Proj{mn, mx};
}

// Circle.Project after axis.Unit() (circle.go:41-54).
__device__ __forceinline__ Proj project_circle(V2 pos, double radius, V2 u) {
    double pc = vdot(u, pos);
    double pr = dmul(vmag(u), radius);
    double mn = dsub(pc, pr), mx = dadd(pc, pr);
    if (mn > mx) { double t = mn; mn = mx; mx = t; }
    return Proj{mn, mx};
}

// Projection.Overlap (utils.go:75-77).
__device__ __forceinline__ double overlap(Proj a, Proj b) { return go_min(dsub(a.mx, b.mn), dsub(b.mx, a.mn)); }

// Necessary condition for 0 < fl(num/det) < 1: same sign, both non-zero, |num| < |det|. Lets almost every edge
// pair skip the two IEEE divisions; pairs that pass still take the exact reference test below.
__device__ __forceinline__ bool may_be_inside(double num, double det) {
    return (num > 0 && det > 0 && num < det) || (num < 0 && det < 0 && num > det);
}

// collidingLine.IntersectionPointsLine (convexPolygon.go:715-744), including the "+ 1" fudge.
__device__ __forceinline__ bool line_x_line(V2 ls, V2 le, V2 os, V2 oe, V2& out) {
    double ldx = dsub(le.x, ls.x), ldy = dsub(le.y, ls.y);
    double odx = dsub(oe.x, os.x), ody = dsub(oe.y, os.y);
    double det = dsub(dmul(ldx, ody), dmul(odx, ldy));
    if (!(det != 0)) return false;
    double sy = dsub(ls.y, os.y), sx = dsub(ls.x, os.x);
    double numL = dadd(dsub(dmul(sy, odx), dmul(sx, ody)), 1.0);
    if (!may_be_inside(numL, det)) return false;
    double numG = dadd(dsub(dmul(sy, ldx), dmul(sx, ldy)), 1.0);
    if (!may_be_inside(numG, det)) return false;
    double lambda = numL / det, gamma = numG / det;
    if ((0 < lambda && lambda < 1) && (0 < gamma && gamma < 1)) {
        out = V2{dadd(ls.x, dmul(lambda, ldx)), dadd(ls.y, dmul(lambda, ldy))};
        return true;
    }
    return false;
}

// collidingLine.IntersectionPointsCircle (convexPolygon.go:747-789). Returns 0..2 points in the reference's order.
__device__ __forceinline__ int line_x_circle(V2 s, V2 e, V2 c, double radius, V2 out[2]) {
    V2 ls = vsub(s, c), le = vsub(e, c), diff = vsub(le, ls);
    double a = dadd(dmul(diff.x, diff.x), dmul(diff.y, diff.y));
    double b = dmul(2.0, dadd(dmul(diff.x, ls.x), dmul(diff.y, ls.y)));
    double cc = dsub(dadd(dmul(ls.x, ls.x), dmul(ls.y, ls.y)), dmul(radius, radius));
    double det = dsub(dmul(b, b), dmul(dmul(4.0, a), cc));
    int n = 0;
    if (det < 0) {
    } else if (det == 0) {
        double t = -b / dmul(2.0, a);
        if (t >= 0 && t <= 1) out[n++] = V2{dadd(s.x, dmul(t, diff.x)), dadd(s.y, dmul(t, diff.y))};
    } else {
        // NaN det falls here too, like the reference's else branch; the comparisons below then fail.
        double sq = sqrt(det), a2 = dmul(2.0, a);
        double t = dadd(-b, sq) / a2;
        if (t >= 0 && t <= 1) out[n++] = V2{dadd(s.x, dmul(t, diff.x)), dadd(s.y, dmul(t, diff.y))};
        t = dsub(-b, sq) / a2;
        if (t >= 0 && t <= 1) out[n++] = V2{dadd(s.x, dmul(t, diff.x)), dadd(s.y, dmul(t, diff.y))};
    }
    return n;
}

// Bounds.Center of the world AABB (utils.go:101-103) == ConvexPolygon.Center (convexPolygon.go:200-215).
__device__ __forceinline__ V2 box_center(const Box& b) {
    return V2{dadd(b.minx, dmul(dsub(b.maxx, b.minx), 0.5)), dadd(b.miny, dmul(dsub(b.maxy, b.miny), 0.5))};
}

// One SAT axis step of calculateMTV (convexPolygon.go:430-441 and twins). axis is a line normal.
// smag caches smallest.Magnitude(), which the reference recomputes on every axis: it only changes with `smallest`.
template <class Poly, class ProjOther>
__device__ __forceinline__ bool sat_axis(const Poly& cp, V2 axis, ProjOther projOther, V2& smallest, double& smag) {
    V2 u = vunit(axis);  // Project() re-normalises its argument
    double ov = overlap(project_poly(cp, u), projOther(u));
    if (ov <= 0) return false;
    if (smag > ov) {
        smallest = vscale(axis, ov);
        smag = vmag(smallest);
    }
    return true;
}

// ConvexPolygon.calculateMTV, *ConvexPolygon case (convexPolygon.go:426-465, 503-513).
template <class Poly>
__device__ __forceinline__ bool mtv_poly_poly(const Poly& cp, const Poly& other, V2 centerCp, V2 centerOther, V2& out) {
    V2 smallest{DBL_MAX, 0.0};
    double smag = vmag(smallest);
    auto projOther = [&](V2 u) { return project_poly(other, u); };
    for (int i = 0; i < cp.nl; i++)
        if (!sat_axis(cp, edge_normal(cp.get(i), cp.line_end(i)), projOther, smallest, smag)) return false;
    for (int i = 0; i < other.nl; i++)
        if (!sat_axis(cp, edge_normal(other.get(i), other.line_end(i)), projOther, smallest, smag)) return false;
    if (vdot(vsub(centerCp, centerOther), smallest) < 0) smallest = vinv(smallest);
    V2 delta = smallest;
    if (vdot(vsub(other.pos, cp.pos), delta) > 0) delta = vinv(delta);
    out = delta;
    return true;
}

// ConvexPolygon.calculateMTV, *Circle case (convexPolygon.go:467-513). The reference sorts a copy of the
// vertices by distance to the circle centre only to take element 0; under its (stable, n <= 12) insertion sort
// that is the first vertex attaining the strict minimum.
template <class Poly>
__device__ __forceinline__ bool mtv_poly_circle(const Poly& cp, V2 centerCp, V2 cpos, double radius, V2& out) {
    V2 closest = cp.get(0);
    double best = vmag(vsub(closest, cpos));
    for (int i = 1; i < cp.n; i++) {
        V2 v = cp.get(i);
        double m = vmag(vsub(v, cpos));
        if (m < best) { best = m; closest = v; }
    }
    auto projOther = [&](V2 u) { return project_circle(cpos, radius, u); };
    V2 axis{dsub(cpos.x, closest.x), dsub(cpos.y, closest.y)};
    V2 u = vunit(axis);
    double ov = overlap(project_poly(cp, u), projOther(u));
    if (ov <= 0) return false;
    V2 smallest = vscale(u, ov);  // axis.Unit().Scale(overlap)
    double smag = vmag(smallest);
    for (int i = 0; i < cp.nl; i++)
        if (!sat_axis(cp, edge_normal(cp.get(i), cp.line_end(i)), projOther, smallest, smag)) return false;
    if (vdot(vsub(centerCp, cpos), smallest) < 0) smallest = vinv(smallest);
    V2 delta = smallest;
    if (vdot(vsub(cpos, cp.pos), delta) > 0) delta = vinv(delta);
    out = delta;
    return true;
}

// Stable insertion sort of contacts by squared distance to `c` (sort.Slice with n <= 12 is an insertion sort;
// shape.go:385-387, 467-469). Beyond 12 items sort.Slice is Go's pdqsort_func — unstable, its result depends on the
// input order — which the host layers run (csrc/host/resolv.hpp: GoSortSlice; the Go package calls sort.Slice itself):
// such sets are delivered in GENERATION order, the input the reference's sort sees (include/resolv_b200.h: rsv_contact).
constexpr int kInsertionSortMax = 12;
__device__ __forceinline__ void sort_contacts(rsv_contact* a, int n, V2 c) {
    if (n > kInsertionSortMax) return;
    for (int i = 1; i < n; i++) {
        rsv_contact key = a[i];
        double kd = vdist2(V2{key.px, key.py}, c);
        int j = i;
        while (j > 0 && kd < vdist2(V2{a[j - 1].px, a[j - 1].py}, c)) { a[j] = a[j - 1]; j--; }
        a[j] = key;
    }
}

__device__ __forceinline__ void store_contact(rsv_contact* p, V2 pt, V2 nrm) {
    double2* q = reinterpret_cast<double2*>(p);
    q[0] = make_double2(pt.x, pt.y);
    q[1] = make_double2(nrm.x, nrm.y);
}

// Visits the contacts of convexConvexTest in generation order: for otherLine in B.Lines(): for line in A.Lines()
// (shape.go:446-461). f(point, j, i) with j the index of B's line and i of A's.
template <class Poly, class F>
__device__ __forceinline__ void visit_poly_poly(const Poly& A, const Poly& B, F f) {
    for (int j = 0; j < B.nl; j++) {
        V2 os = B.get(j), oe = B.line_end(j);
            for (int i = 0; i < A.nl; i++) {
            V2 pt;
            if (line_x_line(A.get(i), A.line_end(i), os, oe, pt)) f(pt, j, i);
        }
    }
}

// Visits line/circle contacts for each line of the polygon (shape.go:328-341, 366-383). f(point, lineIndex).
template <class Poly, class F>
__device__ __forceinline__ void visit_poly_circle(const Poly& P, V2 cpos, double radius, F f) {
    for (int i = 0; i < P.nl; i++) {
        V2 pts[2];
        V2 s = P.get(i), e = P.line_end(i);
        int n = line_x_circle(s, e, cpos, radius, pts);
        for (int k = 0; k < n; k++) f(pts[k], i);
    }
}

template <bool STAGED>
__device__ __forceinline__ PolyT<STAGED> make_poly(const DevView& d, uint32_t slot, uint32_t meta, V2 pos, double2* stage) {
    PolyT<STAGED> p;
    p.n = (int)(meta >> RSV_META_NV_SHIFT);
    p.nl = ((meta & RSV_META_CLOSED) && p.n > 2) ? p.n : max(p.n - 1, 0);
    p.pos = pos;
    p.gl = d.verts + __ldg(d.vert_off + slot);
    p.sm = stage;
    if (STAGED) {
        for (int i = 0; i < p.n; i++) {
            double2 v = __ldg(p.gl + i);
            stage[i * kNarrowBlock] = make_double2(dadd(v.x, pos.x), dadd(v.y, pos.y));
        }
    }
    return p;
}

// Pair kinds: bit 1 = tester is a ConvexPolygon, bit 0 = other is a ConvexPolygon (circle.go:69-82, convexPolygon.go:288-300).
enum { kCircleCircle = 0, kCircleConvex = 1, kConvexCircle = 2, kConvexConvex = 3 };
// Narrowphase bins: kNarrowBins, rsv_device.cuh. (Round 1 binned by vertex-count CLASS, <= 4 / <= 6 / <= 8: a warp of
// 5-gon and 6-gon pairs ran every lane for 36 edge pairs and 12 axes, 12 of 32 threads active per instruction on the
// mixed world.)
constexpr uint32_t kNoBin = 0xffffffffu;
constexpr int kBinUnstaged = 80;   // + kind
// the staged convex-convex bins by falling nA * nB (processing order of k_narrow<0>: heaviest first)
__constant__ uint8_t kOrderCC[64] = {80, 79, 72, 71, 78, 64, 70, 63, 77, 56, 62, 69, 55, 76, 48, 61, 54, 68, 47, 53, 75, 60, 46, 40, 67, 39, 52, 45, 59, 38, 74, 44, 32, 51, 37, 66, 31, 58, 43, 36, 30, 50, 29, 35, 73, 42, 28, 24, 65, 23, 57, 34, 27, 22, 49, 21, 41, 26, 20, 33, 19, 25, 18, 17};
static_assert(kStageVerts == 8 && kNarrowBins == kBinUnstaged + 4, "bin layout");

// Per-record pass in front of the narrowphase, one thread per record:
//  * the exact float64 Bounds gate (utils.go:146-173) — k_pairs only applied the conservative float32 one; records
//    that fail it (and, in RSV_FRAME_ALL_CANDIDATES dumps, records that failed the float gate) are finalised here as
//    empty sets;
//  * computes the record's narrowphase bin (parked in the still unused first_contact field) and counts the bins.
// A mixed world otherwise diverges four ways per warp, times the spread of polygon sizes: 3 active lanes per warp
// were measured before binning.
// Narrowphase bin of an ordered pair (see kNarrowBins).
__device__ __forceinline__ uint32_t narrow_bin_of(const DevView& d, uint32_t A, uint32_t B) {
    const uint32_t ma = __ldg(d.meta + A), mb = __ldg(d.meta + B);
    const bool polyA = ma & RSV_META_POLYGON, polyB = mb & RSV_META_POLYGON;
    const uint32_t na = ma >> RSV_META_NV_SHIFT, nb = mb >> RSV_META_NV_SHIFT;
    const int kind = (polyA ? 2 : 0) | (polyB ? 1 : 0);
    if ((polyA && na > kStageVerts) || (polyB && nb > kStageVerts)) return kBinUnstaged + kind;
    const uint32_t ca = max(na, 1u) - 1u, cb = max(nb, 1u) - 1u;   // (a polygon without vertices shares the 1-vertex bins)
    if (kind == kConvexConvex) return 17 + 8 * ca + cb;
    if (kind == kConvexCircle) return 9 + ca;
    if (kind == kCircleConvex) return 1 + cb;
    return 0;
}

// Cell-pair frames (rsv_cellpairs.cuh): k_place and k_bin in one pass over the grouped records. A record's place within
// its tester's group is the number of the group's records with a smaller rank (generation order); the record is
// written there with its narrowphase bin, and the bins are counted. (The exact Bounds gate was applied by k_cellpairs.)
__global__ void __launch_bounds__(256) k_place_bin(DevView d) {
    if (d.ctr->overflow & RSV_OVF_PAIRS) return;
    const uint32_t n_rec = min(d.ctr->pair_cursor, d.cap_pairs);
    const int lane = threadIdx.x & 31;
    __shared__ uint32_t s_cnt[kNarrowBins];
    if (threadIdx.x < kNarrowBins) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t base = blockIdx.x * blockDim.x; base < n_rec; base += gridDim.x * blockDim.x) {
        const uint32_t g = base + threadIdx.x;
        uint32_t bin = kNoBin;
        if (g < n_rec) {
            const uint4 rec = d.pgrp[g];
            const uint32_t s = d.tester_start[rec.x], n = d.tester_count[rec.x];
            bin = narrow_bin_of(d, rec.x, rec.y);
            uint32_t pos = 0;
            if (n > 1)
                for (uint32_t q = s; q < s + n; q++) pos += d.pgrp[q].z < rec.z;
            *reinterpret_cast<uint4*>(d.hits + s + pos) = make_uint4(rec.x, rec.y, bin, rec.z << 8);
        }
        const uint32_t peers = __match_any_sync(0xffffffffu, bin);
        if (bin != kNoBin && lane == __ffs(peers) - 1) atomicAdd(&s_cnt[bin], __popc(peers));
    }
    __syncthreads();
    if (threadIdx.x < kNarrowBins && s_cnt[threadIdx.x]) atomicAdd(&d.ctr->kind_count[threadIdx.x], s_cnt[threadIdx.x]);
}

// pregated: the records come from a path that already applied the exact gate.
__global__ void __launch_bounds__(256) k_bin(DevView d, bool pregated) {
    if (d.ctr->overflow & RSV_OVF_PAIRS) return;  // records past the capacity were never written: frame is retried
    const uint32_t n_rec = min(d.ctr->pair_cursor, d.cap_pairs);
    const int lane = threadIdx.x & 31;
    const uint32_t stride = gridDim.x * blockDim.x;
    // bin counts are collected per block first: a uniform world puts every record into ONE bin, and one global word
    // cannot take an atomic per warp
    __shared__ uint32_t s_cnt[kNarrowBins];
    if (threadIdx.x < kNarrowBins) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t base = blockIdx.x * blockDim.x; base < n_rec; base += stride) {
        const uint32_t p = base + threadIdx.x;
        uint32_t bin = kNoBin;
        if (p < n_rec) {
            const uint4 rec = *reinterpret_cast<const uint4*>(d.hits + p);
            bool gated = !(rec.w & RSV_HIT_NOT_GATED);
            if (gated && !pregated) {
                const Box a = load_box_nc(d.aabb + rec.x), o = load_box_nc(d.aabb + rec.y);
                gated = bounds_gate(a.minx, a.miny, a.maxx, a.maxy, o.minx, o.miny, o.maxx, o.maxy);
                if (!gated) d.hits[p].count_rank = rec.w | RSV_HIT_NOT_GATED;
            }
            if (!gated) {
                reinterpret_cast<double2*>(d.hits + p)[1] = make_double2(0.0, 0.0);
            } else {
                bin = narrow_bin_of(d, rec.x, rec.y);
            }
            d.hits[p].first_contact = bin;
        }
        const uint32_t peers = __match_any_sync(0xffffffffu, bin);
        if (bin != kNoBin && lane == __ffs(peers) - 1) atomicAdd(&s_cnt[bin], __popc(peers));
    }
    __syncthreads();
    if (threadIdx.x < kNarrowBins && s_cnt[threadIdx.x]) atomicAdd(&d.ctr->kind_count[threadIdx.x], s_cnt[threadIdx.x]);
}

// Second pass: scatters the record indices into one queue, bin after bin (offsets = prefix sums of the bin counts).
__global__ void __launch_bounds__(256) k_bin_scatter(DevView d) {
    if (d.ctr->overflow & RSV_OVF_PAIRS) return;
    __shared__ uint32_t s_off[kNarrowBins];
    __shared__ int s_nonempty;
    if (threadIdx.x < kNarrowBins) s_off[threadIdx.x] = d.ctr->kind_count[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        int ne = 0;
        for (int k = 0; k < kNarrowBins; k++) { const uint32_t c = s_off[k]; s_off[k] = run; run += c; ne += c != 0; }
        s_nonempty = ne;
    }
    __syncthreads();
    if (s_nonempty <= 1) return;  // one bin only (e.g. a world of rectangles): k_narrow walks the records directly
    const uint32_t n_rec = min(d.ctr->pair_cursor, d.cap_pairs);
    const int lane = threadIdx.x & 31;
    const uint32_t stride = gridDim.x * blockDim.x;
    // A block first counts its records per bin and reserves its share of every bin with ONE global atomic per bin
    // (a mixed world has ~10 bins: an atomic per warp step and bin on 10 hot words took 15 us for 170 k records),
    // then places the records through shared-memory cursors.
    __shared__ uint32_t s_cnt[kNarrowBins], s_cur[kNarrowBins];
    if (threadIdx.x < kNarrowBins) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t base = blockIdx.x * blockDim.x; base < n_rec; base += stride) {
        const uint32_t p = base + threadIdx.x;
        const uint32_t bin = p < n_rec ? d.hits[p].first_contact : kNoBin;
        const uint32_t peers = __match_any_sync(0xffffffffu, bin);
        if (bin != kNoBin && lane == __ffs(peers) - 1) atomicAdd(&s_cnt[bin], __popc(peers));
    }
    __syncthreads();
    if (threadIdx.x < kNarrowBins)
        s_cur[threadIdx.x] = s_off[threadIdx.x] + (s_cnt[threadIdx.x] ? atomicAdd(&d.ctr->kind_cursor[threadIdx.x], s_cnt[threadIdx.x]) : 0u);
    __syncthreads();
    for (uint32_t base = blockIdx.x * blockDim.x; base < n_rec; base += stride) {
        const uint32_t p = base + threadIdx.x;
        const uint32_t bin = p < n_rec ? d.hits[p].first_contact : kNoBin;
        const uint32_t peers = __match_any_sync(0xffffffffu, bin);
        const int leader = __ffs(peers) - 1;
        uint32_t at = 0;
        if (bin != kNoBin && lane == leader) at = atomicAdd(&s_cur[bin], __popc(peers));
        at = __shfl_sync(0xffffffffu, at, leader);
        if (bin != kNoBin) d.kind_queue[at + __popc(peers & ((1u << lane) - 1u))] = p;
        else if (p < n_rec) d.hits[p].first_contact = 0u;  // an empty set owns no contacts
    }
}

// One ordered pair of kind KIND, owned by one lane. Two passes around the warp-aggregated contact allocation.
template <int KIND, bool STAGED>
__device__ __forceinline__ void narrow_one(const DevView& d, double2* sva, double2* svb, bool live, uint32_t p) {
    const int lane = threadIdx.x & 31;
    {
        uint32_t A = 0, B = 0, cr = 0;
        if (live) {
            const uint4 rec = *reinterpret_cast<const uint4*>(d.hits + p);
            A = rec.x; B = rec.y; cr = rec.w;
        }
        V2 pa{0, 0}, pb{0, 0};
        double ra = 0, rb = 0;
        PolyT<STAGED> PA, PB;
        PA.n = PB.n = PA.nl = PB.nl = 0;
        uint32_t n = 0;
        unsigned long long xmask = 0;  // crossing edge pairs (bit j*nlA + i) when nlA*nlB <= 64
        bool use_mask = false;
        if (live) {
            double2 t = __ldg(d.pos + A); pa = V2{t.x, t.y};
            t = __ldg(d.pos + B); pb = V2{t.x, t.y};
            if (KIND & 2) PA = make_poly<STAGED>(d, A, __ldg(d.meta + A), pa, sva); else ra = __ldg(d.radius + A);
            if (KIND & 1) PB = make_poly<STAGED>(d, B, __ldg(d.meta + B), pb, svb); else rb = __ldg(d.radius + B);
            // ---- pass 1: count contacts
            if (KIND == kConvexConvex) {
                use_mask = PA.nl * PB.nl <= 64;
                visit_poly_poly(PA, PB, [&](V2, int j, int i) {
                    n++;
                    if (use_mask) xmask |= 1ull << (j * PA.nl + i);
                });
            } else if (KIND == kConvexCircle) {
                visit_poly_circle(PA, pb, rb, [&](V2, int) { n++; });
            } else if (KIND == kCircleConvex) {
                visit_poly_circle(PB, pa, ra, [&](V2, int) { n++; });
            } else {
                // circleCircleTest (shape.go:407-411)
                double dx = dsub(pb.x, pa.x), dy = dsub(pb.y, pa.y);
                double dist = sqrt(dadd(dmul(dx, dx), dmul(dy, dy)));
                if (!(dist > dadd(ra, rb) || dist < fabs(dsub(ra, rb)) || dist == 0)) n = 2;
            }
        }
        // ---- warp-aggregated allocation in the contact buffer
        uint32_t inc = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
        uint32_t nhits = __popc(__ballot_sync(0xffffffffu, n > 0));
        uint32_t first = 0;
        if (lane == 0 && total) {
            first = atomicAdd(&d.ctr->contact_cursor, total);
            atomicAdd(&d.ctr->n_contacts_needed, (unsigned long long)total);
            atomicAdd(&d.ctr->n_hits, (unsigned long long)nhits);
            if (first + total > d.cap_contacts || first + total < first) atomicOr(&d.ctr->overflow, RSV_OVF_CONTACTS);
        }
        first = __shfl_sync(0xffffffffu, first, 0) + (inc - n);
        if (!live) return;
        V2 mtv{0, 0};
        if (n > 127) { atomicOr(&d.ctr->overflow, RSV_OVF_PAIR_CONTACTS); n = 127; }
        if (n && (unsigned long long)first + n <= d.cap_contacts) {
            rsv_contact* out = d.contacts + first;
            uint32_t w = 0;
            if (KIND == kConvexConvex) {
                // convexConvexTest (shape.go:438-479)
                if (use_mask) {
                    // second pass only over the edge pairs that crossed in the first
                    for (unsigned long long m = xmask; m && w < n; m &= m - 1) {
                        int idx = __ffsll((long long)m) - 1;
                        int j = idx / PA.nl, i = idx - j * PA.nl;
                        V2 pt;
                        line_x_line(PA.get(i), PA.line_end(i), PB.get(j), PB.line_end(j), pt);
                        store_contact(out + w, pt, edge_normal(PB.get(j), PB.line_end(j)));
                        w++;
                    }
                } else {
                    visit_poly_poly(PA, PB, [&](V2 pt, int j, int) {
                        if (w < n) store_contact(out + w, pt, edge_normal(PB.get(j), PB.line_end(j)));
                        w++;
                    });
                }
                Box ba = load_box_nc(d.aabb + A), bb = load_box_nc(d.aabb + B);
                V2 ca = box_center(ba);
                sort_contacts(out, (int)n, ca);
                V2 m;
                if (mtv_poly_poly(PA, PB, ca, box_center(bb), m)) mtv = m;
            } else if (KIND == kConvexCircle) {
                // convexCircleTest (shape.go:356-397)
                visit_poly_circle(PA, pb, rb, [&](V2 pt, int) {
                    if (w < n) store_contact(out + w, pt, vunit(vsub(pt, pb)));
                    w++;
                });
                sort_contacts(out, (int)n, pb);
                V2 m;
                if (mtv_poly_circle(PA, box_center(load_box_nc(d.aabb + A)), pb, rb, m)) mtv = m;
            } else if (KIND == kCircleConvex) {
                // circleConvexTest (shape.go:318-354): edge normals, unsorted, inverted MTV
                visit_poly_circle(PB, pa, ra, [&](V2 pt, int i) {
                    if (w < n) store_contact(out + w, pt, edge_normal(PB.get(i), PB.line_end(i)));
                    w++;
                });
                V2 m;
                if (mtv_poly_circle(PB, box_center(load_box_nc(d.aabb + B)), pa, ra, m)) mtv = vinv(m);
            } else {
                // circleCircleTest (shape.go:413-433)
                double dx = dsub(pb.x, pa.x), dy = dsub(pb.y, pa.y);
                double dist = sqrt(dadd(dmul(dx, dx), dmul(dy, dy)));
                double a = dadd(dsub(dmul(ra, ra), dmul(rb, rb)), dmul(dist, dist)) / dmul(2.0, dist);
                double h = sqrt(dsub(dmul(ra, ra), dmul(a, a)));
                double x2 = dadd(pa.x, dmul(a, dx) / dist);
                double y2 = dadd(pa.y, dmul(a, dy) / dist);
                double hx = dmul(h, dy) / dist, hy = dmul(h, dx) / dist;
                V2 p0{dadd(x2, hx), dsub(y2, hy)}, p1{dsub(x2, hx), dadd(y2, hy)};
                store_contact(out, p0, vunit(vsub(p0, pa)));
                store_contact(out + 1, p1, vunit(vsub(p1, pa)));
                V2 m{dsub(pa.x, pb.x), dsub(pa.y, pb.y)};
                double md = vmag(m);
                mtv = vscale(vunit(m), dsub(dadd(ra, rb), md));
            }
        } else if (n) {
            n = 0;  // contact buffer overflow: report as capacity error, never write out of bounds
        }
        uint4* hp = reinterpret_cast<uint4*>(d.hits + p);
        hp[0] = make_uint4(A, B, first, (cr & ~0x7fu) | n);
        reinterpret_cast<double2*>(hp)[1] = make_double2(mtv.x, mtv.y);
    }
}

// One thread per pair. GROUP 0: the convex-convex bins with staged vertices (compiled on their own so that the other
// pair tests do not inflate this kernel's registers and code). GROUP 1: every other bin.
// (Tried twice in round 2 and removed. Eight lanes per pair — vertices, edge normals, edge pairs and SAT axes dealt over the
// lanes, `smallest` folded in axis order. Bit-identical, but the sequential tail of a pair (contact sort, fold, sign
// flips) runs on one lane of eight: 380 warp instructions per pair against 144, and slower on every config —
// 385 / 158 / 41 us against 125 / 110 / 34 us on configs 2 / 3 / 4; profiles/r2_narrow_coop_unroll_variants.log.
// Partial unrolling of the vertex / edge loops for more independent fp64 chains: 127 / 114 us at x2, worse at x4.
// Four lanes per pair for the polygon bins above 4 x 4 vertices only, per-axis results in the idle columns of the
// vertex stage, loops rolled, fold and sort on the first lane (profiles/r3a_narrow_groups.patch): bit-identical
// again, k_narrow<0> 80 -> 76 us, k_narrow<1> 68 -> 115 us (launch lists profiles/r2z_*, r3a_*); with the per-lane
// slots in registers the code tripled to 170 KB of SASS and both kernels got slower.
// A division-free pre-test of every edge pair / polygon line followed by a dense loop over the survivors (so that the
// IEEE divisions do not diverge inside the full loop): 0.3210 -> 0.3195 ms on config 3, nothing on config 2; not kept.)
template <int GROUP>
// (the convex-convex instantiation fits 96 registers without spilling: five blocks per SM; the mixed one does not)
__global__ void __launch_bounds__(kNarrowBlock, GROUP == 0 ? RSV_NARROW_MIN_BLOCKS + 1 : RSV_NARROW_MIN_BLOCKS) k_narrow(DevView d) {
    __shared__ double2 s_va[kStageVerts][kNarrowBlock];
    __shared__ double2 s_vb[kStageVerts][kNarrowBlock];
    if (d.ctr->overflow & RSV_OVF_PAIRS) return;
    double2* sva = &s_va[0][threadIdx.x];
    double2* svb = &s_vb[0][threadIdx.x];
    // The bins of this group form one chunk space of 32 records per chunk (a chunk never straddles two bins); warps
    // take chunks round-robin, so every warp runs one pair test with similar trip counts and the load is spread.
    constexpr int kMaxOrder = 64;
    __shared__ uint32_t s_off[kMaxOrder], s_cnt[kMaxOrder], s_chunk0[kMaxOrder + 1], s_kc[kNarrowBins], s_boff[kNarrowBins];
    __shared__ uint8_t s_bin[kMaxOrder];
    __shared__ int s_n, s_single;
    if (threadIdx.x < kNarrowBins) s_kc[threadIdx.x] = d.ctr->kind_count[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t off = 0;
        int ne = 0, only = 0;
        for (int k = 0; k < kNarrowBins; k++) {
            s_boff[k] = off; off += s_kc[k];
            if (s_kc[k]) { ne++; only = k; }
        }
        // a single non-empty bin was not scattered: its "queue" is the record buffer itself (other records: kNoBin)
        s_single = ne == 1 ? only : -1;
        // processing order: heaviest bins first. GROUP 0: the staged convex-convex bins by falling nA * nB;
        // GROUP 1: convex-circle and circle-convex by falling vertex count, circle-circle, the unstaged kinds.
        const uint32_t pair_cursor = min(d.ctr->pair_cursor, d.cap_pairs);
        uint32_t run = 0;
        int n = 0;
        auto add = [&](int b) {
            s_off[n] = s_boff[b];
            s_bin[n] = (uint8_t)b;
            s_cnt[n] = s_single == b ? pair_cursor : s_kc[b];   // (single bin: walk every record, skip the others)
            s_chunk0[n] = run;
            run += (s_cnt[n] + 31) / 32;
            n++;
        };
        if (GROUP == 0) {
            for (int i = 0; i < 64; i++)
                if (s_kc[kOrderCC[i]] || s_single == kOrderCC[i]) add(kOrderCC[i]);
        } else {
            for (int v = 8; v >= 1; v--) { add(9 + v - 1); add(1 + v - 1); }
            add(0);
            add(kBinUnstaged + kConvexConvex); add(kBinUnstaged + kConvexCircle); add(kBinUnstaged + kCircleConvex);
        }
        s_chunk0[n] = run;
        s_n = n;
    }
    __syncthreads();
    const uint32_t n_chunks = s_chunk0[s_n];
    const uint32_t warp_id = blockIdx.x * (kNarrowBlock / 32) + (threadIdx.x >> 5), n_warps = gridDim.x * (kNarrowBlock / 32);
    const int lane = threadIdx.x & 31;
    for (uint32_t c = warp_id; c < n_chunks; c += n_warps) {
        int i = 0;
        while (c >= s_chunk0[i + 1]) i++;
        const uint32_t idx = (c - s_chunk0[i]) * 32 + lane;
        bool live = idx < s_cnt[i];
        const uint32_t boff = s_off[i];
        uint32_t p = 0;
        if (s_single >= 0) {
            p = idx;
            live = live && d.hits[p].first_contact == (uint32_t)s_single;
        } else if (live) {
            p = d.kind_queue[boff + idx];
        }
        if (GROUP == 0) {
            narrow_one<kConvexConvex, true>(d, sva, svb, live, p);
        } else {
            const int bin = (int)s_bin[i];
            if (bin >= kBinUnstaged) {
                if (bin == kBinUnstaged + kConvexConvex) narrow_one<kConvexConvex, false>(d, sva, svb, live, p);
                else if (bin == kBinUnstaged + kConvexCircle) narrow_one<kConvexCircle, false>(d, sva, svb, live, p);
                else if (bin == kBinUnstaged + kCircleConvex) narrow_one<kCircleConvex, false>(d, sva, svb, live, p);
            } else if (bin >= 9) narrow_one<kConvexCircle, true>(d, sva, svb, live, p);
            else if (bin >= 1) narrow_one<kCircleConvex, true>(d, sva, svb, live, p);
            else narrow_one<kCircleCircle, true>(d, sva, svb, live, p);
        }
    }
}

}  // namespace rsv
