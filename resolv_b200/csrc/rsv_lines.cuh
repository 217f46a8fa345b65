// rsv_lines.cuh — batched LineTest: rays versus the cell grid.
//
// Replaces LineTest (utils.go:276-370) for a batch of rays at once:
//   cast line = (Start - 0.01 * Unit(End - Start)) -> End                               utils.go:278-282
//   Circle:  line.IntersectionPointsCircle, normals Unit(point - centre)                utils.go:300-311
//   Polygon: line.IntersectionPointsLine against every line of the polygon,
//            normal = that line's Normal()                                              utils.go:313-324
//   per set: Center = mean of the contacts (summed in generation order), contacts sorted by d2 to settings.Start,
//            MTV = contacts[0].Point - Start - 0.01 * vu                                utils.go:328-346
//   sets are ordered by d2(contacts[0].Point, cast line start)                          utils.go:355-357
// OnIntersect is replayed by the host from the compacted buffer (sort key and generation rank are stored).
//
// The reference has no spatial acceleration for rays (TestAgainst is whatever iterator the caller passes). Two
// candidate iterators are offered; both read the grid of the last rsv_frame:
//   RECT (default): every shape registered in the cell rectangle covering the cast line's bounds — exactly the
//        reference-expressible  space.FilterCells(Bounds{cast line}).FilterShapes()  (space.go:181-195, cell.go:73-121),
//        so results equal the reference's for that TestAgainst, bit for bit, in the same generation order.
//   DDA  (RSV_LINE_DDA): only the cells a 4-connected grid traversal of the cast line (extended by kDdaSlack world
//        units at both ends) passes through, in traversal order. This is the fast path for long rays. It visits every
//        cell the segment geometrically crosses; it does NOT reproduce the reference's "phantom" contacts, where the
//        "+1" fudge of IntersectionPointsLine (convexPolygon.go:721-726) reports a hit on a far-away, nearly parallel
//        edge.
// Dedupe (a shape registered in several visited cells is tested once, at its first visited cell) uses the
// first/last row/column flags of the grid entries: no per-ray seen-set and no gather of the other shape's rect.
#pragma once
#include "rsv_device.cuh"
#include "rsv_grid.cuh"
#include "rsv_narrow.cuh"

namespace rsv {

constexpr int kLinesBlock = 128;
constexpr double kDdaSlack = 1.0;

struct LineState {
    // pinned host staging
    double* h_seg = nullptr;               // f64[4 * max_rays]: sx, sy, ex, ey
    unsigned long long* h_mask = nullptr;  // u64[3 * max_rays]: ByTags mask, NotByTags mask, use-ByTags flag
    // device
    double* seg = nullptr;
    unsigned long long* mask = nullptr;
    uint32_t* ray_first = nullptr;  // first hit record of ray i
    uint32_t* ray_count = nullptr;  // number of hit records of ray i
    rsv_ray_hit* hits = nullptr;
    rsv_contact* contacts = nullptr;
    uint32_t max_rays = 0, cap_hits = 0, cap_contacts = 0;
    // pinned result mirrors
    rsv_ray_hit* h_hits = nullptr;
    rsv_contact* h_contacts = nullptr;
    uint32_t h_cap_hits = 0, h_cap_contacts = 0;
    uint32_t* h_ray_first = nullptr;
    uint32_t* h_ray_count = nullptr;
    uint64_t last_hits = 0, last_contacts = 0;
    uint32_t last_rays = 0, launched_flags = 0, launched_rays = 0;
    bool pending = false;
    cudaEvent_t ev[2] = {};
};

inline void lines_free(LineState& L) {
    if (L.h_seg) cudaFreeHost(L.h_seg);
    if (L.h_mask) cudaFreeHost(L.h_mask);
    if (L.h_hits) cudaFreeHost(L.h_hits);
    if (L.h_contacts) cudaFreeHost(L.h_contacts);
    if (L.h_ray_first) cudaFreeHost(L.h_ray_first);
    if (L.h_ray_count) cudaFreeHost(L.h_ray_count);
    void* dev[] = {L.seg, L.mask, L.ray_first, L.ray_count, L.hits, L.contacts};
    for (void* p : dev) if (p) cudaFree(p);
    for (auto& e : L.ev) if (e) cudaEventDestroy(e);
    L = LineState{};
}

struct LinesView {
    const double* seg;
    const unsigned long long* mask;
    uint32_t* ray_first;
    uint32_t* ray_count;
    rsv_ray_hit* hits;
    rsv_contact* contacts;
    uint32_t n_rays, cap_hits, cap_contacts;
};

struct Ray {
    V2 S, E;       // settings.Start / settings.End
    V2 vu;         // Unit(End - Start)
    V2 start;      // cast line start (pulled back by castMargin)
    unsigned long long any, none;
    bool useAny;
    bool homeOnly;  // RSV_LINE_HOME_ONLY
};

// Result of the count pass for one (ray, shape) item, 16 bits: for a polygon of at most kRayMaskEdges lines the set
// of lines the cast line hits (bit i = line i), for a circle 0 / 1 / 3 (= 0, 1 or 2 points); otherwise
// kRayEncCount | number of contacts. The write pass then evaluates only the lines that hit, so its cost no longer
// depends on the polygon's size (items of different shapes diverged there, and most of a polygon's lines miss).
constexpr int kRayMaskEdges = 15;
constexpr uint32_t kRayEncCount = 0x8000u;
__device__ __forceinline__ uint32_t ray_enc_contacts(uint32_t enc) {
    return (enc & kRayEncCount) ? (enc & 0x7fffu) : (uint32_t)__popc(enc);
}

// Tests one shape against the cast line. PASS 0 counts contacts and returns the encoding above; PASS 1 is given that
// encoding, writes the contacts to out[0..n), sorts them, fills the hit record and returns the number of contacts.
template <int PASS>
__device__ __forceinline__ uint32_t ray_vs_shape(const DevView& d, const Ray& ray, uint32_t B, rsv_contact* out,
                                                 rsv_ray_hit* rec, uint32_t enc = kRayEncCount) {
    const uint32_t meta = __ldg(d.meta + B);
    const double2 pb2 = __ldg(d.pos + B);
    const V2 pb{pb2.x, pb2.y};
    uint32_t n = 0, mask = 0;
    V2 sum{0.0, 0.0};
    if (!(meta & RSV_META_POLYGON)) {
        const double rad = __ldg(d.radius + B);
        V2 pts[2];
        const int k = line_x_circle(ray.start, ray.E, pb, rad, pts);
        for (int i = 0; i < k; i++) {
            if (PASS == 1) {
                store_contact(out + n, pts[i], vunit(vsub(pts[i], pb)));
                sum = vadd(sum, pts[i]);
            }
            n++;
        }
        mask = k == 2 ? 3u : (k == 1 ? 1u : 0u);
    } else {
        const int nv = (int)(meta >> RSV_META_NV_SHIFT);
        const int nl = ((meta & RSV_META_CLOSED) && nv > 2) ? nv : max(nv - 1, 0);
        const double2* lv = d.verts + __ldg(d.vert_off + B);
        if (PASS == 1 && !(enc & kRayEncCount)) {
            // only the lines that hit (ascending line index == generation order, utils.go:313-324)
            for (uint32_t m = enc; m; m &= m - 1) {
                const int i = __ffs(m) - 1;
                const double2 qs = __ldg(lv + i), qe = __ldg(lv + (i + 1 < nv ? i + 1 : 0));
                const V2 s{dadd(qs.x, pb.x), dadd(qs.y, pb.y)}, e{dadd(qe.x, pb.x), dadd(qe.y, pb.y)};
                V2 pt;
                if (line_x_line(ray.start, ray.E, s, e, pt)) {
                    store_contact(out + n, pt, edge_normal(s, e));
                    sum = vadd(sum, pt);
                    n++;
                }
            }
        } else if (nl > 0) {
            double2 q = __ldg(lv);
            const V2 first{dadd(q.x, pb.x), dadd(q.y, pb.y)};
            V2 s = first;
            for (int i = 0; i < nl; i++) {
                V2 e = first;
                if (i + 1 < nv) {
                    q = __ldg(lv + i + 1);
                    e = V2{dadd(q.x, pb.x), dadd(q.y, pb.y)};
                }
                V2 pt;
                if (line_x_line(ray.start, ray.E, s, e, pt)) {
                    if (PASS == 1) {
                        store_contact(out + n, pt, edge_normal(s, e));
                        sum = vadd(sum, pt);
                    }
                    if (i < kRayMaskEdges) mask |= 1u << i;
                    n++;
                }
                s = e;
            }
        }
        if (nl > kRayMaskEdges) mask = kRayEncCount | min(n, 0x7fffu);
    }
    if (PASS == 1 && n) {
        // Center (utils.go:332-337), then sort by d2 to settings.Start (utils.go:340-342; n <= 12 -> insertion sort)
        const V2 center{sum.x / (double)n, sum.y / (double)n};
        sort_contacts(out, (int)n, ray.S);
        const V2 p0{out[0].px, out[0].py};
        const V2 mtv = vsub(vsub(p0, ray.S), vscale(ray.vu, 0.01));
        rec->mtv_x = mtv.x; rec->mtv_y = mtv.y;
        rec->center_x = center.x; rec->center_y = center.y;
        rec->key = vdist2(p0, ray.start);
    }
    return PASS == 0 ? mask : n;
}

__device__ __forceinline__ bool ray_filter(const DevView& d, const Ray& ray, uint32_t B) {
    if (ray.homeOnly && !is_home(d, d.crect[B])) return false;   // strips: another device reports this shape
    if (!ray.useAny && ray.none == 0) return true;
    const unsigned long long t = __ldg(d.tags + B);
    if (ray.useAny && (t & ray.any) == 0) return false;   // ByTags (tags.go:29-31)
    return (t & ray.none) == 0;                           // NotByTags
}

// DDA iterator only: conservative cast-line-versus-AABB test on the entry's float box grown by kDdaGate world units
// (separating axes: the segment's bounds, then the box against the segment's supporting line). A shape whose grown
// box the cast line misses cannot touch it geometrically; what is skipped are only contacts the "+1" fudge of
// IntersectionPointsLine invents for lines that pass a shape by (the DDA deviation). Saves the random
// gathers and the per-edge tests for ~80 % of the entries a traversal visits.
constexpr double kDdaGate = 1.0;
__device__ __forceinline__ bool ray_box_gate(const Ray& ray, float4 fb) {
    const double minx = (double)fb.x - kDdaGate, miny = (double)fb.y - kDdaGate;
    const double maxx = (double)fb.z + kDdaGate, maxy = (double)fb.w + kDdaGate;
    const double ax = ray.start.x, ay = ray.start.y, bx = ray.E.x, by = ray.E.y;
    if (fmax(ax, bx) < minx || fmin(ax, bx) > maxx || fmax(ay, by) < miny || fmin(ay, by) > maxy) return false;
    const double dx = bx - ax, dy = by - ay;
    const double cx = 0.5 * (minx + maxx), cy = 0.5 * (miny + maxy), hx = 0.5 * (maxx - minx), hy = 0.5 * (maxy - miny);
    const double s = dx * (cy - ay) - dy * (cx - ax);
    const double r = fabs(dx) * hy + fabs(dy) * hx;
    return !(fabs(s) > r);   // (NaN anywhere: keep the candidate)
}

// Visits every candidate of the ray in generation order and calls f(B, k) (k = index of the grid entry).
template <class F>
__device__ __forceinline__ void ray_candidates(const DevView& d, const Ray& ray, bool dda, uint32_t space, F f) {
    const uint32_t base = space * d.cells_per_space;
    if (!dda) {
        // Space.FilterCells(bounds of the cast line) (space.go:181-195) walked like CellSelection.ForEach (cell.go:86-121)
        const double minx = ray.start.x < ray.E.x ? ray.start.x : ray.E.x, maxx = ray.start.x < ray.E.x ? ray.E.x : ray.start.x;
        const double miny = ray.start.y < ray.E.y ? ray.start.y : ray.E.y, maxy = ray.start.y < ray.E.y ? ray.E.y : ray.start.y;
        const int qx0 = max(cell_of(minx, d.cell_w), 0), qx1 = min(cell_of(maxx, d.cell_w), d.W - 1);
        const int qy0 = max(cell_of(miny, d.cell_h), d.row_lo), qy1 = min(cell_of(maxy, d.cell_h), d.row_hi - 1);
        for (int y = qy0; y <= qy1; y++) {
            const uint32_t* row = d.cell + base + (uint32_t)(y - d.row_lo) * (uint32_t)d.W;
            if (qx0 > qx1) break;
            const uint32_t s0 = row[qx0], s1 = row[qx0 + 1], e = min(row[qx1 + 1], d.cap_entries);
            for (uint32_t k = s0; k < e; k++) {
                const uint32_t ent = d.entries[k];
                if (!((ent & kEntFirstCol) || k < s1)) continue;      // first occurrence in the row-major walk
                if (!((ent & kEntFirstRow) || y == qy0)) continue;
                f(ent >> kEntShift, k);
            }
        }
        return;
    }
    // 4-connected traversal of the cast line, extended by kDdaSlack at both ends.
    const V2 a{dsub(ray.start.x, dmul(ray.vu.x, kDdaSlack)), dsub(ray.start.y, dmul(ray.vu.y, kDdaSlack))};
    const V2 b{dadd(ray.E.x, dmul(ray.vu.x, kDdaSlack)), dadd(ray.E.y, dmul(ray.vu.y, kDdaSlack))};
    // clip the parameter range to the grid box grown by one cell (slab method; conservative)
    const double gx0 = -d.cell_w, gx1 = (double)(d.W + 1) * d.cell_w;
    const double gy0 = (double)(d.row_lo - 1) * d.cell_h, gy1 = (double)(d.row_hi + 1) * d.cell_h;
    const double dx = b.x - a.x, dy = b.y - a.y;
    double t0 = 0.0, t1 = 1.0;
    if (dx != 0.0) {
        double u0 = (gx0 - a.x) / dx, u1 = (gx1 - a.x) / dx;
        if (u0 > u1) { double t = u0; u0 = u1; u1 = t; }
        t0 = fmax(t0, u0); t1 = fmin(t1, u1);
    } else if (a.x < gx0 || a.x > gx1) return;
    if (dy != 0.0) {
        double u0 = (gy0 - a.y) / dy, u1 = (gy1 - a.y) / dy;
        if (u0 > u1) { double t = u0; u0 = u1; u1 = t; }
        t0 = fmax(t0, u0); t1 = fmin(t1, u1);
    } else if (a.y < gy0 || a.y > gy1) return;
    if (!(t0 <= t1)) return;
    const double ax = a.x + t0 * dx, ay = a.y + t0 * dy, bx = a.x + t1 * dx, by = a.y + t1 * dy;
    int ix = cell_of(ax, d.cell_w), iy = cell_of(ay, d.cell_h);
    const int ex = cell_of(bx, d.cell_w), ey = cell_of(by, d.cell_h);
    const int stepx = ex > ix ? 1 : -1, stepy = ey > iy ? 1 : -1;
    int nx = abs(ex - ix), ny = abs(ey - iy);
    // parameter (along a->b clipped) at which the line crosses the next vertical / horizontal grid line
    const double ddx = bx - ax, ddy = by - ay;
    const double tdx = ddx != 0.0 ? d.cell_w / fabs(ddx) : 0.0, tdy = ddy != 0.0 ? d.cell_h / fabs(ddy) : 0.0;
    double tmx = ddx != 0.0 ? (((double)(ix + (stepx > 0 ? 1 : 0)) * d.cell_w - ax) / ddx) : 0.0;
    double tmy = ddy != 0.0 ? (((double)(iy + (stepy > 0 ? 1 : 0)) * d.cell_h - ay) / ddy) : 0.0;
    int came = 0;  // how the current cell was entered: 0 start, 1 step in x, 2 step in y
    // The walk is a chain of dependent loads (cell offsets -> entry -> box); the next cell's offsets do not depend on
    // the current cell's entries, so they are requested before those are processed.
    auto cell_range = [&](int x, int y, uint32_t& s0, uint32_t& e0) {
        s0 = e0 = 0;
        if (x >= 0 && x < d.W && y >= d.row_lo && y < d.row_hi) {
            const uint32_t c = base + (uint32_t)(y - d.row_lo) * (uint32_t)d.W + (uint32_t)x;
            s0 = d.cell[c];
            e0 = min(d.cell[c + 1], d.cap_entries);
        }
    };
    uint32_t s, e;
    cell_range(ix, iy, s, e);
    for (;;) {
        const bool last = nx == 0 && ny == 0;
        const int came_now = came;
        uint32_t s2 = 0, e2 = 0;
        if (!last) {
            if (ny == 0 || (nx != 0 && tmx < tmy)) { ix += stepx; tmx += tdx; nx--; came = 1; }
            else { iy += stepy; tmy += tdy; ny--; came = 2; }
            cell_range(ix, iy, s2, e2);
        }
        // a shape is new here iff the cell we came from is outside its (in-grid) cell rect
        const uint32_t need = came_now == 0 ? 0u : (came_now == 1 ? (stepx > 0 ? kEntFirstCol : kEntLastCol)
                                                                  : (stepy > 0 ? kEntFirstRow : kEntLastRow));
        for (uint32_t k = s; k < e; k++) {
            const uint32_t ent = d.entries[k];
            if ((ent & need) != need) continue;
            f(ent >> kEntShift, k);
        }
        if (last) break;
        s = s2; e = e2;
    }
}

// ---- the kernel -----------------------------------------------------------------------------------------------
// A warp takes 32 rays. The cell walks are per lane (they are short and light once no exact test sits inside
// them), but everything heavy is done one QUEUE ITEM per lane: the walk only appends (shape, owner ray, rank) items
// to a per-warp queue in shared memory — per ray contiguous and in generation order — and the exact tests
// (ray_vs_shape) then run over the queue with all 32 lanes busy, no matter how candidates are distributed over rays.
// The first version tested inside the walk and ran at 2.7-3.9 active threads per warp instruction (ncu, profiles/).
//   W1  walk: count this ray's items (DDA: after the box gate)            -> prefix sums, rounds of <= kRayQueue items
//   W2  write the items of the round's rays: from the first walk's per-lane cache if the ray has <= kRaySeg items
//       (almost always with the DDA iterator), otherwise by walking again
//   T1  per item: count contacts                                           (balanced)
//       per ray: number of hits / contacts, positions of its items; one warp-aggregated allocation per round
//   T2  per hit item: contacts + hit record at its final position          (balanced)
// A single ray with more than kRayQueue items is handled by its lane alone (ray_serial), like the first version.
// The kernel is instantiated per iterator, each with its own occupancy target and queue size (measured, 4 M rays of
// 16-512 px over the 1 M-shape mixed world, profiles/README.md): the cell walks are chains of dependent loads, so
// resident warps matter more than spill-free code. RECT visits ~11 candidates per ray and wants a long queue; DDA
// queues ~1.3 gated items per ray and runs best at 8 blocks per SM with 64 registers.
#ifndef RSV_LINES_MINB_RECT
#define RSV_LINES_MINB_RECT 4
#endif
#ifndef RSV_RAY_QUEUE_RECT
#define RSV_RAY_QUEUE_RECT 448
#endif
#ifndef RSV_LINES_MINB_DDA
#define RSV_LINES_MINB_DDA 8
#endif
#ifndef RSV_RAY_QUEUE_DDA
#define RSV_RAY_QUEUE_DDA 96
#endif
template <bool DDA>
struct LinesCfg {
    static constexpr int kMinBlocks = DDA ? RSV_LINES_MINB_DDA : RSV_LINES_MINB_RECT;
    static constexpr int kQueue = DDA ? RSV_RAY_QUEUE_DDA : RSV_RAY_QUEUE_RECT;   // (qPos packs the hit index in 10 bits)
    static_assert(kQueue <= 1024, "qPos packing");
};
constexpr int kLinesWarps = kLinesBlock / 32;
constexpr int kRaySeg = 8;

template <int kRayQueue>
struct RayWarpSmem {
    double rs[8][32];            // per ray: S.x S.y E.x E.y vu.x vu.y start.x start.y
    uint32_t qB[kRayQueue];      // item: shape slot
    uint32_t qMeta[kRayQueue];   // item: owner lane | generation rank << 5
    uint16_t qN[kRayQueue];      // item: hit lines / contacts (T1, ray_enc_contacts); its class while the processing order is built
    uint32_t qPos[kRayQueue];    // item: index of its hit record within the ray's records | first contact within the ray's << 10
    uint32_t hbase[32], cbase[32];  // per ray: first hit record / first contact of the ray
    uint32_t segB[kRaySeg][32];  // the first items of each ray, kept from the count walk (saves the second walk)
    uint32_t segRank[kRaySeg][32];
    uint16_t perm[kRayQueue];    // processing order of the round's items: binned by shape kind / edge count
    uint32_t bin[16];            // bin cursors of that counting sort
};

// Processing class of a shape for the exact tests: 0 = circle, otherwise min(number of lines, kRayBins - 1). Lanes of a
// warp step that work on items of one class run the same trip counts.
constexpr int kRayBins = 10;
__device__ __forceinline__ uint32_t ray_item_class(const DevView& d, uint32_t B) {
    const uint32_t meta = __ldg(d.meta + B);
    if (!(meta & RSV_META_POLYGON)) return 0u;
    const int nv = (int)(meta >> RSV_META_NV_SHIFT);
    const int nl = ((meta & RSV_META_CLOSED) && nv > 2) ? nv : max(nv - 1, 0);
    return (uint32_t)min(nl + 1, kRayBins - 1);
}

template <int Q>
__device__ __forceinline__ Ray ray_from_smem(const RayWarpSmem<Q>& sm, int o) {
    Ray r{};
    r.S = V2{sm.rs[0][o], sm.rs[1][o]}; r.E = V2{sm.rs[2][o], sm.rs[3][o]};
    r.vu = V2{sm.rs[4][o], sm.rs[5][o]}; r.start = V2{sm.rs[6][o], sm.rs[7][o]};
    return r;
}

__device__ __forceinline__ void load_ray(const DevView& d, const LinesView& L, uint32_t i, Ray& ray, uint32_t& space,
                                         bool home_only) {
    ray.homeOnly = home_only;
    const double2* sp = reinterpret_cast<const double2*>(L.seg + 4 * (size_t)i);
    const double2 s = __ldg(sp), e = __ldg(sp + 1);
    ray.S = V2{s.x, s.y}; ray.E = V2{e.x, e.y};
    ray.vu = vunit(vsub(ray.E, ray.S));
    ray.start = vsub(ray.S, vscale(ray.vu, 0.01));
    ray.any = __ldg(L.mask + 3 * (size_t)i);
    ray.none = __ldg(L.mask + 3 * (size_t)i + 1);
    const unsigned long long w = __ldg(L.mask + 3 * (size_t)i + 2);
    ray.useAny = (w & 1ull) != 0;
    space = (uint32_t)(w >> 32);
    if (space >= d.n_spaces) space = 0;
}

// The whole LineTest of one ray by one thread (count pass, allocation, write pass). Returns the number of hits.
__device__ __noinline__ uint32_t ray_serial(const DevView d, const LinesView L, bool dda, uint32_t i, bool home_only) {  // (by value: a reference would make the kernel keep an addressable copy of its parameters)
    Ray ray{};
    uint32_t space = 0;
    load_ray(d, L, i, ray, space, home_only);
    uint32_t nh = 0, nc = 0;
    ray_candidates(d, ray, dda, space, [&](uint32_t B, uint32_t k) {
        if (!ray_filter(d, ray, B)) return;
        if (dda && !ray_box_gate(ray, d.ent_box[k])) return;
        const uint32_t n = ray_enc_contacts(ray_vs_shape<0>(d, ray, B, nullptr, nullptr));
        if (n) { nh++; nc += n; }
    });
    uint32_t bh = 0, bc = 0;
    if (nh) {
        bh = atomicAdd(&d.ctr->ray_hit_cursor, nh);
        bc = atomicAdd(&d.ctr->ray_contact_cursor, nc);
        if (bh + nh > L.cap_hits || bh + nh < bh) atomicOr(&d.ctr->overflow, RSV_OVF_RAY_HITS);
        if (bc + nc > L.cap_contacts || bc + nc < bc) atomicOr(&d.ctr->overflow, RSV_OVF_RAY_CONTACTS);
    }
    L.ray_first[i] = bh;
    L.ray_count[i] = nh;
    if (nh && (unsigned long long)bh + nh <= L.cap_hits && (unsigned long long)bc + nc <= L.cap_contacts) {
        uint32_t wh = 0, wc = 0, rank = 0;
        ray_candidates(d, ray, dda, space, [&](uint32_t B, uint32_t k) {
            if (!ray_filter(d, ray, B)) return;
            uint32_t enc = 0;
            if (!(dda && !ray_box_gate(ray, d.ent_box[k])) && (enc = ray_vs_shape<0>(d, ray, B, nullptr, nullptr)) != 0 &&
                ray_enc_contacts(enc)) {
                rsv_ray_hit rec;
                rec.ray = i; rec.b = B; rec.first_contact = bc + wc; rec.pad = 0.0;
                const uint32_t n = ray_vs_shape<1>(d, ray, B, L.contacts + bc + wc, &rec, enc);
                rec.count_rank = (n & 0xffu) | (rank << 8);
                L.hits[bh + wh] = rec;
                wh++; wc += n;
            }
            rank++;
        });
    }
    return nh;
}

template <bool DDA>
__global__ void __launch_bounds__(kLinesBlock, LinesCfg<DDA>::kMinBlocks) k_linetest(DevView d, LinesView L, uint32_t flags) {
    constexpr int kRayQueue = LinesCfg<DDA>::kQueue;
    __shared__ RayWarpSmem<kRayQueue> s_warp[kLinesWarps];
    const int lane = threadIdx.x & 31;
    RayWarpSmem<kRayQueue>& sm = s_warp[threadIdx.x >> 5];
    constexpr bool dda = DDA;
    const bool home_only = (flags & RSV_LINE_HOME_ONLY) != 0;
    unsigned long long my_cands = 0, my_hits = 0;
    const uint32_t stride = gridDim.x * kLinesBlock;
    for (uint32_t base = blockIdx.x * kLinesBlock + (threadIdx.x & ~31u); base < L.n_rays; base += stride) {
        const uint32_t i = base + lane;
        const bool live = i < L.n_rays;
        Ray ray{};
        uint32_t space = 0, cnt = 0;
        if (live) {
            load_ray(d, L, i, ray, space, home_only);
            // ---- W1: count the items of this ray (and keep the first few)
            uint32_t rank = 0;
            ray_candidates(d, ray, dda, space, [&](uint32_t B, uint32_t k) {
                if (!ray_filter(d, ray, B)) return;
                if (!dda || ray_box_gate(ray, d.ent_box[k])) {
                    if (cnt < (uint32_t)kRaySeg) { sm.segB[cnt][lane] = B; sm.segRank[cnt][lane] = rank; }
                    cnt++;
                }
                rank++;
            });
            my_cands += rank;
        }
        sm.rs[0][lane] = ray.S.x; sm.rs[1][lane] = ray.S.y; sm.rs[2][lane] = ray.E.x; sm.rs[3][lane] = ray.E.y;
        sm.rs[4][lane] = ray.vu.x; sm.rs[5][lane] = ray.vu.y; sm.rs[6][lane] = ray.start.x; sm.rs[7][lane] = ray.start.y;
        uint32_t incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        __syncwarp();
        // ---- rounds over ranges of rays [r0, r1) whose items fit the queue
        int r0 = 0;
        while (r0 < 32) {
            const uint32_t before = r0 ? __shfl_sync(0xffffffffu, incl, r0 - 1) : 0u;
            const uint32_t fits = __ballot_sync(0xffffffffu, lane >= r0 && incl - before <= (uint32_t)kRayQueue);
            // rays r0.. fit as long as the mask is contiguous from r0 (incl is monotone, so it is)
            const int r1 = fits ? 32 - __clz(fits) : r0;
            if (r1 == r0) {  // this ray alone overflows the queue
                if (lane == r0 && live) my_hits += ray_serial(d, L, dda, i, home_only);
                __syncwarp();
                r0++;
                continue;
            }
            const bool mine = lane >= r0 && lane < r1;
            const uint32_t off = incl - cnt - before;  // first queue slot of this ray
            const uint32_t n_items = __shfl_sync(0xffffffffu, incl, r1 - 1) - before;
            // ---- W2: write the items
            if (mine && cnt && cnt <= (uint32_t)kRaySeg) {
                for (uint32_t q = 0; q < cnt; q++) {
                    sm.qB[off + q] = sm.segB[q][lane];
                    sm.qMeta[off + q] = (uint32_t)lane | (sm.segRank[q][lane] << 5);
                }
            } else if (mine && cnt) {
                uint32_t pos = off, rank = 0;
                ray_candidates(d, ray, dda, space, [&](uint32_t B, uint32_t k) {
                    if (!ray_filter(d, ray, B)) return;
                    if (!dda || ray_box_gate(ray, d.ent_box[k])) {
                        sm.qB[pos] = B;
                        sm.qMeta[pos] = (uint32_t)lane | (rank << 5);
                        pos++;
                    }
                    rank++;
                });
            }
            if (lane < 16) sm.bin[lane] = 0u;
            __syncwarp();
            // ---- order of processing: a counting sort of the items by shape class (the queue itself stays in
            //      generation order; mixed circles / 3..8-gons ran the tests at ~5 active threads per instruction)
            for (uint32_t j = lane; j < n_items; j += 32) {
                const uint32_t c = ray_item_class(d, sm.qB[j]);
                sm.qN[j] = (uint16_t)c;
                atomicAdd(&sm.bin[c], 1u);
            }
            __syncwarp();
            {
                const uint32_t v = lane < 16 ? sm.bin[lane] : 0u;
                uint32_t inc = v;
#pragma unroll
                for (int o = 1; o < 16; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += t;
                }
                __syncwarp();
                if (lane < 16) sm.bin[lane] = inc - v;
            }
            __syncwarp();
            for (uint32_t j = lane; j < n_items; j += 32) sm.perm[atomicAdd(&sm.bin[sm.qN[j]], 1u)] = (uint16_t)j;
            __syncwarp();
            // ---- T1: contacts per item
            for (uint32_t x = lane; x < n_items; x += 32) {
                const uint32_t j = sm.perm[x];
                const Ray r = ray_from_smem(sm, (int)(sm.qMeta[j] & 31u));
                sm.qN[j] = (uint16_t)ray_vs_shape<0>(d, r, sm.qB[j], nullptr, nullptr);   // hit-line mask / contact count (ray_enc_contacts)
            }
            __syncwarp();
            // the items that hit, compacted in place (still grouped by class): T2 runs with full warps
            uint32_t n_hit_items = 0;
            for (uint32_t x0 = 0; x0 < n_items; x0 += 32) {
                const uint32_t x = x0 + lane;
                const uint32_t j = x < n_items ? sm.perm[x] : 0u;
                const bool hit = x < n_items && ray_enc_contacts(sm.qN[j]) != 0u;
                const uint32_t m = __ballot_sync(0xffffffffu, hit);
                __syncwarp();
                if (hit) sm.perm[n_hit_items + __popc(m & ((1u << lane) - 1u))] = (uint16_t)j;
                n_hit_items += __popc(m);
            }
            __syncwarp();
            // ---- per ray: hits, contacts, positions of its items
            uint32_t nh = 0, nc = 0;
            if (mine) {
                for (uint32_t j = off; j < off + cnt; j++) {
                    const uint32_t n = ray_enc_contacts(sm.qN[j]);
                    sm.qPos[j] = nh | (nc << 10);   // nh < kRayQueue <= 1024, nc <= 255 * kRayQueue < 2^22
                    nh += n ? 1u : 0u; nc += n;
                }
            }
            uint32_t ih = nh, ic = nc;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t th = __shfl_up_sync(0xffffffffu, ih, o), tc = __shfl_up_sync(0xffffffffu, ic, o);
                if (lane >= o) { ih += th; ic += tc; }
            }
            const uint32_t toth = __shfl_sync(0xffffffffu, ih, 31), totc = __shfl_sync(0xffffffffu, ic, 31);
            uint32_t bh = 0, bc = 0;
            if (lane == 0 && toth) {
                bh = atomicAdd(&d.ctr->ray_hit_cursor, toth);
                bc = atomicAdd(&d.ctr->ray_contact_cursor, totc);
                if (bh + toth > L.cap_hits || bh + toth < bh) atomicOr(&d.ctr->overflow, RSV_OVF_RAY_HITS);
                if (bc + totc > L.cap_contacts || bc + totc < bc) atomicOr(&d.ctr->overflow, RSV_OVF_RAY_CONTACTS);
            }
            bh = __shfl_sync(0xffffffffu, bh, 0);
            bc = __shfl_sync(0xffffffffu, bc, 0);
            const bool room = (unsigned long long)bh + toth <= L.cap_hits && (unsigned long long)bc + totc <= L.cap_contacts;
            sm.hbase[lane] = bh + (ih - nh);
            sm.cbase[lane] = bc + (ic - nc);
            if (mine && live) {
                L.ray_first[i] = bh + (ih - nh);
                L.ray_count[i] = nh;
                my_hits += nh;
            }
            __syncwarp();
            // ---- T2: contacts and hit records
            if (room) {
                for (uint32_t x = lane; x < n_hit_items; x += 32) {
                    const uint32_t j = sm.perm[x];
                    const uint32_t m = sm.qMeta[j];
                    const int o = (int)(m & 31u);
                    const Ray r = ray_from_smem(sm, o);
                    const uint32_t qp = sm.qPos[j];
                    const uint32_t c0 = sm.cbase[o] + (qp >> 10);
                    rsv_ray_hit rec;
                    rec.ray = base + (uint32_t)o; rec.b = sm.qB[j]; rec.first_contact = c0; rec.pad = 0.0;
                    const uint32_t n = ray_vs_shape<1>(d, r, rec.b, L.contacts + c0, &rec, sm.qN[j]);
                    rec.count_rank = (n & 0xffu) | ((m >> 5) << 8);
                    L.hits[sm.hbase[o] + (qp & 1023u)] = rec;
                }
            }
            __syncwarp();
            r0 = r1;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        my_cands += __shfl_down_sync(0xffffffffu, my_cands, o);
        my_hits += __shfl_down_sync(0xffffffffu, my_hits, o);
    }
    if (lane == 0) {
        if (my_cands) atomicAdd(&d.ctr->n_ray_candidates, my_cands);
        if (my_hits) atomicAdd(&d.ctr->n_ray_hits, my_hits);
    }
}

}  // namespace rsv
