// rsv_device.cuh — device-side view of one rsv_space (SoA buffers resident in HBM) shared by all kernels.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/resolv_b200.h"
#include "rsv_math.cuh"

namespace rsv {

struct alignas(32) Box {  // (minx, miny, maxx, maxy); moved with ONE 256-bit load / store (LDG.E.256 / STG.E.256, sm_100)
    double minx, miny, maxx, maxy;
};

// Narrowphase bins (rsv_narrow.cuh): pair kind x EXACT vertex counts of the staged polygons, so that the lanes of a warp
// run the same pair test with the same loop trip counts.
//   0        circle-circle
//   1..8     circle-convex, vertices of the polygon              9..16   convex-circle, vertices of the polygon
//   17..80   convex-convex, 17 + 8 (nA - 1) + (nB - 1)
//   80+kind  some polygon has more than kStageVerts (8) vertices (read through L1 instead of the shared-memory stage)
constexpr int kNarrowBins = 84;

// Device counters of one frame. Cursors are 32-bit (capacities are < 2^32); statistics are 64-bit.
struct Counters {
    unsigned long long n_testers;
    unsigned long long n_entries;
    unsigned long long n_candidates;
    unsigned long long n_hits;
    unsigned long long n_contacts_needed;  // total contacts produced, even past capacity
    unsigned int pair_cursor;              // allocation cursor of the pair/hit record buffer (== records needed)
    unsigned int chunk_cursor;             // next chunk of the entry array for k_pairs
    unsigned int orphan_cursor;            // next batch of the orphan tester list for k_pairs
    unsigned int n_orphans;                // testers that touch no stored cell but whose margin reaches the grid
    unsigned int contact_cursor;           // allocation cursor of the contact buffer
    unsigned int overflow;                 // RSV_OVF_*
    unsigned int scan_ticket;              // dynamic tile ids of the chained scan
    unsigned int n_multi;                  // cells holding two or more entries (k_scan -> k_cellsort / k_cellpairs)
    unsigned int grp_cursor;               // allocation cursor of the per-tester record groups (cell-pair path, k_alloc)
    unsigned int pad_;
    unsigned int kind_count[kNarrowBins];  // gated records per narrowphase bin (k_bin / k_place_bin)
    unsigned int kind_cursor[kNarrowBins]; // fill cursors of the bins (k_bin_scatter)
    // line tests
    unsigned int ray_hit_cursor;
    unsigned int ray_contact_cursor;
    unsigned long long n_ray_candidates;
    unsigned long long n_ray_hits;
    unsigned long long n_ray_hits_needed;
    unsigned long long n_ray_contacts_needed;
};

struct DevView {
    // ---- host-committed shape state (RSV_BUF_*)
    const double2* pos;
    const Box* laabb;
    const double* radius;
    const unsigned long long* tags;
    const uint32_t* meta;
    const uint32_t* vert_off;
    const double2* verts;
    const ulonglong2* qmask;
    const uint32_t* space_idx;
    // ---- rebuilt every frame
    Box* aabb;            // world Bounds() per shape (circle.go:33-39, convexPolygon.go:158-161)
    int4* crect;          // Bounds.toCellSpace() per shape (utils.go:90-98), clamped to +-2^30
    uint32_t* cell;       // [n_cells+1]: after the frame, cell c holds entries[cell[c] .. cell[c+1])
    uint32_t* cnt;        // [n_cells+1]: per-cell registration counters of the running frame; zero between frames (k_scan)
    uint32_t* entries;    // shape slot << 6 | flag bits (rsv_grid.cuh), ascending per cell
    float4* ent_box;      // per entry: outward-rounded float32 copy of the shape's world AABB (minx,miny,maxx,maxy)
    unsigned long long* ent_tags;  // per entry: the shape's tags (written only in frames that use tag filters)
    uint2* ent_home;      // per HOME entry: (global index of the home cell, columns-1 | rows-1 << 16 of the clamped rect)
    unsigned long long* tile_state;  // chained-scan tile descriptors
    uint2* multi;                    // (first entry, entry count) of the cells with >= 2 entries
    uint32_t cap_multi;
    uint32_t* tester_start;          // first pair record of tester i
    uint32_t* tester_count;          // number of pair records of tester i
    uint32_t* orphans;               // see Counters::n_orphans
    // ---- cell-pair path (rsv_cellpairs.cuh)
    uint32_t* grp_count;             // [n_shapes]: records per tester of the running frame; zero between frames (k_alloc)
    uint4* ptmp;                     // [cap_pairs]: (A, B, index within A's records) in arrival order
    uint4* pgrp;                     // [cap_pairs]: (A, B, rank) grouped by tester
    // ---- strip partitioning (multi-GPU): which slots this device processes and which rows it tests
    uint32_t own_lo, own_hi;         // owned slot range (default [0, n_shapes))
    uint32_t* ghosts;                // halo slots received this frame
    uint32_t* n_ghosts;              // device-side count of `ghosts` (nullptr: none)
    uint32_t cap_ghosts;
    int32_t home_lo, home_hi;        // rows whose shapes run IntersectionTest here (default [0, H))
    // ---- incremental grid (rsv_grid.cuh, RSV_FRAME_INCREMENTAL): which shapes the grid kernels of this launch visit
    uint32_t subset;                 // kSubsetAll, kSubsetStaticBuild or kSubsetDynamic
    uint32_t part;                   // kPartAll, or kPartOwned for the k_bounds launch of a frame with a fused halo exchange
    uint32_t* dyn_list;              // slots that are re-registered every frame (not RSV_META_STATIC, or static orphan testers)
    uint32_t* n_dyn_dev;             // device-side length of dyn_list (written by the static build)
    uint32_t n_dyn;                  // its value on the host (valid in kSubsetDynamic frames)
    uint32_t static_entries;         // registrations / testers held by the cached static grid
    uint32_t static_testers;
    // ---- results
    rsv_hit* hits;
    uint32_t* kind_queue;            // [cap_pairs]: record indices grouped by narrowphase bin
    rsv_contact* contacts;
    Counters* ctr;
    // ---- geometry of the grid
    uint32_t n_shapes;
    int32_t W, H;            // grid size in cells (space.go:23)
    int32_t row_lo, row_hi;  // stored strip of rows, [row_lo,row_hi) within [0,H)
    double cell_w, cell_h;   // float64(cellWidth), float64(cellHeight)
    uint32_t cells_per_space;  // W * (row_hi-row_lo)
    uint32_t n_cells;          // cells_per_space * n_spaces
    uint32_t n_spaces;
    uint32_t cap_entries, cap_pairs, cap_contacts;
};

// A gather of a Box by shape slot is one L1 wavefront per lane with a 256-bit access instead of two with 128-bit halves
// (every lane hits a different line), and these gathers are what the candidate and narrowphase kernels wait on.
__device__ __forceinline__ Box load_box(const Box* p) {      // host-committed data: read-only path
    Box b;
    asm("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(b.minx), "=d"(b.miny), "=d"(b.maxx), "=d"(b.maxy) : "l"(p));
    return b;
}
__device__ __forceinline__ Box load_box_nc(const Box* p) {   // written by an EARLIER kernel of the same frame: plain load
    Box b;                                                   // (no kernel reads a Box it wrote itself, so the asm needs no memory clobber)
    asm("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(b.minx), "=d"(b.miny), "=d"(b.maxx), "=d"(b.maxy) : "l"(p));
    return b;
}
__device__ __forceinline__ void store_box(Box* p, Box b) {
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(b.minx), "d"(b.miny), "d"(b.maxx), "d"(b.maxy) : "memory");
}

// Outward-rounded float32 copy of a float64 box: contains the exact box, so "disjoint in float" implies
// "disjoint in float64" and can be used as a cheap pre-reject in front of the exact Bounds gate.
__device__ __forceinline__ float4 float_box(const Box& b) {
    return make_float4(__double2float_rd(b.minx), __double2float_rd(b.miny), __double2float_ru(b.maxx),
                       __double2float_ru(b.maxy));
}

constexpr uint32_t kSubsetAll = 0u, kSubsetStaticBuild = 1u, kSubsetDynamic = 2u;
constexpr uint32_t kPartAll = 0u, kPartOwned = 1u;

constexpr int kClampCell = 1 << 30;

// int(math.Floor(v / cell)) (utils.go:92-95), saturated to +-2^30 so that "+- margin" cannot overflow int32.
// Saturation never changes which in-grid cells a rect covers because grids are < 2^30 cells wide.
__device__ __forceinline__ int cell_of(double v, double cell) {
    double f = floor(v / cell);
    if (!(f > -(double)kClampCell)) return -kClampCell;  // also catches NaN
    if (f > (double)kClampCell) return kClampCell;
    return (int)f;
}

}  // namespace rsv
