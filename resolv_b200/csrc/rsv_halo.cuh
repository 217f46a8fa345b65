// rsv_halo.cuh — strip partitioning of one large Space across GPUs: halo pack / apply kernels.
//
// The reference has no distributed layer at all. The candidate relation is local in cell space (a tester only
// sees shapes whose cell rect meets its own rect grown by `margin`, shape.go:220-237 + cell.go:86-121), so a Space
// can be cut into horizontal strips of cell rows: each device stores its strip plus a halo of rows, owns a range of
// shape slots, and once per frame tells its two neighbours the positions of the owned shapes that reach into their
// stored rows. Static shape data (local bounds, vertices, tags, ...) of a device's own and its neighbours' slot
// ranges is resident everywhere, so a halo record is just (slot, position): 24 bytes.
#pragma once
#include <climits>

#include "rsv_device.cuh"
#include "rsv_grid.cuh"

namespace rsv {

struct HaloRec {
    uint32_t slot;  // record 0 of a buffer is a header: slot = number of records that follow
    uint32_t pad;
    double x, y;
};
static_assert(sizeof(HaloRec) == 24, "HaloRec layout");

// Owned shapes whose cell rows reach the upper neighbour's stored rows (first row <= up_row_max) or the lower
// neighbour's (last row >= dn_row_min) are appended to the matching send buffer.
__global__ void __launch_bounds__(kGridBlock) k_halo_pack(DevView d, int up_row_max, int dn_row_min, HaloRec* up,
                                                          HaloRec* dn, uint32_t cap) {
    for (uint32_t i = d.own_lo + blockIdx.x * blockDim.x + threadIdx.x; i < d.own_hi; i += gridDim.x * blockDim.x) {
        const uint32_t meta = __ldg(d.meta + i);
        if (meta & RSV_META_ABSENT) continue;
        const double2 p = __ldg(d.pos + i);
        const Box lb = load_box(d.laabb + i);
        const int cy = cell_of(dadd(lb.miny, p.y), d.cell_h), ey = cell_of(dadd(lb.maxy, p.y), d.cell_h);
        if (up && cy <= up_row_max) {
            const uint32_t k = atomicAdd(&up[0].slot, 1u);
            if (k < cap) up[1 + k] = HaloRec{i, 0u, p.x, p.y};
        }
        if (dn && ey >= dn_row_min) {
            const uint32_t k = atomicAdd(&dn[0].slot, 1u);
            if (k < cap) dn[1 + k] = HaloRec{i, 0u, p.x, p.y};
        }
    }
}

// Writes received positions into the resident position buffer and appends the slots to this frame's ghost list.
__global__ void __launch_bounds__(kGridBlock) k_halo_apply(DevView d, const HaloRec* recv, int slot_off, uint32_t cap) {
    const uint32_t n = min(recv[0].slot, cap);
    double2* pos = const_cast<double2*>(d.pos);
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const HaloRec r = recv[1 + k];
        const long long slot = (long long)r.slot + slot_off;
        if (slot < 0 || slot >= (long long)d.n_shapes) continue;
        pos[slot] = make_double2(r.x, r.y);
        const uint32_t g = atomicAdd(d.n_ghosts, 1u);
        if (g < d.cap_ghosts) d.ghosts[g] = (uint32_t)slot;  // (overflow is reported by the next frame's k_bounds)
    }
}

}  // namespace rsv
