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
// cnt_up / cnt_dn: the append counters (the header words of the buffers themselves, or local words when the buffers
// are a neighbour's mailbox written through NVLink).
__global__ void __launch_bounds__(kGridBlock) k_halo_pack(DevView d, int up_row_max, int dn_row_min, HaloRec* up,
                                                          HaloRec* dn, uint32_t cap, uint32_t* cnt_up, uint32_t* cnt_dn) {
    for (uint32_t i = d.own_lo + blockIdx.x * blockDim.x + threadIdx.x; i < d.own_hi; i += gridDim.x * blockDim.x) {
        const uint32_t meta = __ldg(d.meta + i);
        if (meta & RSV_META_ABSENT) continue;
        const double2 p = __ldg(d.pos + i);
        const Box lb = load_box(d.laabb + i);
        const int cy = cell_of(dadd(lb.miny, p.y), d.cell_h), ey = cell_of(dadd(lb.maxy, p.y), d.cell_h);
        if (up && cy <= up_row_max) {
            const uint32_t k = atomicAdd(cnt_up, 1u);
            if (k < cap) up[1 + k] = HaloRec{i, 0u, p.x, p.y};
        }
        if (dn && ey >= dn_row_min) {
            const uint32_t k = atomicAdd(cnt_dn, 1u);
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

// ---- peer-memory transport ------------------------------------------------------------------------------------
// Every device owns a MAILBOX (one cudaMalloc, exported through CUDA IPC so that the neighbours' processes can map
// it): two receive buffers per side (epoch parity) plus flag words. A device pushes its packed records straight
// into the neighbour's mailbox over NVLink (plain peer stores from a copy kernel), raises the neighbour's `ready`
// flag, and waits for its own `ready` flags with stream-ordered memory waits (cuStreamWaitValue32): no collective
// library and no spinning kernel on the data path.
struct HaloMailboxHeader {
    uint32_t ready[2][2];  // [side the data came from: 0 upper, 1 lower][parity]: epoch whose records are in recv[side][parity]
    uint32_t done[2][2];   // [side][parity]: last epoch the neighbour on `side` has finished applying out of ITS [.][parity] buffer
    uint32_t cap;
    uint32_t cnt[2];       // append counters of the running pack (records go straight into the neighbours' mailboxes)
    uint32_t pad[5];
};
static_assert(sizeof(HaloMailboxHeader) == 64, "mailbox header layout");

__host__ __device__ inline size_t halo_mailbox_bytes(uint32_t cap) { return sizeof(HaloMailboxHeader) + 4 * (size_t)(cap + 1) * sizeof(HaloRec); }
__host__ __device__ inline HaloRec* halo_mailbox_recv(void* mailbox, uint32_t cap, int side, int parity) {
    return reinterpret_cast<HaloRec*>(static_cast<unsigned char*>(mailbox) + sizeof(HaloMailboxHeader)) +
           (size_t)(side * 2 + parity) * (cap + 1);
}

// After the pack kernel has stored its records into the neighbours' mailboxes: publish the counts, raise the
// neighbours' `ready` flags, and reset the local words for the next epoch (one thread).
__global__ void k_halo_publish(uint32_t* cnt, HaloRec* peer_up, HaloRec* peer_dn, uint32_t* ready_up, uint32_t* ready_dn,
                               uint32_t epoch, uint32_t* n_ghosts) {
    __threadfence_system();
    if (peer_up) *reinterpret_cast<volatile uint32_t*>(&peer_up[0].slot) = cnt[0];
    if (peer_dn) *reinterpret_cast<volatile uint32_t*>(&peer_dn[0].slot) = cnt[1];
    __threadfence_system();
    if (ready_up) *reinterpret_cast<volatile uint32_t*>(ready_up) = epoch;
    if (ready_dn) *reinterpret_cast<volatile uint32_t*>(ready_dn) = epoch;
    cnt[0] = 0; cnt[1] = 0;
    *n_ghosts = 0;
}

// k_halo_apply for both sides in one launch (blockIdx.y = side).
__global__ void __launch_bounds__(kGridBlock) k_halo_apply2(DevView d, const HaloRec* recv_up, int off_up, const HaloRec* recv_dn,
                                                            int off_dn, uint32_t cap) {
    const HaloRec* recv = blockIdx.y ? recv_dn : recv_up;
    const int slot_off = blockIdx.y ? off_dn : off_up;
    if (!recv) return;
    const uint32_t n = min(*reinterpret_cast<const volatile uint32_t*>(&recv[0].slot), cap);
    double2* pos = const_cast<double2*>(d.pos);
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const HaloRec r = recv[1 + k];
        const long long slot = (long long)r.slot + slot_off;
        if (slot < 0 || slot >= (long long)d.n_shapes) continue;
        pos[slot] = make_double2(r.x, r.y);
        const uint32_t g = atomicAdd(d.n_ghosts, 1u);
        if (g < d.cap_ghosts) d.ghosts[g] = (uint32_t)slot;
    }
}

// One word to a peer, after everything this stream wrote before it.
__global__ void k_halo_signal(uint32_t* flag_a, uint32_t* flag_b, uint32_t value) {
    __threadfence_system();
    if (flag_a) *reinterpret_cast<volatile uint32_t*>(flag_a) = value;
    if (flag_b) *reinterpret_cast<volatile uint32_t*>(flag_b) = value;
    __threadfence_system();
}

}  // namespace rsv
