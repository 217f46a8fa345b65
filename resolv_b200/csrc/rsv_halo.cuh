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
    // (n_ghosts[1]: sticky "a batch exceeded its buffer" word, turned into RSV_OVF_GHOSTS by the next frame's k_bounds)
    if (recv[0].slot > cap && blockIdx.x == 0 && threadIdx.x == 0) d.n_ghosts[1] = recv[0].slot;
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
    uint32_t epoch;        // fused path (k_bounds<true>): exchanges this device has published, kept on the device
    uint32_t pack_done;    // blocks of the running k_bounds<true> that have finished
    uint32_t apply_done;   // blocks of the running k_halo_apply_bounds that have finished
    uint32_t sticky_ovf;   // a received batch exceeded the mailbox capacity (reported by the next frame: RSV_OVF_GHOSTS)
    uint32_t pad[1];
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
                                                            int off_dn, uint32_t cap, uint32_t* sticky_ovf) {
    const HaloRec* recv = blockIdx.y ? recv_dn : recv_up;
    const int slot_off = blockIdx.y ? off_dn : off_up;
    if (!recv) return;
    const uint32_t n_pub = *reinterpret_cast<const volatile uint32_t*>(&recv[0].slot);
    const uint32_t n = min(n_pub, cap);
    // the sender dropped the records past the capacity: the frame that follows must not pass for complete
    if (n_pub > cap && blockIdx.x == 0 && threadIdx.x == 0 && sticky_ovf) *sticky_ovf = n_pub;
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

// ---- fused path: pack inside k_bounds, apply + ghost registration in one kernel, epochs and waits on the device -----
// k_bounds over the frame's shapes. HALO: the launch covers the OWNED shapes only (DevView::part == kPartOwned) and
// appends every owned shape whose cell rows reach a neighbour's stored rows (first row <= up_row_max / last row >=
// dn_row_min) to that neighbour's mailbox; the last block to finish publishes counts + `ready` flags and advances the
// device-side epoch.
template <bool HALO>
__global__ void __launch_bounds__(kGridBlock) k_bounds(DevView d, int margin, HaloArgs h) {
    unsigned long long my_entries = 0, my_testers = 0;
    const uint32_t n = n_items(d);
    if (blockIdx.x == 0 && threadIdx.x == 0 && d.n_ghosts) {
        if (!HALO && d.part == kPartAll && *d.n_ghosts > d.cap_ghosts) atomicOr(&d.ctr->overflow, RSV_OVF_GHOSTS);
        if (d.n_ghosts[1]) { atomicOr(&d.ctr->overflow, RSV_OVF_GHOSTS); d.n_ghosts[1] = 0u; }
    }
    uint32_t e = 0;
    int par = 0;
    __shared__ int s_wrote;   // some thread of this block stored a record into a neighbour's mailbox
    if (HALO) {
        if (threadIdx.x == 0) s_wrote = 0;
        e = ld_vol_u32(&h.mine->epoch) + 1u;
        par = (int)(e & 1u);
        // the neighbours must have applied what we stored into this parity two epochs ago
        if (threadIdx.x == 0 && e > 2u)
            for (int side = 0; side < 2; side++)
                if (h.dst[side][par])
                    while (ld_vol_u32(&h.mine->done[side][par]) < e - 2u) {}
        __syncthreads();
    }
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t j0 = blockIdx.x * blockDim.x; j0 < n; j0 += stride) {  // (warp-uniform trip count)
        const uint32_t j = j0 + threadIdx.x;
        const bool valid = j < n;
        const uint32_t i = valid ? item_slot(d, j) : 0u;
        double2 p = make_double2(0.0, 0.0);
        if (valid) p = __ldg(d.pos + i);
        int4 r = make_int4(0, 0, -1, -1);
        uint32_t meta = RSV_META_ABSENT;
        bounds_item(d, margin, valid, i, p, my_entries, my_testers, r, meta);
        if (HALO && valid && !(meta & RSV_META_ABSENT)) {
            if (h.dst[0][par] && r.y <= h.up_row_max) {
                const uint32_t k = atomicAdd(&h.mine->cnt[0], 1u);
                if (k < h.cap) h.dst[0][par][1 + k] = HaloRec{i, 0u, p.x, p.y};
                s_wrote = 1;
            }
            if (h.dst[1][par] && r.w >= h.dn_row_min) {
                const uint32_t k = atomicAdd(&h.mine->cnt[1], 1u);
                if (k < h.cap) h.dst[1][par][1 + k] = HaloRec{i, 0u, p.x, p.y};
                s_wrote = 1;
            }
        }
    }
    bounds_stats(d, my_entries, my_testers, d.subset == kSubsetDynamic && blockIdx.x == 0);
    if (HALO) {
        // this block's records must be in the neighbours' memory before it counts as finished: ONE system-scope fence per
        // block, by the thread that counts (cumulative over the block's stores through the barrier) — a fence in every
        // thread cost 40 us per frame here
        __syncthreads();
        if (threadIdx.x == 0 && s_wrote) __threadfence_system();   // (most blocks hold no boundary shape)
        if (threadIdx.x == 0 && atomicAdd(&h.mine->pack_done, 1u) == gridDim.x - 1u) {
            __threadfence_system();
            const uint32_t c0 = atomicAdd(&h.mine->cnt[0], 0u), c1 = atomicAdd(&h.mine->cnt[1], 0u);
            if (h.dst[0][par]) st_vol_u32(&h.dst[0][par][0].slot, c0);
            if (h.dst[1][par]) st_vol_u32(&h.dst[1][par][0].slot, c1);
            __threadfence_system();
            if (h.peer_ready[0][par]) st_vol_u32(h.peer_ready[0][par], e);
            if (h.peer_ready[1][par]) st_vol_u32(h.peer_ready[1][par], e);
            h.mine->cnt[0] = 0u; h.mine->cnt[1] = 0u; h.mine->pack_done = 0u;
            *d.n_ghosts = 0u;
            st_vol_u32(&h.mine->epoch, e);
            __threadfence();
        }
    }
}

// Waits (on the device) for both neighbours' batches of the current epoch, writes the received positions, appends the
// slots to this frame's ghost list and registers the ghosts exactly as k_bounds registers owned shapes. The last block
// tells the neighbours that their buffers of this parity are free again. blockIdx.y = side.
__global__ void __launch_bounds__(kGridBlock) k_halo_apply_bounds(DevView d, int margin, HaloArgs h) {
    unsigned long long my_entries = 0, my_testers = 0;
    const uint32_t e = ld_vol_u32(&h.mine->epoch);
    const int par = (int)(e & 1u), side = (int)blockIdx.y;
    const HaloRec* recv = h.recv[side][par];
    uint32_t n = 0;
    if (h.dst[side][par]) {   // a neighbour exists on this side
        if (threadIdx.x == 0)
            while (ld_vol_u32(&h.mine->ready[side][par]) < e) {}
        __syncthreads();
        const uint32_t n_pub = ld_vol_u32(&recv[0].slot);
        n = min(n_pub, h.cap);
        if (n_pub > h.cap && blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&d.ctr->overflow, RSV_OVF_GHOSTS);
    }
    double2* pos = const_cast<double2*>(d.pos);
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t k0 = blockIdx.x * blockDim.x; k0 < n; k0 += stride) {   // (warp-uniform trip count)
        const uint32_t k = k0 + threadIdx.x;
        bool valid = k < n;
        uint32_t slot = 0;
        double2 p = make_double2(0.0, 0.0);
        if (valid) {
            const HaloRec r = recv[1 + k];
            const long long sl = (long long)r.slot + h.slot_off[side];
            valid = sl >= 0 && sl < (long long)d.n_shapes;
            slot = (uint32_t)sl;
            p = make_double2(r.x, r.y);
            if (valid) {
                pos[slot] = p;
                const uint32_t g = atomicAdd(d.n_ghosts, 1u);
                if (g < d.cap_ghosts) d.ghosts[g] = slot;
                else { atomicOr(&d.ctr->overflow, RSV_OVF_GHOSTS); valid = false; }   // (not listed: must not be registered either)
            }
        }
        int4 r = make_int4(0, 0, -1, -1);
        uint32_t meta = 0;
        bounds_item(d, margin, valid, slot, p, my_entries, my_testers, r, meta);
    }
    bounds_stats(d, my_entries, my_testers, false);
    __syncthreads();
    if (threadIdx.x == 0) __threadfence();
    if (threadIdx.x == 0 && atomicAdd(&h.mine->apply_done, 1u) == gridDim.x * gridDim.y - 1u) {
        __threadfence_system();
        if (h.peer_done[0][par]) st_vol_u32(h.peer_done[0][par], e);
        if (h.peer_done[1][par]) st_vol_u32(h.peer_done[1][par], e);
        h.mine->apply_done = 0u;
        __threadfence_system();
    }
}

}  // namespace rsv
