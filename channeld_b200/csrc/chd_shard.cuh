// chd_shard.cuh — multi-GPU X-slab sharding kernels: border export / halo import records.
#pragma once
#include "chd_types.cuh"

namespace chd {

// ---- X-slab sharding.  A record is (global entity id, cell index).
// An own entity is exported when another rank may need it: its column is not strictly interior to this slab.
__global__ void border_flag_kernel(GridDev g, const uint32_t* __restrict__ key, uint32_t n, uint32_t* __restrict__ flag,
                                   unsigned long long* bump_epoch) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *bump_epoch = chd_next_epoch(*bump_epoch);  // border stage epoch
    if (i >= n) return;
    const uint32_t k = key[i];
    uint32_t f = 0;
    if (k < g.cells) {
        const uint32_t col = k % g.cols;
        const bool interior = col >= g.col_lo + g.halo && col + g.halo < g.col_hi;
        const bool left_edge_open = g.col_lo > 0, right_edge_open = g.col_hi < g.cols;
        if (!interior) {
            // columns near a world edge with no neighbour beyond need no export
            const bool near_left = col < g.col_lo + g.halo, near_right = col + g.halo >= g.col_hi;
            f = (near_left && left_edge_open) || (near_right && right_edge_open) || col < g.col_lo || col >= g.col_hi;
        }
    }
    flag[i] = f;
}

// writes the flagged records and pads the rest of the caller's buffer with 0xFFFFFFFF (no separate fill needed)
__global__ void border_write_kernel(const uint32_t* __restrict__ key, const uint32_t* __restrict__ gid, uint32_t n,
                                    const uint32_t* __restrict__ flag, const uint32_t* __restrict__ off, uint32_t* __restrict__ out,
                                    uint32_t cap, Counters* __restrict__ ctr) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t count = off[n];
    if (i == 0 && count > cap) atomicOr(&ctr->overflow, (uint32_t)CHD_OVF_BORDER);
    if (i < cap && i >= count) {  // padding slot
        out[2 * i] = 0xFFFFFFFFu;
        out[2 * i + 1] = 0xFFFFFFFFu;
    }
    if (i >= n || !flag[i]) return;
    const uint32_t o = off[i];
    if (o < cap) {
        out[2 * o] = gid ? gid[i] : i;
        out[2 * o + 1] = key[i];
    }
}

// keep gathered records whose column lies in this rank's extended range and which another rank exported
__global__ void halo_flag_kernel(GridDev g, const uint32_t* __restrict__ rec, uint32_t n, uint32_t skip_first, uint32_t skip_count,
                                 uint32_t* __restrict__ flag, unsigned long long* bump_epoch) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *bump_epoch = chd_next_epoch(*bump_epoch);  // border stage epoch
    if (i >= n) return;
    const uint32_t cell = rec[2 * i + 1];
    uint32_t f = 0;
    if (cell < g.cells && !(i >= skip_first && i - skip_first < skip_count)) {
        const uint32_t col = cell % g.cols;
        f = (col + g.halo >= g.col_lo) && (col < g.col_hi + g.halo);
    }
    flag[i] = f;
}

// appends the kept records after the `base` own entities and publishes the build length (own + halo) on the device:
// the host never needs the halo count, so the whole multi-GPU tick is free of host round trips.
__global__ void halo_append_kernel(const uint32_t* __restrict__ rec, uint32_t n, const uint32_t* __restrict__ flag,
                                   const uint32_t* __restrict__ off, uint32_t base, uint32_t cap_total, uint32_t* __restrict__ key,
                                   uint32_t* __restrict__ gid, uint32_t* __restrict__ n_build, Counters* __restrict__ ctr) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        const uint64_t total = (uint64_t)base + off[n];
        if (total > cap_total) atomicOr(&ctr->overflow, (uint32_t)CHD_OVF_BORDER);
        *n_build = (uint32_t)min(total, (uint64_t)cap_total);
    }
    if (i >= n || !flag[i]) return;
    const uint32_t o = base + off[i];
    if (o >= cap_total) return;
    gid[o] = rec[2 * i];
    key[o] = rec[2 * i + 1];
}

}  // namespace chd
