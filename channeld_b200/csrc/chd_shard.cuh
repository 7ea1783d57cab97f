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

// Gathered records: rank r's `per_rank` records start at word r * stride_words (the rest of a rank's contribution is its
// migration blob); with stride_words == 2 * per_rank the records are simply contiguous.
struct RecView {
    const uint32_t* base;
    uint32_t per_rank;
    uint64_t stride_words;
    __device__ __forceinline__ const uint32_t* at(uint32_t i) const {
        return base + (uint64_t)(i / per_rank) * stride_words + 2ull * (i % per_rank);
    }
};

// keep gathered records whose column lies in this rank's extended range and which another rank exported
__global__ void halo_flag_kernel(GridDev g, RecView rec, uint32_t n, uint32_t skip_first, uint32_t skip_count,
                                 uint32_t* __restrict__ flag, unsigned long long* bump_epoch) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *bump_epoch = chd_next_epoch(*bump_epoch);  // border stage epoch
    if (i >= n) return;
    const uint32_t cell = rec.at(i)[1];
    uint32_t f = 0;
    if (cell < g.cells && !(i >= skip_first && i - skip_first < skip_count)) {
        const uint32_t col = cell % g.cols;
        f = (col + g.halo >= g.col_lo) && (col < g.col_hi + g.halo);
    }
    flag[i] = f;
}

// appends the kept records after the `base` own entities and publishes the build length (own + halo) on the device:
// the host never needs the halo count, so the whole multi-GPU tick is free of host round trips.
__global__ void halo_append_kernel(RecView rec, uint32_t n, const uint32_t* __restrict__ flag,
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
    const uint32_t* r = rec.at(i);
    gid[o] = r[0];
    key[o] = r[1];
}

// chd_get_rehome: own entities whose column now belongs to another rank (a set: block-aggregated append)
__global__ void __launch_bounds__(256) rehome_kernel(GridDev g, const uint32_t* __restrict__ key, const uint32_t* __restrict__ gid, uint32_t n,
                                                     uint32_t world, uint32_t* __restrict__ out_gid, uint32_t* __restrict__ out_rank,
                                                     uint32_t cap, uint32_t* __restrict__ count) {
    __shared__ uint32_t s_cnt, s_base;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    bool away = false;
    uint32_t col = 0, my = 0;
    if (i < n && key[i] < g.cells) {
        col = key[i] % g.cols;
        away = col < g.col_lo || col >= g.col_hi;
    }
    if (away) my = atomicAdd(&s_cnt, 1u);
    __syncthreads();
    if (threadIdx.x == 0 && s_cnt) s_base = atomicAdd(count, s_cnt);
    __syncthreads();
    if (away && s_base + my < cap) {
        out_gid[s_base + my] = gid ? gid[i] : i;
        // owner of a column = the largest rank r with floor(r * cols / world) <= col
        out_rank[s_base + my] = min((uint32_t)((((uint64_t)col + 1) * world - 1) / g.cols), world - 1);
    }
}

__global__ void __launch_bounds__(256) slot_ctl_mark_kernel(const uint32_t* __restrict__ slot, uint32_t n, uint8_t ctl, uint8_t* __restrict__ slot_ctl) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) slot_ctl[slot[i]] = ctl;
}
// chd_migrate_in: slot[i] becomes connection conn[i]; its previous run is record (first_index + i) of src_rank's blob
__global__ void __launch_bounds__(256) slot_import_mark_kernel(const uint32_t* __restrict__ slot, const uint32_t* __restrict__ conn, uint32_t n,
                                                               uint32_t src_rank, uint32_t first_index, uint8_t* __restrict__ slot_ctl,
                                                               uint32_t* __restrict__ slot_src, uint32_t* __restrict__ conn_id) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = slot[i];
    slot_ctl[s] = SLOT_IMPORT;
    slot_src[s] = (src_rank << 20) | ((first_index + i) & 0xFFFFFu);
    conn_id[s] = conn[i];
}

// chd_migrate_out: packs the subscriptions + fan-out state of the listed slots into this rank's migration blob (MigView layout).
// One block: the lists are control-plane sized (a per-cent of the subscribers cross a slab boundary per tick).
__global__ void __launch_bounds__(256) mig_pack_kernel(const uint32_t* __restrict__ slot, uint32_t n, PairBuf pb, const uint32_t* __restrict__ conn_id,
                                                       uint32_t* __restrict__ blob, MigView lay, Counters* __restrict__ ctr) {
    __shared__ uint32_t s_warp[8], s_run;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    uint32_t* off = blob + lay.o_off();
    uint32_t* oconn = blob + lay.o_conn();
    uint32_t *ocell = blob + lay.o_cell(), *odist = blob + lay.o_dist(), *oiv = blob + lay.o_interval(), *ofl = blob + lay.o_flags();
    int64_t* olast = reinterpret_cast<int64_t*>(blob + lay.o_last());
    uint64_t* olidx = reinterpret_cast<uint64_t*>(blob + lay.o_lidx());
    const uint32_t n_fit = min(n, lay.max_subs);
    for (uint32_t base = 0; base < n_fit; base += blockDim.x) {
        const uint32_t i = base + threadIdx.x;
        uint32_t s = 0, p0 = 0, k = 0;
        if (i < n_fit) {
            s = slot[i];
            p0 = pb.off[s];
            k = pb.off[s + 1] - p0;
        }
        uint32_t incl = k;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t a = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += a;
        }
        if (lane == 31) s_warp[w] = incl;
        __syncthreads();
        uint32_t o = s_run + incl - k;
        for (int j = 0; j < w; j++) o += s_warp[j];
        if (i < n_fit) {
            off[i] = o;
            oconn[i] = conn_id[s];
            for (uint32_t j = 0; j < k && o + j < lay.max_pairs; j++) {
                ocell[o + j] = pb.cell[p0 + j]; odist[o + j] = pb.dist[p0 + j]; oiv[o + j] = pb.interval[p0 + j];
                ofl[o + j] = pb.flags[p0 + j]; olast[o + j] = pb.last[p0 + j]; olidx[o + j] = pb.last_index[p0 + j];
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t tot = 0;
            for (int j = 0; j < 8; j++) tot += s_warp[j];
            s_run += tot;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        off[n_fit] = s_run;
        blob[0] = n_fit;
        blob[1] = min(s_run, lay.max_pairs);
        if (n > lay.max_subs || s_run > lay.max_pairs) atomicOr(&ctr->overflow, (uint32_t)CHD_OVF_BORDER);  // truncated: those arrive without state
    }
}

}  // namespace chd
