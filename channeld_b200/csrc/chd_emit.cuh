// chd_emit.cuh — expanded per-subscriber visible-entity lists: for every subscription pair
// (subscriber, cell), in pair order, copy the cell's entity list out of the cell CSR.  This is the dominant
// HBM term of a tick (8V bytes algorithmic: read 4V + write 4V; V ~ 4.9e8 on the benchmark config).
//
// Decomposition: the OUTPUT array is cut into fixed tiles of EMIT_TILE entries (load-balanced regardless of
// how entities are distributed over cells).  A partition pass records the first pair of each tile; a
// persistent grid (multiple of the SM count) then walks tiles round-robin.  Stores are fully coalesced and
// 128 B-aligned (tile bases are multiples of 4096 entries); loads are coalesced runs out of the L2-resident
// cell CSR (sorted_entity is 4 B x N: 4 MB at 1 M entities, far below the 126 MB L2).
#pragma once
#include "chd_interest.cuh"

namespace chd {

constexpr int EMIT_THREADS = 256;
constexpr int EMIT_ITEMS = 16;
constexpr int EMIT_TILE = EMIT_THREADS * EMIT_ITEMS;  // 4096 entries = 16 KB of output per tile
constexpr int EMIT_SMEM_PAIRS = 1024;

// per pair: number of visible entities = size of the cell's list
__global__ void __launch_bounds__(256)
    pair_vcount_kernel(const uint32_t* __restrict__ n_pairs_ptr, uint64_t pair_cap, const uint32_t* __restrict__ pair_cell,
                       const uint32_t* __restrict__ cell_start, uint32_t* __restrict__ vcnt) {
    const uint64_t n = min((uint64_t)*n_pairs_ptr, pair_cap);
    for (uint64_t p = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; p < n; p += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t c = pair_cell[p];
        vcnt[p] = cell_start[c + 1] - cell_start[c];
    }
}

// vis_off[s] = voff[pair_off[s]]; also records V and the overflow flag
__global__ void __launch_bounds__(256)
    vis_off_kernel(uint32_t n_slots, const uint32_t* __restrict__ pair_off, uint64_t pair_cap, const uint64_t* __restrict__ voff,
                   uint64_t* __restrict__ vis_off, uint64_t vis_cap, Counters* __restrict__ ctr) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s > n_slots) return;
    const uint64_t np = min((uint64_t)pair_off[n_slots], pair_cap);
    const uint64_t p = min((uint64_t)pair_off[s], np);
    vis_off[s] = voff[p];
    if (s == n_slots) {
        const uint64_t V = voff[np];
        ctr->n_visible = V;
        ctr->required_visible = V;
        if (V > vis_cap) atomicOr(&ctr->overflow, (uint32_t)CHD_OVF_VISIBLE);
    }
}

// first_pair[t] = the pair whose output interval [voff[p], voff[p+1]) contains entry t*EMIT_TILE
__global__ void __launch_bounds__(256)
    emit_partition_kernel(const uint32_t* __restrict__ n_pairs_ptr, uint64_t pair_cap, const uint64_t* __restrict__ voff,
                          uint32_t* __restrict__ first_pair, uint64_t max_tiles) {
    const uint64_t n = min((uint64_t)*n_pairs_ptr, pair_cap);
    for (uint64_t p = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; p < n; p += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t b = voff[p], e = voff[p + 1];
        if (e == b) continue;
        for (uint64_t t = (b + EMIT_TILE - 1) / EMIT_TILE; t * EMIT_TILE < e && t < max_tiles; t++) first_pair[t] = (uint32_t)p;
    }
}

__global__ void __launch_bounds__(EMIT_THREADS)
    emit_visible_kernel(const uint32_t* __restrict__ n_pairs_ptr, uint64_t pair_cap, const uint64_t* __restrict__ voff,
                        const uint32_t* __restrict__ pair_cell, const uint32_t* __restrict__ cell_start,
                        const uint32_t* __restrict__ sorted_entity, const uint32_t* __restrict__ first_pair,
                        uint32_t* __restrict__ vis_entity, uint64_t vis_cap) {
    __shared__ uint32_t s_end[EMIT_SMEM_PAIRS];  // end of pair (p0+k) relative to the tile base, clamped to EMIT_TILE
    __shared__ uint32_t s_src[EMIT_SMEM_PAIRS];  // source index of the pair's entry that lands on max(voff[p], tile base)
    const uint64_t np = min((uint64_t)*n_pairs_ptr, pair_cap);
    if (np == 0) return;
    const uint64_t V = voff[np];
    if (V > vis_cap) return;
    const uint64_t n_tiles = (V + EMIT_TILE - 1) / EMIT_TILE;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (uint64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const uint64_t base = t * EMIT_TILE;
        const uint32_t p0 = first_pair[t];
        const uint32_t p1 = (t + 1 < n_tiles) ? first_pair[t + 1] : (uint32_t)(np - 1);
        const uint32_t cnt = p1 - p0 + 1;
        const uint32_t tile_len = (uint32_t)min((uint64_t)EMIT_TILE, V - base);
        if (cnt <= EMIT_SMEM_PAIRS) {
            __syncthreads();  // previous tile's readers are done
            for (uint32_t k = threadIdx.x; k < cnt; k += EMIT_THREADS) {
                const uint64_t b = voff[p0 + k], e = voff[p0 + k + 1];
                const uint32_t c = pair_cell[p0 + k];
                s_end[k] = (uint32_t)min((uint64_t)EMIT_TILE, e > base ? e - base : 0);
                // source position for relative output position max(b,base)-base
                s_src[k] = cell_start[c] + (uint32_t)(b < base ? base - b : 0);
            }
            __syncthreads();
            // warp w covers relative outputs [w*512, w*512+512)
            const uint32_t seg = w * (32 * EMIT_ITEMS);
            // uniform binary search: first k with s_end[k] > seg
            uint32_t lo = 0, hi = cnt;
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (s_end[mid] > seg) hi = mid; else lo = mid + 1;
            }
            uint32_t k = lo;
#pragma unroll 4
            for (int it = 0; it < EMIT_ITEMS; it++) {
                const uint32_t o = seg + it * 32 + lane;
                if (o < tile_len) {
                    while (s_end[k] <= o) k++;  // o < tile_len guarantees termination (s_end[cnt-1] >= tile_len)
                    const uint32_t beg = k == 0 ? 0u : s_end[k - 1];  // relative start of pair k inside the tile
                    vis_entity[base + o] = sorted_entity[s_src[k] + (o - beg)];
                }
            }
        } else {
            // many tiny/empty cells inside one tile: per-lane binary search over the global offsets
            for (int it = 0; it < EMIT_ITEMS; it++) {
                const uint32_t o = threadIdx.x + it * EMIT_THREADS;
                if (o < tile_len) {
                    const uint64_t go = base + o;
                    uint64_t lo = p0, hi = p1;  // last p in [p0,p1] with voff[p] <= go
                    while (lo < hi) {
                        const uint64_t mid = (lo + hi + 1) >> 1;
                        if (voff[mid] <= go) lo = mid; else hi = mid - 1;
                    }
                    const uint32_t c = pair_cell[lo];
                    vis_entity[go] = sorted_entity[cell_start[c] + (uint32_t)(go - voff[lo])];
                }
            }
        }
    }
}

}  // namespace chd
