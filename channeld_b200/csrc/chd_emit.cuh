// chd_emit.cuh — expanded per-subscriber visible-entity lists: for every subscription pair
// (subscriber, cell), in pair order, copy the cell's entity list out of the cell CSR.  This is the dominant
// HBM term of a tick (8V bytes algorithmic: read 4V + write 4V; V ~ 4.9e8 on the benchmark config).
//
// Decomposition: the OUTPUT array is cut into fixed tiles of EMIT_TILE entries (load-balanced regardless of
// how entities are distributed over cells).  A partition pass records the first pair of each tile; a
// persistent grid (multiple of the SM count) then walks tiles round-robin, one tile per WARP.  Every lane moves
// 16-byte chunks: stores are fully coalesced and 128 B-aligned (tile bases are multiples of 1024 entries); loads are co-aligned
// 16-byte reads out of the L2-resident phase copies of the cell CSR (4 x 4 B x N: 16 MB at 1 M entities, far
// below the 126 MB L2).  v1 of this kernel moved 4 bytes per thread-iteration and was instruction-issue bound
// (ncu: 73 % issue-active, 30 % DRAM): see profiles/r1_v1_emit_ncu_details.txt.
#pragma once
#include "chd_types.cuh"

namespace chd {


// scan input functor: visible entries of pair p = size of its cell's list (computed on the fly by the offset scan)
struct PairVcountIn {
    const uint32_t* pair_cell;
    const uint32_t* cell_start;
    __device__ __forceinline__ uint64_t operator()(uint64_t p) const {
        const uint32_t c = pair_cell[p];
        return (uint64_t)(cell_start[c + 1] - cell_start[c]);
    }
};

// first_pair[t] = the pair whose output interval [voff[p], voff[p+1]) contains entry t*EMIT_TILE;
// also vis_off[s] = voff[pair_off[s]] and the V / overflow bookkeeping (one launch instead of two)
__global__ void __launch_bounds__(256)
    emit_partition_kernel(const uint32_t* __restrict__ n_pairs_ptr, uint64_t pair_cap, const uint64_t* __restrict__ voff,
                          uint32_t* __restrict__ first_pair, uint64_t max_tiles, uint32_t n_slots, const uint32_t* __restrict__ pair_off,
                          uint64_t* __restrict__ vis_off, uint64_t vis_cap, Counters* __restrict__ ctr, unsigned long long* bump_epoch) {
    const uint64_t n = min((uint64_t)*n_pairs_ptr, pair_cap);
    // last kernel of the emit preparation: open the NEXT execution's scan epoch (the scan of this one has completed)
    if (bump_epoch && blockIdx.x == 0 && threadIdx.x == 0) *bump_epoch = chd_next_epoch(*bump_epoch);
    for (uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; s <= n_slots; s += (uint64_t)gridDim.x * blockDim.x) {
        vis_off[s] = voff[min((uint64_t)pair_off[s], n)];
        if (s == n_slots) {
            const uint64_t V = voff[n];
            ctr->n_visible = V;
            ctr->required_visible = V;
            if (V > vis_cap) atomicOr(&ctr->overflow, (uint32_t)CHD_OVF_VISIBLE);
        }
    }
    for (uint64_t p = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; p < n; p += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t b = voff[p], e = voff[p + 1];
        if (e == b) continue;
        for (uint64_t t = (b + EMIT_TILE - 1) / EMIT_TILE; t * EMIT_TILE < e && t < max_tiles; t++) first_pair[t] = (uint32_t)p;
    }
}

// v3: every WARP owns whole tiles (no block barrier anywhere): the segment table of a tile lives in a warp-private
// slice of shared memory guarded by __syncwarp.  Output stores are streaming (st.global.cs) so the 1.95 GB write
// stream does not evict the L2-resident phase copies the loads come from.
__global__ void __launch_bounds__(EMIT_THREADS, 4)
    emit_visible_kernel(const uint32_t* __restrict__ n_pairs_ptr, uint64_t pair_cap, const uint64_t* __restrict__ voff,
                        const uint32_t* __restrict__ pair_cell, const uint32_t* __restrict__ cell_start,
                        const uint32_t* __restrict__ sorted4, uint32_t stride, const uint32_t* __restrict__ first_pair,
                        uint32_t* __restrict__ vis_entity, uint64_t vis_cap) {
    __shared__ uint32_t s_end_all[EMIT_WARPS][EMIT_SMEM_PAIRS];  // end of pair (p0+k) relative to the tile base, clamped to EMIT_TILE
    __shared__ uint32_t s_src_all[EMIT_WARPS][EMIT_SMEM_PAIRS];  // source index of the pair's entry that lands on max(voff[p], tile base)
    const uint64_t np = min((uint64_t)*n_pairs_ptr, pair_cap);
    if (np == 0) return;
    const uint64_t V = voff[np];
    if (V > vis_cap) return;
    const uint64_t n_tiles = (V + EMIT_TILE - 1) / EMIT_TILE;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t* s_end = s_end_all[w];
    uint32_t* s_src = s_src_all[w];
    const uint64_t warp_id = (uint64_t)blockIdx.x * EMIT_WARPS + w, n_warps = (uint64_t)gridDim.x * EMIT_WARPS;
    for (uint64_t t = warp_id; t < n_tiles; t += n_warps) {
        const uint64_t base = t * EMIT_TILE;
        const uint32_t p0 = first_pair[t];
        const uint32_t p1 = (t + 1 < n_tiles) ? first_pair[t + 1] : (uint32_t)(np - 1);
        const uint32_t cnt = p1 - p0 + 1;
        const uint32_t tile_len = (uint32_t)min((uint64_t)EMIT_TILE, V - base);
        uint32_t* __restrict__ out = vis_entity + base;
        if (cnt <= EMIT_SMEM_PAIRS) {
            __syncwarp();  // the previous tile's readers are done
            for (uint32_t k = lane; k < cnt; k += 32) {
                const uint64_t b = voff[p0 + k], e = voff[p0 + k + 1];
                const uint32_t c = pair_cell[p0 + k];
                s_end[k] = (uint32_t)min((uint64_t)EMIT_TILE, e > base ? e - base : 0);
                s_src[k] = cell_start[c] + (uint32_t)(b < base ? base - b : 0);
            }
            __syncwarp();
            uint32_t k = 0;  // segment cursor of this lane (its chunks ascend)
            uint32_t src[EMIT_CHUNKS];  // element index into sorted4 (multiple of 4), or 0xFFFFFFFF
            // phase 1: resolve every chunk's source (nullptr = handled element-wise / out of range)
#pragma unroll
            for (int it = 0; it < EMIT_CHUNKS; it++) {
                const uint32_t o = (it * 32 + lane) * 4;  // first entry of this 16-byte chunk
                src[it] = 0xFFFFFFFFu;
                if (o < tile_len) {
                    while (s_end[k] <= o) k++;  // terminates: s_end[cnt-1] >= tile_len > o
                    const uint32_t beg = k == 0 ? 0u : s_end[k - 1];
                    const uint32_t sidx = s_src[k] + (o - beg);
                    if (o + 4 <= s_end[k]) {
                        // whole chunk inside one segment: co-aligned 16-byte move out of phase copy (-sidx)&3
                        const uint32_t ph = (0u - sidx) & 3u;
                        src[it] = ph * stride + ph + sidx;
                    } else {
                        // the chunk straddles a segment boundary (or the end of the list): entry by entry
                        uint32_t kk = k;
                        for (uint32_t j = 0; j < 4 && o + j < tile_len; j++) {
                            while (s_end[kk] <= o + j) kk++;
                            const uint32_t bb = kk == 0 ? 0u : s_end[kk - 1];
                            out[o + j] = __ldg(sorted4 + s_src[kk] + (o + j - bb));
                        }
                    }
                }
            }
            // phase 2: all loads in flight, phase 3: streaming stores
            uint4 v[EMIT_CHUNKS];
#pragma unroll
            for (int it = 0; it < EMIT_CHUNKS; it++)
                if (src[it] != 0xFFFFFFFFu) v[it] = __ldg(reinterpret_cast<const uint4*>(sorted4 + src[it]));
#pragma unroll
            for (int it = 0; it < EMIT_CHUNKS; it++)
                if (src[it] != 0xFFFFFFFFu) __stcs(reinterpret_cast<uint4*>(out + (it * 32 + lane) * 4), v[it]);
        } else {
            // more than EMIT_SMEM_PAIRS pairs inside one tile (tiny / empty cells): per-entry binary search
            for (uint32_t o = lane; o < tile_len; o += 32) {
                const uint64_t go = base + o;
                uint64_t lo = p0, hi = p1;  // last p in [p0,p1] with voff[p] <= go
                while (lo < hi) {
                    const uint64_t mid = (lo + hi + 1) >> 1;
                    if (voff[mid] <= go) lo = mid; else hi = mid - 1;
                }
                const uint32_t c = pair_cell[lo];
                vis_entity[go] = sorted4[cell_start[c] + (uint32_t)(go - voff[lo])];
            }
        }
    }
}

}  // namespace chd
