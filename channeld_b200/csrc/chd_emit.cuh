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
#include "chd_interest.cuh"

namespace chd {

constexpr int EMIT_THREADS = 256;
constexpr int EMIT_WARPS = EMIT_THREADS / 32;
constexpr int EMIT_CHUNKS = 8;                  // 16-byte chunks per lane per tile
constexpr int EMIT_TILE = 32 * EMIT_CHUNKS * 4;  // 1024 entries = 4 KB of output per WARP tile
constexpr int EMIT_SMEM_PAIRS = 64;              // pairs per warp tile staged in shared memory

// The cell CSR's entity array is kept in FOUR phase-shifted copies: copy k stores element i at index
// k*stride + k + i (stride % 4 == 0), i.e. at 16-byte phase (k + i) % 4.  Output chunks are 16-byte aligned, so
// for a run that starts at source index s the copy k = (-s) & 3 makes source and destination co-aligned and
// the whole run moves as LDG.128 -> STG.128 with no realignment shuffles.  Cost: 12 extra bytes per entity
// written once per build (L2-resident), against 8 bytes per VISIBLE entry saved from 4-byte accesses.
__global__ void __launch_bounds__(256)
    replicate_phases_kernel(const uint32_t* src, uint32_t n, const uint32_t* __restrict__ n_ptr, uint32_t stride, uint32_t* dst4) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (n_ptr) n = min(n, *n_ptr);
    if (i >= n) return;
    const uint32_t v = src[i];
#pragma unroll
    for (uint32_t k = 1; k < 4; k++) dst4[(size_t)k * stride + k + i] = v;  // copy 0 is src itself (dst4 == src)
}

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
    if (bump_epoch && blockIdx.x == 0 && threadIdx.x == 0) *bump_epoch = (*bump_epoch + 1) & ((1ull << 22) - 1);
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

// ---------------------------------------------------------------------------------------------------------
// v5 (experimental, CHD_EMIT_VARIANT=5): cell-grouped work order.  v3 is bound by the L2 -> HBM write stream with every
// visible entry also READ from L2; here the copy units are ordered by CELL (the by_cell permutation of the pairs) so the
// warps of an SM copy the same cell's entity list over and over and the reads become L1 hits.  Work unit = up to
// EMIT_UNIT consecutive entries of one pair, described by a 16-byte descriptor written by the preparation pass.
// Measured (profiles/README.md): L1 hit rate 11 % -> 46 %, L2 traffic and instruction count down — and the kernel
// SLOWER (0.42-0.46 ms vs 0.355 ms), because each warp now sweeps its own distant region of the output: thousands of
// scattered write streams instead of v3's single contiguous sweep, and the write stream is the scarce resource.
constexpr int EMIT_UNIT = 1024;

struct EmitUnit {
    uint32_t src;  // first source entry (index into phase copy 0)
    uint32_t len;  // entries (1..EMIT_UNIT)
    uint64_t dst;  // first destination entry
};

// scan input functor: units of the pair at by-cell position i
struct UnitCountIn {
    const uint32_t* pair_cell;
    const uint32_t* by_cell;
    const uint32_t* cell_start;
    __device__ __forceinline__ uint64_t operator()(uint64_t i) const {
        const uint32_t c = pair_cell[by_cell[i]];
        return (uint64_t)((cell_start[c + 1] - cell_start[c] + EMIT_UNIT - 1) / EMIT_UNIT);
    }
};

// descriptors + vis_off + V bookkeeping (last kernel of the v5 emit preparation: also opens the next scan epoch)
__global__ void __launch_bounds__(256)
    emit_units_kernel(const uint32_t* __restrict__ n_pairs_ptr, uint64_t pair_cap, const uint64_t* __restrict__ voff,
                      const uint32_t* __restrict__ uoff, const uint32_t* __restrict__ by_cell, const uint32_t* __restrict__ pair_cell,
                      const uint32_t* __restrict__ cell_start, EmitUnit* __restrict__ units, uint64_t unit_cap, uint32_t n_slots,
                      const uint32_t* __restrict__ pair_off, uint64_t* __restrict__ vis_off, uint64_t vis_cap, Counters* __restrict__ ctr,
                      unsigned long long* bump_epoch) {
    const uint64_t n = min((uint64_t)*n_pairs_ptr, pair_cap);
    if (bump_epoch && blockIdx.x == 0 && threadIdx.x == 0) *bump_epoch = (*bump_epoch + 1) & ((1ull << 22) - 1);
    for (uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; s <= n_slots; s += (uint64_t)gridDim.x * blockDim.x) {
        vis_off[s] = voff[min((uint64_t)pair_off[s], n)];
        if (s == n_slots) {
            const uint64_t V = voff[n];
            ctr->n_visible = V;
            ctr->required_visible = V;
            if (V > vis_cap || uoff[n] > unit_cap) atomicOr(&ctr->overflow, (uint32_t)CHD_OVF_VISIBLE);
        }
    }
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t p = by_cell[i];
        const uint32_t c = pair_cell[p];
        const uint32_t cs = cell_start[c], len = cell_start[c + 1] - cs;
        const uint64_t d0 = voff[p];
        const uint32_t u0 = uoff[i], nu = uoff[i + 1] - u0;
        for (uint32_t k = 0; k < nu; k++) {
            if ((uint64_t)u0 + k >= unit_cap) break;
            EmitUnit u;
            u.src = cs + k * EMIT_UNIT;
            u.len = min((uint32_t)EMIT_UNIT, len - k * EMIT_UNIT);
            u.dst = d0 + (uint64_t)k * EMIT_UNIT;
            units[u0 + k] = u;
        }
    }
}

__global__ void __launch_bounds__(EMIT_THREADS, 4)
    emit_visible_v5_kernel(const uint32_t* __restrict__ n_pairs_ptr, uint64_t pair_cap, const uint64_t* __restrict__ voff,
                           const uint32_t* __restrict__ uoff, const EmitUnit* __restrict__ units, uint64_t unit_cap,
                           const uint32_t* __restrict__ sorted4, uint32_t stride, uint32_t* __restrict__ vis_entity, uint64_t vis_cap,
                           uint32_t sm_count) {
    const uint64_t np = min((uint64_t)*n_pairs_ptr, pair_cap);
    if (np == 0) return;
    if (voff[np] > vis_cap) return;
    const uint32_t U = uoff[np];
    if (U > unit_cap) return;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    // blocks b, b+sm_count, b+2*sm_count, ... are typically co-resident on one SM: give them adjacent slices
    const uint32_t per_sm = gridDim.x / sm_count ? gridDim.x / sm_count : 1;
    const uint32_t lb = (blockIdx.x % sm_count) * per_sm + (blockIdx.x / sm_count);
    const uint64_t n_warps = (uint64_t)gridDim.x * EMIT_WARPS, wid = (uint64_t)(lb < gridDim.x ? lb : blockIdx.x) * EMIT_WARPS + w;
    const uint32_t u0 = (uint32_t)((uint64_t)U * wid / n_warps), u1 = (uint32_t)((uint64_t)U * (wid + 1) / n_warps);
    if (u0 >= u1) return;
    const uint4* __restrict__ desc = reinterpret_cast<const uint4*>(units);
    uint4 nd = __ldg(desc + u0);  // software pipeline: the next descriptor is always in flight
    for (uint32_t u = u0; u < u1; u++) {
        const uint4 cd = nd;
        if (u + 1 < u1) nd = __ldg(desc + u + 1);
        uint32_t s = cd.x, m = cd.y;
        uint64_t d = (uint64_t)cd.z | ((uint64_t)cd.w << 32);
        const uint32_t head = min(m, (uint32_t)((4 - (d & 3)) & 3));  // entries up to the next 16-byte boundary
        if (lane < head) vis_entity[d + lane] = __ldg(sorted4 + s + lane);
        s += head; d += head; m -= head;
        const uint32_t ph = (0u - s) & 3u;  // d is 16-byte aligned now: pick the co-aligned phase copy
        const uint4* __restrict__ src = reinterpret_cast<const uint4*>(sorted4 + (size_t)ph * stride + ph + s);
        uint4* __restrict__ dst = reinterpret_cast<uint4*>(vis_entity + d);
        const uint32_t nv = m >> 2;
        uint4 v[EMIT_UNIT / 128];
#pragma unroll
        for (int it = 0; it < EMIT_UNIT / 128; it++)
            if (it * 32 + lane < nv) v[it] = __ldg(src + it * 32 + lane);
#pragma unroll
        for (int it = 0; it < EMIT_UNIT / 128; it++)
            if (it * 32 + lane < nv) __stcs(dst + it * 32 + lane, v[it]);
        const uint32_t tail = m & 3u;
        if (lane < tail) vis_entity[d + 4 * nv + lane] = __ldg(sorted4 + s + 4 * nv + lane);
    }
}

}  // namespace chd
