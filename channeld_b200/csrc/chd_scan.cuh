// chd_scan.cuh — device-wide exclusive prefix sums (u32 counts -> u32/u64 offsets), reduce-then-scan.
// out has n+1 entries (out[n] = total).  Deterministic.  Scratch comes from the caller's arena.
// exclusive_scan returns the number of kernels it launched.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "chd_device.cuh"

namespace chd {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;  // 4096

template <typename T>
__device__ __forceinline__ T warp_incl_scan(T v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = __shfl_up_sync(0xffffffffu, v, d);
        if ((threadIdx.x & 31) >= d) v += o;
    }
    return v;
}

// block-wide exclusive scan of one value per thread; returns exclusive prefix, total in `total`
template <typename T>
__device__ __forceinline__ T block_excl_scan(T v, T& total) {
    __shared__ T warp_sums[SCAN_THREADS / 32];
    __shared__ T tot;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    T incl = warp_incl_scan(v);
    if (lane == 31) warp_sums[w] = incl;
    __syncthreads();
    if (w == 0) {
        T s = lane < SCAN_THREADS / 32 ? warp_sums[lane] : T(0);
        T si = warp_incl_scan(s);
        if (lane < SCAN_THREADS / 32) warp_sums[lane] = si - s;
        if (lane == SCAN_THREADS / 32 - 1) tot = si;
    }
    __syncthreads();
    T res = incl - v + warp_sums[w];
    total = tot;
    __syncthreads();
    return res;
}

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(SCAN_THREADS)
    scan_tile_sums(const TIn* __restrict__ in, TOut* __restrict__ sums, uint64_t n, const uint32_t* __restrict__ n_ptr) {
    if (n_ptr) n = min(n, (uint64_t)*n_ptr);
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE;
    TOut acc = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        const uint64_t i = base + (uint64_t)k * SCAN_THREADS + threadIdx.x;
        if (i < n) acc += (TOut)in[i];
    }
    TOut total;
    block_excl_scan<TOut>(acc, total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// Scans tile `blockIdx.x`; adds tile_base[blockIdx.x] (exclusive tile offsets) when given.
// The block that owns element n-1 also writes out[n] = total.
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(SCAN_THREADS)
    scan_tiles(const TIn* __restrict__ in, TOut* __restrict__ out, uint64_t n, const TOut* __restrict__ tile_base,
               const uint32_t* __restrict__ n_ptr) {
    if (n_ptr) n = min(n, (uint64_t)*n_ptr);
    if ((uint64_t)blockIdx.x * SCAN_TILE >= n && !(n == 0 && blockIdx.x == 0)) return;
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
    TOut v[SCAN_ITEMS];
    TOut acc = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        const uint64_t i = base + k;
        v[k] = i < n ? (TOut)in[i] : TOut(0);
        acc += v[k];
    }
    TOut total;
    TOut pre = block_excl_scan<TOut>(acc, total);
    if (tile_base) pre += tile_base[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        const uint64_t i = base + k;
        if (i < n) out[i] = pre;
        pre += v[k];
        if (i + 1 == n) out[n] = pre;
    }
    if (n == 0 && blockIdx.x == 0 && threadIdx.x == 0) out[0] = 0;
}

// scratch must hold scan_scratch_elems<TOut>(n) elements of TOut
inline uint64_t scan_scratch_elems(uint64_t n) {
    uint64_t total = 0;
    uint64_t t = (n + SCAN_TILE - 1) / SCAN_TILE;
    while (t > 1) {
        total += t + 1;
        t = (t + SCAN_TILE - 1) / SCAN_TILE;
    }
    return total + 2;
}

template <typename TIn, typename TOut>
inline int exclusive_scan(const TIn* in, TOut* out, uint64_t n, TOut* scratch, cudaStream_t st,
                          const uint32_t* n_ptr = nullptr) {
    // n is the host-known upper bound; when n_ptr is given the live length is min(n, *n_ptr) (device side),
    // so no host sync is needed to size the launch.  Tiles beyond the live length contribute zeros.
    const uint64_t tiles = n == 0 ? 1 : (n + SCAN_TILE - 1) / SCAN_TILE;
    if (tiles == 1) {
        scan_tiles<TIn, TOut><<<1, SCAN_THREADS, 0, st>>>(in, out, n, nullptr, n_ptr);
        return 1;
    }
    TOut* sums = scratch;             // [tiles+1]: after the recursive scan, exclusive tile offsets
    TOut* next = scratch + tiles + 1;
    scan_tile_sums<TIn, TOut><<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(in, sums, n, n_ptr);
    const int inner = exclusive_scan<TOut, TOut>(sums, sums, tiles, next, st);  // in-place is safe: items are read before written
    scan_tiles<TIn, TOut><<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(in, out, n, sums, n_ptr);
    return inner + 2;  // kernels launched
}

// ---------------------------------------------------------------------------------------------------------
// Single-pass scan (decoupled look-back): ONE kernel per prefix sum instead of three.  Each scan call site owns
// a ScanSite: tile descriptors in device memory plus a pointer to its stage's epoch counter, which an earlier
// kernel of the same stage increments once per stage execution (bump_epoch_kernel / stage_begin_kernel).  A launch
// therefore carries no per-call host state and can be replayed from a CUDA graph.  A descriptor is one 64-bit
// word [epoch:22 | flag:2 | value:40]; descriptors written by an earlier stage execution carry an older epoch
// and read as "not ready"; a valid descriptor also needs a non-zero flag, and the engine clears every site's descriptors
// long before the 22-bit epoch can wrap around to a value a stale descriptor still carries (chd_epoch_tick, chd_engine.cu).  Tile id = blockIdx.x (lower-indexed blocks of a 1-D grid are dispatched first, the
// usual forward-progress assumption of decoupled look-back).  No atomics.  Sums must stay below 2^40; a site may
// be used once per stage execution.
struct ScanSite {
    unsigned long long* desc;         // [tiles]
    const unsigned long long* epoch;  // the owning stage's epoch counter
    uint32_t* error;                  // bit 31 is set if a look-back ever times out (reported as an overflow bit)
    uint64_t tiles;                   // descriptor capacity
    int stage;                        // index of the owning stage epoch (host bookkeeping)
};

static __global__ void bump_epoch_kernel(unsigned long long* epoch) { *epoch = chd_next_epoch(*epoch); }

constexpr unsigned long long SCAN_FLAG_AGG = 1ull, SCAN_FLAG_PREFIX = 2ull;
__device__ __forceinline__ unsigned long long scan_pack(unsigned long long epoch, unsigned long long flag, unsigned long long v) {
    return (epoch << 42) | (flag << 40) | (v & ((1ull << 40) - 1));
}

// scan input: either a plain array or a functor computing element i on the fly (saves the kernel that would have
// materialised the counts)
template <typename T>
struct ScanPtrIn {
    const T* p;
    __device__ __forceinline__ uint64_t operator()(uint64_t i) const { return (uint64_t)p[i]; }
};

// The look-back itself (first warp of the block): publishes this tile's aggregate, sums the predecessors' and publishes the
// inclusive prefix; every thread gets the tile's exclusive prefix back (one __syncthreads inside).
__device__ __forceinline__ unsigned long long scan_lookback_prefix(const ScanSite& site, uint64_t tile, unsigned long long epoch, unsigned long long total) {
    __shared__ unsigned long long s_prefix;
    volatile unsigned long long* desc = site.desc;
    if (threadIdx.x < 32) {
        // warp-parallel look-back: 32 predecessor descriptors per hop
        const int lane = threadIdx.x;
        if (tile == 0) {
            if (lane == 0) {
                s_prefix = 0;
                desc[0] = scan_pack(epoch, SCAN_FLAG_PREFIX, total);
            }
        } else {
            if (lane == 0) desc[tile] = scan_pack(epoch, SCAN_FLAG_AGG, total);
            unsigned long long run = 0;
            int64_t start = (int64_t)tile - 1;
            uint32_t spins = 0;
            for (;;) {
                const int64_t j = start - lane;
                // tiles before 0 act as an (always valid) zero prefix
                const unsigned long long d = j >= 0 ? desc[j] : scan_pack(epoch, SCAN_FLAG_PREFIX, 0ull);
                const bool valid = (d >> 42) == epoch && ((d >> 40) & 3ull) != 0;
                const bool is_prefix = valid && ((d >> 40) & 3ull) == SCAN_FLAG_PREFIX;
                const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
                const uint32_t pmask = __ballot_sync(0xffffffffu, is_prefix);
                // lanes 0..fp are needed (fp = nearest published prefix), or all 32 when none is visible yet
                const int fp = pmask ? __ffs(pmask) - 1 : 31;
                const uint32_t need = fp == 31 ? 0xffffffffu : ((2u << fp) - 1u);
                if ((vmask & need) != need) {  // a needed predecessor has not published yet: re-read
                    // bounded: a scheduling pathology must surface as an error, never as a hung GPU
                    if (++spins > (1u << 24)) {
                        if (lane == 0 && site.error) atomicOr(site.error, 0x80000000u);
                        break;
                    }
                    continue;
                }
                unsigned long long x = (lane <= fp) ? (d & ((1ull << 40) - 1)) : 0ull;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
                run += x;
                if (pmask) break;
                start -= 32;
            }
            if (lane == 0) {
                s_prefix = run;
                desc[tile] = scan_pack(epoch, SCAN_FLAG_PREFIX, run + total);
            }
        }
    }
    __syncthreads();
    return s_prefix;
}

template <typename InFn, typename TOut>
__global__ void __launch_bounds__(SCAN_THREADS)
    scan_lookback_kernel(InFn in, TOut* __restrict__ out, uint64_t n, const uint32_t* __restrict__ n_ptr, ScanSite site) {
    if (n_ptr) n = min(n, (uint64_t)*n_ptr);
    const uint64_t tile = blockIdx.x;
    if (!(tile * SCAN_TILE < n || (n == 0 && tile == 0))) return;  // beyond the live length: nothing to publish
    const unsigned long long epoch = *site.epoch;
    const uint64_t base = tile * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
    TOut v[SCAN_ITEMS];
    TOut acc = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        const uint64_t i = base + k;
        v[k] = i < n ? (TOut)in(i) : TOut(0);
        acc += v[k];
    }
    TOut total;
    TOut pre = block_excl_scan<TOut>(acc, total);
    pre += (TOut)scan_lookback_prefix(site, tile, epoch, (unsigned long long)total);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        const uint64_t i = base + k;
        if (i < n) out[i] = pre;
        pre += v[k];
        if (i + 1 == n) out[n] = pre;
    }
    if (n == 0 && threadIdx.x == 0) out[0] = 0;
}

template <typename TIn, typename TOut>
inline int exclusive_scan_1p(const TIn* in, TOut* out, uint64_t n, const ScanSite& site, cudaStream_t st, const uint32_t* n_ptr = nullptr) {
    const uint64_t tiles = n == 0 ? 1 : (n + SCAN_TILE - 1) / SCAN_TILE;  // <= site.tiles by construction of the site
    scan_lookback_kernel<ScanPtrIn<TIn>, TOut><<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(ScanPtrIn<TIn>{in}, out, n, n_ptr, site);
    return 1;
}

template <typename InFn, typename TOut>
inline int exclusive_scan_fn(InFn in, TOut* out, uint64_t n, const ScanSite& site, cudaStream_t st, const uint32_t* n_ptr = nullptr) {
    const uint64_t tiles = n == 0 ? 1 : (n + SCAN_TILE - 1) / SCAN_TILE;
    scan_lookback_kernel<InFn, TOut><<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(in, out, n, n_ptr, site);
    return 1;
}

// Single-pass stream compaction on the same machinery: flag(i) in {0,1} is computed on the fly, sink(i, k) is called for every
// flagged element with its rank k among the flagged ones (input order: deterministic), sink.total(count) once.  Replaces the
// flag kernel + prefix sum + scatter kernel triple of a classic compaction by ONE launch.
template <typename FlagFn, typename SinkFn>
__global__ void __launch_bounds__(SCAN_THREADS)
    compact_lookback_kernel(FlagFn flag, SinkFn sink, uint64_t n, const uint32_t* __restrict__ n_ptr, ScanSite site) {
    if (n_ptr) n = min(n, (uint64_t)*n_ptr);
    const uint64_t tile = blockIdx.x;
    if (!(tile * SCAN_TILE < n || (n == 0 && tile == 0))) return;
    const unsigned long long epoch = *site.epoch;
    const uint64_t base = tile * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t bits = 0, acc = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        const uint64_t i = base + k;
        const uint32_t f = (i < n && flag(i)) ? 1u : 0u;
        bits |= f << k;
        acc += f;
    }
    uint32_t total;
    uint32_t pre = block_excl_scan<uint32_t>(acc, total);
    pre += (uint32_t)scan_lookback_prefix(site, tile, epoch, (unsigned long long)total);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        const uint64_t i = base + k;
        if ((bits >> k) & 1u) sink(i, pre++);
        if (i + 1 == n) sink.total(pre);
    }
    if (n == 0 && threadIdx.x == 0) sink.total(0u);
}

template <typename FlagFn, typename SinkFn>
inline int compact_1p(FlagFn flag, SinkFn sink, uint64_t n, const ScanSite& site, cudaStream_t st, const uint32_t* n_ptr = nullptr) {
    const uint64_t tiles = n == 0 ? 1 : (n + SCAN_TILE - 1) / SCAN_TILE;  // <= site.tiles by construction of the site
    compact_lookback_kernel<FlagFn, SinkFn><<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(flag, sink, n, n_ptr, site);
    return 1;
}

}  // namespace chd
