// emit_copy_probe.cu — standalone microbenchmark for the design question DESIGN.md §9 leaves open: the emit kernel
// (channeld_b200/csrc/chd_emit.cuh) writes V x 4 B of visible-entity ids per tick, every byte read from an L2-resident
// 16 MB source and written once; it is bound by L2 slice throughput (read + write = 8 B per entry), not by HBM.
// This probe measures, for the same traffic pattern (16 KB source segments out of a 16 MB pool -> one contiguous
// 1.9 GB output), four ways of moving the bytes:
//   A  ldg_stg      warp-owned 4 KB tiles, LDG.128 -> STG.128 streaming stores (what emit v3 does)
//   B  bulk         per-warp double-buffered cp.async.bulk global -> shared -> global (TMA engine moves the bytes, SMs idle)
//   C  stage_once   a CTA stages one 16 KB segment in shared memory once and bulk-stores it to R consecutive outputs
//                   (= the cell-major output order: L2 read traffic / R)
//   D  fill         write-only stream (the DRAM-side ceiling)
// Build + run on a B200:   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/emit_copy_probe.cu -o /tmp/probe && /tmp/probe
// NOT part of the product and not built by __graft_entry__.build(); compile-checked only so far (no GPU in the build
// container) — its numbers decide whether an ABI-visible re-ordering of the visible lists is worth it.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t r_ = (x);                                                                  \
        if (r_ != cudaSuccess) {                                                               \
            fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(r_)); \
            exit(1);                                                                           \
        }                                                                                      \
    } while (0)

constexpr uint32_t SEG_ENTRIES = 4096;               // 16 KB per source segment ("cell")
constexpr uint32_t TILE_ENTRIES = 1024;              // 4 KB per warp tile
constexpr uint32_t N_CELLS = 1024;                   // 16 MB source pool
constexpr uint32_t THREADS = 256, WARPS = THREADS / 32;

__host__ __device__ inline uint32_t cell_of_segment(uint64_t seg) {  // pseudo-random source cell of output segment `seg`
    uint64_t x = seg * 0x9E3779B97F4A7C15ull;
    x ^= x >> 29;
    x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 32;
    return (uint32_t)(x % N_CELLS);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(phase)
            : "memory");
    }
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- A: LDG.128 -> STG.128, warp-owned 4 KB tiles (the shape of emit v3)
__global__ void __launch_bounds__(THREADS, 4) ldg_stg_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, uint64_t n_tiles) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint64_t warp_id = (uint64_t)blockIdx.x * WARPS + w, n_warps = (uint64_t)gridDim.x * WARPS;
    for (uint64_t t = warp_id; t < n_tiles; t += n_warps) {
        const uint64_t seg = t / (SEG_ENTRIES / TILE_ENTRIES);
        const uint32_t within = (uint32_t)(t % (SEG_ENTRIES / TILE_ENTRIES)) * TILE_ENTRIES;
        const uint4* s4 = reinterpret_cast<const uint4*>(src + (uint64_t)cell_of_segment(seg) * SEG_ENTRIES + within);
        uint4* d4 = reinterpret_cast<uint4*>(dst + t * TILE_ENTRIES);
        uint4 v[8];
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = __ldg(s4 + i * 32 + lane);
#pragma unroll
        for (int i = 0; i < 8; i++) __stcs(d4 + i * 32 + lane, v[i]);
    }
}

// ---- B: per-warp double-buffered bulk copies global -> shared -> global (one lane drives the TMA engine)
constexpr int B_STAGES = 2;
__global__ void __launch_bounds__(THREADS) bulk_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, uint64_t n_tiles) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint32_t* buf = reinterpret_cast<uint32_t*>(smem_raw);                                    // [WARPS][B_STAGES][TILE_ENTRIES]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)WARPS * B_STAGES * TILE_ENTRIES * 4);  // [WARPS][B_STAGES]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0)
        for (int s = 0; s < B_STAGES; s++) mbar_init(&bars[w * B_STAGES + s], 1);
    __syncthreads();
    if (lane != 0) return;  // the copy engine does the work
    const uint64_t warp_id = (uint64_t)blockIdx.x * WARPS + w, n_warps = (uint64_t)gridDim.x * WARPS;
    uint32_t phase[B_STAGES] = {0, 0};
    int st = 0;
    for (uint64_t t = warp_id; t < n_tiles; t += n_warps) {
        uint32_t* b = buf + ((size_t)w * B_STAGES + st) * TILE_ENTRIES;
        uint64_t* bar = &bars[w * B_STAGES + st];
        // the store that last read this stage must have finished reading shared memory
        bulk_wait_read<B_STAGES - 1>();
        const uint64_t seg = t / (SEG_ENTRIES / TILE_ENTRIES);
        const uint32_t within = (uint32_t)(t % (SEG_ENTRIES / TILE_ENTRIES)) * TILE_ENTRIES;
        mbar_expect_tx(bar, TILE_ENTRIES * 4);
        bulk_g2s(b, src + (uint64_t)cell_of_segment(seg) * SEG_ENTRIES + within, TILE_ENTRIES * 4, bar);
        mbar_wait(bar, phase[st]);
        phase[st] ^= 1;
        bulk_s2g(dst + t * TILE_ENTRIES, b, TILE_ENTRIES * 4);
        bulk_commit();
        st = (st + 1) % B_STAGES;
    }
    bulk_wait_read<0>();
}

// ---- C: stage a 16 KB segment once per CTA, bulk-store it to `repeat` consecutive output segments
__global__ void __launch_bounds__(THREADS) stage_once_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, uint64_t n_groups,
                                                             uint32_t repeat) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint32_t* buf = reinterpret_cast<uint32_t*>(smem_raw);  // [2][SEG_ENTRIES]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + 2 * SEG_ENTRIES * 4);
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    uint32_t phase[2] = {0, 0};
    int st = 0;
    for (uint64_t gidx = blockIdx.x; gidx < n_groups; gidx += gridDim.x) {
        uint32_t* b = buf + (size_t)st * SEG_ENTRIES;
        bulk_wait_read<1>();  // stores of the group that used this stage two iterations ago are done with it
        mbar_expect_tx(&bars[st], SEG_ENTRIES * 4);
        bulk_g2s(b, src + (uint64_t)cell_of_segment(gidx) * SEG_ENTRIES, SEG_ENTRIES * 4, &bars[st]);
        mbar_wait(&bars[st], phase[st]);
        phase[st] ^= 1;
        for (uint32_t r = 0; r < repeat; r++) bulk_s2g(dst + (gidx * repeat + r) * SEG_ENTRIES, b, SEG_ENTRIES * 4);
        bulk_commit();
        st ^= 1;
    }
    bulk_wait_read<0>();
}


// ---- C2: like C, but the lane 0 of every warp issues a share of the bulk stores (is one issuing thread a bottleneck?)
__global__ void __launch_bounds__(THREADS) stage_once_mw_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, uint64_t n_groups,
                                                                uint32_t repeat) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint32_t* buf = reinterpret_cast<uint32_t*>(smem_raw);  // [2][SEG_ENTRIES]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + 2 * SEG_ENTRIES * 4);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
    }
    __syncthreads();
    uint32_t phase[2] = {0, 0};
    int st = 0;
    for (uint64_t gidx = blockIdx.x; gidx < n_groups; gidx += gridDim.x) {
        uint32_t* b = buf + (size_t)st * SEG_ENTRIES;
        if (lane == 0) bulk_wait_read<1>();  // this warp's stores out of this stage (two groups ago) have read shared memory
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_expect_tx(&bars[st], SEG_ENTRIES * 4);
            bulk_g2s(b, src + (uint64_t)cell_of_segment(gidx) * SEG_ENTRIES, SEG_ENTRIES * 4, &bars[st]);
        }
        mbar_wait(&bars[st], phase[st]);
        phase[st] ^= 1;
        if (lane == 0) {
            for (uint32_t r = w; r < repeat; r += WARPS) bulk_s2g(dst + (gidx * repeat + r) * SEG_ENTRIES, b, SEG_ENTRIES * 4);
            bulk_commit();
        }
        st ^= 1;
    }
    if (lane == 0) bulk_wait_read<0>();
}

// ---- E: stage once (bulk load), then the THREADS store the staged segment: LDS.128 -> STG.128 streaming stores, no L2 reads
__global__ void __launch_bounds__(THREADS) stage_once_stg_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, uint64_t n_groups,
                                                                 uint32_t repeat) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint32_t* buf = reinterpret_cast<uint32_t*>(smem_raw);  // [2][SEG_ENTRIES]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + 2 * SEG_ENTRIES * 4);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
    }
    __syncthreads();
    uint32_t phase[2] = {0, 0};
    int st = 0;
    uint64_t gidx = blockIdx.x;
    if (threadIdx.x == 0 && gidx < n_groups) {
        mbar_expect_tx(&bars[0], SEG_ENTRIES * 4);
        bulk_g2s(buf, src + (uint64_t)cell_of_segment(gidx) * SEG_ENTRIES, SEG_ENTRIES * 4, &bars[0]);
    }
    for (; gidx < n_groups; gidx += gridDim.x) {
        const uint64_t nxt = gidx + gridDim.x;
        __syncthreads();  // everybody is done reading the other stage
        if (threadIdx.x == 0 && nxt < n_groups) {
            mbar_expect_tx(&bars[st ^ 1], SEG_ENTRIES * 4);
            bulk_g2s(buf + (size_t)(st ^ 1) * SEG_ENTRIES, src + (uint64_t)cell_of_segment(nxt) * SEG_ENTRIES, SEG_ENTRIES * 4, &bars[st ^ 1]);
        }
        mbar_wait(&bars[st], phase[st]);
        phase[st] ^= 1;
        const uint4* s4 = reinterpret_cast<const uint4*>(buf + (size_t)st * SEG_ENTRIES);
        for (uint32_t r = w; r < repeat; r += WARPS) {
            uint4* d4 = reinterpret_cast<uint4*>(dst + (gidx * repeat + r) * SEG_ENTRIES);
#pragma unroll 8
            for (uint32_t i = lane; i < SEG_ENTRIES / 4; i += 32) __stcs(d4 + i, s4[i]);
        }
        st ^= 1;
    }
}

// ---- D: write-only
__global__ void __launch_bounds__(THREADS, 4) fill_kernel(uint32_t* __restrict__ dst, uint64_t n_tiles) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint64_t warp_id = (uint64_t)blockIdx.x * WARPS + w, n_warps = (uint64_t)gridDim.x * WARPS;
    const uint4 v = make_uint4(1, 2, 3, 4);
    for (uint64_t t = warp_id; t < n_tiles; t += n_warps) {
        uint4* d4 = reinterpret_cast<uint4*>(dst + t * TILE_ENTRIES);
#pragma unroll
        for (int i = 0; i < 8; i++) __stcs(d4 + i * 32 + lane, v);
    }
}

template <typename F>
static double best_ms(F&& launch, int reps = 6) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    double best = 1e30;
    for (int i = 0; i < reps; i++) {
        CK(cudaEventRecord(a));
        launch();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        CK(cudaGetLastError());
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (i > 0 && ms < best) best = ms;  // first run = warm-up
    }
    return best;
}

static uint64_t checksum(const uint32_t* d_dst, uint64_t n) {  // first / middle / last tile against the source pattern
    uint32_t h[3];
    CK(cudaMemcpy(&h[0], d_dst, 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&h[1], d_dst + n / 2, 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&h[2], d_dst + n - 1, 4, cudaMemcpyDeviceToHost));
    return ((uint64_t)h[0] << 40) ^ ((uint64_t)h[1] << 20) ^ h[2];
}

int main() {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const uint64_t n_out = 486ull << 20;  // ~4.9e8 entries = 1.95 GB, config #2's visible list
    const uint64_t n_segs = n_out / SEG_ENTRIES, n_tiles = n_segs * (SEG_ENTRIES / TILE_ENTRIES);
    const uint64_t out_entries = n_segs * SEG_ENTRIES;
    uint32_t *src = nullptr, *dst = nullptr;
    CK(cudaMalloc(&src, (size_t)N_CELLS * SEG_ENTRIES * 4));
    CK(cudaMalloc(&dst, out_entries * 4));
    {
        uint32_t* h = (uint32_t*)malloc((size_t)N_CELLS * SEG_ENTRIES * 4);
        for (uint32_t i = 0; i < N_CELLS * SEG_ENTRIES; i++) h[i] = i * 2654435761u;
        CK(cudaMemcpy(src, h, (size_t)N_CELLS * SEG_ENTRIES * 4, cudaMemcpyHostToDevice));
        free(h);
    }
    const double gb = out_entries * 4.0 / 1e9;
    printf("{\"sms\": %d, \"output_gb\": %.3f", sms, gb);

    double ms = best_ms([&] { fill_kernel<<<sms * 4, THREADS>>>(dst, n_tiles); });
    printf(", \"fill_ms\": %.4f, \"fill_write_gbs\": %.0f", ms, gb / ms * 1e3);

    ms = best_ms([&] { ldg_stg_kernel<<<sms * 4, THREADS>>>(src, dst, n_tiles); });
    const uint64_t ref = checksum(dst, out_entries);
    printf(", \"ldg_stg_ms\": %.4f, \"ldg_stg_write_gbs\": %.0f", ms, gb / ms * 1e3);

    const size_t smem_b = (size_t)WARPS * B_STAGES * TILE_ENTRIES * 4 + WARPS * B_STAGES * 8;
    CK(cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
    CK(cudaMemset(dst, 0, out_entries * 4));
    ms = best_ms([&] { bulk_kernel<<<sms * 3, THREADS, smem_b>>>(src, dst, n_tiles); });
    printf(", \"bulk_ms\": %.4f, \"bulk_write_gbs\": %.0f, \"bulk_matches_ldg_stg\": %s", ms, gb / ms * 1e3,
           checksum(dst, out_entries) == ref ? "true" : "false");

    const size_t smem_c = 2 * SEG_ENTRIES * 4 + 16;
    CK(cudaFuncSetAttribute(stage_once_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
    for (uint32_t repeat : {1u, 4u, 16u}) {
        const uint64_t groups = n_segs / repeat;
        ms = best_ms([&] { stage_once_kernel<<<sms * 6, THREADS, smem_c>>>(src, dst, groups, repeat); });
        printf(", \"stage_once_r%u_ms\": %.4f, \"stage_once_r%u_write_gbs\": %.0f", repeat, ms, repeat, groups * repeat * SEG_ENTRIES * 4.0 / 1e9 / ms * 1e3);
    }

    for (uint32_t repeat : {16u, 64u}) {
        const uint64_t groups = n_segs / repeat;
        for (int bps : {1, 2, 4}) {
            ms = best_ms([&] { stage_once_kernel<<<sms * bps, THREADS, smem_c>>>(src, dst, groups, repeat); });
            printf(", \"stage_once_r%u_g%d_ms\": %.4f", repeat, bps, ms);
        }
    }
    CK(cudaFuncSetAttribute(stage_once_mw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
    CK(cudaFuncSetAttribute(stage_once_stg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
    for (uint32_t repeat : {8u, 16u, 64u}) {
        const uint64_t groups = n_segs / repeat;
        for (int bps : {1, 2, 4}) {
            ms = best_ms([&] { stage_once_mw_kernel<<<sms * bps, THREADS, smem_c>>>(src, dst, groups, repeat); });
            printf(", \"stage_once_mw_r%u_g%d_ms\": %.4f", repeat, bps, ms);
            ms = best_ms([&] { stage_once_stg_kernel<<<sms * bps, THREADS, smem_c>>>(src, dst, groups, repeat); });
            printf(", \"stage_once_stg_r%u_g%d_ms\": %.4f", repeat, bps, ms);
        }
    }
    ms = best_ms([&] { CK(cudaMemcpyAsync(dst, dst + out_entries / 2, out_entries * 2, cudaMemcpyDeviceToDevice)); });
    printf(", \"memcpy_d2d_ms\": %.4f, \"memcpy_d2d_copy_gbs\": %.0f}\n", ms, out_entries * 2 * 2.0 / 1e9 / ms * 1e3);
    return 0;
}
