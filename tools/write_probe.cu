// write_probe.cu — which store pattern reaches the HBM write ceiling on B200?  The emit kernel (chd_emit.cuh) is a pure write
// stream on the DRAM side (its reads are L2 hits); round 1 measured 5.3 TB/s for it against 7.2-7.4 TB/s for torch's fill_
// and 6.1 TB/s for a persistent-grid streaming-store fill (tools/emit_copy_probe.cu).  This probe isolates the store side:
// cache operator (wb / cg / cs / wt), grid shape (persistent vs one CTA per chunk), bytes in flight per thread, chunk
// geometry (warp tiles vs CTA rows), TMA bulk stores, cudaMemset; then the best shapes are repeated as COPIES out of an
// L2-resident 16 MB pool (the emit kernel's traffic pattern).
// Build + run:  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/write_probe.cu -o tools/_bin/write_probe
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

enum { OP_WB = 0, OP_CG = 1, OP_CS = 2, OP_WT = 3 };
template <int OP>
__device__ __forceinline__ void st16(uint4* p, uint4 v) {
    if (OP == OP_WB) *p = v;
    else if (OP == OP_CG) __stcg(p, v);
    else if (OP == OP_CS) __stcs(p, v);
    else __stwt(p, v);
}

// persistent grid, warp-owned tiles of TILE_V uint4 per lane (tile bytes = 32 * TILE_V * 16)
template <int OP, int TILE_V>
__global__ void __launch_bounds__(256) fill_warp_tiles(uint4* __restrict__ dst, uint64_t n_tiles) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint64_t warp_id = (uint64_t)blockIdx.x * 8 + w, n_warps = (uint64_t)gridDim.x * 8;
    const uint4 v = make_uint4(1, 2, 3, 4);
    for (uint64_t t = warp_id; t < n_tiles; t += n_warps) {
        uint4* d = dst + t * (32 * TILE_V);
#pragma unroll
        for (int i = 0; i < TILE_V; i++) st16<OP>(d + i * 32 + lane, v);
    }
}

// persistent grid, CTA-owned chunks: 256 threads write ROWS rows of 4 KB (thread i -> uint4 i of the row)
template <int OP, int ROWS>
__global__ void __launch_bounds__(256) fill_cta_rows(uint4* __restrict__ dst, uint64_t n_chunks) {
    const uint4 v = make_uint4(1, 2, 3, 4);
    for (uint64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        uint4* d = dst + c * (256 * ROWS);
#pragma unroll
        for (int i = 0; i < ROWS; i++) st16<OP>(d + i * 256 + threadIdx.x, v);
    }
}

// non-persistent: one CTA per chunk of ROWS x 4 KB
template <int OP, int ROWS>
__global__ void __launch_bounds__(256) fill_grid(uint4* __restrict__ dst) {
    const uint4 v = make_uint4(1, 2, 3, 4);
    uint4* d = dst + (uint64_t)blockIdx.x * (256 * ROWS);
#pragma unroll
    for (int i = 0; i < ROWS; i++) st16<OP>(d + i * 256 + threadIdx.x, v);
}

// ---- TMA bulk store out of shared memory (constant contents), CHUNK bytes per op
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
template <int CHUNK>
__global__ void __launch_bounds__(256) fill_tma(uint4* __restrict__ dst, uint64_t n_chunks) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint4* buf = reinterpret_cast<uint4*>(smem_raw);
    for (int i = threadIdx.x; i < CHUNK / 16; i += 256) buf[i] = make_uint4(1, 2, 3, 4);
    fence_async_smem();
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane != 0) return;
    int inflight = 0;
    for (uint64_t c = (uint64_t)blockIdx.x * 8 + w; c < n_chunks; c += (uint64_t)gridDim.x * 8) {
        bulk_s2g(dst + c * (CHUNK / 16), buf, CHUNK);
        bulk_commit();
        if (++inflight >= 4) bulk_wait_read<3>();
    }
    bulk_wait_read<0>();
}

// ---- copies out of an L2-resident pool (16 MB), segment = 16 KB "cell", pseudo-random cell per output segment
constexpr uint32_t SEG_V = 1024;   // uint4 per segment (16 KB)
constexpr uint32_t N_CELLS = 1024;
__host__ __device__ inline uint32_t cell_of_segment(uint64_t seg) {
    uint64_t x = seg * 0x9E3779B97F4A7C15ull;
    x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
    return (uint32_t)(x % N_CELLS);
}
template <int OP, int TILE_V>
__global__ void __launch_bounds__(256) copy_warp_tiles(const uint4* __restrict__ src, uint4* __restrict__ dst, uint64_t n_tiles) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint64_t warp_id = (uint64_t)blockIdx.x * 8 + w, n_warps = (uint64_t)gridDim.x * 8;
    constexpr uint32_t TPS = SEG_V / (32 * TILE_V);  // tiles per segment
    for (uint64_t t = warp_id; t < n_tiles; t += n_warps) {
        const uint64_t seg = t / TPS;
        const uint4* s = src + (uint64_t)cell_of_segment(seg) * SEG_V + (t % TPS) * (32 * TILE_V);
        uint4* d = dst + t * (32 * TILE_V);
        uint4 v[TILE_V];
#pragma unroll
        for (int i = 0; i < TILE_V; i++) v[i] = __ldg(s + i * 32 + lane);
#pragma unroll
        for (int i = 0; i < TILE_V; i++) st16<OP>(d + i * 32 + lane, v[i]);
    }
}
template <int OP, int ROWS>
__global__ void __launch_bounds__(256) copy_grid(const uint4* __restrict__ src, uint4* __restrict__ dst) {
    constexpr uint32_t CPS = SEG_V / (256 * ROWS);  // chunks per segment
    const uint64_t c = blockIdx.x;
    const uint4* s = src + (uint64_t)cell_of_segment(c / CPS) * SEG_V + (c % CPS) * (256 * ROWS);
    uint4* d = dst + c * (256 * ROWS);
    uint4 v[ROWS];
#pragma unroll
    for (int i = 0; i < ROWS; i++) v[i] = __ldg(s + i * 256 + threadIdx.x);
#pragma unroll
    for (int i = 0; i < ROWS; i++) st16<OP>(d + i * 256 + threadIdx.x, v[i]);
}
template <int OP, int ROWS>
__global__ void __launch_bounds__(256) copy_cta_rows(const uint4* __restrict__ src, uint4* __restrict__ dst, uint64_t n_chunks) {
    constexpr uint32_t CPS = SEG_V / (256 * ROWS);
    for (uint64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const uint4* s = src + (uint64_t)cell_of_segment(c / CPS) * SEG_V + (c % CPS) * (256 * ROWS);
        uint4* d = dst + c * (256 * ROWS);
        uint4 v[ROWS];
#pragma unroll
        for (int i = 0; i < ROWS; i++) v[i] = __ldg(s + i * 256 + threadIdx.x);
#pragma unroll
        for (int i = 0; i < ROWS; i++) st16<OP>(d + i * 256 + threadIdx.x, v[i]);
    }
}

// persistent CTAs that pull chunk numbers from a global ticket (dynamic order, like the hardware block scheduler)
template <int OP, int ROWS>
__global__ void __launch_bounds__(256) fill_ticket(uint4* __restrict__ dst, uint64_t n_chunks, unsigned long long* ticket) {
    __shared__ unsigned long long s_c;
    const uint4 v = make_uint4(1, 2, 3, 4);
    for (;;) {
        if (threadIdx.x == 0) s_c = atomicAdd(ticket, 1ull);
        __syncthreads();
        const uint64_t c = s_c;
        __syncthreads();
        if (c >= n_chunks) return;
        uint4* d = dst + c * (256 * ROWS);
#pragma unroll
        for (int i = 0; i < ROWS; i++) st16<OP>(d + i * 256 + threadIdx.x, v);
    }
}
// the same as a copy out of the L2-resident pool; the next ticket is requested before the current chunk is copied
template <int OP, int ROWS>
__global__ void __launch_bounds__(256) copy_ticket(const uint4* __restrict__ src, uint4* __restrict__ dst, uint64_t n_chunks, unsigned long long* ticket) {
    __shared__ unsigned long long s_c[2];
    constexpr uint32_t CPS = SEG_V / (256 * ROWS);
    if (threadIdx.x == 0) s_c[0] = atomicAdd(ticket, 1ull);
    __syncthreads();
    int b = 0;
    for (;;) {
        const uint64_t c = s_c[b];
        if (threadIdx.x == 0) s_c[b ^ 1] = atomicAdd(ticket, 1ull);  // in flight while this chunk is copied
        if (c >= n_chunks) return;
        const uint4* s = src + (uint64_t)cell_of_segment(c / CPS) * SEG_V + (c % CPS) * (256 * ROWS);
        uint4* d = dst + c * (256 * ROWS);
        uint4 v[ROWS];
#pragma unroll
        for (int i = 0; i < ROWS; i++) v[i] = __ldg(s + i * 256 + threadIdx.x);
#pragma unroll
        for (int i = 0; i < ROWS; i++) st16<OP>(d + i * 256 + threadIdx.x, v[i]);
        __syncthreads();
        b ^= 1;
    }
}
// non-persistent copy whose source comes from a per-chunk descriptor in memory (one dependent load, like the emit kernel)
template <int OP, int ROWS>
__global__ void __launch_bounds__(256) copy_grid_desc(const uint4* __restrict__ src, uint4* __restrict__ dst, const uint4* __restrict__ desc) {
    const uint64_t c = blockIdx.x;
    const uint4 dw = __ldg(desc + c);
    const uint4* s = src + dw.x;
    uint4* d = dst + c * (256 * ROWS);
    uint4 v[ROWS];
#pragma unroll
    for (int i = 0; i < ROWS; i++) v[i] = __ldg(s + i * 256 + threadIdx.x);
#pragma unroll
    for (int i = 0; i < ROWS; i++) st16<OP>(d + i * 256 + threadIdx.x, v[i]);
}
// the same with the source shifted by SHIFT uint4 (16-byte aligned like the emit kernel's phase-matched reads, but not 128-byte
// aligned: a warp's 512-byte load then spans 5 cache lines instead of 4)
template <int OP, int ROWS, int SHIFT>
__global__ void __launch_bounds__(256) copy_grid_shift(const uint4* __restrict__ src, uint4* __restrict__ dst) {
    constexpr uint32_t CPS = SEG_V / (256 * ROWS);
    const uint64_t c = blockIdx.x;
    const uint4* s = src + ((uint64_t)cell_of_segment(c / CPS) * SEG_V + (c % CPS) * (256 * ROWS) + SHIFT) % ((uint64_t)(N_CELLS - 1) * SEG_V);
    uint4* d = dst + c * (256 * ROWS);
    uint4 v[ROWS];
#pragma unroll
    for (int i = 0; i < ROWS; i++) v[i] = __ldg(s + i * 256 + threadIdx.x);
#pragma unroll
    for (int i = 0; i < ROWS; i++) st16<OP>(d + i * 256 + threadIdx.x, v[i]);
}
// per-chunk pseudo-random shift (what the emit kernel sees: the phase is different for every pair)
template <int OP, int ROWS>
__global__ void __launch_bounds__(256) copy_grid_rshift(const uint4* __restrict__ src, uint4* __restrict__ dst) {
    constexpr uint32_t CPS = SEG_V / (256 * ROWS);
    const uint64_t c = blockIdx.x;
    const uint32_t shift = cell_of_segment(c * 7 + 3) & 7;
    const uint4* s = src + ((uint64_t)cell_of_segment(c / CPS) * SEG_V + (c % CPS) * (256 * ROWS) + shift) % ((uint64_t)(N_CELLS - 1) * SEG_V);
    uint4* d = dst + c * (256 * ROWS);
    uint4 v[ROWS];
#pragma unroll
    for (int i = 0; i < ROWS; i++) v[i] = __ldg(s + i * 256 + threadIdx.x);
#pragma unroll
    for (int i = 0; i < ROWS; i++) st16<OP>(d + i * 256 + threadIdx.x, v[i]);
}

template <int ROWS>
__global__ void make_desc(uint4* desc, uint64_t n_chunks) {
    constexpr uint32_t CPS = SEG_V / (256 * ROWS);
    const uint64_t c = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (c < n_chunks) desc[c] = make_uint4((uint32_t)((uint64_t)cell_of_segment(c / CPS) * SEG_V + (c % CPS) * (256 * ROWS)), 0, 0, 0);
}

template <typename F>
static double best_ms(F&& launch, int reps = 7) {
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
        if (i > 0 && ms < best) best = ms;
    }
    return best;
}

static bool first = true;
static void out(const char* name, double ms, double gb) {
    printf("%s\"%s\": {\"ms\": %.4f, \"gbs\": %.0f}", first ? "" : ", ", name, ms, gb / ms * 1e3);
    first = false;
    fflush(stdout);
}

int main() {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const uint64_t n_v = (486ull << 20) / 4;  // uint4 count: 1.95 GB... rounded to 2.04 GB like the first probe
    const uint64_t bytes = n_v * 16;
    const double gb = bytes / 1e9;
    uint4 *dst = nullptr, *src = nullptr;
    CK(cudaMalloc(&dst, bytes));
    CK(cudaMalloc(&src, (size_t)N_CELLS * SEG_V * 16));
    CK(cudaMemset(src, 0x5a, (size_t)N_CELLS * SEG_V * 16));
    printf("{\"sms\": %d, \"gb\": %.3f, ", sms, gb);
#define WT(OP, V, BPS) out("fill_warp_" #OP "_v" #V "_b" #BPS, best_ms([&] { fill_warp_tiles<OP, V><<<sms * BPS, 256>>>(dst, n_v / (32 * V)); }), gb)
    WT(OP_CS, 8, 4); WT(OP_WB, 8, 4); WT(OP_CG, 8, 4); WT(OP_WT, 8, 4);
    WT(OP_WB, 8, 8); WT(OP_WB, 8, 2); WT(OP_WB, 4, 8); WT(OP_WB, 16, 4); WT(OP_CS, 8, 8); WT(OP_CS, 16, 4);
#define CR(OP, R, BPS) out("fill_cta_" #OP "_r" #R "_b" #BPS, best_ms([&] { fill_cta_rows<OP, R><<<sms * BPS, 256>>>(dst, n_v / (256 * R)); }), gb)
    CR(OP_WB, 1, 8); CR(OP_WB, 4, 8); CR(OP_WB, 8, 4); CR(OP_CS, 4, 8); CR(OP_WB, 4, 4); CR(OP_WB, 2, 8);
#define GR(OP, R) out("fill_grid_" #OP "_r" #R, best_ms([&] { fill_grid<OP, R><<<(unsigned)(n_v / (256 * R)), 256>>>(dst); }), gb)
    GR(OP_WB, 1); GR(OP_WB, 2); GR(OP_WB, 4); GR(OP_WB, 8); GR(OP_CS, 1); GR(OP_CS, 4); GR(OP_CG, 4); GR(OP_WT, 4);
    out("memset", best_ms([&] { CK(cudaMemsetAsync(dst, 7, bytes)); }), gb);
    CK(cudaFuncSetAttribute(fill_tma<16384>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384));
    CK(cudaFuncSetAttribute(fill_tma<4096>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096));
    CK(cudaFuncSetAttribute(fill_tma<65536>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    out("fill_tma_16k_b1", best_ms([&] { fill_tma<16384><<<sms, 256, 16384>>>(dst, bytes / 16384); }), gb);
    out("fill_tma_16k_b2", best_ms([&] { fill_tma<16384><<<sms * 2, 256, 16384>>>(dst, bytes / 16384); }), gb);
    out("fill_tma_4k_b2", best_ms([&] { fill_tma<4096><<<sms * 2, 256, 4096>>>(dst, bytes / 4096); }), gb);
    out("fill_tma_64k_b1", best_ms([&] { fill_tma<65536><<<sms, 256, 65536>>>(dst, bytes / 65536); }), gb);
    // copies
#define CW(OP, V, BPS) out("copy_warp_" #OP "_v" #V "_b" #BPS, best_ms([&] { copy_warp_tiles<OP, V><<<sms * BPS, 256>>>(src, dst, n_v / (32 * V)); }), gb)
    CW(OP_CS, 8, 4); CW(OP_WB, 8, 4); CW(OP_CG, 8, 4); CW(OP_WB, 4, 8); CW(OP_CS, 4, 8); CW(OP_WB, 8, 8);
#define CG_(OP, R) out("copy_grid_" #OP "_r" #R, best_ms([&] { copy_grid<OP, R><<<(unsigned)(n_v / (256 * R)), 256>>>(src, dst); }), gb)
    CG_(OP_WB, 1); CG_(OP_WB, 2); CG_(OP_WB, 4); CG_(OP_CS, 1); CG_(OP_CS, 4); CG_(OP_CG, 2);
#define CC(OP, R, BPS) out("copy_cta_" #OP "_r" #R "_b" #BPS, best_ms([&] { copy_cta_rows<OP, R><<<sms * BPS, 256>>>(src, dst, n_v / (256 * R)); }), gb)
    CC(OP_WB, 4, 8); CC(OP_WB, 2, 8); CC(OP_CS, 4, 8); CC(OP_WB, 4, 4);
    unsigned long long* ticket = nullptr;
    CK(cudaMalloc(&ticket, 8));
#define FT(OP, R, BPS) out("fill_ticket_" #OP "_r" #R "_b" #BPS, best_ms([&] { CK(cudaMemsetAsync(ticket, 0, 8)); fill_ticket<OP, R><<<sms * BPS, 256>>>(dst, n_v / (256 * R), ticket); }), gb)
    FT(OP_WB, 4, 8); FT(OP_WB, 4, 4); FT(OP_WB, 16, 4); FT(OP_CS, 4, 8);
#define CT(OP, R, BPS) out("copy_ticket_" #OP "_r" #R "_b" #BPS, best_ms([&] { CK(cudaMemsetAsync(ticket, 0, 8)); copy_ticket<OP, R><<<sms * BPS, 256>>>(src, dst, n_v / (256 * R), ticket); }), gb)
    CT(OP_CS, 4, 8); CT(OP_CS, 4, 6); CT(OP_CS, 4, 4); CT(OP_CS, 2, 8);
    {
        uint4* desc = nullptr;
        CK(cudaMalloc(&desc, (n_v / 256) * 16));
        make_desc<4><<<(unsigned)((n_v / 1024 + 255) / 256), 256>>>(desc, n_v / 1024);
        out("copy_grid_desc_OP_CS_r4", best_ms([&] { copy_grid_desc<OP_CS, 4><<<(unsigned)(n_v / 1024), 256>>>(src, dst, desc); }), gb);
        make_desc<2><<<(unsigned)((n_v / 512 + 255) / 256), 256>>>(desc, n_v / 512);
        out("copy_grid_desc_OP_CS_r2", best_ms([&] { copy_grid_desc<OP_CS, 2><<<(unsigned)(n_v / 512), 256>>>(src, dst, desc); }), gb);
        make_desc<1><<<(unsigned)((n_v / 256 + 255) / 256), 256>>>(desc, n_v / 256);
        out("copy_grid_desc_OP_CS_r1", best_ms([&] { copy_grid_desc<OP_CS, 1><<<(unsigned)(n_v / 256), 256>>>(src, dst, desc); }), gb);
    }
    CG_(OP_CS, 2);
    out("copy_grid_shift1_OP_CS_r4", best_ms([&] { copy_grid_shift<OP_CS, 4, 1><<<(unsigned)(n_v / 1024), 256>>>(src, dst); }), gb);
    out("copy_grid_shift4_OP_CS_r4", best_ms([&] { copy_grid_shift<OP_CS, 4, 4><<<(unsigned)(n_v / 1024), 256>>>(src, dst); }), gb);
    out("copy_grid_shift8_OP_CS_r4", best_ms([&] { copy_grid_shift<OP_CS, 4, 8><<<(unsigned)(n_v / 1024), 256>>>(src, dst); }), gb);
    out("copy_grid_rshift_OP_CS_r4", best_ms([&] { copy_grid_rshift<OP_CS, 4><<<(unsigned)(n_v / 1024), 256>>>(src, dst); }), gb);
    out("memcpy_d2d_half", best_ms([&] { CK(cudaMemcpyAsync(dst, dst + n_v / 2, bytes / 2, cudaMemcpyDeviceToDevice)); }), gb);
    printf("}\n");
    return 0;
}
