// chd_misc.cuh — small bookkeeping kernels (counters, id conversion).  Internal linkage: several translation units use them.
#pragma once
#include "chd_types.cuh"

namespace chd {

static __global__ void cell_key_to_id_kernel(uint32_t* __restrict__ k, uint32_t n, uint32_t cells, uint32_t id_start, uint8_t* __restrict__ valid) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool ok = k[i] < cells;
    k[i] = ok ? k[i] + id_start : 0u;
    if (valid) valid[i] = ok ? 1 : 0;
}

static __global__ void narrow_offsets_kernel(const uint64_t* __restrict__ in, uint32_t n, uint32_t* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint32_t)in[i];
}

static __global__ void add_const_kernel(const uint32_t* __restrict__ in, uint32_t n, uint32_t c, uint32_t* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i] + c;
}

static __global__ void set_i64_kernel(int64_t* dst, int64_t v) { *dst = v; }
// first kernel of the interest / fan-out stages: publishes the tick time, opens a new scan epoch (never 0: zero-initialised
// descriptors must read as stale) and zeroes the stage's counters / cursors
static __global__ void stage_begin_kernel(int64_t* dst, int64_t v, unsigned long long* epoch, uint32_t* zero_u32, uint32_t n_zero,
                                          unsigned long long* zero_u64) {
    *dst = v;
    *epoch = chd_next_epoch(*epoch);
    for (uint32_t k = 0; k < n_zero; k++) zero_u32[k] = 0;
    if (zero_u64) *zero_u64 = 0;
}
static __global__ void set_u32_kernel(uint32_t* dst, uint32_t v) { *dst = v; }
// the slot table grew from old_n to new_n slots: the new slots hold no pairs (CSR offsets continue at the total)
static __global__ void extend_offsets_kernel(uint32_t* __restrict__ off, uint32_t old_n, uint32_t new_n) {
    const uint32_t i = old_n + 1 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= new_n) off[i] = off[old_n];
}

}  // namespace chd
