// chd_types.cuh — plain structs shared between the kernels (csrc/*.cuh) and the host-side launch code (csrc/*.cu).
#pragma once
#include "chd_device.cuh"

namespace chd {

// tile geometry the host-side sizing code shares with the kernels
constexpr int BUILD_THREADS = 256;
constexpr int BUILD_WARPS = BUILD_THREADS / 32;
constexpr int BUILD_ROUNDS = 8;
constexpr int BUILD_TILE = BUILD_THREADS * BUILD_ROUNDS;  // 2048 entities per tile
constexpr int BUILD_MAX_BINS = 1024;
constexpr int EMIT_THREADS = 256;
constexpr int EMIT_WARPS = EMIT_THREADS / 32;
constexpr int EMIT_CHUNKS = 8;                   // 16-byte chunks per lane per tile
constexpr int EMIT_TILE = 32 * EMIT_CHUNKS * 4;  // 1024 entries = 4 KB of output per WARP tile
constexpr int EMIT_SMEM_PAIRS = 64;              // pairs per warp tile staged in shared memory

struct QueryDev {
    uint32_t n;
    const uint32_t* sub;
    const uint8_t* kind;
    const double *sph_cx, *sph_cz, *sph_r;
    const double *box_cx, *box_cz, *box_ex, *box_ez;
    const double *cone_cx, *cone_cz, *cone_dx, *cone_dz, *cone_angle, *cone_r;
    const uint32_t *spot_off, *spot_ndist;
    const double *spot_x, *spot_z;
    const uint32_t* spot_dist;
};

struct Bbox {
    uint32_t gx0, gy0, bw, bh;  // bw == 0 => empty
};

struct PairBuf {
    uint32_t* off;         // [S+1]
    uint32_t* sub;         // [P] owning subscriber slot
    uint32_t* cell;        // [P] cell index
    uint32_t* dist;        // [P]
    uint32_t* interval;    // [P] FanOutIntervalMs
    uint8_t* flags;        // [P]
    int64_t* last;         // [P] lastFanOutTime (ns)
    uint64_t* last_index;  // [P] lastMessageIndex
};
enum : uint8_t { PF_HAD_FIRST = 1, PF_NEW = 2, PF_SKIP_SELF = 4 };

struct Counters {  // device mirror of chd_tick_summary's counters
    unsigned long long n_pairs, n_visible;
    uint32_t n_entities_in_world, n_query_errors, n_sub_new, n_unsub, n_kept, n_due, n_handover, overflow;
    unsigned long long required_pairs, required_window_cells, required_visible;
    uint32_t required_due, reserved;
};

struct DiffOut {  // the two interest-diff lists of a tick: (subscriber slot, channel id)
    uint32_t *new_sub, *new_ch, *gone_sub, *gone_ch;
};

struct RingDev {
    const uint32_t* off;        // [C+1]
    const int64_t* arrival;     // insertion order per cell
    const uint32_t* sender;
    const uint64_t* index;
    const uint64_t* channel_msg_index;  // [C] or nullptr
    const uint32_t* total;              // entries uploaded (device scalar): offsets are clamped to it
};

// Payload identity of a decision (window classes, chd_classes.cuh): two decisions of one channel carry the same merged
// payload if they are both FULL, or if they start from the same lastFanOutTime (`lo`), end at the same nextFanOutTime
// (the record's window_hi) and neither subscriber had an own update left out of that window; a decision with
// self-skipped updates is its own class.  word = cell << 34 | kind << 33 | skipped << 32 | (skipped ? subscriber slot : 0).
struct DueKey {
    int64_t lo;
    uint64_t word;
};

struct HandoverOut {
    uint32_t* entity;
    uint32_t* src_cell;
    uint32_t* dst_cell;
    uint32_t* count;  // device counter (may exceed cap: required size)
    uint32_t cap;
};

// Extras of the FINAL pass of the entity build, fused into the scatter:
//   phase_stride != 0 : also write the three phase-shifted copies of the payload (chd_emit.cuh)
//   cell_start != null: single-pass sorts only (digit == key): block 0 publishes the cell CSR offsets straight from
//                       the scanned histogram (cell_start[c] = #keys < c), replacing a separate boundaries kernel
struct ScatterExtras {
    uint32_t phase_stride;
    uint32_t* cell_start;
    uint32_t cells;
    uint32_t* n_in_world;
};

struct BcastDev {
    uint32_t n;
    const uint32_t *channel, *flags, *sender, *client;
};

}  // namespace chd
