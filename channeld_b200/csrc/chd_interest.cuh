// chd_interest.cuh — handleUpdateSpatialInterest for a batch (message_spatial.go:41-129): turn each query's
// result into the subscriber's new spatial-subscription set, diff it against the current set
// (util.go:105-113 Difference) and carry / initialise the per-(subscriber, cell) fan-out state the way
// SubscribeToChannel does (subscription.go:34-102).
//
// Subscriptions live in HBM as a CSR by subscriber slot with SoA state, double-buffered (prev -> cur).
// Within a slot pairs are sorted by cell, so the diff is a linear merge of two sorted runs.
#pragma once
#include "chd_query.cuh"

namespace chd {

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

__global__ void __launch_bounds__(256) slot_scatter_kernel(const uint32_t* __restrict__ sub, uint32_t nq, uint32_t n_slots, int32_t* __restrict__ slot_query) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq && sub[i] < n_slots) slot_query[sub[i]] = (int32_t)i;
}

// new pair count per slot
__global__ void __launch_bounds__(256)
    slot_count_kernel(uint32_t n_slots, const int32_t* __restrict__ slot_query, const uint32_t* __restrict__ status,
                      const uint32_t* __restrict__ qcount, const uint32_t* __restrict__ prev_off, uint32_t* __restrict__ cnt,
                      Counters* __restrict__ ctr) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t err = 0;
    if (s < n_slots) {
        const int32_t q = slot_query[s];
        if (q >= 0 && status[q] == CHD_Q_OK)
            cnt[s] = qcount[q];
        else {
            cnt[s] = prev_off[s + 1] - prev_off[s];
            err = q >= 0;
        }
    }
    const uint32_t nerr = __syncthreads_count(err);
    if (threadIdx.x == 0 && nerr) atomicAdd(&ctr->n_query_errors, nerr);
}

// Builds the new subscription set of each slot + diff flags.
//   new_flag[p]  (cur index)  = 1 if pair p was subscribed by this update
//   gone_flag[p] (prev index) = 1 if prev pair p was unsubscribed by this update
__global__ void __launch_bounds__(128)
    interest_fill_kernel(GridDev g, uint32_t n_slots, const int32_t* __restrict__ slot_query, const uint32_t* __restrict__ status,
                         const Bbox* __restrict__ bbox, const uint64_t* __restrict__ win_off, const uint32_t* __restrict__ window,
                         const uint32_t* __restrict__ side_cell, const uint32_t* __restrict__ side_dist,
                         const uint32_t* __restrict__ side_cnt, const uint32_t* __restrict__ spot_off, PairBuf prev, PairBuf cur,
                         uint64_t pair_cap, const int64_t* __restrict__ now_ptr, uint32_t* __restrict__ new_flag, uint32_t* __restrict__ gone_flag,
                         Counters* __restrict__ ctr) {
    __shared__ uint32_t s_new, s_gone, s_kept;
    if (threadIdx.x == 0) s_new = s_gone = s_kept = 0;
    __syncthreads();
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t now_ns = *now_ptr;  // device-resident so the launch can be replayed from a CUDA graph
    uint32_t n_new = 0, n_gone = 0, n_kept = 0;
    if (s == 0) {
        const unsigned long long p = cur.off[n_slots];
        ctr->n_pairs = p;
        ctr->required_pairs = p;
        if (p > pair_cap) atomicOr(&ctr->overflow, (uint32_t)CHD_OVF_PAIRS);
    }
    if (s < n_slots && cur.off[n_slots] <= pair_cap) {
        const int32_t q = slot_query[s];
        uint32_t pp = prev.off[s];
        const uint32_t pe = prev.off[s + 1];
        uint32_t o = cur.off[s];
        if (q >= 0 && status[q] == CHD_Q_OK) {
            ResultIter it;
            it.init(window, win_off, bbox, side_cell, side_dist, side_cnt, spot_off, (uint32_t)q, g.cols);
            uint32_t c, d;
            while (it.next(c, d)) {
                while (pp < pe && prev.cell[pp] < c) {  // existing \ wanted -> unsubscribe (message_spatial.go:88-108)
                    gone_flag[pp] = 1;
                    n_gone++;
                    pp++;
                }
                const uint32_t interval = damping_interval_ms(d, g.default_interval_ms);  // message_spatial.go:65-80
                cur.sub[o] = s;
                cur.cell[o] = c;
                cur.dist[o] = d;
                cur.interval[o] = interval;
                if (pp < pe && prev.cell[pp] == c) {
                    // already subscribed: options merged, fan-out state untouched (subscription.go:43-58)
                    cur.flags[o] = prev.flags[pp] & ~PF_NEW;
                    cur.last[o] = prev.last[pp];
                    cur.last_index[o] = prev.last_index[pp];
                    gone_flag[pp] = 0;
                    new_flag[o] = 0;
                    n_kept++;
                    pp++;
                } else {
                    // new subscription (subscription.go:60-87): hadFirstFanOut = SkipFirstFanOut(false),
                    // lastFanOutTime = now + FanOutDelayMs, SkipSelfUpdateFanOut = true
                    cur.flags[o] = PF_NEW | PF_SKIP_SELF;
                    cur.last[o] = now_ns + (int64_t)g.default_delay_ms * 1000000ll;
                    cur.last_index[o] = 0;
                    new_flag[o] = 1;
                    n_new++;
                }
                o++;
            }
            while (pp < pe) {
                gone_flag[pp] = 1;
                n_gone++;
                pp++;
            }
        } else {
            // no query this batch, or the query errored: subscriptions stay (message_spatial.go:60-63)
            for (; pp < pe; pp++, o++) {
                cur.sub[o] = s;
                cur.cell[o] = prev.cell[pp];
                cur.dist[o] = prev.dist[pp];
                cur.interval[o] = prev.interval[pp];
                cur.flags[o] = prev.flags[pp] & ~PF_NEW;
                cur.last[o] = prev.last[pp];
                cur.last_index[o] = prev.last_index[pp];
                gone_flag[pp] = 0;
                new_flag[o] = 0;
            }
        }
    }
    if (n_new) atomicAdd(&s_new, n_new);
    if (n_gone) atomicAdd(&s_gone, n_gone);
    if (n_kept) atomicAdd(&s_kept, n_kept);
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_new) atomicAdd(&ctr->n_sub_new, s_new);
        if (s_gone) atomicAdd(&ctr->n_unsub, s_gone);
        if (s_kept) atomicAdd(&ctr->n_kept, s_kept);
    }
}

// (sub, channel id) of flagged pairs, compacted in pair order: new subscriptions index the cur buffer,
// unsubscriptions the prev buffer.
__global__ void __launch_bounds__(256)
    diff_compact_kernel(const uint32_t* __restrict__ new_flag, const uint32_t* __restrict__ new_off, const uint32_t* __restrict__ n_cur_ptr,
                        const uint32_t* __restrict__ gone_flag, const uint32_t* __restrict__ gone_off, const uint32_t* __restrict__ n_prev_ptr,
                        uint64_t cap, PairBuf cur, PairBuf prev, uint32_t id_start, uint32_t* __restrict__ new_sub,
                        uint32_t* __restrict__ new_ch, uint32_t* __restrict__ gone_sub, uint32_t* __restrict__ gone_ch) {
    const uint64_t nc = min((uint64_t)*n_cur_ptr, cap), np = min((uint64_t)*n_prev_ptr, cap);
    const uint64_t n = max(nc, np);
    for (uint64_t p = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; p < n; p += (uint64_t)gridDim.x * blockDim.x) {
        if (p < nc && new_flag[p]) {
            const uint32_t o = new_off[p];
            new_sub[o] = cur.sub[p];
            new_ch[o] = cur.cell[p] + id_start;
        }
        if (p < np && gone_flag[p]) {
            const uint32_t o = gone_off[p];
            gone_sub[o] = prev.sub[p];
            gone_ch[o] = prev.cell[p] + id_start;
        }
    }
}

}  // namespace chd
