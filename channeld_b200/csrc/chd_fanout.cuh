// chd_fanout.cuh — Channel.tickData (data.go:175-291) for ALL spatial channels in one launch: one thread per
// (subscriber, cell) subscription pair runs the reference's per-connection state machine
//   while t >= lastFanOutTime + interval:           (the queue re-sort at data.go:270-286 re-visits a
//       first time  -> FULL, last = t                connection until it is no longer due: an effective loop)
//       otherwise   -> scan the cell's update ring in insertion order, select entries with
//                      lastUpdateTime <= arrival <= nextFanOutTime (skipping own updates), lastUpdateTime
//                      advancing to every picked arrival; last += interval
// Evaluate pass (state machine once per pair, up to two decisions kept in per-pair slots, state committed) ->
// exclusive scan of the decision counts -> gather pass (moves the kept decisions to (slot, cell, step) order;
// pairs that are more than two intervals behind are re-evaluated there).
#pragma once
#include "chd_interest.cuh"

namespace chd {

struct RingDev {
    const uint32_t* off;        // [C+1]
    const int64_t* arrival;     // insertion order per cell
    const uint32_t* sender;
    const uint64_t* index;
    const uint64_t* channel_msg_index;  // [C] or nullptr
    const uint32_t* total;              // entries uploaded (device scalar): offsets are clamped to it
};

constexpr uint32_t FANOUT_MAX_STEPS = 1u << 16;
constexpr uint32_t FANOUT_SLOTS = 2;            // decisions per pair kept by the evaluation pass

// The per-pair state machine of tickData.  Decisions j < n_keep are handed to `emit(j, decision)`.
// Returns the number of decisions; the final state is left in (last, flags, last_index).
template <typename Emit>
__device__ __forceinline__ uint32_t fanout_eval(const RingDev& ring, uint32_t ring_total, int64_t t, uint32_t interval, uint32_t c, uint32_t s,
                                                uint32_t me, uint32_t id_start, int64_t& last, uint8_t& flags, uint64_t& last_index, Emit&& emit) {
    const int64_t step_ns = (int64_t)interval * 1000000ll;  // ChannelTime.AddMs (channel.go:30-32)
    const bool skip_self = flags & PF_SKIP_SELF;
    const uint32_t r0 = min(ring.off[c], ring_total), r1 = min(ring.off[c + 1], ring_total);
    const uint32_t max_steps = interval ? FANOUT_MAX_STEPS : 1u;  // interval 0: the reference never terminates
    uint32_t n_out = 0;
    for (uint32_t step = 0; step < max_steps; step++) {
        const int64_t next = last + step_ns;  // data.go:205
        if (t < next) break;
        int64_t latest = next;
        if (!(flags & PF_HAD_FIRST)) {  // data.go:218-224: whole channel data
            flags |= PF_HAD_FIRST;
            last_index = ring.channel_msg_index ? ring.channel_msg_index[c] : 0ull;
            latest = t;
            chd_due d;
            d.sub = s; d.channel_id = c + id_start; d.kind = 0; d.n_selected = 0; d.first_sel = 0; d.last_sel = 0;
            d.sel_hash = 0; d.last_message_index = last_index; d.window_hi = next;
            emit(n_out, d);
            n_out++;
        } else if (r1 > r0) {  // data.go:225-265
            int64_t last_update = 0;
            if (last >= last_update) last_update = last;
            uint32_t nsel = 0, first = 0, lastsel = 0;
            uint64_t hash = 0;
            for (uint32_t k = r0; k < r1; k++) {
                if (skip_self && ring.sender[k] == me) continue;
                const int64_t a = ring.arrival[k];
                if (a >= last_update && a <= next) {
                    if (!nsel) first = k - r0;
                    lastsel = k - r0;
                    nsel++;
                    const uint64_t mi = ring.index[k];
                    hash += mi;
                    last_update = a;
                    last_index = mi;
                }
            }
            if (nsel) {
                chd_due d;
                d.sub = s; d.channel_id = c + id_start; d.kind = 1; d.n_selected = nsel; d.first_sel = first; d.last_sel = lastsel;
                d.sel_hash = hash; d.last_message_index = last_index; d.window_hi = next;
                emit(n_out, d);
                n_out++;
            }
        }
        last = latest;  // data.go:268
    }
    return n_out;
}

// Pass 1 (evaluate): runs the state machine ONCE per pair, keeps up to FANOUT_SLOTS decisions in slots[] and commits
// the state.  Pairs with more decisions (far behind) are left uncommitted.
__global__ void __launch_bounds__(128)
    fanout_eval_kernel(const uint32_t* __restrict__ n_pairs_ptr, uint64_t pair_cap, PairBuf pb, const uint32_t* __restrict__ conn_id,
                       RingDev ring, const int64_t* __restrict__ t_ptr, uint32_t id_start, uint32_t* __restrict__ due_cnt,
                       chd_due* __restrict__ slots, const uint32_t* __restrict__ by_cell) {
    const uint64_t n = min((uint64_t)*n_pairs_ptr, pair_cap);
    const int64_t t = *t_ptr;  // device-resident so the launch can be replayed from a CUDA graph
    const uint32_t ring_total = *ring.total;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        // pairs are visited grouped by cell: neighbouring lanes walk the same ring (uniform-address loads); results are
        // stored by the canonical pair index p, so the output order (slot, channel, step) does not depend on the grouping
        const uint64_t p = by_cell[i];
        const uint32_t interval = pb.interval[p];
        int64_t last = pb.last[p];
        if (t < last + (int64_t)interval * 1000000ll) {  // not due: the common case, no further state is read
            due_cnt[p] = 0;
            continue;
        }
        uint8_t flags = pb.flags[p];
        uint64_t last_index = pb.last_index[p];
        const uint32_t c = pb.cell[p], s = pb.sub[p];
        chd_due* my = slots + p * FANOUT_SLOTS;
        const uint32_t n_out = fanout_eval(ring, ring_total, t, interval, c, s, conn_id[s], id_start, last, flags, last_index,
                                           [&](uint32_t j, const chd_due& d) { if (j < FANOUT_SLOTS) my[j] = d; });
        if (n_out <= FANOUT_SLOTS) {  // otherwise: left uncommitted, the gather pass re-evaluates and commits
            pb.last[p] = last;
            pb.flags[p] = flags;
            pb.last_index[p] = last_index;
        }
        due_cnt[p] = n_out;
    }
}

// Pass 2 (gather): copies the kept decisions to their final positions (slot, channel, step order); re-evaluates the
// rare pairs with more than FANOUT_SLOTS decisions from their untouched state, writing directly.
__global__ void __launch_bounds__(128)
    fanout_gather_kernel(const uint32_t* __restrict__ n_pairs_ptr, uint64_t pair_cap, PairBuf pb, const uint32_t* __restrict__ conn_id,
                         RingDev ring, const int64_t* __restrict__ t_ptr, uint32_t id_start, const uint32_t* __restrict__ due_cnt,
                         const uint32_t* __restrict__ due_off, const chd_due* __restrict__ slots, chd_due* __restrict__ due, uint32_t due_cap,
                         Counters* __restrict__ ctr) {
    const uint64_t n = min((uint64_t)*n_pairs_ptr, pair_cap);
    const uint32_t total = due_off[n];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        ctr->n_due = total;
        ctr->required_due = total;
        if (total > due_cap) atomicOr(&ctr->overflow, (uint32_t)CHD_OVF_DUE);
    }
    const bool can_write = total <= due_cap;
    for (uint64_t p = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; p < n; p += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t cnt = due_cnt[p];
        if (cnt == 0) continue;
        const uint32_t o = due_off[p];
        if (cnt <= FANOUT_SLOTS) {
            if (can_write)
                for (uint32_t j = 0; j < cnt; j++) due[o + j] = slots[p * FANOUT_SLOTS + j];
        } else {
            int64_t last = pb.last[p];
            uint8_t flags = pb.flags[p];
            uint64_t last_index = pb.last_index[p];
            const uint32_t c = pb.cell[p], s = pb.sub[p];
            fanout_eval(ring, *ring.total, *t_ptr, pb.interval[p], c, s, conn_id[s], id_start, last, flags, last_index,
                        [&](uint32_t j, const chd_due& d) { if (can_write) due[o + j] = d; });
            pb.last[p] = last;
            pb.flags[p] = flags;
            pb.last_index[p] = last_index;
        }
    }
}

}  // namespace chd
