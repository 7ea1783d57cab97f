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

__global__ void __launch_bounds__(256) slot_scatter_kernel(const uint32_t* __restrict__ sub, uint32_t nq, uint32_t n_slots, int32_t* __restrict__ slot_query) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq && sub[i] < n_slots) slot_query[sub[i]] = (int32_t)i;
}

// chd_add_subscribers / chd_remove_subscribers / chd_migrate_*: record the pending change of each listed slot
__global__ void __launch_bounds__(256) slot_ctl_set_kernel(const uint32_t* __restrict__ slot, const uint32_t* __restrict__ aux, uint32_t n, uint8_t ctl,
                                                           uint32_t src_rank, uint8_t* __restrict__ slot_ctl, uint32_t* __restrict__ slot_src,
                                                           uint32_t* __restrict__ conn_id) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = slot[i];
    slot_ctl[s] = ctl;
    if (ctl == SLOT_IMPORT) slot_src[s] = (src_rank << 20) | (i & 0xFFFFFu);  // record i of src_rank's blob
    if (conn_id && aux) conn_id[s] = aux[i];
}

// The previous run of a slot: normally a slice of the prev pair buffer; for an immigrant (SLOT_IMPORT) a slice of the
// gathered migration blob of the rank it came from.
struct PrevRun {
    const uint32_t *cell, *dist, *interval;
    const uint8_t* flags8;    // prev pair buffer
    const uint32_t* flags32;  // migration blob (flags travel as words)
    const int64_t* last;
    const uint64_t* last_index;
    uint32_t pb, pe;
    __device__ __forceinline__ uint8_t flag(uint32_t pp) const { return flags8 ? flags8[pp] : (uint8_t)flags32[pp]; }
};
__device__ __forceinline__ PrevRun prev_run_of(const PairBuf& prev, uint32_t s, uint8_t ctl, const uint32_t* __restrict__ slot_src, const MigView& mig) {
    PrevRun r;
    if (ctl == SLOT_IMPORT && mig.base) {
        const uint32_t src = slot_src[s];
        const uint32_t* b = mig.base + (uint64_t)(src >> 20) * mig.stride_words;
        const uint32_t idx = src & 0xFFFFFu;
        const uint32_t nsub = min(b[0], mig.max_subs), npairs = min(b[1], mig.max_pairs);
        r.cell = b + mig.o_cell(); r.dist = b + mig.o_dist(); r.interval = b + mig.o_interval();
        r.flags8 = nullptr; r.flags32 = b + mig.o_flags();
        r.last = reinterpret_cast<const int64_t*>(b + mig.o_last());
        r.last_index = reinterpret_cast<const uint64_t*>(b + mig.o_lidx());
        if (idx < nsub) {
            r.pb = min(b[mig.o_off() + idx], npairs);
            r.pe = min(b[mig.o_off() + idx + 1], npairs);
            if (r.pe < r.pb) r.pe = r.pb;
        } else {
            r.pb = r.pe = 0;  // the record never arrived (capacity overflow at the sender): the subscriber starts afresh
        }
        return r;
    }
    r.cell = prev.cell; r.dist = prev.dist; r.interval = prev.interval; r.flags8 = prev.flags; r.flags32 = nullptr;
    r.last = prev.last; r.last_index = prev.last_index;
    r.pb = prev.off[s];
    r.pe = (ctl == SLOT_IMPORT) ? r.pb : prev.off[s + 1];  // (import without data: empty)
    return r;
}

// scan input functor: new pair count of subscriber slot s (computed on the fly by the offset scan)
struct SlotCountIn {
    const int32_t* slot_query;  // nullptr = identity batch
    uint32_t n_queries;
    const uint32_t* status;
    const uint32_t* qcount;
    PairBuf prev;
    const uint8_t* slot_ctl;    // pending lifecycle changes (nullptr: none ever requested)
    const uint32_t* slot_src;
    MigView mig;
    __device__ __forceinline__ uint64_t operator()(uint64_t s) const {
        const uint8_t ctl = slot_ctl ? slot_ctl[s] : (uint8_t)SLOT_NORMAL;
        if (ctl == SLOT_REMOVE || ctl == SLOT_DROP) return 0;
        const int32_t q = slot_query ? slot_query[s] : (s < n_queries ? (int32_t)s : -1);
        if (q >= 0 && status[q] == CHD_Q_OK) return qcount[q];
        if (ctl == SLOT_IMPORT) {
            const PrevRun r = prev_run_of(prev, (uint32_t)s, ctl, slot_src, mig);
            return r.pe - r.pb;
        }
        return prev.off[s + 1] - prev.off[s];
    }
};

// Builds the new subscription set of each slot and appends the interest diff:
//   new list  = wanted \ existing  -> handleSubToChannel   (message_spatial.go:110-128)
//   gone list = existing \ wanted  -> handleUnsubFromChannel (message_spatial.go:88-108, util.go:105-113)
// Two sweeps of the same sorted merge per subscriber: sweep 1 counts, a block-wide scan turns the counts into
// offsets, ONE atomicAdd per block and list reserves the block's output range (the running totals double as the
// n_sub_new / n_unsub counters of the summary), sweep 2 writes pairs + list entries.  The lists are therefore grouped
// by block, ordered within a block, and unordered across blocks: they are SETS (the reference issues these messages
// in Go map order, i.e. in no order at all).
__global__ void __launch_bounds__(128)
    interest_fill_kernel(GridDev g, uint32_t n_slots, const int32_t* __restrict__ slot_query, uint32_t n_queries, const uint32_t* __restrict__ status,
                         const Bbox* __restrict__ bbox, const uint64_t* __restrict__ win_off, const uint32_t* __restrict__ window,
                         const uint32_t* __restrict__ side_cell, const uint32_t* __restrict__ side_dist,
                         const uint32_t* __restrict__ side_cnt, const uint32_t* __restrict__ spot_off, PairBuf prev, PairBuf cur,
                         const uint32_t* __restrict__ new_off, uint64_t pair_cap, const int64_t* __restrict__ now_ptr, DiffOut diff,
                         uint32_t* __restrict__ pair_channel, const unsigned long long* __restrict__ win_cursor, Counters* __restrict__ ctr,
                         uint8_t* __restrict__ slot_ctl, const uint32_t* __restrict__ slot_src, uint32_t* __restrict__ conn_id, MigView mig,
                         const int64_t* __restrict__ cell_start_ns, uint32_t* __restrict__ cell_max_interval) {
    __shared__ uint32_t s_warp_new[4], s_warp_gone[4], s_base_new, s_base_gone, s_kept;
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t now_ns = *now_ptr;  // device-resident so the launch can be replayed from a CUDA graph
    if (threadIdx.x == 0) s_kept = 0;
    // Transactional capacity rule: the new offsets were scanned into scratch (new_off).  If the new pair total does not fit,
    // NOTHING changes: every subscriber keeps its current subscriptions (cur := prev, diff lists empty) and CHD_OVF_PAIRS is
    // raised with required_pairs = the total that would have been needed.
    const unsigned long long total_new = new_off[n_slots];
    const bool ovf = total_new > pair_cap;
    if (s == 0) {
        ctr->n_pairs = ovf ? (unsigned long long)prev.off[n_slots] : total_new;
        ctr->required_pairs = total_new;
        ctr->required_window_cells = *win_cursor;
        if (ovf) atomicOr(&ctr->overflow, (uint32_t)CHD_OVF_PAIRS);
    }
    const bool active = s < n_slots;
    int32_t q = -1;
    bool queried = false;
    // lifecycle (SLOT_*): a removed slot is diffed against the empty set (all its subscriptions are reported as unsubscribed), an
    // emigrated slot's run vanishes silently, an immigrant's previous run is read out of the migration blob
    uint8_t ctl = (active && slot_ctl) ? slot_ctl[s] : (uint8_t)SLOT_NORMAL;
    if (ovf && ctl != SLOT_REMOVE) ctl = SLOT_NORMAL;  // (a removal stays pending; a migration that meets CHD_OVF_PAIRS is lost)
    const bool leaving = !ovf && (ctl == SLOT_REMOVE || ctl == SLOT_DROP);
    if (s < n_slots) {
        q = slot_query ? slot_query[s] : (s < n_queries ? (int32_t)s : -1);
        queried = q >= 0 && status[q] == CHD_Q_OK && !ovf && !leaving;
    }
    {   // queries that failed (message_spatial.go:60-63): counted for the summary
        const uint32_t nerr = __syncthreads_count(s < n_slots && q >= 0 && !queried);
        if (threadIdx.x == 0 && nerr) atomicAdd(&ctr->n_query_errors, nerr);
    }
    PrevRun pr;
    pr.pb = pr.pe = 0;
    if (active) {
        pr = prev_run_of(prev, s, ovf ? (uint8_t)SLOT_NORMAL : ctl, slot_src, mig);
        if (slot_ctl && !ovf && ctl != SLOT_NORMAL) {  // consumed
            slot_ctl[s] = SLOT_NORMAL;
            if (leaving) conn_id[s] = 0;  // the slot is free
        }
    } else {
        queried = false;
    }
    const uint32_t pb = pr.pb, pe = pr.pe;
    // ---- sweep 1: count
    uint32_t n_new = 0, n_gone = 0, n_kept = 0;
    if (queried) {
        ResultIter it;
        it.init(window, win_off, bbox, side_cell, side_dist, side_cnt, spot_off, (uint32_t)q, g.cols);
        uint32_t pp = pb, c, d;
        while (it.next(c, d)) {
            while (pp < pe && pr.cell[pp] < c) { n_gone++; pp++; }
            if (pp < pe && pr.cell[pp] == c) { n_kept++; pp++; } else n_new++;
        }
        n_gone += pe - pp;
    } else if (ctl == SLOT_REMOVE && !ovf) {
        n_gone = pe - pb;
    }
    // ---- block-wide exclusive offsets (128 threads = 4 warps)
    uint32_t in_new = n_new, in_gone = n_gone;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t a = __shfl_up_sync(0xffffffffu, in_new, o), b2 = __shfl_up_sync(0xffffffffu, in_gone, o);
        if (lane >= o) { in_new += a; in_gone += b2; }
    }
    if (lane == 31) { s_warp_new[w] = in_new; s_warp_gone[w] = in_gone; }
    __syncthreads();
    uint32_t off_new = in_new - n_new, off_gone = in_gone - n_gone;
    for (int k = 0; k < w; k++) { off_new += s_warp_new[k]; off_gone += s_warp_gone[k]; }
    if (threadIdx.x == 0) {
        const uint32_t tn = s_warp_new[0] + s_warp_new[1] + s_warp_new[2] + s_warp_new[3];
        const uint32_t tg = s_warp_gone[0] + s_warp_gone[1] + s_warp_gone[2] + s_warp_gone[3];
        s_base_new = tn ? atomicAdd(&ctr->n_sub_new, tn) : 0u;
        s_base_gone = tg ? atomicAdd(&ctr->n_unsub, tg) : 0u;
    }
    if (n_kept) atomicAdd(&s_kept, n_kept);
    __syncthreads();
    if (threadIdx.x == 0 && s_kept) atomicAdd(&ctr->n_kept, s_kept);
    if (!active) return;
    uint32_t on = s_base_new + off_new, og = s_base_gone + off_gone;
    // ---- sweep 2: write the subscriber's new pair run (+ diff entries)
    uint32_t pp = pb;
    uint32_t o = ovf ? prev.off[s] : new_off[s];
    const uint32_t o_end = ovf ? prev.off[s + 1] : new_off[s + 1];  // (defensive bound: a run never grows past its reserved range)
    cur.off[s] = o;
    if (s + 1 == n_slots) cur.off[n_slots] = o_end;
    if (queried) {
        ResultIter it;
        it.init(window, win_off, bbox, side_cell, side_dist, side_cnt, spot_off, (uint32_t)q, g.cols);
        uint32_t c, d;
        while (o < o_end && it.next(c, d)) {
            while (pp < pe && pr.cell[pp] < c) {  // existing \ wanted -> unsubscribe
                if (og < pair_cap) { diff.gone_sub[og] = s; diff.gone_ch[og] = pr.cell[pp] + g.id_start; }
                og++;
                pp++;
            }
            const uint32_t interval = damping_interval_ms(d, g.default_interval_ms);  // message_spatial.go:65-80
            cur.sub[o] = s;
            cur.cell[o] = c;
            pair_channel[o] = c + g.id_start;  // host-facing copy (chd_fetch_results reads it back without a conversion pass)
            cur.dist[o] = d;
            cur.interval[o] = interval;
            if (pp < pe && pr.cell[pp] == c) {
                // already subscribed: options merged, fan-out state untouched (subscription.go:43-58)
                cur.flags[o] = pr.flag(pp) & ~PF_NEW;
                cur.last[o] = pr.last[pp];
                cur.last_index[o] = pr.last_index[pp];
                pp++;
            } else {
                // new subscription (subscription.go:60-87): hadFirstFanOut = SkipFirstFanOut(false),
                // lastFanOutTime = now + FanOutDelayMs, SkipSelfUpdateFanOut = true
                cur.flags[o] = PF_NEW | PF_SKIP_SELF;
                cur.last[o] = now_ns - (cell_start_ns ? cell_start_ns[c] : 0ll) + (int64_t)g.default_delay_ms * 1000000ll;  // ch.GetTime() + delay
                cur.last_index[o] = 0;
                // ChannelData.maxFanOutIntervalMs only ever grows, and only when a subscription is created (subscription.go:84-86)
                if (cell_max_interval && cell_max_interval[c] < interval) atomicMax(&cell_max_interval[c], interval);
                if (on < pair_cap) { diff.new_sub[on] = s; diff.new_ch[on] = c + g.id_start; }
                on++;
            }
            o++;
        }
        for (; pp < pe; pp++) {
            if (og < pair_cap) { diff.gone_sub[og] = s; diff.gone_ch[og] = pr.cell[pp] + g.id_start; }
            og++;
        }
    } else if (leaving) {
        if (ctl == SLOT_REMOVE)  // UnsubscribeFromChannel for every channel of the slot (subscription.go:104-125)
            for (; pp < pe; pp++) {
                if (og < pair_cap) { diff.gone_sub[og] = s; diff.gone_ch[og] = pr.cell[pp] + g.id_start; }
                og++;
            }
    } else {
        // no query this batch, or the query errored: subscriptions stay (message_spatial.go:60-63)
        for (; pp < pe && o < o_end; pp++, o++) {
            cur.sub[o] = s;
            cur.cell[o] = pr.cell[pp];
            pair_channel[o] = pr.cell[pp] + g.id_start;
            cur.dist[o] = pr.dist[pp];
            cur.interval[o] = pr.interval[pp];
            cur.flags[o] = pr.flag(pp) & ~PF_NEW;
            cur.last[o] = pr.last[pp];
            cur.last_index[o] = pr.last_index[pp];
        }
    }
}

}  // namespace chd
