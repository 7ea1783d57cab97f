// chd_broadcast.cuh — BroadcastType_ADJACENT_CHANNELS recipient sets (message.go:188-239) for a batch of messages:
// for a message sent to spatial channel ch, the recipients are the union, WITHOUT duplicates, of the connections
// subscribed to ch's 3x3 neighbours (GetAdjacentChannels, spatial.go:358-381) and — unless ALL_BUT_OWNER is set — to ch
// itself (message.go:203-205), minus the sender (ALL_BUT_SENDER), the clients (ALL_BUT_CLIENT), the servers
// (ALL_BUT_SERVER) and the connection named by ServerForwardMessage.ClientConnId (message.go:220-237).
//
// Every channel's subscriber list already exists on the device: the pairs grouped by cell (by_cell, the fan-out pass's
// work order).  One warp handles one (message, neighbour k) with k = 0..8 in the reference's y-outer / x-inner order.
// A subscriber that is subscribed to several of the <= 9 cells is reported for the first of them in that order only
// (its own pair run is sorted by cell, so "first" = no smaller member cell in the run): that is the de-duplication the
// reference does with a Go map.  The result is a SET per message (the reference iterates a map: no order).
#pragma once
#include "chd_types.cuh"

namespace chd {

enum : uint32_t {  // channeldpb.BroadcastType (channeld.proto), Check() = any bit in common (channeldpb/extension.go:5-7)
    BC_ALL_BUT_SENDER = 4, BC_ALL_BUT_OWNER = 8, BC_ALL_BUT_CLIENT = 16, BC_ALL_BUT_SERVER = 32
};
enum : uint8_t { CONN_SERVER = 1, CONN_CLIENT = 2 };  // channeldpb.ConnectionType

// cell of neighbour k (0..8, row-major around the centre) of the message's channel, or 0xFFFFFFFF if it is outside the
// grid / excluded (the centre under ALL_BUT_OWNER) / the channel id is not a cell of this grid
__device__ __forceinline__ uint32_t bcast_cell(const GridDev& g, uint32_t channel, uint32_t flags, int k) {
    const uint32_t idx = channel - g.id_start;
    if (channel < g.id_start || idx >= g.cells) return 0xFFFFFFFFu;
    const int gx = (int)(idx % g.cols), gy = (int)(idx / g.cols);
    const int x = gx + (k % 3) - 1, y = gy + (k / 3) - 1;
    if (x < 0 || y < 0 || x >= (int)g.cols || y >= (int)g.rows) return 0xFFFFFFFFu;
    if (k == 4 && (flags & BC_ALL_BUT_OWNER)) return 0xFFFFFFFFu;
    return (uint32_t)x + (uint32_t)y * g.cols;
}

__device__ __forceinline__ bool bcast_member(const GridDev& g, uint32_t channel, uint32_t flags, uint32_t cell) {
    const uint32_t idx = channel - g.id_start;
    const int dx = (int)(cell % g.cols) - (int)(idx % g.cols), dy = (int)(cell / g.cols) - (int)(idx / g.cols);
    if (dx < -1 || dx > 1 || dy < -1 || dy > 1) return false;
    return (dx | dy) != 0 || !(flags & BC_ALL_BUT_OWNER);
}

// first position in the by-cell order whose pair's cell is >= c
__device__ __forceinline__ uint32_t bcast_lower_bound(const uint32_t* __restrict__ by_cell, const uint32_t* __restrict__ pair_cell, uint32_t n,
                                                      uint32_t c) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (pair_cell[by_cell[mid]] < c) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// WRITE = false: count[m * 9 + k] = recipients contributed by neighbour k.  WRITE = true: writes them at off[m * 9 + k].
template <bool WRITE>
__global__ void __launch_bounds__(256)
    bcast_kernel(GridDev g, BcastDev b, const uint32_t* __restrict__ n_pairs_ptr, uint64_t pair_cap, PairBuf pb,
                 const uint32_t* __restrict__ by_cell, const uint32_t* __restrict__ conn_id, const uint8_t* __restrict__ conn_type,
                 uint32_t* __restrict__ count, const uint32_t* __restrict__ off, uint32_t* __restrict__ out_slot, uint64_t out_cap,
                 unsigned long long* bump_epoch) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (!WRITE && bump_epoch && blockIdx.x == 0 && threadIdx.x == 0) *bump_epoch = chd_next_epoch(*bump_epoch);
    if (warp >= b.n * 9u) return;
    const uint32_t m = warp / 9u;
    const int k = (int)(warp % 9u);
    const uint32_t channel = b.channel[m], flags = b.flags[m];
    const uint32_t c = bcast_cell(g, channel, flags, k);
    uint32_t total = 0;
    if (c != 0xFFFFFFFFu) {
        const uint32_t np = (uint32_t)min((uint64_t)*n_pairs_ptr, pair_cap);
        uint32_t lo = 0, hi = 0;
        if (lane == 0) {
            lo = bcast_lower_bound(by_cell, pb.cell, np, c);
            hi = bcast_lower_bound(by_cell, pb.cell, np, c + 1);
        }
        lo = __shfl_sync(0xffffffffu, lo, 0);
        hi = __shfl_sync(0xffffffffu, hi, 0);
        const uint32_t sender = b.sender[m], client = b.client[m];
        const uint32_t base = WRITE ? off[warp] : 0u;
        for (uint32_t i0 = lo; i0 < hi; i0 += 32) {
            const uint32_t i = i0 + lane;
            bool keep = false;
            uint32_t s = 0;
            if (i < hi) {
                s = pb.sub[by_cell[i]];
                keep = true;
                for (uint32_t q = pb.off[s]; pb.cell[q] < c; q++)  // the run contains c itself, so this stops
                    if (bcast_member(g, channel, flags, pb.cell[q])) { keep = false; break; }  // reported for an earlier neighbour
                const uint32_t cid = conn_id[s];
                const uint8_t type = conn_type ? conn_type[s] : (uint8_t)0;
                if ((flags & BC_ALL_BUT_SENDER) && cid == sender) keep = false;  // message.go:222-224
                if ((flags & BC_ALL_BUT_CLIENT) && type == CONN_CLIENT) keep = false;  // :226-228
                if ((flags & BC_ALL_BUT_SERVER) && type == CONN_SERVER) keep = false;  // :230-232
                if (cid == client) keep = false;                                      // :234-236
            }
            const uint32_t vote = __ballot_sync(0xffffffffu, keep);
            if (WRITE && keep) {
                const uint64_t o = (uint64_t)base + total + __popc(vote & ((1u << lane) - 1u));
                if (o < out_cap) out_slot[o] = s;
            }
            total += __popc(vote);
        }
    }
    if (!WRITE && lane == 0) count[warp] = total;
}

}  // namespace chd
