// chd_results.cu — results: getters, the one-call read-back of a tick (chd_fetch_results), device views.
#include "chd_engine.h"

#include "chd_misc.cuh"

extern "C" {

chd_status chd_get_cells(chd_engine* e, uint32_t* cell_start, uint32_t* sorted_entity) {
    if (!e) return CHD_ERR_INVALID;
    CU(e, cudaSetDevice(e->device));
    const uint32_t C = e->g.cells;
    if (cell_start) CU(e, cudaMemcpyAsync(cell_start, e->d_cell_start, sizeof(uint32_t) * ((uint64_t)C + 1), cudaMemcpyDefault, e->stream));
    if (sorted_entity) {
        uint32_t nin = 0;
        chd_status st = chd_read_u32(e, e->d_cell_start + C, &nin);
        if (st != CHD_OK) return st;
        CU(e, cudaMemcpyAsync(sorted_entity, e->d_sorted_ent, sizeof(uint32_t) * nin, cudaMemcpyDefault, e->stream));
    }
    CU(e, cudaStreamSynchronize(e->stream));
    return CHD_OK;
}

chd_status chd_get_pairs(chd_engine* e, uint32_t* pair_off, uint32_t* channel_id, uint32_t* dist, uint32_t* interval_ms, uint8_t* flags,
                         int64_t* last_fanout_ns, uint64_t* last_message_index) {
    if (!e) return CHD_ERR_INVALID;
    CU(e, cudaSetDevice(e->device));
    PairBuf& pb = e->pairs[e->cur];
    const uint32_t S = e->n_slots;
    uint32_t P = 0;
    chd_status st = chd_read_u32(e, pb.off + S, &P);
    if (st != CHD_OK) return st;
    if (P > e->lim.max_pairs) return CHD_ERR_CAPACITY;
    cudaStream_t s = e->stream;
    if (pair_off) CU(e, cudaMemcpyAsync(pair_off, pb.off, sizeof(uint32_t) * ((uint64_t)S + 1), cudaMemcpyDefault, s));
    if (channel_id) {
        add_const_kernel<<<blocks_for(P ? P : 1, 256), 256, 0, s>>>(pb.cell, P, e->g.id_start, e->d_vcnt);
        KCHECK(e);
        CU(e, cudaMemcpyAsync(channel_id, e->d_vcnt, sizeof(uint32_t) * P, cudaMemcpyDefault, s));
    }
    if (dist) CU(e, cudaMemcpyAsync(dist, pb.dist, sizeof(uint32_t) * P, cudaMemcpyDefault, s));
    if (interval_ms) CU(e, cudaMemcpyAsync(interval_ms, pb.interval, sizeof(uint32_t) * P, cudaMemcpyDefault, s));
    if (flags) CU(e, cudaMemcpyAsync(flags, pb.flags, P, cudaMemcpyDefault, s));
    if (last_fanout_ns) CU(e, cudaMemcpyAsync(last_fanout_ns, pb.last, sizeof(int64_t) * P, cudaMemcpyDefault, s));
    if (last_message_index) CU(e, cudaMemcpyAsync(last_message_index, pb.last_index, sizeof(uint64_t) * P, cudaMemcpyDefault, s));
    CU(e, cudaStreamSynchronize(s));
    return CHD_OK;
}

chd_status chd_get_query_status(chd_engine* e, uint32_t* status, uint32_t n) {
    if (!e || !status) return CHD_ERR_INVALID;
    CU(e, cudaSetDevice(e->device));
    if (n > e->last_nq) n = e->last_nq;
    CU(e, cudaMemcpyAsync(status, e->d_status, sizeof(uint32_t) * n, cudaMemcpyDefault, e->stream));
    CU(e, cudaStreamSynchronize(e->stream));
    return CHD_OK;
}

chd_status chd_get_diff(chd_engine* e, uint32_t* new_sub, uint32_t* new_channel, uint32_t* unsub_sub, uint32_t* unsub_channel) {
    if (!e) return CHD_ERR_INVALID;
    CU(e, cudaSetDevice(e->device));
    CU(e, cudaMemcpyAsync(e->h_ctr, e->d_ctr, sizeof(Counters), cudaMemcpyDeviceToHost, e->stream));
    CU(e, cudaStreamSynchronize(e->stream));
    const uint32_t nn = e->h_ctr->n_sub_new, nu = e->h_ctr->n_unsub;
    cudaStream_t s = e->stream;
    if (new_sub) CU(e, cudaMemcpyAsync(new_sub, e->d_new_sub, sizeof(uint32_t) * nn, cudaMemcpyDefault, s));
    if (new_channel) CU(e, cudaMemcpyAsync(new_channel, e->d_new_ch, sizeof(uint32_t) * nn, cudaMemcpyDefault, s));
    if (unsub_sub) CU(e, cudaMemcpyAsync(unsub_sub, e->d_gone_sub, sizeof(uint32_t) * nu, cudaMemcpyDefault, s));
    if (unsub_channel) CU(e, cudaMemcpyAsync(unsub_channel, e->d_gone_ch, sizeof(uint32_t) * nu, cudaMemcpyDefault, s));
    CU(e, cudaStreamSynchronize(s));
    return CHD_OK;
}

chd_status chd_get_visible(chd_engine* e, uint64_t* vis_off, uint32_t* vis_entity) {
    if (!e) return CHD_ERR_INVALID;
    CU(e, cudaSetDevice(e->device));
    const uint32_t S = e->n_slots;
    cudaStream_t s = e->stream;
    if (vis_off) CU(e, cudaMemcpyAsync(vis_off, e->d_vis_off, sizeof(uint64_t) * ((uint64_t)S + 1), cudaMemcpyDefault, s));
    if (vis_entity) {
        uint64_t* h64 = (uint64_t*)e->h_get;
        CU(e, cudaMemcpyAsync(h64, e->d_vis_off + S, 8, cudaMemcpyDeviceToHost, s));
        CU(e, cudaStreamSynchronize(s));
        const uint64_t V = *h64;
        if (V > e->lim.max_visible) return CHD_ERR_CAPACITY;
        CU(e, cudaMemcpyAsync(vis_entity, e->d_vis, sizeof(uint32_t) * V, cudaMemcpyDefault, s));
    }
    CU(e, cudaStreamSynchronize(s));
    return CHD_OK;
}

chd_status chd_get_visible_slot(chd_engine* e, uint32_t slot, uint32_t* out, uint64_t cap, uint64_t* count) {
    if (!e || !count || slot >= e->n_slots) return CHD_ERR_INVALID;
    CU(e, cudaSetDevice(e->device));
    uint64_t* h64 = (uint64_t*)e->h_get;
    CU(e, cudaMemcpyAsync(h64, e->d_vis_off + slot, 16, cudaMemcpyDeviceToHost, e->stream));
    CU(e, cudaStreamSynchronize(e->stream));
    const uint64_t b = h64[0], n = h64[1] - h64[0];
    *count = n;
    if (h64[1] > e->lim.max_visible) return CHD_ERR_CAPACITY;
    const uint64_t m = n < cap ? n : cap;
    if (m && out) {
        CU(e, cudaMemcpyAsync(out, e->d_vis + b, sizeof(uint32_t) * m, cudaMemcpyDefault, e->stream));
        CU(e, cudaStreamSynchronize(e->stream));
    }
    return CHD_OK;
}

chd_status chd_get_due(chd_engine* e, chd_due* out, uint32_t cap) {
    if (!e || !out) return CHD_ERR_INVALID;
    CU(e, cudaSetDevice(e->device));
    CU(e, cudaMemcpyAsync(e->h_ctr, e->d_ctr, sizeof(Counters), cudaMemcpyDeviceToHost, e->stream));
    CU(e, cudaStreamSynchronize(e->stream));
    uint32_t n = e->h_ctr->n_due;
    if (n > e->lim.max_due) n = e->lim.max_due;  // the list overflowed (CHD_OVF_DUE): what fitted + CHD_DUE_VOID holes
    if (n > cap) n = cap;
    CU(e, cudaMemcpyAsync(out, e->d_due, sizeof(chd_due) * n, cudaMemcpyDefault, e->stream));
    CU(e, cudaStreamSynchronize(e->stream));
    return CHD_OK;
}

chd_status chd_get_handover(chd_engine* e, uint32_t* entity, uint32_t* src_channel, uint32_t* dst_channel, uint32_t cap) {
    if (!e) return CHD_ERR_INVALID;
    CU(e, cudaSetDevice(e->device));
    uint32_t n = 0;
    chd_status st = chd_read_u32(e, &e->d_ctr->n_handover, &n);
    if (st != CHD_OK) return st;
    if (n > e->ho_cap) n = e->ho_cap;
    if (n > cap) n = cap;
    cudaStream_t s = e->stream;
    if (entity) CU(e, cudaMemcpyAsync(entity, e->d_ho_entity, sizeof(uint32_t) * n, cudaMemcpyDefault, s));
    if (src_channel) CU(e, cudaMemcpyAsync(src_channel, e->d_ho_src, sizeof(uint32_t) * n, cudaMemcpyDefault, s));
    if (dst_channel) CU(e, cudaMemcpyAsync(dst_channel, e->d_ho_dst, sizeof(uint32_t) * n, cudaMemcpyDefault, s));
    CU(e, cudaStreamSynchronize(s));
    return CHD_OK;
}

chd_status chd_fetch_results(chd_engine* e, const chd_result_buffers* b, chd_tick_summary* summary) {
    if (!e || !b || !summary) return CHD_ERR_INVALID;
    chd_status st;
    cudaStream_t s = e->stream;
    cudaStream_t main_stream = e->stream;
    const bool early = e->early_ready && e->dl_stream;
    // error exits: no copy into the caller's buffers may still be in flight when this returns
    auto drain = [&](chd_status r) {
        if (e->dl_stream) cudaStreamSynchronize(e->dl_stream);
        if (e->dl_stream_b) cudaStreamSynchronize(e->dl_stream_b);
        cudaStreamSynchronize(main_stream);
        return r;
    };
    Counters ca{};  // counters as of phase A
    if (early) {
        // The tick is probably still running.  Read back on a separate stream, in the order in which results become final:
        //   phase A  after the interest fill (ev_pairs) and the build: pairs, interest diff, query statuses, handover list, cell CSR
        //   phase B  after the aux chain (ev_join: fan-out) and the emit preparation (ev_prep_done): due list, visible offsets
        // while the emit kernel is still writing the expanded list.
        CU(e, cudaSetDevice(e->device));
        s = e->dl_stream;
        CU(e, cudaStreamWaitEvent(s, e->ev_pairs, 0));
        if (e->build_done_recorded) CU(e, cudaStreamWaitEvent(s, e->ev_build_done, 0));
        CU(e, cudaMemcpyAsync(e->h_ctr, e->d_ctr, sizeof(Counters), cudaMemcpyDeviceToHost, s));
        CU(e, cudaStreamSynchronize(s));  // sync A
        ca = *e->h_ctr;
        if (ca.overflow & CHD_OVF_PAIRS) {  // the pair arrays are incomplete: report through the full summary
            CU(e, cudaStreamSynchronize(main_stream));
            return chd_summary(e, summary);
        }
    } else {
        st = chd_summary(e, summary);  // sync #1 (also surfaces capacity overflows)
        if (st != CHD_OK) return st;
        ca.n_pairs = summary->n_pairs; ca.n_sub_new = summary->n_sub_new; ca.n_unsub = summary->n_unsub;
        ca.n_handover = summary->n_handover; ca.n_entities_in_world = summary->n_entities_in_world;
    }
    PairBuf& pb = e->pairs[e->cur];
    const uint32_t S = e->n_slots;
    const uint64_t P = ca.n_pairs;
    if ((b->pair_channel || b->pair_dist || b->pair_interval_ms) && P > b->pair_cap) {
        e->fail("chd_fetch_results: %llu pairs > pair_cap %llu", (unsigned long long)P, (unsigned long long)b->pair_cap);
        return drain(CHD_ERR_CAPACITY);
    }
    if (b->pair_off) CU(e, cudaMemcpyAsync(b->pair_off, pb.off, sizeof(uint32_t) * ((uint64_t)S + 1), cudaMemcpyDefault, s));
    if (b->pair_channel && e->pair_ch_valid) {
        CU(e, cudaMemcpyAsync(b->pair_channel, e->d_pair_ch, sizeof(uint32_t) * P, cudaMemcpyDefault, s));
    } else if (b->pair_channel) {
        add_const_kernel<<<blocks_for(P ? P : 1, 256), 256, 0, main_stream>>>(pb.cell, (uint32_t)P, e->g.id_start, e->d_vcnt);
        KCHECK(e);
        CU(e, cudaMemcpyAsync(b->pair_channel, e->d_vcnt, sizeof(uint32_t) * P, cudaMemcpyDefault, main_stream));
    }
    if (b->pair_dist) CU(e, cudaMemcpyAsync(b->pair_dist, pb.dist, sizeof(uint32_t) * P, cudaMemcpyDefault, s));
    if (b->pair_interval_ms) CU(e, cudaMemcpyAsync(b->pair_interval_ms, pb.interval, sizeof(uint32_t) * P, cudaMemcpyDefault, s));
    const uint64_t nn = ca.n_sub_new, nu = ca.n_unsub;
    if ((b->new_sub || b->new_channel) && nn > b->diff_cap) return drain(CHD_ERR_CAPACITY);
    if ((b->unsub_sub || b->unsub_channel) && nu > b->diff_cap) return drain(CHD_ERR_CAPACITY);
    if (b->new_sub) CU(e, cudaMemcpyAsync(b->new_sub, e->d_new_sub, sizeof(uint32_t) * nn, cudaMemcpyDefault, s));
    if (b->new_channel) CU(e, cudaMemcpyAsync(b->new_channel, e->d_new_ch, sizeof(uint32_t) * nn, cudaMemcpyDefault, s));
    if (b->unsub_sub) CU(e, cudaMemcpyAsync(b->unsub_sub, e->d_gone_sub, sizeof(uint32_t) * nu, cudaMemcpyDefault, s));
    if (b->unsub_channel) CU(e, cudaMemcpyAsync(b->unsub_channel, e->d_gone_ch, sizeof(uint32_t) * nu, cudaMemcpyDefault, s));
    if (b->handover_entity || b->handover_src || b->handover_dst) {
        uint32_t nh = ca.n_handover;
        if (nh > e->ho_cap) nh = e->ho_cap;
        if (nh > b->handover_cap) return drain(CHD_ERR_CAPACITY);
        if (b->handover_entity) CU(e, cudaMemcpyAsync(b->handover_entity, e->d_ho_entity, sizeof(uint32_t) * nh, cudaMemcpyDefault, s));
        if (b->handover_src) CU(e, cudaMemcpyAsync(b->handover_src, e->d_ho_src, sizeof(uint32_t) * nh, cudaMemcpyDefault, s));
        if (b->handover_dst) CU(e, cudaMemcpyAsync(b->handover_dst, e->d_ho_dst, sizeof(uint32_t) * nh, cudaMemcpyDefault, s));
    }
    if (b->query_status) {
        const uint32_t n = e->last_nq < b->status_cap ? e->last_nq : b->status_cap;
        CU(e, cudaMemcpyAsync(b->query_status, e->d_status, sizeof(uint32_t) * n, cudaMemcpyDefault, s));
    }
    if (b->cell_start) CU(e, cudaMemcpyAsync(b->cell_start, e->d_cell_start, sizeof(uint32_t) * ((uint64_t)e->g.cells + 1), cudaMemcpyDefault, s));
    if (b->sorted_entity) {
        if (ca.n_entities_in_world > b->entity_cap) return drain(CHD_ERR_CAPACITY);
        CU(e, cudaMemcpyAsync(b->sorted_entity, e->d_sorted_ent, sizeof(uint32_t) * (uint64_t)ca.n_entities_in_world, cudaMemcpyDefault, s));
    }
    cudaStream_t sa = s;  // stream carrying the phase-A copies
    if (early) {  // phase B, on its own stream: its counter read-back must not queue behind the phase-A copies
        s = e->dl_stream_b;
        CU(e, cudaStreamWaitEvent(s, e->ev_join, 0));
        CU(e, cudaStreamWaitEvent(s, e->ev_prep_done, 0));
        CU(e, cudaMemcpyAsync(e->h_ctr, e->d_ctr, sizeof(Counters), cudaMemcpyDeviceToHost, s));
        CU(e, cudaStreamSynchronize(s));  // sync B
        st = chd_decode_summary(e, summary);
        if (st != CHD_OK) return drain(st);
    }
    if (b->due) {
        if (summary->n_due > b->due_cap) return drain(CHD_ERR_CAPACITY);
        CU(e, cudaMemcpyAsync(b->due, e->d_due, sizeof(chd_due) * (uint64_t)summary->n_due, cudaMemcpyDefault, s));
    }
    if (b->vis_off) CU(e, cudaMemcpyAsync(b->vis_off, e->d_vis_off, sizeof(uint64_t) * ((uint64_t)S + 1), cudaMemcpyDefault, s));
    if (b->vis_entity) {
        if (summary->n_visible > b->vis_cap) return drain(CHD_ERR_CAPACITY);
        CU(e, cudaMemcpyAsync(b->vis_entity, e->d_vis, sizeof(uint32_t) * summary->n_visible, cudaMemcpyDefault, main_stream));
    }
    CU(e, cudaStreamSynchronize(s));  // last sync of the read-back stream(s)
    if (sa != s) CU(e, cudaStreamSynchronize(sa));
    if (s != main_stream) CU(e, cudaStreamSynchronize(main_stream));  // the tick itself (expanded list) has finished
    return CHD_OK;
}

chd_status chd_device_view(chd_engine* e, int which, void** d_ptr, uint64_t* count) {
    if (!e || !d_ptr) return CHD_ERR_INVALID;
    PairBuf& pb = e->pairs[e->cur];
    uint64_t c = 0;
    void* p = nullptr;
    switch (which) {
        case CHD_VIEW_CELL_START: p = e->d_cell_start; c = (uint64_t)e->g.cells + 2; break;
        case CHD_VIEW_SORTED_ENTITY: p = e->d_sorted_ent; c = e->n_sorted; break;
        case CHD_VIEW_ENT_CELL: p = e->d_key; c = e->n_own + e->n_halo; break;
        case CHD_VIEW_PAIR_OFF: p = pb.off; c = (uint64_t)e->n_slots + 1; break;
        case CHD_VIEW_PAIR_CHANNEL: p = pb.cell; c = e->lim.max_pairs; break;
        case CHD_VIEW_PAIR_DIST: p = pb.dist; c = e->lim.max_pairs; break;
        case CHD_VIEW_VIS_OFF: p = e->d_vis_off; c = (uint64_t)e->n_slots + 1; break;
        case CHD_VIEW_VIS_ENTITY: p = e->d_vis; c = e->lim.max_visible; break;
        case CHD_VIEW_DUE: p = e->d_due; c = e->lim.max_due; break;
        default: return CHD_ERR_INVALID;
    }
    *d_ptr = p;
    if (count) *count = c;
    return CHD_OK;
}
}  // extern "C"
