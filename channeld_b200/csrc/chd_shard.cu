// chd_shard.cu — multi-GPU X-slab sharding: border export, halo import.
#include "chd_engine.h"

#include "chd_shard.cuh"

extern "C" {

chd_status chd_set_slab(chd_engine* e, uint32_t col_lo, uint32_t col_hi, uint32_t halo) {
    if (!e || col_lo >= col_hi || col_hi > e->g.cols) return CHD_ERR_INVALID;
    e->g.col_lo = col_lo;
    e->g.col_hi = col_hi;
    e->g.halo = halo;
    e->halo_on_device = false;  // set again by chd_import_halo
    return CHD_OK;
}

chd_status chd_export_border(chd_engine* e, uint32_t* d_records, uint32_t cap_records, uint32_t* out_count) {
    if (!e || !d_records) return CHD_ERR_INVALID;
    CU(e, cudaSetDevice(e->device));
    cudaStream_t s = e->stream;
    const uint32_t n = e->n_own;
    const uint32_t n_launch = n > cap_records ? n : cap_records;  // the write pass also pads the caller's buffer
    auto enqueue = [&]() -> chd_status {
        chd_status st = chd_assign_cells_impl(e);
        if (st != CHD_OK) return st;
        border_flag_kernel<<<blocks_for(n ? n : 1, 256), 256, 0, s>>>(e->g, e->d_key, n, e->d_bflag, e->d_epoch + EP_BORDER);
        KCHECK(e);
        SCAN(e, exclusive_scan_1p<uint32_t, uint32_t>(e->d_bflag, e->d_boff, n, e->site_border, s));
        border_write_kernel<<<blocks_for(n_launch ? n_launch : 1, 256), 256, 0, s>>>(e->d_key, e->have_gid ? e->d_gid : nullptr, n, e->d_bflag,
                                                                                      e->d_boff, d_records, cap_records, e->d_ctr);
        KCHECK(e);
        return CHD_OK;
    };
    chd_status st = chd_epoch_tick(e, EP_BORDER);
    if (st != CHD_OK) return st;
    if (!e->assigned) {
        // replayable: cell assignment + border selection of one tick (two variants: the key buffers ping-pong)
        uint32_t* target = e->have_prev_key ? e->d_prev_key : e->d_key;
        const int slot = (target == e->d_key_a ? 0 : 1) + 2 * e->pos_buf;
        uint64_t key = mix_key(mix_key(mix_key(0x6578706full, n), e->have_gid), e->have_prev_key);
        key = mix_key(mix_key(key, (uint64_t)(uintptr_t)target), (uint64_t)(uintptr_t)e->pos_x ^ ((uint64_t)(uintptr_t)e->pos_z << 1));
        key = mix_key(mix_key(key, (uint64_t)(uintptr_t)d_records), cap_records);
        key = mix_key(mix_key(mix_key(key, e->g.col_lo), e->g.col_hi), e->g.halo);
        st = run_stage(e, e->g_export[slot], key, enqueue);
        if (st == CHD_OK && !e->assigned) {  // replayed graph: mirror the host-side bookkeeping of chd_assign_cells
            if (e->have_prev_key) {
                uint32_t* t = e->d_key;
                e->d_key = e->d_prev_key;
                e->d_prev_key = t;
            }
            if (e->n_own) e->have_prev_key = true;
            e->n_halo = 0;
            e->assigned = true;
        }
        if (st == CHD_OK) st = chd_note_pos_read(e);
    } else {
        st = enqueue();
    }
    if (st != CHD_OK) return st;
    if (out_count) {
        st = chd_read_u32(e, e->d_boff + n, out_count);
        if (st != CHD_OK) return st;
        if (*out_count > cap_records) {
            e->fail("border export needs %u records > capacity %u", *out_count, cap_records);
            return CHD_ERR_CAPACITY;
        }
    }
    return CHD_OK;
}

chd_status chd_import_halo(chd_engine* e, const uint32_t* d_records, uint32_t n_records, uint32_t skip_first, uint32_t skip_count) {
    if (!e || (n_records && !d_records)) return CHD_ERR_INVALID;
    CU(e, cudaSetDevice(e->device));
    if (!e->assigned) {
        e->fail("chd_import_halo before chd_export_border / chd_assign_cells");
        return CHD_ERR_STATE;
    }
    if (!e->have_gid) {
        e->fail("chd_import_halo requires global entity ids (chd_set_entity_ids)");
        return CHD_ERR_STATE;
    }
    cudaStream_t s = e->stream;
    if (n_records > e->lim.max_entities) {
        e->fail("halo import of %u records > max_entities scratch %u", n_records, e->lim.max_entities);
        return CHD_ERR_CAPACITY;
    }
    {
        chd_status st0 = chd_epoch_tick(e, EP_BORDER);
        if (st0 != CHD_OK) return st0;
        const int slot = e->d_key == e->d_key_a ? 0 : 1;  // the halo keys are appended to the current key buffer
        uint64_t key = mix_key(mix_key(mix_key(0x696d706full, n_records), skip_first), skip_count);
        key = mix_key(mix_key(mix_key(key, (uint64_t)(uintptr_t)d_records), e->n_own), (uint64_t)(uintptr_t)e->d_key);
        key = mix_key(mix_key(mix_key(key, e->g.col_lo), e->g.col_hi), e->g.halo);
        chd_status st = run_stage(e, e->g_import[slot], key, [&]() -> chd_status {
            halo_flag_kernel<<<blocks_for(n_records ? n_records : 1, 256), 256, 0, s>>>(e->g, d_records, n_records, skip_first, skip_count,
                                                                                        e->d_bflag, e->d_epoch + EP_BORDER);
            KCHECK(e);
            SCAN(e, exclusive_scan_1p<uint32_t, uint32_t>(e->d_bflag, e->d_boff, n_records, e->site_border, s));
            // no host round trip: the kept count and the build length stay on the device (overflow -> CHD_OVF_BORDER)
            halo_append_kernel<<<blocks_for(n_records ? n_records : 1, 256), 256, 0, s>>>(d_records, n_records, e->d_bflag, e->d_boff, e->n_own,
                                                                                          e->lim.max_entities, e->d_key, e->d_gid, e->d_n_build,
                                                                                          e->d_ctr);
            KCHECK(e);
            return CHD_OK;
        });
        if (st != CHD_OK) return st;
    }
    e->halo_on_device = true;
    e->n_halo = 0;
    e->entities_dirty = true;
    return CHD_OK;
}
}  // extern "C"
