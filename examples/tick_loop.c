/* tick_loop.c — the C ABI as a host sees it (plain C99; what the cgo shim in go/gpucontroller.go does, minus Go).
 * One engine, the pipelined tick loop of INTEGRATION.md §2: while tick k runs on the GPU the inputs of tick k+1 are
 * uploaded, and the results of tick k-1 are consumed (chd_fetch_results_async / chd_fetch_wait: the host never waits for a tick
 * before it has enqueued the next one).
 *
 *   gcc -std=c99 -Iinclude examples/tick_loop.c -Lchanneld_b200 -lchd_b200 -Wl,-rpath,$PWD/channeld_b200 -lm -o tick_loop
 *
 * tests/test_abi.py compiles and links this file on every run (it needs a B200 to actually execute).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "chd_gpu.h"

#define CHECK(call)                                                              \
    do {                                                                         \
        chd_status st_ = (call);                                                 \
        if (st_ != CHD_OK) {                                                     \
            fprintf(stderr, "%s -> %d: %s\n", #call, (int)st_, chd_last_error(e)); \
            return 1;                                                            \
        }                                                                        \
    } while (0)

/* two sets of page-locked staging buffers: the engine reads set (k+1)&1 while the host may already refill set k&1 */
typedef struct staging {
    double *x, *z;             /* entity positions */
    double *cx, *cz, *r;       /* one sphere query per subscriber (identity batch: query i belongs to subscriber slot i) */
    uint32_t* ring_off;        /* update rings: per-cell CSR of (arrival, sender, message index) */
    int64_t* arrival;
    uint32_t* sender;
    uint64_t *index, *channel_msg_index;
    uint32_t n_ring;
} staging;

static void fill_inputs(staging* s, uint32_t n_entities, uint32_t n_subs, uint32_t cells, int tick) {
    /* a real host copies what arrived since the last tick; here: entities on a slowly rotating lattice */
    for (uint32_t i = 0; i < n_entities; i++) {
        s->x[i] = -14000.0 + (double)((i * 37u + (uint32_t)tick * 11u) % 28000u);
        s->z[i] = -14000.0 + (double)((i * 101u + (uint32_t)tick * 7u) % 28000u);
    }
    for (uint32_t j = 0; j < n_subs; j++) {
        s->cx[j] = s->x[j * (n_entities / n_subs)];
        s->cz[j] = s->z[j * (n_entities / n_subs)];
        s->r[j] = 50.0;
    }
    for (uint32_t c = 0; c <= cells; c++) s->ring_off[c] = 0; /* no channel data updates in this sketch */
    s->n_ring = 0;
    memset(s->channel_msg_index, 0, sizeof(uint64_t) * cells);
}

int main(void) {
    const uint32_t n_entities = 100000, n_subs = 10000;
    const int64_t tick_ns = 33000000;
    chd_grid_cfg cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.world_offset_x = cfg.world_offset_z = -15000.0; /* config/spatial_static_benchmark.json */
    cfg.grid_width = cfg.grid_height = 2000.0;
    cfg.grid_cols = cfg.grid_rows = 15;
    cfg.server_cols = cfg.server_rows = 3;
    cfg.channel_id_start = 65536;
    const uint32_t cells = cfg.grid_cols * cfg.grid_rows;

    chd_limits lim;
    chd_default_limits(&cfg, n_entities, n_subs, &lim);
    chd_engine* e = NULL;
    if (chd_create(&cfg, &lim, 0, &e) != CHD_OK) {
        fprintf(stderr, "chd_create: %s\n", chd_last_error(NULL));
        return 2; /* no GPU: there is no CPU fallback */
    }

    staging st[2];
    for (int k = 0; k < 2; k++) {
        st[k].x = chd_alloc_pinned(8ull * n_entities);
        st[k].z = chd_alloc_pinned(8ull * n_entities);
        st[k].cx = chd_alloc_pinned(8ull * n_subs);
        st[k].cz = chd_alloc_pinned(8ull * n_subs);
        st[k].r = chd_alloc_pinned(8ull * n_subs);
        st[k].ring_off = chd_alloc_pinned(4ull * (cells + 1));
        st[k].arrival = chd_alloc_pinned(8);
        st[k].sender = chd_alloc_pinned(4);
        st[k].index = chd_alloc_pinned(8);
        st[k].channel_msg_index = chd_alloc_pinned(8ull * cells);
    }
    uint32_t* conn = malloc(4ull * n_subs);
    for (uint32_t j = 0; j < n_subs; j++) conn[j] = j + 1; /* connection id of subscriber slot j */
    CHECK(chd_set_subscribers(e, conn, n_subs));

    /* two sets of PINNED result buffers: the read-back of tick k (chd_fetch_results_async) overlaps tick k+1 */
    chd_result_buffers rb[2];
    void* hdr[2];
    for (int k = 0; k < 2; k++) {
        memset(&rb[k], 0, sizeof rb[k]);
        rb[k].pair_cap = rb[k].diff_cap = lim.max_pairs;
        rb[k].due_cap = lim.max_due;
        rb[k].pair_off = chd_alloc_pinned(4ull * (n_subs + 1));
        rb[k].pair_channel = chd_alloc_pinned(4ull * rb[k].pair_cap);
        rb[k].pair_interval_ms = chd_alloc_pinned(4ull * rb[k].pair_cap);
        rb[k].new_sub = chd_alloc_pinned(4ull * rb[k].diff_cap);
        rb[k].new_channel = chd_alloc_pinned(4ull * rb[k].diff_cap);
        rb[k].unsub_sub = chd_alloc_pinned(4ull * rb[k].diff_cap);
        rb[k].unsub_channel = chd_alloc_pinned(4ull * rb[k].diff_cap);
        rb[k].due = chd_alloc_pinned(sizeof(chd_due) * (uint64_t)rb[k].due_cap);
        rb[k].handover_cap = n_entities;
        rb[k].handover_entity = chd_alloc_pinned(4ull * n_entities);
        rb[k].handover_src = chd_alloc_pinned(4ull * n_entities);
        rb[k].handover_dst = chd_alloc_pinned(4ull * n_entities);
        /* the cell CSR makes the result lossless: visible(s) = concatenation of sorted_entity[cell_start[c] .. cell_start[c+1]) over
         * the subscriber's pairs */
        rb[k].cell_start = chd_alloc_pinned(4ull * (cells + 1));
        rb[k].entity_cap = n_entities;
        rb[k].sorted_entity = chd_alloc_pinned(4ull * n_entities);
        hdr[k] = chd_alloc_pinned(CHD_FETCH_HEADER_BYTES);
    }

#define PREFETCH(s)                                                                                                       \
    do {                                                                                                                  \
        chd_query_batch q_;                                                                                               \
        memset(&q_, 0, sizeof q_);                                                                                        \
        q_.n = n_subs; /* sub == NULL: identity batch; kind == NULL: all sphere */                                        \
        q_.sph_cx = (s)->cx; q_.sph_cz = (s)->cz; q_.sph_r = (s)->r;                                                      \
        CHECK(chd_prefetch_rings(e, (s)->ring_off, (s)->n_ring, (s)->arrival, (s)->sender, (s)->index, (s)->channel_msg_index)); \
        CHECK(chd_prefetch_queries(e, &q_));                                                                              \
        CHECK(chd_prefetch_entities(e, (s)->x, (s)->z, n_entities));                                                      \
    } while (0)

    fill_inputs(&st[0], n_entities, n_subs, cells, 0);
    PREFETCH(&st[0]);
    chd_tick_summary sum;
    for (int tick = 0; tick < 100; tick++) {
        const int64_t now = (int64_t)(tick + 1) * tick_ns;
        CHECK(chd_adopt_prefetched(e));            /* inputs of this tick were uploaded during the previous one */
        CHECK(chd_begin_interest(e, NULL, now, 1)); /* interest diff + fan-out pass start on the second stream */
        CHECK(chd_tick(e, NULL, now, CHD_TICK_ALL, NULL)); /* asynchronous */
        staging* next = &st[(tick + 1) & 1];
        fill_inputs(next, n_entities, n_subs, cells, tick + 1); /* the host collects tick k+1 while tick k runs */
        PREFETCH(next);
        CHECK(chd_fetch_results_async(e, &rb[tick & 1], hdr[tick & 1])); /* returns at once: device-side sizes, pinned targets */
        if (tick == 0) continue;
        /* results of tick - 1, while tick runs on the GPU */
        CHECK(chd_fetch_wait(e, &sum));
        /* apply from rb[(tick - 1) & 1]: sum.n_sub_new x handleSubToChannel(new_sub[i], new_channel[i]), sum.n_unsub x
         * handleUnsubFromChannel, sum.n_due x fanOutDataUpdate(due[i]), sum.n_handover x the orchestration half of Notify */
        if ((tick - 1) % 25 == 0)
            printf("tick %d: %llu pairs, %llu visible entries, +%u/-%u subscriptions, %u sends, %u handovers\n", tick - 1,
                   (unsigned long long)sum.n_pairs, (unsigned long long)sum.n_visible, sum.n_sub_new, sum.n_unsub, sum.n_due, sum.n_handover);
    }
    CHECK(chd_fetch_wait(e, &sum)); /* the last tick */
    chd_destroy(e);
    free(conn);
    return 0;
}
