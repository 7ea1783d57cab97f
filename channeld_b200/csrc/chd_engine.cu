// chd_engine.cu — the engine behind include/chd_gpu.h: device memory, launch sequences, C ABI.
// All device buffers are allocated once in chd_create; the tick path allocates nothing and — apart from
// chd_summary / the chd_get_* copies — never synchronises with the host.
#include <cuda_runtime.h>

#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "chd_broadcast.cuh"
#include "chd_build.cuh"
#include "chd_classes.cuh"
#include "chd_emit.cuh"
#include "chd_fanout.cuh"
#include "chd_misc.cuh"
#include "chd_scan.cuh"

using namespace chd;

static thread_local std::string g_create_error;

struct chd_engine {
    chd_grid_cfg cfg;
    chd_limits lim;
    GridDev g;
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    // Guards (a) the shared query scratch between the stateless query entry points and the tick driver and (b) the
    // temporary redirection of `stream` to aux_stream while the interest / fan-out chain is being enqueued.
    std::recursive_mutex mu;
    mutable std::string err;
    std::vector<void*> allocs;
    int sm_count = 148;
    uint64_t n_launch = 0;  // kernels launched by this engine (bench.py reports it as gpu_launches)
    // CUDA graphs: the launch-bound small-kernel stages are captured once per (shape, parity) and replayed.
    struct GraphSlot {
        cudaGraphExec_t exec = nullptr;
        uint64_t key = 0, pending_key = 0;
        uint64_t nodes = 0;
    };
    bool use_graphs = true;
    bool overlap_fanout = true;        // chd_tick runs interest + fan-out on aux_stream concurrently with build + emit
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_interest = nullptr, ev_pairs = nullptr;
    bool interest_pending = false, pending_fanout = false;  // chd_begin_interest issued, not yet joined by chd_tick
    GraphSlot g_build[4], g_interest[4], g_interest_b[4], g_emit_prep[2], g_fanout[4], g_export[4], g_import[2];  // build / export: [key buffer][position buffer]
    uint64_t graph_launches = 0, graph_captures = 0;
    uint32_t* d_key_a = nullptr;  // identity of the first key buffer (graph slot selection)
    int64_t* d_time = nullptr;      // [0] = now_ns of the last update_interest, [1] = t_ns of the last fanout_tick
    uint32_t* d_ring_total = nullptr;
    // optional per-stage CUDA-event timing (chd_profile_*): [stage][0=start,1=stop]
    bool profiling = false;
    static constexpr int EV_RING = 1024;
    cudaEvent_t* ev = nullptr;  // [CHD_STAGE_COUNT][EV_RING][2]
    uint64_t stage_n[CHD_STAGE_COUNT] = {};
    cudaEvent_t& evt(int stage, uint64_t i, int which) { return ev[((size_t)stage * EV_RING + (size_t)(i % EV_RING)) * 2 + which]; }

    // ---- entities
    uint32_t n_own = 0, n_halo = 0;  // entities with positions / appended halo records
    bool halo_on_device = false;     // multi-GPU: the build length (own + halo) lives in d_n_build
    uint32_t* d_n_build = nullptr;
    bool have_gid = false;
    double *d_x = nullptr, *d_z = nullptr;        // engine-owned staging for host inputs (the FRONT buffers)
    // chd_prefetch_entities: the BACK buffers receive the next tick's positions on `up_stream` while the current tick runs;
    // chd_adopt_prefetched swaps front and back.  ev_pos_read[b] = last assign_cells that read buffer pair b.
    double *d_xb[2] = {nullptr, nullptr}, *d_zb[2] = {nullptr, nullptr};
    int pos_buf = 0;
    cudaStream_t up_stream = nullptr;
    cudaEvent_t ev_upload = nullptr, ev_pos_read[2] = {nullptr, nullptr};
    bool pos_read_recorded[2] = {false, false};
    bool staged = false;
    uint32_t staged_n = 0;
    const double *pos_x = nullptr, *pos_z = nullptr;  // what the kernels read: the staging buffers, or the caller's device arrays
    uint32_t *d_gid = nullptr;            // [max_entities] global ids (multi-GPU) of own + halo
    uint32_t *d_key = nullptr, *d_prev_key = nullptr;  // [max_entities] cell key per entity
    uint32_t *d_tmp_key = nullptr, *d_tmp_val = nullptr, *d_sorted_key = nullptr, *d_sorted_ent = nullptr;
    uint32_t *d_cell_start = nullptr;     // [C+2]
    uint32_t *d_sorted4 = nullptr;        // 4 phase-shifted copies of d_sorted_ent (chd_emit.cuh), stride = phase_stride
    uint32_t phase_stride = 0;
    uint32_t *d_hist = nullptr;           // [BUILD_MAX_BINS * nblocks + 1]
    // one look-back scan site per call site: stages run concurrently on two streams and must not share scan state
    unsigned long long* d_epoch = nullptr;  // [EP_COUNT] stage epochs for the look-back scans
    ScanSite site_hist_b{}, site_pchist_b{}, site_hist{}, site_win{}, site_qoff{}, site_slot{}, site_pchist{}, site_voff{}, site_uoff{}, site_border{};
    uint32_t build_blocks = 0;
    bool assigned = false, built = false, have_prev_key = false, entities_dirty = false;
    uint32_t n_sorted = 0;
    // handover
    uint32_t *d_ho_entity = nullptr, *d_ho_src = nullptr, *d_ho_dst = nullptr;
    uint32_t ho_cap = 0;

    // ---- subscribers / pairs
    uint32_t n_slots = 0;
    uint32_t* d_conn = nullptr;
    PairBuf pairs[2];
    int cur = 0;
    // ---- query scratch
    struct QStage {
        uint32_t *sub; uint8_t* kind;
        double *sph_cx, *sph_cz, *sph_r, *box_cx, *box_cz, *box_ex, *box_ez, *cone_cx, *cone_cz, *cone_dx, *cone_dz, *cone_angle, *cone_r;
        uint32_t *spot_off, *spot_ndist; double *spot_x, *spot_z; uint32_t* spot_dist;
    } dq{};
    // chd_prefetch_queries / chd_prefetch_rings: two dedicated staging sets each (allocated on first use), filled on
    // up_stream while a tick is in flight and handed to the next tick by chd_adopt_prefetched
    QStage dq_pf[2] = {};
    bool dq_pf_alloc[2] = {false, false};
    int q_next = 0;                       // set the next chd_prefetch_queries fills
    bool staged_q = false, have_adopted_q = false, wait_q = false;
    int staged_q_set = 0, adopted_q_set = 0;
    QueryDev staged_qd{}, adopted_qd{};
    cudaEvent_t ev_upload_q = nullptr, ev_q_read[2] = {nullptr, nullptr};
    bool q_read_recorded[2] = {false, false};
    struct RStage {
        uint32_t *off, *sender; int64_t* arrival; uint64_t *index, *cmi;
    } ring_pf[2] = {};
    bool ring_pf_alloc[2] = {false, false};
    int ring_next = 0, staged_ring_set = 0, ring_set = -1;  // ring_set: prefetch set the current ring pointers refer to (-1: none)
    bool staged_rings = false, staged_ring_cmi = false, wait_rings = false;
    uint32_t staged_ring_total = 0;
    cudaEvent_t ev_upload_rings = nullptr, ev_ring_read[2] = {nullptr, nullptr};
    bool ring_read_recorded[2] = {false, false};
    Bbox* d_bbox = nullptr;
    uint32_t *d_win_size = nullptr, *d_window = nullptr, *d_side_cell = nullptr, *d_side_dist = nullptr, *d_side_cnt = nullptr;
    uint64_t* d_win_off = nullptr;
    uint32_t *d_status = nullptr, *d_qcount = nullptr;
    uint64_t* d_qoff = nullptr;  // stateless query CSR offsets
    uint32_t *d_qout_id = nullptr, *d_qout_dist = nullptr;
    int32_t* d_slot_query = nullptr;
    uint32_t last_nq = 0;
    // diff
    uint32_t *d_new_off = nullptr;  // scratch for u64 -> u32 offset narrowing (stateless query path)
    uint32_t *d_new_sub = nullptr, *d_new_ch = nullptr, *d_gone_sub = nullptr, *d_gone_ch = nullptr;
    // emit
    uint32_t *d_vcnt = nullptr, *d_first_pair = nullptr, *d_vis = nullptr, *d_uoff = nullptr;
    uint32_t* d_pair_ch = nullptr;  // channel id of every current pair (written by interest_fill_kernel)
    // ---- window classes of the due list (chd_due_classes): keys written by the fan-out kernel, scratch allocated on first use
    DueKey* d_due_key = nullptr;
    uint32_t *d_cls_table = nullptr, *d_cls_slot = nullptr, *d_cls_rep = nullptr, *d_cls_cnt = nullptr, *d_cls_flag = nullptr, *d_cls_rank = nullptr,
             *d_cls_of = nullptr, *d_cls_out_rep = nullptr, *d_cls_out_cnt = nullptr;
    uint32_t cls_table_size = 0;
    ScanSite site_class{};
    // ---- ADJACENT_CHANNELS broadcast sets (chd_adjacent_broadcast): scratch allocated on first use, grown on demand
    uint8_t* d_conn_type = nullptr;
    bool have_conn_type = false, by_cell_valid = false;
    uint32_t *d_bc_in = nullptr, *d_bc_cnt = nullptr, *d_bc_off = nullptr, *d_bc_msgoff = nullptr, *d_bc_out = nullptr;
    uint64_t bc_msg_cap = 0, bc_out_cap = 0;
    ScanSite site_bcast{};
    bool pair_ch_valid = false;
    // chd_fetch_results reads back on its own stream as soon as the aux chain (pairs, diff, due list) and the emit
    // preparation (visible offsets) are done, i.e. while the emit kernel is still streaming
    cudaStream_t dl_stream = nullptr, dl_stream_b = nullptr;  // phase A / phase B of the early read-back
    cudaEvent_t ev_prep_done = nullptr, ev_build_done = nullptr;
    bool build_done_recorded = false;  // this tick ran a build (its end is ev_build_done)
    bool early_ready = false;  // ev_join + ev_prep_done of the last tick are recorded
    bool trace_fetch = false;  // CHD_TRACE_FETCH=1: host-side phase times of chd_fetch_results, printed by chd_destroy
    double fetch_t[4] = {0, 0, 0, 0};
    uint64_t fetch_n = 0;
    bool early_results_tick = false;  // CHD_TICK_EARLY_RESULTS of the tick being enqueued
    EmitUnit* d_units = nullptr;  // v5 copy-unit descriptors
    uint64_t unit_cap = 0;
    // 3 = output-ordered warp tiles (default: 0.355 ms on config #2); 5 = cell-grouped copy units with L1-resident sources
    // (experimental: fewer instructions and less L2 traffic, but it scatters the write stream: 0.42-0.46 ms, profiles/README.md)
    int emit_variant = 3;
    int emit_blocks_per_sm = 4;
    // CTAs withheld from the emit grid so that the concurrent aux-stream kernels (fan-out, pair grouping) find free SM
    // slots instead of queueing behind the saturating emit kernel (measured: profiles/README.md)
    int emit_grid_reduce = 0;
    // grid = waves x the resident capacity: with more than one wave the emit CTAs retire as they go, so the high-priority
    // aux-stream kernels get SM slots at the first wave boundary instead of after the whole kernel
    int emit_waves = 1;
    // Where the aux chain (interest part 1 + fan-out) is joined: before the emit kernel (it then never competes with the
    // saturating emit kernel for SM slots) or after it (overlap).  Measured: profiles/README.md.
    bool join_before_emit = false;
    // CHD_TICK_EARLY_RESULTS strategy (CHD_EARLY_MODE): 1 = emit kernel after the aux chain, 0 = emit in waves,
    // 2 = emit grid short of `early_reduce` CTAs so the aux chain finds free SM slots next to it (CHD_EARLY_REDUCE)
    int early_mode = 1;
    int early_reduce = 32;
    cudaEvent_t wait_before_emit_kernel = nullptr;  // 4 x 256 threads x 64 registers fill an SM; 3 leaves room for the aux-stream kernels
    uint64_t *d_voff = nullptr, *d_vis_off = nullptr;
    uint64_t max_tiles = 0;
    // fanout
    uint32_t *d_ring_off = nullptr, *d_ring_sender = nullptr;
    int64_t* d_ring_arrival = nullptr;
    uint64_t *d_ring_index = nullptr, *d_ch_msg_index = nullptr;
    bool have_ch_msg_index = false;
    // what the fan-out kernel reads: the staging copies above or the caller's device arrays (zero-copy)
    const uint32_t *ring_off_p = nullptr, *ring_sender_p = nullptr;
    const int64_t* ring_arrival_p = nullptr;
    const uint64_t *ring_index_p = nullptr, *ch_msg_index_p = nullptr;

    uint32_t *d_by_cell = nullptr, *d_pc_hist = nullptr, *d_pc_tmp_key = nullptr, *d_pc_tmp_val = nullptr;  // pairs grouped by cell
    uint32_t pc_blocks = 0;
    chd_due* d_due = nullptr;
    // counters
    Counters* d_ctr = nullptr;
    Counters* h_ctr = nullptr;  // pinned
    // border export scratch
    uint32_t *d_bflag = nullptr, *d_boff = nullptr;
    uint32_t* h_u32 = nullptr;  // pinned scalar
    // scratch of the tick driver's getters (chd_get_visible*, export counts): separate from h_u32, which the stateless entry
    // points use under the engine mutex from other threads
    uint32_t* h_get = nullptr;

    bool fail(const char* fmt, ...) const {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        err = buf;
        return false;
    }
};

#define CU(e, call)                                                                                  \
    do {                                                                                             \
        cudaError_t _r = (call);                                                                     \
        if (_r != cudaSuccess) {                                                                     \
            (e)->fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_r), __FILE__, __LINE__);   \
            return CHD_ERR_CUDA;                                                                     \
        }                                                                                            \
    } while (0)

// one kernel launch precedes every KCHECK; scans report their own launch count through SCAN()
#define KCHECK(e)                    \
    do {                             \
        (e)->n_launch++;             \
        CU(e, cudaGetLastError());   \
    } while (0)
#define SCAN(e, ...)                                  \
    do {                                              \
        (e)->n_launch += (uint64_t)(__VA_ARGS__);     \
        CU(e, cudaGetLastError());                    \
    } while (0)

struct StageTimer {  // records a CUDA-event pair around a stage on the engine stream when profiling is on
    chd_engine* e;
    int stage;
    StageTimer(chd_engine* e_, int stage_) : e(e_), stage(stage_) {
        if (e->profiling) cudaEventRecord(e->evt(stage, e->stage_n[stage], 0), e->stream);
    }
    ~StageTimer() {
        if (e->profiling) {
            cudaEventRecord(e->evt(stage, e->stage_n[stage], 1), e->stream);
            e->stage_n[stage]++;
        }
    }
};

template <typename T>
static bool dalloc(chd_engine* e, T** p, uint64_t count) {
    void* q = nullptr;
    const uint64_t bytes = (count ? count : 1) * sizeof(T);
    cudaError_t r = cudaMalloc(&q, bytes);
    if (r != cudaSuccess) {
        e->fail("cudaMalloc(%llu bytes) failed: %s", (unsigned long long)bytes, cudaGetErrorString(r));
        return false;
    }
    e->allocs.push_back(q);
    *p = (T*)q;
    return true;
}

enum { EP_BUILD = 0, EP_QUERY, EP_EMIT, EP_FANOUT, EP_BORDER, EP_BCAST, EP_CLASS, EP_COUNT };  // d_epoch has 8 slots

// epochs start at 1 so that the zero-initialised descriptors (epoch 0) read as stale on first use
static bool init_epochs(chd_engine* e) {
    const unsigned long long ones[8] = {1, 1, 1, 1, 1, 1, 1, 1};
    return cudaMemcpy(e->d_epoch, ones, sizeof ones, cudaMemcpyHostToDevice) == cudaSuccess;
}

static bool make_site(chd_engine* e, ScanSite& site, uint64_t n_max, int stage) {
    site.tiles = n_max == 0 ? 1 : (n_max + SCAN_TILE - 1) / SCAN_TILE;
    site.epoch = e->d_epoch + stage;
    site.error = nullptr;  // set once d_ctr exists
    if (!dalloc(e, &site.desc, site.tiles)) return false;
    return cudaMemset(site.desc, 0, site.tiles * 8) == cudaSuccess;
}

// Zero-copy inputs: a pointer into this device's memory is consumed in place (no staging copy); the caller keeps it valid
// and unmodified until the work that reads it has finished (chd_summary / chd_sync / any chd_get_*).
static bool is_device_ptr(const chd_engine* e, const void* p) {
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice && a.device == e->device;
}

static inline unsigned blocks_for(uint64_t n, unsigned threads) { return (unsigned)((n + threads - 1) / threads); }

__global__ void set_i64_kernel(int64_t* dst, int64_t v) { *dst = v; }
// first kernel of the interest / fan-out stages: publishes the tick time and opens a new scan epoch
__global__ void stage_begin_kernel(int64_t* dst, int64_t v, unsigned long long* epoch, uint32_t* zero_me) {
    *dst = v;
    *epoch = (*epoch + 1) & ((1ull << 22) - 1);
    if (zero_me) *zero_me = 0;  // fan-out: the due-list cursor (n_due)
}
__global__ void set_u32_kernel(uint32_t* dst, uint32_t v) { *dst = v; }

static inline uint64_t mix_key(uint64_t h, uint64_t v) {
    h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    return h;
}

// Runs `enqueue` (kernel launches + memsets only, no host sync) either directly or as a replayed CUDA graph.
// A graph is captured only after the same key was seen twice in a row, so workloads whose batch shape changes
// every tick simply run direct launches.
template <typename F>
static chd_status run_stage(chd_engine* e, chd_engine::GraphSlot& slot, uint64_t key, F&& enqueue) {
    key |= 1;  // never 0
    if (!e->use_graphs || e->stream == nullptr) return enqueue();
    if (slot.exec && slot.key == key) {
        CU(e, cudaGraphLaunch(slot.exec, e->stream));
        e->n_launch += slot.nodes;
        e->graph_launches++;
        return CHD_OK;
    }
    if (slot.pending_key != key) {  // first sighting: run direct, capture next time
        slot.pending_key = key;
        return enqueue();
    }
    if (slot.exec) {
        cudaGraphExecDestroy(slot.exec);
        slot.exec = nullptr;
    }
    const uint64_t l0 = e->n_launch;
    // Stateless entry points (chd_cell_of, chd_query_channel_ids) may be called from other threads and launch into the same
    // stream: they hold the engine mutex for their whole call, so taking it here keeps their work out of the capture.
    std::lock_guard<std::recursive_mutex> capture_lock(e->mu);
    CU(e, cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
    chd_status st = enqueue();
    cudaGraph_t g = nullptr;
    cudaError_t r = cudaStreamEndCapture(e->stream, &g);
    if (st != CHD_OK || r != cudaSuccess || !g) {
        if (g) cudaGraphDestroy(g);
        if (st == CHD_OK) {
            e->fail("graph capture failed: %s", cudaGetErrorString(r));
            st = CHD_ERR_CUDA;
        }
        e->use_graphs = false;  // fall back to direct launches for the rest of this engine's life
        cudaGetLastError();
        e->n_launch = l0;
        return st == CHD_OK ? enqueue() : st;
    }
    r = cudaGraphInstantiate(&slot.exec, g, 0);
    cudaGraphDestroy(g);
    if (r != cudaSuccess) {
        slot.exec = nullptr;
        e->use_graphs = false;
        cudaGetLastError();
        e->n_launch = l0;
        return enqueue();
    }
    slot.nodes = e->n_launch - l0;
    e->n_launch = l0;
    slot.key = key;
    e->graph_captures++;
    CU(e, cudaGraphLaunch(slot.exec, e->stream));
    e->n_launch += slot.nodes;
    e->graph_launches++;
    return CHD_OK;
}

extern "C" {

uint32_t chd_abi_version(void) { return CHD_ABI_VERSION; }

uint32_t chd_damping_interval_ms(uint32_t dist, uint32_t default_ms) { return damping_interval_ms(dist, default_ms); }

void chd_default_limits(const chd_grid_cfg* cfg, uint32_t n_entities, uint32_t n_subscribers, chd_limits* lim) {
    memset(lim, 0, sizeof *lim);
    const uint64_t cells = (uint64_t)cfg->grid_cols * cfg->grid_rows;
    lim->max_entities = n_entities ? n_entities : 1;
    lim->max_subscribers = n_subscribers ? n_subscribers : 1;
    lim->max_queries = lim->max_subscribers;
    lim->max_spots = 1024;
    lim->max_pairs = (uint64_t)lim->max_subscribers * 16 + 1024;
    lim->max_window_cells = (uint64_t)lim->max_queries * 32 + 65536;
    const double per_cell = (double)n_entities / (double)(cells ? cells : 1);
    double v = (double)n_subscribers * 1.6 * per_cell * 1.5 + 1048576.0;
    if (v > 3.0e9) v = 3.0e9;
    lim->max_visible = (uint64_t)v;
    lim->max_ring_entries = (uint32_t)(cells * 64 < 1048576 ? 1048576 : (cells * 64 > 0x7fffffffull ? 0x7fffffffull : cells * 64));
    lim->max_due = (uint32_t)(lim->max_pairs * 2 > 0x7fffffffull ? 0x7fffffffull : lim->max_pairs * 2);
    lim->default_fanout_interval_ms = 20;  // GLOBAL defaults, settings.go:97-103
    lim->default_fanout_delay_ms = 0;
}

const char* chd_last_error(const chd_engine* e) { return e ? e->err.c_str() : g_create_error.c_str(); }

void* chd_alloc_pinned(uint64_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
void chd_free_pinned(void* p) {
    if (p) cudaFreeHost(p);
}

void chd_destroy(chd_engine* e) {
    if (!e) return;
    if (e->trace_fetch && e->fetch_n)
        fprintf(stderr, "[chd] fetch_results x%llu: phase-A ready %.1f us, phase A+B enqueue (incl. wait for B) %.1f us, copies done %.1f us, tick done %.1f us\n",
                (unsigned long long)e->fetch_n, e->fetch_t[0] / e->fetch_n, e->fetch_t[1] / e->fetch_n, e->fetch_t[2] / e->fetch_n,
                e->fetch_t[3] / e->fetch_n);
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    if (e->up_stream) cudaStreamSynchronize(e->up_stream);
    for (void* p : e->allocs) cudaFree(p);
    for (auto* arr : {e->g_emit_prep, e->g_import})
        for (int i = 0; i < 2; i++)
            if (arr[i].exec) cudaGraphExecDestroy(arr[i].exec);
    for (auto* arr : {e->g_build, e->g_export, e->g_interest, e->g_interest_b, e->g_fanout})
        for (int i = 0; i < 4; i++)
            if (arr[i].exec) cudaGraphExecDestroy(arr[i].exec);
    for (cudaEvent_t ev : {e->ev_upload_q, e->ev_q_read[0], e->ev_q_read[1], e->ev_upload_rings, e->ev_ring_read[0], e->ev_ring_read[1]})
        if (ev) cudaEventDestroy(ev);
    if (e->up_stream) {
        cudaStreamSynchronize(e->up_stream);
        cudaStreamDestroy(e->up_stream);
    }
    if (e->ev_upload) cudaEventDestroy(e->ev_upload);
    for (int i = 0; i < 2; i++)
        if (e->ev_pos_read[i]) cudaEventDestroy(e->ev_pos_read[i]);
    if (e->ev) {
        for (size_t i = 0; i < (size_t)CHD_STAGE_COUNT * chd_engine::EV_RING * 2; i++)
            if (e->ev[i]) cudaEventDestroy(e->ev[i]);
        delete[] e->ev;
    }
    if (e->h_ctr) cudaFreeHost(e->h_ctr);
    if (e->h_u32) cudaFreeHost(e->h_u32);
    if (e->h_get) cudaFreeHost(e->h_get);
    if (e->aux_stream) cudaStreamDestroy(e->aux_stream);
    if (e->dl_stream) cudaStreamDestroy(e->dl_stream);
    if (e->dl_stream_b) cudaStreamDestroy(e->dl_stream_b);
    if (e->ev_prep_done) cudaEventDestroy(e->ev_prep_done);
    if (e->ev_build_done) cudaEventDestroy(e->ev_build_done);
    if (e->ev_fork) cudaEventDestroy(e->ev_fork);
    if (e->ev_join) cudaEventDestroy(e->ev_join);
    if (e->ev_interest) cudaEventDestroy(e->ev_interest);
    if (e->ev_pairs) cudaEventDestroy(e->ev_pairs);
    if (e->own_stream && e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

static bool alloc_pairbuf(chd_engine* e, PairBuf& pb) {
    const uint64_t P = e->lim.max_pairs;
    return dalloc(e, &pb.off, (uint64_t)e->lim.max_subscribers + 1) && dalloc(e, &pb.sub, P) && dalloc(e, &pb.cell, P) &&
           dalloc(e, &pb.dist, P) && dalloc(e, &pb.interval, P) && dalloc(e, &pb.flags, P) && dalloc(e, &pb.last, P) &&
           dalloc(e, &pb.last_index, P);
}

chd_status chd_create(const chd_grid_cfg* cfg, const chd_limits* lim_in, int device, chd_engine** out) {
    if (!cfg || !out) {
        g_create_error = "null argument";
        return CHD_ERR_INVALID;
    }
    *out = nullptr;
    // LoadConfig validation (spatial.go:146-154); ServerInterestBorderSize == 0 tolerated (see header)
    if (!(cfg->grid_width > 0) || !(cfg->grid_height > 0)) {
        g_create_error = "GridWidth and GridHeight should be positive";
        return CHD_ERR_INVALID;
    }
    if (cfg->grid_cols == 0 || cfg->grid_rows == 0) {
        g_create_error = "GridCols and GridRows should be positive";
        return CHD_ERR_INVALID;
    }
    if (cfg->server_cols == 0 || cfg->server_rows == 0) {
        g_create_error = "ServerCols and ServerRows should be positive";
        return CHD_ERR_INVALID;
    }
    const uint64_t cells64 = (uint64_t)cfg->grid_cols * cfg->grid_rows;
    if (cells64 >= (1ull << 20)) {  // spatial id space is [0x10000, 0x80000) (settings.go:94-95): < 2^19 cells
        g_create_error = "too many cells (GridCols*GridRows must be < 2^20)";
        return CHD_ERR_INVALID;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        g_create_error = "no CUDA device: this engine has no CPU fallback";
        return CHD_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) {
        g_create_error = "bad device ordinal";
        return CHD_ERR_INVALID;
    }
    chd_engine* e = new (std::nothrow) chd_engine();
    if (!e) return CHD_ERR_INVALID;
    e->cfg = *cfg;
    if (lim_in)
        e->lim = *lim_in;
    else
        chd_default_limits(cfg, 1u << 20, 1u << 17, &e->lim);
    chd_limits& L = e->lim;
    if (!L.max_entities) L.max_entities = 1;
    if (!L.max_subscribers) L.max_subscribers = 1;
    if (!L.max_queries) L.max_queries = L.max_subscribers;
    if (!L.max_pairs) L.max_pairs = 1024;
    if (!L.max_window_cells) L.max_window_cells = 65536;
    if (!L.max_visible) L.max_visible = 1u << 20;
    if (!L.max_ring_entries) L.max_ring_entries = 1u << 16;
    if (!L.max_due) L.max_due = 1u << 16;
    if (L.max_pairs >= 0xFFFFFFF0ull || L.max_window_cells >= (1ull << 40) || L.max_entities >= (1u << 30)) {
        g_create_error = "limits too large (pairs are indexed with 32 bits)";
        delete e;
        return CHD_ERR_INVALID;
    }
    e->device = device;
    if (const char* v = getenv("CHD_JOIN_BEFORE_EMIT")) e->join_before_emit = atoi(v) != 0;
    if (const char* v = getenv("CHD_EARLY_MODE")) e->early_mode = atoi(v) >= 0 && atoi(v) <= 2 ? atoi(v) : 1;
    if (const char* v = getenv("CHD_EARLY_REDUCE")) e->early_reduce = atoi(v) >= 0 && atoi(v) < 400 ? atoi(v) : 32;
    if (const char* v = getenv("CHD_EMIT_GRID_REDUCE")) e->emit_grid_reduce = atoi(v) >= 0 && atoi(v) < 400 ? atoi(v) : 0;
    if (const char* v = getenv("CHD_EMIT_BPS")) e->emit_blocks_per_sm = atoi(v) >= 1 && atoi(v) <= 4 ? atoi(v) : 4;
    if (const char* v = getenv("CHD_TRACE_FETCH")) e->trace_fetch = atoi(v) != 0;
    if (const char* v = getenv("CHD_EMIT_WAVES")) e->emit_waves = atoi(v) >= 1 && atoi(v) <= 64 ? atoi(v) : 1;
    if (const char* v = getenv("CHD_EMIT_VARIANT")) e->emit_variant = atoi(v) == 5 ? 5 : 3;
#define CCU(call)                                                                       \
    do {                                                                                \
        cudaError_t _r = (call);                                                        \
        if (_r != cudaSuccess) {                                                        \
            g_create_error = std::string(#call " failed: ") + cudaGetErrorString(_r);   \
            chd_destroy(e);                                                             \
            return CHD_ERR_CUDA;                                                        \
        }                                                                               \
    } while (0)
    CCU(cudaSetDevice(device));
    CCU(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    e->own_stream = true;
    {
        // the aux stream carries short latency-bound kernels that should slot in ahead of the long emit kernel
        int lo_prio = 0, hi_prio = 0;
        CCU(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
        CCU(cudaStreamCreateWithPriority(&e->aux_stream, cudaStreamNonBlocking, hi_prio));
    }
    CCU(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
    CCU(cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming));
    CCU(cudaEventCreateWithFlags(&e->ev_interest, cudaEventDisableTiming));
    CCU(cudaEventCreateWithFlags(&e->ev_pairs, cudaEventDisableTiming));
    CCU(cudaEventCreateWithFlags(&e->ev_prep_done, cudaEventDisableTiming));
    CCU(cudaEventCreateWithFlags(&e->ev_build_done, cudaEventDisableTiming));
    {
        int lo_prio = 0, hi_prio2 = 0;
        CCU(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio2));
        CCU(cudaStreamCreateWithPriority(&e->dl_stream, cudaStreamNonBlocking, hi_prio2));
        CCU(cudaStreamCreateWithPriority(&e->dl_stream_b, cudaStreamNonBlocking, hi_prio2));
    }
    CCU(cudaDeviceGetAttribute(&e->sm_count, cudaDevAttrMultiProcessorCount, device));

    GridDev& g = e->g;
    g.off_x = cfg->world_offset_x; g.off_z = cfg->world_offset_z; g.w = cfg->grid_width; g.h = cfg->grid_height;
    g.grid_size = std::sqrt(g.w * g.w + g.h * g.h);                  // spatial.go:134-139
    g.world_x_hi = g.off_x + g.w * (double)cfg->grid_cols;           // spatial.go:126-132,287
    g.world_z_hi = g.off_z + g.h * (double)cfg->grid_rows;
    g.fcols = (double)cfg->grid_cols; g.frows = (double)cfg->grid_rows;
    g.cols = cfg->grid_cols; g.rows = cfg->grid_rows; g.cells = (uint32_t)cells64; g.id_start = cfg->channel_id_start;
    g.col_lo = 0; g.col_hi = g.cols; g.halo = 0;
    g.default_interval_ms = L.default_fanout_interval_ms; g.default_delay_ms = L.default_fanout_delay_ms;

    const uint64_t N = L.max_entities, S = L.max_subscribers, Q = L.max_queries, P = L.max_pairs, C = g.cells;
    e->build_blocks = (uint32_t)e->sm_count * 4;
    e->phase_stride = (uint32_t)(((N + 3) / 4) * 4 + 8);
    e->pc_blocks = (uint32_t)std::min<uint64_t>(4096, (P + BUILD_TILE - 1) / BUILD_TILE);
    if (e->pc_blocks == 0) e->pc_blocks = 1;
    e->ho_cap = L.max_entities;
    e->max_tiles = (L.max_visible + EMIT_TILE - 1) / EMIT_TILE + 1;
    e->unit_cap = L.max_visible / EMIT_UNIT + L.max_pairs + 1;
    uint64_t scan_n = (uint64_t)BUILD_MAX_BINS * e->build_blocks + 1;
    if (P + 1 > scan_n) scan_n = P + 1;
    if (Q + 1 > scan_n) scan_n = Q + 1;
    if (S + 1 > scan_n) scan_n = S + 1;
    if (N + 1 > scan_n) scan_n = N + 1;
    bool ok = true;
    ok = ok && dalloc(e, &e->d_x, N) && dalloc(e, &e->d_z, N) && (e->d_xb[0] = e->d_x, e->d_zb[0] = e->d_z, true) &&
         cudaEventCreateWithFlags(&e->ev_pos_read[0], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&e->ev_pos_read[1], cudaEventDisableTiming) == cudaSuccess && dalloc(e, &e->d_gid, N) && dalloc(e, &e->d_key, N) &&
         dalloc(e, &e->d_prev_key, N) && dalloc(e, &e->d_tmp_key, N) && dalloc(e, &e->d_tmp_val, N) &&
         dalloc(e, &e->d_sorted_key, N) && dalloc(e, &e->d_sorted4, 4 * (((N + 3) / 4) * 4 + 8)) && dalloc(e, &e->d_cell_start, C + 2) &&
         dalloc(e, &e->d_hist, (uint64_t)BUILD_MAX_BINS * e->build_blocks + 2) &&
         dalloc(e, &e->d_epoch, 8) && init_epochs(e) &&
         make_site(e, e->site_hist, (uint64_t)BUILD_MAX_BINS * e->build_blocks + 1, EP_BUILD) &&
         make_site(e, e->site_hist_b, (uint64_t)BUILD_MAX_BINS * e->build_blocks + 1, EP_BUILD) && make_site(e, e->site_win, Q + 1, EP_QUERY) &&
         make_site(e, e->site_qoff, Q + 1, EP_QUERY) && make_site(e, e->site_slot, S + 1, EP_QUERY) &&
         make_site(e, e->site_voff, P + 1, EP_EMIT) && make_site(e, e->site_uoff, P + 1, EP_EMIT) &&
         make_site(e, e->site_border, N + 1, EP_BORDER) &&
         dalloc(e, &e->d_ho_entity, N) && dalloc(e, &e->d_ho_src, N) && dalloc(e, &e->d_ho_dst, N) &&
         dalloc(e, &e->d_bflag, N + 1) && dalloc(e, &e->d_boff, N + 2);
    e->d_sorted_ent = e->d_sorted4;  // phase copy 0 IS the plain sorted entity array
    e->d_key_a = e->d_key;
    ok = ok && dalloc(e, &e->d_conn, S) && alloc_pairbuf(e, e->pairs[0]) && alloc_pairbuf(e, e->pairs[1]);
    ok = ok && dalloc(e, &e->dq.sub, Q) && dalloc(e, &e->dq.kind, Q) && dalloc(e, &e->dq.sph_cx, Q) && dalloc(e, &e->dq.sph_cz, Q) &&
         dalloc(e, &e->dq.sph_r, Q) && dalloc(e, &e->dq.box_cx, Q) && dalloc(e, &e->dq.box_cz, Q) && dalloc(e, &e->dq.box_ex, Q) &&
         dalloc(e, &e->dq.box_ez, Q) && dalloc(e, &e->dq.cone_cx, Q) && dalloc(e, &e->dq.cone_cz, Q) && dalloc(e, &e->dq.cone_dx, Q) &&
         dalloc(e, &e->dq.cone_dz, Q) && dalloc(e, &e->dq.cone_angle, Q) && dalloc(e, &e->dq.cone_r, Q) &&
         dalloc(e, &e->dq.spot_off, Q + 1) && dalloc(e, &e->dq.spot_ndist, Q) && dalloc(e, &e->dq.spot_x, (uint64_t)L.max_spots) &&
         dalloc(e, &e->dq.spot_z, (uint64_t)L.max_spots) && dalloc(e, &e->dq.spot_dist, (uint64_t)L.max_spots);
    ok = ok && dalloc(e, &e->d_bbox, Q) && dalloc(e, &e->d_win_size, Q) && dalloc(e, &e->d_win_off, Q + 1) &&
         dalloc(e, &e->d_window, L.max_window_cells) && dalloc(e, &e->d_side_cell, (uint64_t)L.max_spots) &&
         dalloc(e, &e->d_side_dist, (uint64_t)L.max_spots) && dalloc(e, &e->d_side_cnt, Q) && dalloc(e, &e->d_status, Q) &&
         dalloc(e, &e->d_qcount, Q) && dalloc(e, &e->d_qoff, Q + 1) && dalloc(e, &e->d_qout_id, P) && dalloc(e, &e->d_qout_dist, P) &&
         dalloc(e, &e->d_slot_query, S);
    ok = ok && dalloc(e, &e->d_new_off, (Q > P ? Q : P) + 2) && dalloc(e, &e->d_new_sub, P) && dalloc(e, &e->d_new_ch, P) && dalloc(e, &e->d_gone_sub, P) &&
         dalloc(e, &e->d_gone_ch, P);
    ok = ok && dalloc(e, &e->d_pair_ch, P) && dalloc(e, &e->d_vcnt, P) && dalloc(e, &e->d_uoff, P + 2) && dalloc(e, &e->d_units, e->unit_cap) && dalloc(e, &e->d_voff, P + 1) && dalloc(e, &e->d_first_pair, e->max_tiles + 1) &&
         dalloc(e, &e->d_vis_off, S + 1) && dalloc(e, &e->d_vis, L.max_visible);
    ok = ok && dalloc(e, &e->d_ring_off, C + 1) && dalloc(e, &e->d_ring_arrival, (uint64_t)L.max_ring_entries) &&
         dalloc(e, &e->d_ring_sender, (uint64_t)L.max_ring_entries) && dalloc(e, &e->d_ring_index, (uint64_t)L.max_ring_entries) &&
         dalloc(e, &e->d_ch_msg_index, C) && dalloc(e, &e->d_by_cell, P) &&
         dalloc(e, &e->d_pc_hist, (uint64_t)BUILD_MAX_BINS * e->pc_blocks + 2) && dalloc(e, &e->d_pc_tmp_key, P) && dalloc(e, &e->d_pc_tmp_val, P) &&
         make_site(e, e->site_pchist, (uint64_t)BUILD_MAX_BINS * e->pc_blocks + 1, EP_QUERY) &&
         make_site(e, e->site_pchist_b, (uint64_t)BUILD_MAX_BINS * e->pc_blocks + 1, EP_QUERY) &&
         dalloc(e, &e->d_due, (uint64_t)L.max_due) && dalloc(e, &e->d_due_key, (uint64_t)L.max_due) && dalloc(e, &e->d_ctr, 1) && dalloc(e, &e->d_time, 2) &&
         dalloc(e, &e->d_ring_total, 1) && dalloc(e, &e->d_n_build, 1);
    if (!ok) {
        g_create_error = e->err;
        chd_destroy(e);
        return CHD_ERR_CUDA;
    }
    for (ScanSite* site : {&e->site_hist, &e->site_hist_b, &e->site_pchist, &e->site_pchist_b, &e->site_win, &e->site_qoff, &e->site_slot,
                           &e->site_voff, &e->site_uoff, &e->site_border})
        site->error = &e->d_ctr->overflow;
    CCU(cudaHostAlloc((void**)&e->h_ctr, sizeof(Counters), cudaHostAllocDefault));
    CCU(cudaHostAlloc((void**)&e->h_u32, 64, cudaHostAllocDefault));
    CCU(cudaHostAlloc((void**)&e->h_get, 64, cudaHostAllocDefault));
    CCU(cudaMemsetAsync(e->d_ctr, 0, sizeof(Counters), e->stream));
    CCU(cudaMemsetAsync(e->d_time, 0, 16, e->stream));
    CCU(cudaMemsetAsync(e->d_ring_total, 0, 4, e->stream));
    CCU(cudaMemsetAsync(e->pairs[0].off, 0, (S + 1) * 4, e->stream));
    CCU(cudaMemsetAsync(e->pairs[1].off, 0, (S + 1) * 4, e->stream));
    CCU(cudaMemsetAsync(e->d_ring_off, 0, (C + 1) * 4, e->stream));
    CCU(cudaMemsetAsync(e->d_cell_start, 0, (C + 2) * 4, e->stream));
    CCU(cudaMemsetAsync(e->d_vis_off, 0, (S + 1) * 8, e->stream));
    CCU(cudaStreamSynchronize(e->stream));
#undef CCU
    *out = e;
    return CHD_OK;
}

chd_status chd_set_stream(chd_engine* e, void* cuda_stream) {
    if (!e) return CHD_ERR_INVALID;
    CU(e, cudaSetDevice(e->device));
    CU(e, cudaStreamSynchronize(e->stream));
    if (e->own_stream) {
        cudaStreamDestroy(e->stream);
        e->own_stream = false;
    }
    e->stream = (cudaStream_t)cuda_stream;
    return CHD_OK;
}

chd_status chd_sync(chd_engine* e) {
    if (!e) return CHD_ERR_INVALID;
    CU(e, cudaSetDevice(e->device));
    CU(e, cudaStreamSynchronize(e->stream));
    return CHD_OK;
}

/* ------------------------------------------------------------------ entities / build ---- */

chd_status chd_cell_of(chd_engine* e, const double* x, const double* z, uint32_t n, uint32_t* out) {
    if (!e || (n && (!x || !z || !out))) return CHD_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> lk(e->mu);
    CU(e, cudaSetDevice(e->device));
    // chunked through the tmp buffers (they are free outside chd_build)
    double *dx = nullptr, *dz = nullptr;
    uint32_t* dk = nullptr;
    const uint32_t chunk = 1u << 20;
    CU(e, cudaMallocAsync((void**)&dx, sizeof(double) * chunk, e->stream));
    CU(e, cudaMallocAsync((void**)&dz, sizeof(double) * chunk, e->stream));
    CU(e, cudaMallocAsync((void**)&dk, sizeof(uint32_t) * chunk, e->stream));
    HandoverOut ho{};
    for (uint32_t b = 0; b < n; b += chunk) {
        const uint32_t m = n - b < chunk ? n - b : chunk;
        CU(e, cudaMemcpyAsync(dx, x + b, sizeof(double) * m, cudaMemcpyDefault, e->stream));
        CU(e, cudaMemcpyAsync(dz, z + b, sizeof(double) * m, cudaMemcpyDefault, e->stream));
        assign_cells_kernel<<<blocks_for(m, 256), 256, 0, e->stream>>>(e->g, dx, dz, m, dk, nullptr, ho);
        KCHECK(e);
        cell_key_to_id_kernel<<<blocks_for(m, 256), 256, 0, e->stream>>>(dk, m, e->g.cells, e->g.id_start);
        KCHECK(e);
        CU(e, cudaMemcpyAsync(out + b, dk, sizeof(uint32_t) * m, cudaMemcpyDefault, e->stream));
    }
    CU(e, cudaFreeAsync(dx, e->stream));
    CU(e, cudaFreeAsync(dz, e->stream));
    CU(e, cudaFreeAsync(dk, e->stream));
    CU(e, cudaStreamSynchronize(e->stream));
    return CHD_OK;
}

chd_status chd_set_entities(chd_engine* e, const double* x, const double* z, uint32_t n) {
    if (!e || (n && (!x || !z))) return CHD_ERR_INVALID;
    if (n > e->lim.max_entities) {
        e->fail("chd_set_entities: %u > max_entities %u", n, e->lim.max_entities);
        return CHD_ERR_CAPACITY;
    }
    CU(e, cudaSetDevice(e->device));
    if (n != e->n_own) e->have_prev_key = false;
    if (n && is_device_ptr(e, x) && is_device_ptr(e, z)) {
        e->pos_x = x;  // device-resident producer: read in place
        e->pos_z = z;
    } else {
        CU(e, cudaMemcpyAsync(e->d_x, x, sizeof(double) * n, cudaMemcpyDefault, e->stream));
        CU(e, cudaMemcpyAsync(e->d_z, z, sizeof(double) * n, cudaMemcpyDefault, e->stream));
        e->pos_x = e->d_x;
        e->pos_z = e->d_z;
    }
    e->n_own = n;
    e->n_halo = 0;
    e->halo_on_device = false;
    e->assigned = false;
    e->entities_dirty = true;
    return CHD_OK;
}

static chd_status ensure_upload_stream(chd_engine* e);
static chd_status upload_queries(chd_engine* e, const chd_query_batch* q, QueryDev* out, bool need_sub, chd_engine::QStage* stage = nullptr,
                                 cudaStream_t on_stream = nullptr);

chd_status chd_prefetch_entities(chd_engine* e, const double* x, const double* z, uint32_t n) {
    if (!e || (n && (!x || !z))) return CHD_ERR_INVALID;
    if (n > e->lim.max_entities) {
        e->fail("chd_prefetch_entities: %u > max_entities %u", n, e->lim.max_entities);
        return CHD_ERR_CAPACITY;
    }
    CU(e, cudaSetDevice(e->device));
    const int back = e->pos_buf ^ 1;
    {
        chd_status st = ensure_upload_stream(e);
        if (st != CHD_OK) return st;
    }
    if (!e->d_xb[back]) {  // the second pair of position buffers exists only for hosts that prefetch
        if (!dalloc(e, &e->d_xb[back], e->lim.max_entities) || !dalloc(e, &e->d_zb[back], e->lim.max_entities)) return CHD_ERR_CUDA;
    }
    // the back pair was last read by the cell assignment of an earlier tick
    if (e->pos_read_recorded[back]) CU(e, cudaStreamWaitEvent(e->up_stream, e->ev_pos_read[back], 0));
    CU(e, cudaMemcpyAsync(e->d_xb[back], x, sizeof(double) * n, cudaMemcpyDefault, e->up_stream));
    CU(e, cudaMemcpyAsync(e->d_zb[back], z, sizeof(double) * n, cudaMemcpyDefault, e->up_stream));
    CU(e, cudaEventRecord(e->ev_upload, e->up_stream));
    e->staged = true;
    e->staged_n = n;
    return CHD_OK;
}

static chd_status ensure_upload_stream(chd_engine* e) {
    if (e->up_stream) return CHD_OK;
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    CU(e, cudaStreamCreateWithPriority(&e->up_stream, cudaStreamNonBlocking, hi));
    CU(e, cudaEventCreateWithFlags(&e->ev_upload, cudaEventDisableTiming));
    CU(e, cudaEventCreateWithFlags(&e->ev_upload_q, cudaEventDisableTiming));
    CU(e, cudaEventCreateWithFlags(&e->ev_upload_rings, cudaEventDisableTiming));
    for (int i = 0; i < 2; i++) {
        CU(e, cudaEventCreateWithFlags(&e->ev_q_read[i], cudaEventDisableTiming));
        CU(e, cudaEventCreateWithFlags(&e->ev_ring_read[i], cudaEventDisableTiming));
    }
    return CHD_OK;
}

chd_status chd_prefetch_queries(chd_engine* e, const chd_query_batch* q) {
    if (!e || !q) return CHD_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> lk(e->mu);
    CU(e, cudaSetDevice(e->device));
    chd_status st = ensure_upload_stream(e);
    if (st != CHD_OK) return st;
    const int set = e->q_next;
    chd_engine::QStage& qs = e->dq_pf[set];
    if (!e->dq_pf_alloc[set]) {
        const uint64_t Q = e->lim.max_queries;
        const bool ok = dalloc(e, &qs.sub, Q) && dalloc(e, &qs.kind, Q) && dalloc(e, &qs.sph_cx, Q) && dalloc(e, &qs.sph_cz, Q) && dalloc(e, &qs.sph_r, Q) &&
                        dalloc(e, &qs.box_cx, Q) && dalloc(e, &qs.box_cz, Q) && dalloc(e, &qs.box_ex, Q) && dalloc(e, &qs.box_ez, Q) &&
                        dalloc(e, &qs.cone_cx, Q) && dalloc(e, &qs.cone_cz, Q) && dalloc(e, &qs.cone_dx, Q) && dalloc(e, &qs.cone_dz, Q) &&
                        dalloc(e, &qs.cone_angle, Q) && dalloc(e, &qs.cone_r, Q) && dalloc(e, &qs.spot_off, Q + 1) && dalloc(e, &qs.spot_ndist, Q) &&
                        dalloc(e, &qs.spot_x, (uint64_t)e->lim.max_spots) && dalloc(e, &qs.spot_z, (uint64_t)e->lim.max_spots) &&
                        dalloc(e, &qs.spot_dist, (uint64_t)e->lim.max_spots);
        if (!ok) return CHD_ERR_CUDA;
        e->dq_pf_alloc[set] = true;
    }
    if (e->q_read_recorded[set]) CU(e, cudaStreamWaitEvent(e->up_stream, e->ev_q_read[set], 0));  // its previous batch has been consumed
    QueryDev d;
    st = upload_queries(e, q, &d, true, &qs, e->up_stream);
    if (st != CHD_OK) return st;
    CU(e, cudaEventRecord(e->ev_upload_q, e->up_stream));
    e->staged_qd = d;
    e->staged_q_set = set;
    e->staged_q = true;
    e->q_next = set ^ 1;
    return CHD_OK;
}

chd_status chd_prefetch_rings(chd_engine* e, const uint32_t* ring_off, uint32_t n_entries, const int64_t* arrival, const uint32_t* sender,
                              const uint64_t* index, const uint64_t* ch_msg_index) {
    if (!e || !ring_off) return CHD_ERR_INVALID;
    if (n_entries > e->lim.max_ring_entries) {
        e->fail("%u ring entries > max_ring_entries %u", n_entries, e->lim.max_ring_entries);
        return CHD_ERR_CAPACITY;
    }
    if (n_entries && (!arrival || !sender || !index)) return CHD_ERR_INVALID;
    CU(e, cudaSetDevice(e->device));
    chd_status st = ensure_upload_stream(e);
    if (st != CHD_OK) return st;
    const int set = e->ring_next;
    chd_engine::RStage& r = e->ring_pf[set];
    const uint64_t C = e->g.cells;
    if (!e->ring_pf_alloc[set]) {
        const uint64_t R = e->lim.max_ring_entries;
        if (!(dalloc(e, &r.off, C + 1) && dalloc(e, &r.arrival, R) && dalloc(e, &r.sender, R) && dalloc(e, &r.index, R) && dalloc(e, &r.cmi, C)))
            return CHD_ERR_CUDA;
        e->ring_pf_alloc[set] = true;
    }
    if (e->ring_read_recorded[set]) CU(e, cudaStreamWaitEvent(e->up_stream, e->ev_ring_read[set], 0));
    cudaStream_t us = e->up_stream;
    CU(e, cudaMemcpyAsync(r.off, ring_off, sizeof(uint32_t) * (C + 1), cudaMemcpyDefault, us));
    if (n_entries) {
        CU(e, cudaMemcpyAsync(r.arrival, arrival, sizeof(int64_t) * n_entries, cudaMemcpyDefault, us));
        CU(e, cudaMemcpyAsync(r.sender, sender, sizeof(uint32_t) * n_entries, cudaMemcpyDefault, us));
        CU(e, cudaMemcpyAsync(r.index, index, sizeof(uint64_t) * n_entries, cudaMemcpyDefault, us));
    }
    if (ch_msg_index) CU(e, cudaMemcpyAsync(r.cmi, ch_msg_index, sizeof(uint64_t) * C, cudaMemcpyDefault, us));
    CU(e, cudaEventRecord(e->ev_upload_rings, us));
    e->staged_rings = true;
    e->staged_ring_set = set;
    e->staged_ring_total = n_entries;
    e->staged_ring_cmi = ch_msg_index != nullptr;
    e->ring_next = set ^ 1;
    return CHD_OK;
}

chd_status chd_adopt_prefetched(chd_engine* e) {
    if (!e) return CHD_ERR_INVALID;
    if (!e->staged && !e->staged_q && !e->staged_rings) {
        e->fail("chd_adopt_prefetched without a preceding chd_prefetch_entities / chd_prefetch_queries / chd_prefetch_rings");
        return CHD_ERR_STATE;
    }
    CU(e, cudaSetDevice(e->device));
    if (e->staged_q) {  // consumed by the next chd_begin_interest / chd_update_interest called with q == NULL
        e->adopted_qd = e->staged_qd;
        e->adopted_q_set = e->staged_q_set;
        e->have_adopted_q = true;
        e->wait_q = true;
        e->staged_q = false;
    }
    if (e->staged_rings) {
        const chd_engine::RStage& r = e->ring_pf[e->staged_ring_set];
        e->ring_off_p = r.off; e->ring_arrival_p = r.arrival; e->ring_sender_p = r.sender; e->ring_index_p = r.index;
        e->ch_msg_index_p = e->staged_ring_cmi ? r.cmi : nullptr;
        e->have_ch_msg_index = e->staged_ring_cmi;
        e->ring_set = e->staged_ring_set;
        e->wait_rings = true;
        e->staged_rings = false;
        set_u32_kernel<<<1, 1, 0, e->stream>>>(e->d_ring_total, e->staged_ring_total);
        KCHECK(e);
    }
    if (!e->staged) return CHD_OK;
    CU(e, cudaStreamWaitEvent(e->stream, e->ev_upload, 0));
    e->pos_buf ^= 1;
    e->d_x = e->d_xb[e->pos_buf];
    e->d_z = e->d_zb[e->pos_buf];
    e->pos_x = e->d_x;
    e->pos_z = e->d_z;
    if (e->staged_n != e->n_own) e->have_prev_key = false;
    e->n_own = e->staged_n;
    e->n_halo = 0;
    e->halo_on_device = false;
    e->assigned = false;
    e->entities_dirty = true;
    e->staged = false;
    return CHD_OK;
}

chd_status chd_entity_buffers(chd_engine* e, double** d_x, double** d_z, uint32_t* n) {
    if (!e) return CHD_ERR_INVALID;
    if (d_x) *d_x = e->d_x;
    if (d_z) *d_z = e->d_z;
    if (n) *n = e->n_own;
    e->pos_x = e->d_x;
    e->pos_z = e->d_z;
    e->assigned = false;  // the caller may write positions
    e->entities_dirty = true;
    return CHD_OK;
}

chd_status chd_set_entity_count(chd_engine* e, uint32_t n) {
    if (!e || n > e->lim.max_entities) return CHD_ERR_INVALID;
    if (n != e->n_own) e->have_prev_key = false;
    e->n_own = n;
    e->n_halo = 0;
    e->halo_on_device = false;
    e->assigned = false;
    e->entities_dirty = true;
    return CHD_OK;
}

chd_status chd_set_entity_ids(chd_engine* e, const uint32_t* gid, uint32_t n) {
    if (!e || n > e->lim.max_entities) return CHD_ERR_INVALID;
    CU(e, cudaSetDevice(e->device));
    if (!gid) {
        e->have_gid = false;
        return CHD_OK;
    }
    CU(e, cudaMemcpyAsync(e->d_gid, gid, sizeof(uint32_t) * n, cudaMemcpyDefault, e->stream));
    e->have_gid = true;
    e->entities_dirty = true;
    return CHD_OK;
}

// The engine's position buffers are double-buffered for chd_prefetch_entities: remember (outside any graph capture)
// the last kernel that read the front pair, so an upload into it can be ordered after that read.
static chd_status note_pos_read(chd_engine* e) {
    if (e->pos_x != e->d_x || !e->ev_pos_read[e->pos_buf]) return CHD_OK;
    CU(e, cudaEventRecord(e->ev_pos_read[e->pos_buf], e->stream));
    e->pos_read_recorded[e->pos_buf] = true;
    return CHD_OK;
}

static chd_status assign_cells_impl(chd_engine* e);

chd_status chd_assign_cells(chd_engine* e) {
    if (!e) return CHD_ERR_INVALID;
    const bool was_assigned = e->assigned;
    chd_status st = assign_cells_impl(e);
    if (st == CHD_OK && !was_assigned) st = note_pos_read(e);
    return st;
}

static chd_status assign_cells_impl(chd_engine* e) {
    CU(e, cudaSetDevice(e->device));
    if (e->assigned) return CHD_OK;
    // handover detection compares against the keys of the previous assignment (same entity count):
    // the buffers are swapped, never copied.
    uint32_t* prev = nullptr;
    if (e->have_prev_key) {
        uint32_t* t = e->d_key;
        e->d_key = e->d_prev_key;
        e->d_prev_key = t;
        prev = e->d_prev_key;
    }
    HandoverOut ho{e->d_ho_entity, e->d_ho_src, e->d_ho_dst, &e->d_ctr->n_handover, e->ho_cap};
    CU(e, cudaMemsetAsync(&e->d_ctr->n_handover, 0, 4, e->stream));
    if (e->n_own) {
        assign_cells_kernel<<<blocks_for(e->n_own, 256), 256, 0, e->stream>>>(e->g, e->pos_x ? e->pos_x : e->d_x, e->pos_z ? e->pos_z : e->d_z,
                                                                              e->n_own, e->d_key, prev, ho);
        KCHECK(e);
        e->have_prev_key = true;
    }
    e->n_halo = 0;
    e->assigned = true;
    return CHD_OK;
}

extern "C++" {
template <int BINS>
static chd_status sort_pass(chd_engine* e, uint32_t* hist, const ScanSite& site, const uint32_t* key_in, const uint32_t* val_in, uint32_t n,
                            const uint32_t* n_ptr, uint32_t per_block, uint32_t nblocks, uint32_t shift, uint32_t bits, uint32_t* key_out,
                            uint32_t* val_out, ScatterExtras ex, unsigned long long* bump) {
    const uint32_t mask = (1u << bits) - 1u;
    radix_hist_kernel<BINS><<<nblocks, BUILD_THREADS, 0, e->stream>>>(key_in, n, n_ptr, per_block, shift, mask, hist, nblocks, bump);
    KCHECK(e);
    SCAN(e, exclusive_scan_1p<uint32_t, uint32_t>(hist, hist, (uint64_t)BINS * nblocks, site, e->stream));
    radix_scatter_kernel<BINS><<<nblocks, BUILD_THREADS, 0, e->stream>>>(key_in, val_in, n, n_ptr, per_block, shift, mask, hist, nblocks,
                                                                        key_out, val_out, ex);
    KCHECK(e);
    return CHD_OK;
}
}  // extern "C++"

static chd_status sort_pass_any(chd_engine* e, uint32_t* hist, const ScanSite& site, const uint32_t* key_in, const uint32_t* val_in,
                                uint32_t n, const uint32_t* n_ptr, uint32_t per_block, uint32_t nblocks, uint32_t shift, uint32_t bits,
                                uint32_t* key_out, uint32_t* val_out, ScatterExtras ex = ScatterExtras{0, nullptr, 0, nullptr},
                                unsigned long long* bump = nullptr) {
    if (bits <= 8) return sort_pass<256>(e, hist, site, key_in, val_in, n, n_ptr, per_block, nblocks, shift, bits, key_out, val_out, ex, bump);
    return sort_pass<1024>(e, hist, site, key_in, val_in, n, n_ptr, per_block, nblocks, shift, bits, key_out, val_out, ex, bump);
}

static chd_status build_enqueue(chd_engine* e, bool with_assign) {
    chd_status st;
    if (with_assign) {
        st = assign_cells_impl(e);
        if (st != CHD_OK) return st;
    }
    // multi-GPU: the halo count stays on the device (d_n_build = own + kept halo records); launches are sized for
    // the entity capacity and blocks beyond the live length idle.
    const uint32_t* n_ptr = e->halo_on_device ? e->d_n_build : nullptr;
    const uint32_t n = e->halo_on_device ? e->lim.max_entities : e->n_own + e->n_halo;
    const uint32_t C = e->g.cells;
    uint32_t bits = 1;
    while ((1u << bits) < C + 1) bits++;  // keys are in [0, C]
    const uint32_t passes = bits <= 10 ? 1 : 2;
    const uint32_t bits0 = passes == 1 ? bits : (bits + 1) / 2, bits1 = bits - bits0;
    // contiguous slice per block, multiple of the tile
    uint32_t per_block = (n + e->build_blocks - 1) / e->build_blocks;
    per_block = ((per_block + BUILD_TILE - 1) / BUILD_TILE) * BUILD_TILE;
    if (per_block == 0) per_block = BUILD_TILE;
    uint32_t nblocks = (n + per_block - 1) / per_block;
    if (nblocks == 0) nblocks = 1;
    const uint32_t* vals = e->have_gid ? e->d_gid : nullptr;
    // Phase copies: fused into the final scatter when the build is latency-bound (small N: one launch less), written by
    // a separate fully-coalesced pass when it is bandwidth-bound (large N: the fused variant does 4 scattered 4-byte
    // stores per entity; measured 325 us vs 175 + ~30 us at N = 10 M).
    const bool fuse_phases = n <= (2u << 20);
    const uint32_t fused_stride = fuse_phases ? e->phase_stride : 0u;
    if (passes == 1) {
        // single pass: digit == key, so the scatter also publishes the CSR offsets
        ScatterExtras ex{fused_stride, e->d_cell_start, C, &e->d_ctr->n_entities_in_world};
        st = sort_pass_any(e, e->d_hist, e->site_hist, e->d_key, vals, n, n_ptr, per_block, nblocks, 0, bits0, nullptr, e->d_sorted_ent, ex,
                           e->d_epoch + EP_BUILD);
        if (st != CHD_OK) return st;
    } else {
        st = sort_pass_any(e, e->d_hist, e->site_hist, e->d_key, vals, n, n_ptr, per_block, nblocks, 0, bits0, e->d_tmp_key, e->d_tmp_val,
                           ScatterExtras{0, nullptr, 0, nullptr}, e->d_epoch + EP_BUILD);
        if (st != CHD_OK) return st;
        ScatterExtras ex{fused_stride, nullptr, C, nullptr};
        st = sort_pass_any(e, e->d_hist, e->site_hist_b, e->d_tmp_key, e->d_tmp_val, n, n_ptr, per_block, nblocks, bits0, bits1, e->d_sorted_key,
                           e->d_sorted_ent, ex);
        if (st != CHD_OK) return st;
        cell_bounds_kernel<<<blocks_for((uint64_t)n + 1, 256), 256, 0, e->stream>>>(e->d_sorted_key, n, n_ptr, C, e->d_cell_start,
                                                                                      &e->d_ctr->n_entities_in_world);
        KCHECK(e);
    }
    if (!fuse_phases && n) {
        replicate_phases_kernel<<<blocks_for(n, 256), 256, 0, e->stream>>>(e->d_sorted_ent, n, n_ptr, e->phase_stride, e->d_sorted4);
        KCHECK(e);
    }
    return CHD_OK;
}

chd_status chd_build(chd_engine* e) {
    if (!e) return CHD_ERR_INVALID;
    CU(e, cudaSetDevice(e->device));
    StageTimer timer(e, CHD_STAGE_BUILD);
    chd_status st;
    if (!e->assigned && !e->halo_on_device) {
        // single-GPU flow: assign + sort as one replayable graph.  The key buffers swap every assignment
        // (handover detection compares against the previous keys), so there are two graph variants.
        uint32_t* target = e->have_prev_key ? e->d_prev_key : e->d_key;  // buffer the new keys will be written to
        const int slot = (target == e->d_key_a ? 0 : 1) + 2 * e->pos_buf;
        uint64_t key = mix_key(mix_key(mix_key(0x6275696c64ull, e->n_own), e->have_gid), e->have_prev_key);
        key = mix_key(mix_key(key, (uint64_t)(uintptr_t)target), (uint64_t)(uintptr_t)e->pos_x ^ ((uint64_t)(uintptr_t)e->pos_z << 1));
        st = run_stage(e, e->g_build[slot], key, [&]() { return build_enqueue(e, true); });
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
        if (st == CHD_OK) st = note_pos_read(e);
    } else if (e->assigned && e->halo_on_device) {
        // multi-GPU flow: cells were assigned by chd_export_border and the halo appended on the device; the sort over
        // own + halo entities is sized by capacity (device-side length) and therefore replayable as well.
        const int slot = e->d_key == e->d_key_a ? 0 : 1;
        uint64_t key = mix_key(mix_key(mix_key(0x736f7274ull, e->lim.max_entities), e->have_gid), (uint64_t)(uintptr_t)e->d_key);
        st = run_stage(e, e->g_build[slot], key, [&]() { return build_enqueue(e, false); });
    } else {
        const bool with_assign = !e->assigned;
        st = build_enqueue(e, with_assign);
        if (st == CHD_OK && with_assign) st = note_pos_read(e);
    }
    if (st != CHD_OK) return st;
    e->n_sorted = e->n_own + e->n_halo;
    e->built = true;
    e->entities_dirty = false;
    return CHD_OK;
}

/* ------------------------------------------------------------------ subscribers / queries ---- */

chd_status chd_set_subscribers(chd_engine* e, const uint32_t* conn_id, uint32_t n) {
    if (!e || (n && !conn_id)) return CHD_ERR_INVALID;
    if (n > e->lim.max_subscribers) {
        e->fail("chd_set_subscribers: %u > max_subscribers %u", n, e->lim.max_subscribers);
        return CHD_ERR_CAPACITY;
    }
    CU(e, cudaSetDevice(e->device));
    CU(e, cudaMemcpyAsync(e->d_conn, conn_id, sizeof(uint32_t) * n, cudaMemcpyDefault, e->stream));
    CU(e, cudaMemsetAsync(e->pairs[0].off, 0, ((uint64_t)e->lim.max_subscribers + 1) * 4, e->stream));
    CU(e, cudaMemsetAsync(e->pairs[1].off, 0, ((uint64_t)e->lim.max_subscribers + 1) * 4, e->stream));
    CU(e, cudaMemsetAsync(e->d_vis_off, 0, ((uint64_t)e->lim.max_subscribers + 1) * 8, e->stream));
    e->n_slots = n;
    e->cur = 0;
    return CHD_OK;
}

// copies the batch into the engine's device SoA and returns the device view
static chd_status upload_queries(chd_engine* e, const chd_query_batch* q, QueryDev* out, bool need_sub, chd_engine::QStage* stage,
                                 cudaStream_t on_stream) {
    if (!q) return CHD_ERR_INVALID;
    chd_engine::QStage& dq = stage ? *stage : e->dq;
    const uint32_t n = q->n;
    if (n > e->lim.max_queries) {
        e->fail("query batch of %u > max_queries %u", n, e->lim.max_queries);
        return CHD_ERR_CAPACITY;
    }
    QueryDev d{};
    d.n = n;
    cudaStream_t st = stage ? on_stream : e->stream;
#define UP(field, T)                                                                                      \
    if (q->field) {                                                                                       \
        if (is_device_ptr(e, q->field)) {                                                                 \
            d.field = q->field; /* device-resident batch: consumed in place */                            \
        } else {                                                                                          \
            CU(e, cudaMemcpyAsync(dq.field, q->field, sizeof(T) * n, cudaMemcpyDefault, st));          \
            d.field = dq.field;                                                                        \
        }                                                                                                 \
    }
    if (need_sub) {
        if (!q->sub && n > e->n_slots) {
            e->fail("identity query batch (sub == NULL) of %u queries > %u subscribers", n, e->n_slots);
            return CHD_ERR_INVALID;
        }
        UP(sub, uint32_t);
    }
    UP(kind, uint8_t);
    UP(sph_cx, double); UP(sph_cz, double); UP(sph_r, double);
    UP(box_cx, double); UP(box_cz, double); UP(box_ex, double); UP(box_ez, double);
    UP(cone_cx, double); UP(cone_cz, double); UP(cone_dx, double); UP(cone_dz, double); UP(cone_angle, double); UP(cone_r, double);
    UP(spot_ndist, uint32_t);
#undef UP
    if (!q->kind && n && (!q->sph_cx || !q->sph_cz || !q->sph_r)) {
        e->fail("kind == NULL means all-sphere: sph_cx/sph_cz/sph_r are required");
        return CHD_ERR_INVALID;
    }
    if (q->spot_off) {
        // spot_off may live on the host or on the device; its last element sizes the spot arrays
        uint32_t total = 0;
        CU(e, cudaMemcpyAsync(dq.spot_off, q->spot_off, sizeof(uint32_t) * ((uint64_t)n + 1), cudaMemcpyDefault, st));
        CU(e, cudaMemcpyAsync(e->h_u32, dq.spot_off + n, 4, cudaMemcpyDeviceToHost, st));
        CU(e, cudaStreamSynchronize(st));
        total = *e->h_u32;
        if (total > e->lim.max_spots) {
            e->fail("%u spots > max_spots %u", total, e->lim.max_spots);
            return CHD_ERR_CAPACITY;
        }
        if (total && (!q->spot_x || !q->spot_z)) return CHD_ERR_INVALID;
        CU(e, cudaMemcpyAsync(dq.spot_x, q->spot_x, sizeof(double) * total, cudaMemcpyDefault, st));
        CU(e, cudaMemcpyAsync(dq.spot_z, q->spot_z, sizeof(double) * total, cudaMemcpyDefault, st));
        if (q->spot_dist) CU(e, cudaMemcpyAsync(dq.spot_dist, q->spot_dist, sizeof(uint32_t) * total, cudaMemcpyDefault, st));
        else CU(e, cudaMemsetAsync(dq.spot_dist, 0, sizeof(uint32_t) * total, st));
        d.spot_off = dq.spot_off; d.spot_x = dq.spot_x; d.spot_z = dq.spot_z; d.spot_dist = dq.spot_dist;
        if (!q->spot_ndist) {
            CU(e, cudaMemsetAsync(dq.spot_ndist, 0, sizeof(uint32_t) * n, st));
            d.spot_ndist = dq.spot_ndist;
        }
    }
    *out = d;
    return CHD_OK;
}

// Q1 + scan + Q2: fills bbox / window / side lists / status / qcount for the batch
static chd_status run_query_kernels(chd_engine* e, const QueryDev& d) {
    const uint32_t n = d.n;
    if (n == 0) return CHD_OK;
    query_bbox_kernel<<<blocks_for(n, 256), 256, 0, e->stream>>>(e->g, d, e->d_bbox, e->d_win_size);
    KCHECK(e);
    SCAN(e, exclusive_scan_1p<uint32_t, uint64_t>(e->d_win_size, e->d_win_off, n, e->site_win, e->stream));
    query_sample_kernel<<<blocks_for(n, 128), 128, 0, e->stream>>>(e->g, d, e->d_bbox, e->d_win_off, e->lim.max_window_cells, e->d_window,
                                                                   e->d_side_cell, e->d_side_dist, e->d_side_cnt, e->d_status, e->d_qcount,
                                                                   &e->d_ctr->required_window_cells, &e->d_ctr->overflow);
    KCHECK(e);
    return CHD_OK;
}

chd_status chd_query_channel_ids(chd_engine* e, const chd_query_batch* q, uint32_t* out_status, uint32_t* out_off,
                                 uint32_t* out_channel_id, uint32_t* out_dist, uint64_t cap) {
    if (!e || !q) return CHD_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> lk(e->mu);
    CU(e, cudaSetDevice(e->device));
    if (e->aux_stream) CU(e, cudaStreamSynchronize(e->aux_stream));  // an in-flight interest update shares the query scratch
    QueryDev d;
    chd_status st = upload_queries(e, q, &d, false);
    if (st != CHD_OK) return st;
    const uint32_t n = d.n;
    if (n == 0) {
        if (out_off) {
            const uint32_t zero = 0;
            CU(e, cudaMemcpyAsync(out_off, &zero, 4, cudaMemcpyDefault, e->stream));
            CU(e, cudaStreamSynchronize(e->stream));
        }
        return CHD_OK;
    }
    bump_epoch_kernel<<<1, 1, 0, e->stream>>>(e->d_epoch + EP_QUERY);
    KCHECK(e);
    st = run_query_kernels(e, d);
    if (st != CHD_OK) return st;
    SCAN(e, exclusive_scan_1p<uint32_t, uint64_t>(e->d_qcount, e->d_qoff, n, e->site_qoff, e->stream));
    const uint64_t dev_cap = e->lim.max_pairs;
    query_write_kernel<<<blocks_for(n, 128), 128, 0, e->stream>>>(e->g, n, e->d_status, e->d_bbox, e->d_win_off, e->d_window, e->d_side_cell,
                                                                  e->d_side_dist, e->d_side_cnt, d.spot_off, e->d_qoff, dev_cap,
                                                                  e->d_qout_id, e->d_qout_dist);
    KCHECK(e);
    // totals
    uint64_t* h64 = (uint64_t*)e->h_u32;
    CU(e, cudaMemcpyAsync(h64, e->d_qoff + n, 8, cudaMemcpyDeviceToHost, e->stream));
    CU(e, cudaMemcpyAsync(h64 + 1, e->d_win_off + n, 8, cudaMemcpyDeviceToHost, e->stream));
    CU(e, cudaStreamSynchronize(e->stream));
    const uint64_t total = h64[0], wtotal = h64[1];
    if (wtotal > e->lim.max_window_cells) {
        e->fail("query windows need %llu cells > max_window_cells %llu", (unsigned long long)wtotal,
                (unsigned long long)e->lim.max_window_cells);
        return CHD_ERR_CAPACITY;
    }
    if (total > dev_cap || total > cap) {
        e->fail("query result has %llu entries > capacity %llu", (unsigned long long)total,
                (unsigned long long)(total > dev_cap ? dev_cap : cap));
        return CHD_ERR_CAPACITY;
    }
    if (out_status) CU(e, cudaMemcpyAsync(out_status, e->d_status, sizeof(uint32_t) * n, cudaMemcpyDefault, e->stream));
    if (out_off) {
        // u64 device offsets -> u32 caller offsets
        narrow_offsets_kernel<<<blocks_for((uint64_t)n + 1, 256), 256, 0, e->stream>>>(e->d_qoff, n + 1, e->d_new_off);
        KCHECK(e);
        CU(e, cudaMemcpyAsync(out_off, e->d_new_off, sizeof(uint32_t) * ((uint64_t)n + 1), cudaMemcpyDefault, e->stream));
    }
    if (out_channel_id) CU(e, cudaMemcpyAsync(out_channel_id, e->d_qout_id, sizeof(uint32_t) * total, cudaMemcpyDefault, e->stream));
    if (out_dist) CU(e, cudaMemcpyAsync(out_dist, e->d_qout_dist, sizeof(uint32_t) * total, cudaMemcpyDefault, e->stream));
    CU(e, cudaStreamSynchronize(e->stream));
    return CHD_OK;
}

// part 0: query -> new subscription pairs (everything emit needs); part 1: pairs grouped by cell + diff lists
// (needed by the fan-out pass and the host only).  An event between the two lets emit start early.
static chd_status interest_enqueue(chd_engine* e, const QueryDev& d, int part) {
    const uint32_t n = d.n, S = e->n_slots;
    cudaStream_t s = e->stream;
    chd_status st = CHD_OK;
    PairBuf& prev = e->pairs[e->cur];
    PairBuf& cur = e->pairs[e->cur ^ 1];
    const uint64_t P = e->lim.max_pairs;
    if (part == 0) {
    st = run_query_kernels(e, d);
    if (st != CHD_OK) return st;
    // sub == NULL is the identity batch (query i <-> subscriber slot i): no slot table needed
    const int32_t* slot_query = d.sub ? e->d_slot_query : nullptr;
    if (d.sub) CU(e, cudaMemsetAsync(e->d_slot_query, 0xFF, sizeof(int32_t) * (uint64_t)(S ? S : 1), s));
    CU(e, cudaMemsetAsync(&e->d_ctr->n_query_errors, 0, 4 * 4, s));  // n_query_errors, n_sub_new, n_unsub, n_kept

    if (n && d.sub) {
        slot_scatter_kernel<<<blocks_for(n, 256), 256, 0, s>>>(d.sub, n, S, e->d_slot_query);
        KCHECK(e);
    }
    SCAN(e, exclusive_scan_fn<SlotCountIn, uint32_t>(SlotCountIn{slot_query, n, e->d_status, e->d_qcount, prev.off}, cur.off, S, e->site_slot, s));
    if (S) {
        interest_fill_kernel<<<blocks_for(S, 128), 128, 0, s>>>(e->g, S, slot_query, n, e->d_status, e->d_bbox, e->d_win_off, e->d_window,
                                                                e->d_side_cell, e->d_side_dist, e->d_side_cnt, d.spot_off, prev, cur, P,
                                                                e->d_time, DiffOut{e->d_new_sub, e->d_new_ch, e->d_gone_sub, e->d_gone_ch}, e->d_pair_ch, e->d_ctr);
        e->pair_ch_valid = true;
        e->by_cell_valid = true;
        KCHECK(e);
    }
    return CHD_OK;
    }  // part 0
    // pairs grouped by cell for the fan-out pass (= every channel's subscriber list): a stable radix sort of pair
    // indices by cell with the same kernels as the entity build (no global atomics), device-side length
    {
        const uint32_t C = e->g.cells;
        uint32_t bits = 1;
        while ((1u << bits) < C) bits++;
        const uint32_t passes = bits <= 10 ? 1 : 2;
        const uint32_t bits0 = passes == 1 ? bits : (bits + 1) / 2, bits1 = bits - bits0;
        const uint32_t nb = e->pc_blocks;
        uint32_t per_block = (uint32_t)((P + nb - 1) / nb);
        per_block = ((per_block + BUILD_TILE - 1) / BUILD_TILE) * BUILD_TILE;
        cudaStream_t keep = e->stream;  // sort_pass launches on e->stream, which already is `s`
        (void)keep;
        if (passes == 1) {
            st = sort_pass_any(e, e->d_pc_hist, e->site_pchist, cur.cell, nullptr, (uint32_t)P, cur.off + S, per_block, nb, 0, bits0, nullptr,
                               e->d_by_cell);
            if (st != CHD_OK) return st;
        } else {
            st = sort_pass_any(e, e->d_pc_hist, e->site_pchist, cur.cell, nullptr, (uint32_t)P, cur.off + S, per_block, nb, 0, bits0,
                               e->d_pc_tmp_key, e->d_pc_tmp_val);
            if (st != CHD_OK) return st;
            st = sort_pass_any(e, e->d_pc_hist, e->site_pchist_b, e->d_pc_tmp_key, e->d_pc_tmp_val, (uint32_t)P, cur.off + S, per_block, nb, bits0,
                               bits1, nullptr, e->d_by_cell);
            if (st != CHD_OK) return st;
        }
    }
    return CHD_OK;
}

chd_status chd_update_interest(chd_engine* e, const chd_query_batch* q, int64_t now_ns) {
    if (!e) return CHD_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> lk(e->mu);
    CU(e, cudaSetDevice(e->device));
    StageTimer timer(e, CHD_STAGE_INTEREST);
    QueryDev d;
    chd_status st;
    int pf_set = -1;  // prefetch staging set this batch lives in
    if (!q) {  // the batch uploaded by chd_prefetch_queries and handed over by chd_adopt_prefetched
        if (!e->have_adopted_q) {
            e->fail("interest update without a batch: q == NULL needs chd_prefetch_queries + chd_adopt_prefetched first");
            return CHD_ERR_STATE;
        }
        d = e->adopted_qd;
        pf_set = e->adopted_q_set;
        e->have_adopted_q = false;
        if (d.n > e->n_slots && !d.sub) {
            e->fail("identity query batch (sub == NULL) of %u queries > %u subscribers", d.n, e->n_slots);
            return CHD_ERR_INVALID;
        }
        if (e->wait_q) {
            CU(e, cudaStreamWaitEvent(e->stream, e->ev_upload_q, 0));
            e->wait_q = false;
        }
    } else {
        st = upload_queries(e, q, &d, true);  // H2D / D2D copies into the engine's SoA: outside the graph
        if (st != CHD_OK) return st;
    }
    const int gslot = e->cur + (pf_set == 1 ? 2 : 0);
    stage_begin_kernel<<<1, 1, 0, e->stream>>>(e->d_time, now_ns, e->d_epoch + EP_QUERY, nullptr);
    KCHECK(e);
    // the graph bakes in which staging arrays are live, the batch size and the pair-buffer parity
    uint64_t key = mix_key(mix_key(mix_key(0x696e74ull, d.n), e->n_slots), (uint64_t)e->cur);
    const void* baked[] = {d.sub, d.kind, d.sph_cx, d.sph_cz, d.sph_r, d.box_cx, d.box_cz, d.box_ex, d.box_ez, d.cone_cx, d.cone_cz,
                           d.cone_dx, d.cone_dz, d.cone_angle, d.cone_r, d.spot_off, d.spot_ndist, d.spot_x, d.spot_z, d.spot_dist};
    for (const void* p : baked) key = mix_key(key, (uint64_t)(uintptr_t)p);  // pointers are baked into the captured launches
    st = run_stage(e, e->g_interest[gslot], key, [&]() { return interest_enqueue(e, d, 0); });
    if (st != CHD_OK) return st;
    CU(e, cudaEventRecord(e->ev_pairs, e->stream));  // the new pairs exist: emit may start (chd_tick waits on this)
    st = run_stage(e, e->g_interest_b[gslot], mix_key(key, 0xb), [&]() { return interest_enqueue(e, d, 1); });
    if (st != CHD_OK) return st;
    if (pf_set >= 0) {  // the staging set may be refilled once these kernels have run
        CU(e, cudaEventRecord(e->ev_q_read[pf_set], e->stream));
        e->q_read_recorded[pf_set] = true;
    }
    e->cur ^= 1;
    e->last_nq = d.n;
    return CHD_OK;
}

chd_status chd_emit_visible(chd_engine* e) {
    if (!e) return CHD_ERR_INVALID;
    CU(e, cudaSetDevice(e->device));
    if (!e->built) {
        e->fail("chd_emit_visible before chd_build");
        return CHD_ERR_STATE;
    }
    cudaStream_t s = e->stream;
    PairBuf& pb = e->pairs[e->cur];
    const uint32_t S = e->n_slots;
    const uint64_t P = e->lim.max_pairs;
    StageTimer timer(e, CHD_STAGE_EMIT);
    const unsigned grid = (unsigned)e->sm_count * 8;
    const int variant = e->emit_variant;
    const uint64_t key = mix_key(mix_key(mix_key(0x656d6974ull, S), (uint64_t)e->cur), (uint64_t)variant);
    chd_status st = run_stage(e, e->g_emit_prep[e->cur], key, [&]() -> chd_status {
        if (variant == 5) {
            SCAN(e, exclusive_scan_fn<PairVcountIn, uint64_t>(PairVcountIn{pb.cell, e->d_cell_start}, e->d_voff, P, e->site_voff, s, pb.off + S));
            SCAN(e, exclusive_scan_fn<UnitCountIn, uint32_t>(UnitCountIn{pb.cell, e->d_by_cell, e->d_cell_start}, e->d_uoff, P, e->site_uoff, s,
                                                             pb.off + S));
            emit_units_kernel<<<grid, 256, 0, s>>>(pb.off + S, P, e->d_voff, e->d_uoff, e->d_by_cell, pb.cell, e->d_cell_start, e->d_units,
                                                   e->unit_cap, S, pb.off, e->d_vis_off, e->lim.max_visible, e->d_ctr, e->d_epoch + EP_EMIT);
            KCHECK(e);
        } else {
            // per-pair visible counts are computed by the scan itself; the partition pass opens the next epoch
            SCAN(e, exclusive_scan_fn<PairVcountIn, uint64_t>(PairVcountIn{pb.cell, e->d_cell_start}, e->d_voff, P, e->site_voff, s, pb.off + S));
            emit_partition_kernel<<<grid, 256, 0, s>>>(pb.off + S, P, e->d_voff, e->d_first_pair, e->max_tiles, S, pb.off, e->d_vis_off,
                                                       e->lim.max_visible, e->d_ctr, e->d_epoch + EP_EMIT);
            KCHECK(e);
        }
        return CHD_OK;
    });
    if (st != CHD_OK) return st;
    if (e->wait_before_emit_kernel) {
        CU(e, cudaStreamWaitEvent(s, e->wait_before_emit_kernel, 0));
        e->wait_before_emit_kernel = nullptr;
    }
    CU(e, cudaEventRecord(e->ev_prep_done, s));  // visible offsets + counters are final; only the expanded list is still to come
    {
        StageTimer kt(e, CHD_STAGE_EMIT_KERNEL);
        if (variant == 5)
            emit_visible_v5_kernel<<<(unsigned)e->sm_count * 4, EMIT_THREADS, 0, s>>>(pb.off + S, P, e->d_voff, e->d_uoff, e->d_units, e->unit_cap,
                                                                                      e->d_sorted4, e->phase_stride, e->d_vis, e->lim.max_visible,
                                                                                      (uint32_t)e->sm_count);
        else
            emit_visible_kernel<<<(unsigned)std::max(1, e->sm_count * e->emit_blocks_per_sm * (e->early_results_tick && e->early_mode == 0 ? std::max(e->emit_waves, 8) : e->emit_waves) - e->emit_grid_reduce -
                                                   (e->early_results_tick && e->early_mode == 2 ? e->early_reduce : 0)), EMIT_THREADS, 0, s>>>(pb.off + S, P, e->d_voff, pb.cell, e->d_cell_start, e->d_sorted4,
                                                                                   e->phase_stride, e->d_first_pair, e->d_vis, e->lim.max_visible);
        KCHECK(e);
    }
    return CHD_OK;
}

chd_status chd_set_rings(chd_engine* e, const uint32_t* ring_off, uint32_t n_entries, const int64_t* arrival, const uint32_t* sender,
                         const uint64_t* index, const uint64_t* ch_msg_index) {
    if (!e || !ring_off) return CHD_ERR_INVALID;
    CU(e, cudaSetDevice(e->device));
    const uint32_t C = e->g.cells;
    const uint32_t total = n_entries;
    if (total > e->lim.max_ring_entries) {
        e->fail("%u ring entries > max_ring_entries %u", total, e->lim.max_ring_entries);
        return CHD_ERR_CAPACITY;
    }
    if (total && (!arrival || !sender || !index)) return CHD_ERR_INVALID;
    const bool in_place = is_device_ptr(e, ring_off) && (!total || (is_device_ptr(e, arrival) && is_device_ptr(e, sender) && is_device_ptr(e, index))) &&
                          (!ch_msg_index || is_device_ptr(e, ch_msg_index));
    if (in_place) {  // device-resident rings: consumed in place
        e->ring_off_p = ring_off; e->ring_arrival_p = arrival; e->ring_sender_p = sender; e->ring_index_p = index;
        e->ch_msg_index_p = ch_msg_index;
    } else {
        CU(e, cudaMemcpyAsync(e->d_ring_off, ring_off, sizeof(uint32_t) * ((uint64_t)C + 1), cudaMemcpyDefault, e->stream));
        if (total) {
            CU(e, cudaMemcpyAsync(e->d_ring_arrival, arrival, sizeof(int64_t) * total, cudaMemcpyDefault, e->stream));
            CU(e, cudaMemcpyAsync(e->d_ring_sender, sender, sizeof(uint32_t) * total, cudaMemcpyDefault, e->stream));
            CU(e, cudaMemcpyAsync(e->d_ring_index, index, sizeof(uint64_t) * total, cudaMemcpyDefault, e->stream));
        }
        if (ch_msg_index) CU(e, cudaMemcpyAsync(e->d_ch_msg_index, ch_msg_index, sizeof(uint64_t) * C, cudaMemcpyDefault, e->stream));
        e->ring_off_p = e->d_ring_off; e->ring_arrival_p = e->d_ring_arrival; e->ring_sender_p = e->d_ring_sender;
        e->ring_index_p = e->d_ring_index; e->ch_msg_index_p = ch_msg_index ? e->d_ch_msg_index : nullptr;
    }
    e->have_ch_msg_index = ch_msg_index != nullptr;
    e->ring_set = -1;
    e->wait_rings = false;
    // the fan-out kernel clamps ring_off to this: a lying caller cannot cause out-of-bounds reads
    set_u32_kernel<<<1, 1, 0, e->stream>>>(e->d_ring_total, total);
    KCHECK(e);
    return CHD_OK;
}

chd_status chd_fanout_tick(chd_engine* e, int64_t t_ns) {
    if (!e) return CHD_ERR_INVALID;
    CU(e, cudaSetDevice(e->device));
    cudaStream_t s = e->stream;
    PairBuf& pb = e->pairs[e->cur];
    const uint32_t S = e->n_slots;
    const uint64_t P = e->lim.max_pairs;
    StageTimer timer(e, CHD_STAGE_FANOUT);
    if (e->wait_rings) {  // rings handed over by chd_adopt_prefetched: ordered after their upload
        CU(e, cudaStreamWaitEvent(s, e->ev_upload_rings, 0));
        e->wait_rings = false;
    }
    stage_begin_kernel<<<1, 1, 0, s>>>(e->d_time + 1, t_ns, e->d_epoch + EP_FANOUT, &e->d_ctr->n_due);
    KCHECK(e);
    RingDev ring{e->ring_off_p ? e->ring_off_p : e->d_ring_off, e->ring_arrival_p ? e->ring_arrival_p : e->d_ring_arrival,
                 e->ring_sender_p ? e->ring_sender_p : e->d_ring_sender, e->ring_index_p ? e->ring_index_p : e->d_ring_index,
                 e->have_ch_msg_index ? e->ch_msg_index_p : nullptr, e->d_ring_total};
    const unsigned grid = (unsigned)e->sm_count * 16;
    uint64_t key = mix_key(mix_key(mix_key(0x66616eull, S), (uint64_t)e->cur), e->have_ch_msg_index);
    for (const void* p : {(const void*)ring.off, (const void*)ring.arrival, (const void*)ring.sender, (const void*)ring.index,
                          (const void*)ring.channel_msg_index})
        key = mix_key(key, (uint64_t)(uintptr_t)p);  // pointers are baked into the captured launch
    chd_status st = run_stage(e, e->g_fanout[e->cur + (e->ring_set == 1 ? 2 : 0)], key, [&]() -> chd_status {
        const unsigned blocks = (unsigned)std::min<uint64_t>((P + 127) / 128, (uint64_t)e->sm_count * 16);
        fanout_kernel<<<blocks ? blocks : 1, 128, 0, s>>>(pb.off + S, P, pb, e->d_conn, ring, e->d_time + 1, e->g.id_start, e->d_by_cell, e->d_due,
                                                          e->d_due_key, e->lim.max_due, e->d_ctr);
        KCHECK(e);
        return CHD_OK;
    });
    if (st == CHD_OK && e->ring_set >= 0) {
        CU(e, cudaEventRecord(e->ev_ring_read[e->ring_set], s));
        e->ring_read_recorded[e->ring_set] = true;
    }
    return st;
}

static chd_status decode_summary(chd_engine* e, chd_tick_summary* out);

chd_status chd_summary(chd_engine* e, chd_tick_summary* out) {
    if (!e || !out) return CHD_ERR_INVALID;
    CU(e, cudaSetDevice(e->device));
    CU(e, cudaMemcpyAsync(e->h_ctr, e->d_ctr, sizeof(Counters), cudaMemcpyDeviceToHost, e->stream));
    CU(e, cudaStreamSynchronize(e->stream));
    return decode_summary(e, out);
}

static chd_status decode_summary(chd_engine* e, chd_tick_summary* out) {
    const Counters& c = *e->h_ctr;
    out->n_pairs = c.n_pairs; out->n_visible = c.n_visible; out->n_entities_in_world = c.n_entities_in_world;
    out->n_query_errors = c.n_query_errors; out->n_sub_new = c.n_sub_new; out->n_unsub = c.n_unsub; out->n_kept = c.n_kept;
    out->n_due = c.n_due; out->n_handover = c.n_handover; out->overflow = c.overflow; out->required_pairs = c.required_pairs;
    out->required_window_cells = c.required_window_cells; out->required_visible = c.required_visible; out->required_due = c.n_due;
    out->reserved = 0;
    if (c.overflow) {
        e->fail("capacity overflow mask 0x%x (pairs %llu, window cells %llu, visible %llu, due %u required)", c.overflow,
                (unsigned long long)c.required_pairs, (unsigned long long)c.required_window_cells,
                (unsigned long long)c.required_visible, c.n_due);
        // sticky bits are cleared so the caller can retry after raising limits
        CU(e, cudaMemsetAsync(&e->d_ctr->overflow, 0, 4, e->stream));
        return CHD_ERR_CAPACITY;
    }
    return CHD_OK;
}

chd_status chd_begin_interest(chd_engine* e, const chd_query_batch* q, int64_t t_ns, int with_fanout) {
    if (!e) return CHD_ERR_INVALID;
    if (!q && !e->have_adopted_q) {
        e->fail("chd_begin_interest: q == NULL needs chd_prefetch_queries + chd_adopt_prefetched first");
        return CHD_ERR_STATE;
    }
    if (e->interest_pending) {
        e->fail("chd_begin_interest: the previous one has not been joined by chd_tick yet");
        return CHD_ERR_STATE;
    }
    CU(e, cudaSetDevice(e->device));
    if (!e->aux_stream) {  // no second stream: run in place
        chd_status st = chd_update_interest(e, q, t_ns);
        if (st == CHD_OK && with_fanout) st = chd_fanout_tick(e, t_ns);
        return st;
    }
    std::lock_guard<std::recursive_mutex> lk(e->mu);  // `stream` is redirected below
    cudaStream_t main_stream = e->stream;
    CU(e, cudaEventRecord(e->ev_fork, main_stream));
    CU(e, cudaStreamWaitEvent(e->aux_stream, e->ev_fork, 0));
    e->stream = e->aux_stream;
    chd_status st = chd_update_interest(e, q, t_ns);
    if (st == CHD_OK && cudaEventRecord(e->ev_interest, e->aux_stream) != cudaSuccess) st = CHD_ERR_CUDA;
    if (st == CHD_OK && with_fanout) st = chd_fanout_tick(e, t_ns);
    if (st == CHD_OK && cudaEventRecord(e->ev_join, e->aux_stream) != cudaSuccess) st = CHD_ERR_CUDA;
    e->stream = main_stream;
    if (st != CHD_OK) return st;
    e->interest_pending = true;
    e->pending_fanout = with_fanout != 0;
    return CHD_OK;
}

static chd_status chd_tick_impl(chd_engine* e, const chd_query_batch* q, int64_t t_ns, uint32_t flags, chd_tick_summary* out);

chd_status chd_tick(chd_engine* e, const chd_query_batch* q, int64_t t_ns, uint32_t flags, chd_tick_summary* out) {
    if (!e) return CHD_ERR_INVALID;
    chd_status st;
    {
        StageTimer whole(e, CHD_STAGE_TICK);  // main-stream span of the tick (without the summary read-back)
        st = chd_tick_impl(e, q, t_ns, flags, nullptr);
    }
    if (st != CHD_OK) return st;
    if (out) return chd_summary(e, out);
    return CHD_OK;
}

static chd_status chd_tick_impl(chd_engine* e, const chd_query_batch* q, int64_t t_ns, uint32_t flags, chd_tick_summary* out) {
    chd_status st;
    const bool need_build = (flags & CHD_TICK_BUILD) && (e->entities_dirty || !e->built);
    const bool do_emit = flags & CHD_TICK_EMIT;
    bool do_fanout = flags & CHD_TICK_FANOUT;
    e->early_ready = false;
    e->build_done_recorded = false;
    e->early_results_tick = (flags & CHD_TICK_EARLY_RESULTS) != 0;
    if (e->interest_pending) {
        // interest (+ fan-out) of this tick were started early with chd_begin_interest and are running on aux_stream
        if (q) {
            e->fail("chd_tick: a query batch was given while chd_begin_interest is pending");
            return CHD_ERR_STATE;
        }
        cudaStream_t main_stream = e->stream;
        e->interest_pending = false;
        if (need_build) {
            st = chd_build(e);
            if (st != CHD_OK) return st;
            CU(e, cudaEventRecord(e->ev_build_done, main_stream));
            e->build_done_recorded = true;
        }
        CU(e, cudaStreamWaitEvent(main_stream, e->emit_variant >= 4 ? e->ev_interest : e->ev_pairs, 0));
        if (do_emit) {
            if (e->join_before_emit || (e->early_results_tick && e->early_mode == 1)) e->wait_before_emit_kernel = e->ev_join;
            st = chd_emit_visible(e);
            if (st != CHD_OK) return st;
        }
        CU(e, cudaStreamWaitEvent(main_stream, e->ev_join, 0));
        if (do_fanout && !e->pending_fanout) {
            st = chd_fanout_tick(e, t_ns);
            if (st != CHD_OK) return st;
        } else {
            e->early_ready = do_emit;  // everything but the expanded list is final at ev_join + ev_prep_done
        }
        if (out) return chd_summary(e, out);
        return CHD_OK;
    }
    if (e->overlap_fanout && e->aux_stream && (q || do_fanout) && (need_build || do_emit)) {
        // Dependency graph of a tick:   build ----------------+--> emit
        //                               interest --> fan-out  |      (emit needs the cell CSR and the new pairs)
        // The build / emit chain (HBM-bound) runs on the main stream, the interest / fan-out chain
        // (latency-bound, disjoint state) on aux_stream; they are joined before the summary.
        cudaStream_t main_stream = e->stream;
        CU(e, cudaEventRecord(e->ev_fork, main_stream));
        CU(e, cudaStreamWaitEvent(e->aux_stream, e->ev_fork, 0));
        std::unique_lock<std::recursive_mutex> redirect(e->mu);  // `stream` is redirected until it is restored below
        e->stream = e->aux_stream;
        st = q ? chd_update_interest(e, q, t_ns) : CHD_OK;
        if (st == CHD_OK) {
            cudaError_t r = cudaEventRecord(e->ev_interest, e->aux_stream);
            if (r != cudaSuccess) st = CHD_ERR_CUDA;
        }
        if (st == CHD_OK && do_fanout) st = chd_fanout_tick(e, t_ns);
        e->stream = main_stream;
        redirect.unlock();
        if (st != CHD_OK) return st;
        CU(e, cudaEventRecord(e->ev_join, e->aux_stream));
        if (need_build) {
            st = chd_build(e);
            if (st != CHD_OK) return st;
            CU(e, cudaEventRecord(e->ev_build_done, main_stream));
            e->build_done_recorded = true;
        }
        if (do_emit) {
            CU(e, cudaStreamWaitEvent(main_stream, (e->emit_variant >= 4 || !q) ? e->ev_interest : e->ev_pairs, 0));
            if (e->join_before_emit || (e->early_results_tick && e->early_mode == 1)) e->wait_before_emit_kernel = e->ev_join;
            st = chd_emit_visible(e);
            if (st != CHD_OK) return st;
        }
        CU(e, cudaStreamWaitEvent(main_stream, e->ev_join, 0));
        e->early_ready = do_emit;
    } else {
        if (need_build) {
            st = chd_build(e);
            if (st != CHD_OK) return st;
        }
        if (q) {
            st = chd_update_interest(e, q, t_ns);
            if (st != CHD_OK) return st;
        }
        if (do_emit) {
            st = chd_emit_visible(e);
            if (st != CHD_OK) return st;
        }
        if (do_fanout) {
            st = chd_fanout_tick(e, t_ns);
            if (st != CHD_OK) return st;
        }
    }
    if (out) return chd_summary(e, out);
    return CHD_OK;
}

/* ------------------------------------------------------------------ results ---- */

static chd_status read_u32(chd_engine* e, const uint32_t* d, uint32_t* v) {
    CU(e, cudaMemcpyAsync(e->h_get, d, 4, cudaMemcpyDeviceToHost, e->stream));
    CU(e, cudaStreamSynchronize(e->stream));
    *v = *e->h_get;
    return CHD_OK;
}

chd_status chd_get_cells(chd_engine* e, uint32_t* cell_start, uint32_t* sorted_entity) {
    if (!e) return CHD_ERR_INVALID;
    CU(e, cudaSetDevice(e->device));
    const uint32_t C = e->g.cells;
    if (cell_start) CU(e, cudaMemcpyAsync(cell_start, e->d_cell_start, sizeof(uint32_t) * ((uint64_t)C + 1), cudaMemcpyDefault, e->stream));
    if (sorted_entity) {
        uint32_t nin = 0;
        chd_status st = read_u32(e, e->d_cell_start + C, &nin);
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
    chd_status st = read_u32(e, pb.off + S, &P);
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
    if (n > e->lim.max_due) return CHD_ERR_CAPACITY;
    if (n > cap) n = cap;
    CU(e, cudaMemcpyAsync(out, e->d_due, sizeof(chd_due) * n, cudaMemcpyDefault, e->stream));
    CU(e, cudaStreamSynchronize(e->stream));
    return CHD_OK;
}

chd_status chd_get_handover(chd_engine* e, uint32_t* entity, uint32_t* src_channel, uint32_t* dst_channel, uint32_t cap) {
    if (!e) return CHD_ERR_INVALID;
    CU(e, cudaSetDevice(e->device));
    uint32_t n = 0;
    chd_status st = read_u32(e, &e->d_ctr->n_handover, &n);
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
    const auto now_us = []() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = e->trace_fetch ? now_us() : 0.0;
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
    const double t1 = e->trace_fetch ? now_us() : 0.0;
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
        st = decode_summary(e, summary);
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
    const double t2 = e->trace_fetch ? now_us() : 0.0;
    CU(e, cudaStreamSynchronize(s));  // last sync of the read-back stream(s)
    if (sa != s) CU(e, cudaStreamSynchronize(sa));
    const double t3 = e->trace_fetch ? now_us() : 0.0;
    if (s != main_stream) CU(e, cudaStreamSynchronize(main_stream));  // the tick itself (expanded list) has finished
    if (e->trace_fetch) {
        const double t4 = now_us();
        e->fetch_t[0] += t1 - t0; e->fetch_t[1] += t2 - t1; e->fetch_t[2] += t3 - t2; e->fetch_t[3] += t4 - t3;
        e->fetch_n++;
    }
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

/* ------------------------------------------------------------------ multi-GPU slab ---- */

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
        chd_status st = assign_cells_impl(e);
        if (st != CHD_OK) return st;
        border_flag_kernel<<<blocks_for(n ? n : 1, 256), 256, 0, s>>>(e->g, e->d_key, n, e->d_bflag, e->d_epoch + EP_BORDER);
        KCHECK(e);
        SCAN(e, exclusive_scan_1p<uint32_t, uint32_t>(e->d_bflag, e->d_boff, n, e->site_border, s));
        border_write_kernel<<<blocks_for(n_launch ? n_launch : 1, 256), 256, 0, s>>>(e->d_key, e->have_gid ? e->d_gid : nullptr, n, e->d_bflag,
                                                                                      e->d_boff, d_records, cap_records, e->d_ctr);
        KCHECK(e);
        return CHD_OK;
    };
    chd_status st;
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
        if (st == CHD_OK) st = note_pos_read(e);
    } else {
        st = enqueue();
    }
    if (st != CHD_OK) return st;
    if (out_count) {
        st = read_u32(e, e->d_boff + n, out_count);
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

/* ------------------------------------------------------------------ instrumentation ---- */

uint64_t chd_launch_count(const chd_engine* e) { return e ? e->n_launch : 0; }

chd_status chd_enable_graphs(chd_engine* e, int on) {
    if (!e) return CHD_ERR_INVALID;
    e->use_graphs = on != 0;
    return CHD_OK;
}

uint64_t chd_graph_launch_count(const chd_engine* e) { return e ? e->graph_launches : 0; }

chd_status chd_profile_enable(chd_engine* e, int on) {
    if (e) { e->fetch_n = 0; e->fetch_t[0] = e->fetch_t[1] = e->fetch_t[2] = e->fetch_t[3] = 0; }
    if (!e) return CHD_ERR_INVALID;
    CU(e, cudaSetDevice(e->device));
    if (on && !e->ev) {
        const size_t n = (size_t)CHD_STAGE_COUNT * chd_engine::EV_RING * 2;
        e->ev = new cudaEvent_t[n]();
        for (size_t i = 0; i < n; i++) CU(e, cudaEventCreate(&e->ev[i]));
    }
    CU(e, cudaStreamSynchronize(e->stream));
    for (int s = 0; s < CHD_STAGE_COUNT; s++) e->stage_n[s] = 0;
    e->profiling = on != 0;
    return CHD_OK;
}

chd_status chd_profile_timeline(chd_engine* e, int stage, double* start_ms, double* stop_ms) {
    if (!e || stage < 0 || stage >= CHD_STAGE_COUNT || !start_ms || !stop_ms || !e->ev) return CHD_ERR_INVALID;
    CU(e, cudaSetDevice(e->device));
    CU(e, cudaStreamSynchronize(e->stream));
    if (e->aux_stream) CU(e, cudaStreamSynchronize(e->aux_stream));
    if (!e->stage_n[CHD_STAGE_TICK] || !e->stage_n[stage]) return CHD_ERR_STATE;
    cudaEvent_t t0 = e->evt(CHD_STAGE_TICK, e->stage_n[CHD_STAGE_TICK] - 1, 0);
    float a = 0, b = 0;
    CU(e, cudaEventElapsedTime(&a, t0, e->evt(stage, e->stage_n[stage] - 1, 0)));
    CU(e, cudaEventElapsedTime(&b, t0, e->evt(stage, e->stage_n[stage] - 1, 1)));
    *start_ms = a;
    *stop_ms = b;
    return CHD_OK;
}

chd_status chd_profile_get(chd_engine* e, int stage, double* total_ms, uint64_t* samples) {
    if (!e || stage < 0 || stage >= CHD_STAGE_COUNT || !total_ms || !samples) return CHD_ERR_INVALID;
    CU(e, cudaSetDevice(e->device));
    CU(e, cudaStreamSynchronize(e->stream));
    const uint64_t n = e->stage_n[stage];
    const uint64_t m = n < (uint64_t)chd_engine::EV_RING ? n : (uint64_t)chd_engine::EV_RING;
    double tot = 0;
    for (uint64_t i = n - m; i < n; i++) {
        float ms = 0;
        CU(e, cudaEventElapsedTime(&ms, e->evt(stage, i, 0), e->evt(stage, i, 1)));
        tot += ms;
    }
    *total_ms = tot;
    *samples = m;
    return CHD_OK;
}

/* ------------------------------------------------------------------ host helpers ---- */

/* ------------------------------------------------------------------ window classes of the due list ---- */

chd_status chd_due_classes(chd_engine* e, uint32_t* out_class_of, uint32_t* out_class_rep, uint32_t* out_class_count, uint32_t cap_classes,
                           uint32_t* out_n_classes) {
    if (!e) return CHD_ERR_INVALID;
    CU(e, cudaSetDevice(e->device));
    cudaStream_t s = e->stream;
    uint32_t n_due = 0;
    chd_status st = read_u32(e, &e->d_ctr->n_due, &n_due);
    if (st != CHD_OK) return st;
    if (n_due > e->lim.max_due) n_due = e->lim.max_due;
    if (out_n_classes) *out_n_classes = 0;
    if (n_due == 0) return CHD_OK;
    const uint64_t D = e->lim.max_due;
    if (!e->d_cls_table) {
        uint32_t T = 1024;
        while ((uint64_t)T < 2 * D) T <<= 1;
        if (!(dalloc(e, &e->d_cls_table, (uint64_t)T) && dalloc(e, &e->d_cls_rep, (uint64_t)T) && dalloc(e, &e->d_cls_cnt, (uint64_t)T) &&
              dalloc(e, &e->d_cls_slot, D) && dalloc(e, &e->d_cls_flag, D + 1) && dalloc(e, &e->d_cls_rank, D + 1) && dalloc(e, &e->d_cls_of, D) &&
              dalloc(e, &e->d_cls_out_rep, D) && dalloc(e, &e->d_cls_out_cnt, D) && make_site(e, e->site_class, D + 1, EP_CLASS)))
            return CHD_ERR_CUDA;
        e->site_class.error = &e->d_ctr->overflow;
        e->cls_table_size = T;
    }
    // a table of >= 2 n slots is enough for this call: clear only that much
    uint32_t T = 1024;
    while ((uint64_t)T < 2ull * n_due) T <<= 1;
    CU(e, cudaMemsetAsync(e->d_cls_table, 0xFF, 4ull * T, s));
    CU(e, cudaMemsetAsync(e->d_cls_rep, 0xFF, 4ull * T, s));
    CU(e, cudaMemsetAsync(e->d_cls_cnt, 0, 4ull * T, s));
    const uint32_t* n_ptr = &e->d_ctr->n_due;
    const unsigned blocks = blocks_for(n_due, 256);
    class_insert_kernel<<<blocks, 256, 0, s>>>(e->d_due, e->d_due_key, n_ptr, e->lim.max_due, e->d_cls_table, T - 1, e->d_cls_slot, e->d_cls_rep,
                                               e->d_cls_cnt);
    KCHECK(e);
    class_flag_kernel<<<blocks, 256, 0, s>>>(n_ptr, e->lim.max_due, e->d_cls_slot, e->d_cls_rep, e->d_cls_flag, e->d_epoch + EP_CLASS);
    KCHECK(e);
    SCAN(e, exclusive_scan_1p<uint32_t, uint32_t>(e->d_cls_flag, e->d_cls_rank, n_due, e->site_class, s));
    class_finish_kernel<<<blocks, 256, 0, s>>>(n_ptr, e->lim.max_due, e->d_cls_slot, e->d_cls_rep, e->d_cls_cnt, e->d_cls_rank, e->d_cls_of,
                                               e->d_cls_out_rep, e->d_cls_out_cnt);
    KCHECK(e);
    uint32_t n_classes = 0;
    st = read_u32(e, e->d_cls_rank + n_due, &n_classes);
    if (st != CHD_OK) return st;
    if (out_n_classes) *out_n_classes = n_classes;
    if ((out_class_rep || out_class_count) && n_classes > cap_classes) {
        e->fail("chd_due_classes: %u classes > cap_classes %u", n_classes, cap_classes);
        return CHD_ERR_CAPACITY;
    }
    if (out_class_of) CU(e, cudaMemcpyAsync(out_class_of, e->d_cls_of, 4ull * n_due, cudaMemcpyDefault, s));
    if (out_class_rep) CU(e, cudaMemcpyAsync(out_class_rep, e->d_cls_out_rep, 4ull * n_classes, cudaMemcpyDefault, s));
    if (out_class_count) CU(e, cudaMemcpyAsync(out_class_count, e->d_cls_out_cnt, 4ull * n_classes, cudaMemcpyDefault, s));
    CU(e, cudaStreamSynchronize(s));
    return CHD_OK;
}

/* ------------------------------------------------------------------ ADJACENT_CHANNELS broadcast sets ---- */

__global__ void bcast_msgoff_kernel(const uint32_t* __restrict__ off9, uint32_t n, uint32_t* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n) out[i] = off9[(uint64_t)i * 9];
}

static void dfree(chd_engine* e, void* p) {
    if (!p) return;
    for (size_t i = 0; i < e->allocs.size(); i++)
        if (e->allocs[i] == p) {
            e->allocs.erase(e->allocs.begin() + (long)i);
            break;
        }
    cudaFree(p);
}

chd_status chd_set_subscriber_types(chd_engine* e, const uint8_t* conn_type, uint32_t n) {
    if (!e || n > e->lim.max_subscribers) return CHD_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> lk(e->mu);
    CU(e, cudaSetDevice(e->device));
    if (!conn_type) {
        e->have_conn_type = false;
        return CHD_OK;
    }
    if (!e->d_conn_type && !dalloc(e, &e->d_conn_type, (uint64_t)e->lim.max_subscribers)) return CHD_ERR_CUDA;
    CU(e, cudaMemsetAsync(e->d_conn_type, 0, e->lim.max_subscribers, e->stream));
    CU(e, cudaMemcpyAsync(e->d_conn_type, conn_type, n, cudaMemcpyDefault, e->stream));
    CU(e, cudaStreamSynchronize(e->stream));
    e->have_conn_type = true;
    return CHD_OK;
}

chd_status chd_adjacent_broadcast(chd_engine* e, const chd_broadcast_batch* b, uint32_t* out_status, uint32_t* out_off, uint32_t* out_slot,
                                  uint64_t cap) {
    if (!e || !b || (b->n && (!b->channel_id || !b->broadcast)) || !out_off) return CHD_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> lk(e->mu);
    if (e->interest_pending) {
        e->fail("chd_adjacent_broadcast while a chd_begin_interest is pending (call chd_tick first)");
        return CHD_ERR_STATE;
    }
    CU(e, cudaSetDevice(e->device));
    const uint32_t n = b->n;
    if (out_status)
        for (uint32_t m = 0; m < n; m++)
            out_status[m] = (b->channel_id[m] >= e->g.id_start && b->channel_id[m] - e->g.id_start < e->g.cells) ? CHD_BC_OK : CHD_BC_ERR_NOT_A_CELL;
    if (n == 0 || !e->by_cell_valid || e->n_slots == 0) {  // no messages, or nobody is subscribed to anything yet
        for (uint32_t m = 0; m <= n; m++) out_off[m] = 0;
        return CHD_OK;
    }
    cudaStream_t s = e->stream;
    if (e->aux_stream) CU(e, cudaStreamSynchronize(e->aux_stream));  // an interest update in flight is rewriting the pairs / by-cell order
    if (n > e->bc_msg_cap) {
        const uint64_t c = std::max<uint64_t>(1024, (uint64_t)n + n / 2);
        dfree(e, e->d_bc_in); dfree(e, e->d_bc_cnt); dfree(e, e->d_bc_off); dfree(e, e->d_bc_msgoff); dfree(e, e->site_bcast.desc);
        e->d_bc_in = e->d_bc_cnt = e->d_bc_off = e->d_bc_msgoff = nullptr;
        e->site_bcast.desc = nullptr;
        e->bc_msg_cap = 0;
        if (!dalloc(e, &e->d_bc_in, 4 * c) || !dalloc(e, &e->d_bc_cnt, 9 * c + 1) || !dalloc(e, &e->d_bc_off, 9 * c + 1) ||
            !dalloc(e, &e->d_bc_msgoff, c + 1) || !make_site(e, e->site_bcast, 9 * c + 1, EP_BCAST))
            return CHD_ERR_CUDA;
        e->site_bcast.error = &e->d_ctr->overflow;
        e->bc_msg_cap = c;
    }
    const uint64_t mc = e->bc_msg_cap;
    CU(e, cudaMemcpyAsync(e->d_bc_in, b->channel_id, 4ull * n, cudaMemcpyDefault, s));
    CU(e, cudaMemcpyAsync(e->d_bc_in + mc, b->broadcast, 4ull * n, cudaMemcpyDefault, s));
    if (b->sender_conn_id) CU(e, cudaMemcpyAsync(e->d_bc_in + 2 * mc, b->sender_conn_id, 4ull * n, cudaMemcpyDefault, s));
    else CU(e, cudaMemsetAsync(e->d_bc_in + 2 * mc, 0, 4ull * n, s));
    if (b->client_conn_id) CU(e, cudaMemcpyAsync(e->d_bc_in + 3 * mc, b->client_conn_id, 4ull * n, cudaMemcpyDefault, s));
    else CU(e, cudaMemsetAsync(e->d_bc_in + 3 * mc, 0, 4ull * n, s));
    const BcastDev bd{n, e->d_bc_in, e->d_bc_in + mc, e->d_bc_in + 2 * mc, e->d_bc_in + 3 * mc};
    PairBuf& pb = e->pairs[e->cur];
    const uint32_t S = e->n_slots;
    const uint64_t P = e->lim.max_pairs;
    const uint8_t* types = e->have_conn_type ? e->d_conn_type : nullptr;
    const unsigned blocks = blocks_for(9ull * n * 32, 256);
    bcast_kernel<false><<<blocks, 256, 0, s>>>(e->g, bd, pb.off + S, P, pb, e->d_by_cell, e->d_conn, types, e->d_bc_cnt, nullptr, nullptr, 0,
                                               e->d_epoch + EP_BCAST);
    KCHECK(e);
    SCAN(e, exclusive_scan_1p<uint32_t, uint32_t>(e->d_bc_cnt, e->d_bc_off, 9ull * n, e->site_bcast, s));
    // (callable from any thread under the engine mutex: h_u32 is the mutex-protected scratch, as in chd_query_channel_ids)
    CU(e, cudaMemcpyAsync(e->h_u32, e->d_bc_off + 9ull * n, 4, cudaMemcpyDeviceToHost, s));
    CU(e, cudaStreamSynchronize(s));
    const uint32_t total = *e->h_u32;
    if (total > cap || (total && !out_slot)) {
        e->fail("chd_adjacent_broadcast: %u recipients > capacity %llu", total, (unsigned long long)cap);
        return CHD_ERR_CAPACITY;
    }
    if (total > e->bc_out_cap) {
        dfree(e, e->d_bc_out);
        e->d_bc_out = nullptr;
        e->bc_out_cap = 0;
        const uint64_t c = std::max<uint64_t>(1 << 16, (uint64_t)total + total / 2);
        if (!dalloc(e, &e->d_bc_out, c)) return CHD_ERR_CUDA;
        e->bc_out_cap = c;
    }
    if (total) {
        bcast_kernel<true><<<blocks, 256, 0, s>>>(e->g, bd, pb.off + S, P, pb, e->d_by_cell, e->d_conn, types, nullptr, e->d_bc_off, e->d_bc_out,
                                                  e->bc_out_cap, nullptr);
        KCHECK(e);
        CU(e, cudaMemcpyAsync(out_slot, e->d_bc_out, 4ull * total, cudaMemcpyDefault, s));
    }
    bcast_msgoff_kernel<<<blocks_for((uint64_t)n + 1, 256), 256, 0, s>>>(e->d_bc_off, n, e->d_bc_msgoff);
    KCHECK(e);
    CU(e, cudaMemcpyAsync(out_off, e->d_bc_msgoff, 4ull * ((uint64_t)n + 1), cudaMemcpyDefault, s));
    CU(e, cudaStreamSynchronize(s));
    return CHD_OK;
}

uint32_t chd_get_adjacent_channels(const chd_grid_cfg* cfg, uint32_t channel_id, uint32_t* out8) {  // spatial.go:358-381
    if (!cfg || !out8 || cfg->grid_cols == 0) return 0;
    const uint32_t index = channel_id - cfg->channel_id_start;
    const int64_t gx = index % cfg->grid_cols, gy = index / cfg->grid_cols;
    uint32_t n = 0;
    for (int64_t y = gy - 1; y <= gy + 1; y++) {
        if (y < 0 || y >= (int64_t)cfg->grid_rows) continue;
        for (int64_t x = gx - 1; x <= gx + 1; x++) {
            if (x < 0 || x >= (int64_t)cfg->grid_cols) continue;
            if (x == gx && y == gy) continue;
            out8[n++] = (uint32_t)x + (uint32_t)y * cfg->grid_cols + cfg->channel_id_start;
        }
    }
    return n;
}

chd_status chd_get_regions(const chd_grid_cfg* cfg, double* min_x, double* min_z, double* max_x, double* max_z, uint32_t* channel_id,
                           uint32_t* server_index) {  // spatial.go:319-356
    if (!cfg || cfg->server_cols == 0 || cfg->server_rows == 0) return CHD_ERR_INVALID;
    uint32_t sgc = cfg->grid_cols / cfg->server_cols;
    if (cfg->grid_cols % cfg->server_cols) sgc++;
    uint32_t sgr = cfg->grid_rows / cfg->server_rows;
    if (cfg->grid_rows % cfg->server_rows) sgr++;
    for (uint32_t y = 0; y < cfg->grid_rows; y++)
        for (uint32_t x = 0; x < cfg->grid_cols; x++) {
            const uint32_t i = x + y * cfg->grid_cols;
            if (min_x) min_x[i] = cfg->world_offset_x + cfg->grid_width * (double)x;
            if (min_z) min_z[i] = cfg->world_offset_z + cfg->grid_height * (double)y;
            if (max_x) max_x[i] = cfg->world_offset_x + cfg->grid_width * (double)(x + 1);
            if (max_z) max_z[i] = cfg->world_offset_z + cfg->grid_height * (double)(y + 1);
            if (channel_id) channel_id[i] = cfg->channel_id_start + i;
            if (server_index) server_index[i] = (x / sgc) + (y / sgr) * cfg->server_cols;
        }
    return CHD_OK;
}

}  // extern "C"
