// dlsm.cu -- C-ABI (include/dlsm.h) over the sm_100a kernels in dlsm_kernels.cuh.
// No CPU fallback: every compute entry point needs a CUDA device.
#include "../../include/dlsm.h"
#include "dlsm_kernels.cuh"
#include "dlsm_hdp.cuh"
#include "dlsm_trace.cuh"
#include "dlsm_blk.h"
#include "dlsm_graph.h"
#include "dlsm_cc.h"
#include "dlsm_ccd.h"
#include "dlsm_fullr.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace dlsm;

namespace {

thread_local std::string g_create_error;

struct EventPair { cudaEvent_t a, b; int cat; };

} // namespace

struct dlsm_handle {
    dlsm_config cfg;
    dlsm_hyper hy;
    int lk;                 // Lik
    int W;                  // words per adjacency row
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaStream_t side_stream = nullptr;          // label FFBS + HDP update overlap the intercept MH
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    std::string err;
    // network
    uint32_t *rowbits = nullptr, *colbits = nullptr;
    int32_t *deg = nullptr, *in_edges = nullptr, *out_edges = nullptr;
    int32_t *ctrl_in = nullptr, *ctrl_out = nullptr;
    int max_in = 0, max_out = 0, n_control = 0, ctrl_sets = 0;
    int32_t *d_cc_dep = nullptr;     // [ctrl_sets][T][n] batch schedule of the case-control sweep
    bool cc_dep_valid = false;
    bool have_net = false, have_edges = false, have_ctrl = false;
    // state
    void *field[DLSM_F_COUNT_] = {nullptr};
    size_t field_bytes[DLSM_F_COUNT_] = {0};
    double *rinv = nullptr;
    // scratch
    double *d_eps = nullptr, *d_logu = nullptr, *d_ratio = nullptr, *d_out = nullptr;
    int32_t *d_acc = nullptr;
    double *d_partial = nullptr, *d_bvar = nullptr, *d_prop = nullptr, *d_ll2 = nullptr;
    double *d_rprop = nullptr, *d_rprop_inv = nullptr;
    double *d_small = nullptr;   // [C][4] replay scalars
    int32_t *d_small_i = nullptr;
    unsigned int *d_flags = nullptr;
    int *d_bad = nullptr;
    int full_tiles = 1, full_nblk = 1;
    int *d_progress = nullptr;      // [C][T] wavefront flags of the CTA-per-slice sweep
    unsigned int *d_ticket = nullptr;
    double *d_ffbs_stage = nullptr;  // global (L2-resident) stage of the thread-per-node label kernel
    size_t l2_persist_max = 0, l2_persist_set = 0; // persisting-L2 limits (bytes): device maximum, currently set
    size_t ffbs_stage_bytes = 0;
    int sweep_mode = 0;             // 0 auto, 1 CTA per chain, 2 CTA per (chain, slice)
    int sm_count = 148;
    bool dense_build = false;
    bool no_pipeline = false;       // DLSM_SWEEP_MODE=slice-plain: the unpipelined slice kernel
    // row-sum cache of the exact likelihoods (see node_loglik1 / k_rows): rows[c][t][j], scratch of the
    // proposal's pair terms, and the tile vectors k_rows leaves for k_rows_commit
    double *d_rows = nullptr, *d_scr = nullptr, *d_rows_own = nullptr, *d_rows_part = nullptr;
    int32_t *d_accflag = nullptr;
    bool rows_valid = false;
    int sweeps_since_set = 0;       // device-loop sweeps since the state last changed from outside
    int rows_nb = 0, rows_half = 0, rows_R = 1, rows_L = 0, rows_ipc = 0, rows_ns = 1, rows_grid = 1;
    int cluster_cs = -1;            // CTAs per (chain, slice) cluster (-1: not probed, 0: none)
    int cluster_ncomp = 0;          // compute warps per CTA (per-node cluster kernel) / warps per CTA (block kernel)
    bool cluster_blk = false;       // block-speculative kernel (k_sweep_blk) rather than the per-node one
    bool cluster_win = false;       // ... with the two-block window (k_sweep_blkw)
    double *d_ll_slices = nullptr;  // [C][T] per-slice log-likelihood sums of k_sweep_blkw
    int cc3_ok = -1;                // 2-CTA-cluster case-control sweep: all C*T clusters co-resident? (-1: not probed)
    // developer options (dlsm_set_option; environment defaults are read ONCE, in dlsm_create)
    int64_t opt[DLSM_OPT_COUNT_] = {0};
    // rng
    uint64_t seed = 0, chain_offset = 0;
    uint32_t sweep_idx[5] = {0, 0, 0, 0, 0};
    uint32_t control_draws = 0;     // Philox sweep index of dlsm_resample_controls
    dlsm_hdp_prior hdp_prior;
    bool have_hdp_prior = false;
    // trace pipeline (dlsm_run_traced): two device chunks of trace_R records each, drained to the
    // host on copy_stream while the next chunk fills
    cudaStream_t copy_stream = nullptr;
    struct TraceChunk {
        void *dev[DLSM_F_COUNT_] = {nullptr};
        double *logp = nullptr;
        cudaEvent_t filled = nullptr, drained = nullptr;
        bool used = false;
    } chunk[2];
    size_t trace_slot[DLSM_F_COUNT_] = {0}; // bytes one record takes per field (0 = not traced)
    uint32_t trace_all = 0, trace_first = 0;
    int trace_logp = 0, trace_R = 0;
    // early copy of the positions: X is final once it is centred, so its record can leave for the host
    // while the rest of the sweep (intercept / radii MH, labels, HDP block) still runs
    char *early_x_dst = nullptr;     // host address of this sweep's X record, or null
    cudaEvent_t ev_x_ready = nullptr, ev_x_copied = nullptr;
    cudaStream_t x_stream = nullptr; // its own stream: never queued behind a ring drain
    bool x_copy_pending = false;
    bool early_x_active = false;     // this dlsm_run_traced call bypasses the ring for X
    bool trace_early_x = false;      // ... and the ring was allocated without an X slot
    double *d_logp = nullptr;       // [C] scratch of dlsm_logp
    uint32_t *d_cooc = nullptr;     // [T][n][n] co-clustering counts accumulated by dlsm_run_traced
    uint64_t cooc_samples = 0;
    double *d_gather = nullptr;     // [C][T][n][4] packed {x, y, 1/r, 0} records of the case-control kernels
    CcdWork *ccd = nullptr;         // buffers of the dataflow case-control sweep (dlsm_ccd.cu)
    double *d_radii_terms = nullptr; // [C][chunks][6] Dirichlet terms / gamma totals of the radii MH
    double *d_center = nullptr;     // means [C][8] + partial sums [C][128][8] of the long-chain centring
    double *d_proc_ref = nullptr;   // [C][T][n][d] reference configuration of the in-loop Procrustes
    bool have_proc_ref = false;
    // developer timeline (DLSM_TIMELINE=1): start/stop of every launch group on its own stream
    struct TimelineEntry { cudaEvent_t a, b; const char *name; };
    std::vector<TimelineEntry> timeline;
    bool timeline_on = false;
    // counters
    dlsm_counters ctr;
    bool timing = false;
    std::vector<EventPair> events;
};

#define FAIL(h, code, ...)                                          \
    do {                                                            \
        char _b[512];                                               \
        snprintf(_b, sizeof(_b), __VA_ARGS__);                      \
        (h)->err = _b;                                              \
        return (code);                                              \
    } while (0)

#define CU(h, x)                                                                              \
    do {                                                                                      \
        cudaError_t _e = (x);                                                                 \
        if (_e != cudaSuccess)                                                                \
            FAIL(h, DLSM_ERR_CUDA, "%s failed: %s (%s:%d)", #x, cudaGetErrorString(_e), __FILE__, \
                 __LINE__);                                                                   \
    } while (0)

#define CHECK_LAUNCH(h) CU(h, cudaGetLastError())

namespace {

size_t elem_size(int f)
{
    switch (f) {
    case DLSM_F_Z: case DLSM_F_X_NACC: case DLSM_F_X_NSTEPS: case DLSM_F_X_UNTIL:
    case DLSM_F_B_NACC: case DLSM_F_B_NSTEPS: case DLSM_F_B_UNTIL: case DLSM_F_R_NACC:
    case DLSM_F_R_NSTEPS: case DLSM_F_R_UNTIL: case DLSM_F_NK:
        return 4;
    default:
        return 8;
    }
}

size_t field_elems(const dlsm_config &c, int f)
{
    const size_t C = c.n_chains, T = c.T, n = c.n, d = c.d, K = c.K;
    switch (f) {
    case DLSM_F_X: return C * T * n * d;
    case DLSM_F_INTERCEPT: case DLSM_F_B_STEP: case DLSM_F_B_NACC: case DLSM_F_B_NSTEPS:
    case DLSM_F_B_UNTIL: return C * 2;
    case DLSM_F_RADII: return C * n;
    case DLSM_F_Z: return C * T * n;
    case DLSM_F_MU: return C * K * d;
    case DLSM_F_SIGMA: return C * K;
    case DLSM_F_LAMBDA: return C;
    case DLSM_F_WEIGHTS: case DLSM_F_NCOUNT: return C * T * K * K;
    case DLSM_F_X_STEP: case DLSM_F_X_NACC: case DLSM_F_X_NSTEPS: case DLSM_F_X_UNTIL:
        return C * T * n;
    case DLSM_F_R_STEP: case DLSM_F_R_NACC: case DLSM_F_R_NSTEPS: case DLSM_F_R_UNTIL: return C;
    case DLSM_F_NK: return C * T * K;
    case DLSM_F_BETA: return C * K;
    case DLSM_F_HYPER: return K > 0 ? C * 8 : 0;
    case DLSM_F_LOGLIK: return C;
    default: return 0;
    }
}

template <typename T> T *F(dlsm_handle *h, int f) { return static_cast<T *>(h->field[f]); }

void tl_begin(dlsm_handle *h, const char *name)
{
    if (!h->timeline_on) return;
    dlsm_handle::TimelineEntry e;
    cudaEventCreate(&e.a);
    cudaEventCreate(&e.b);
    e.name = name;
    cudaEventRecord(e.a, h->stream);
    h->timeline.push_back(e);
}

void tl_end(dlsm_handle *h)
{
    if (!h->timeline_on || h->timeline.empty()) return;
    cudaEventRecord(h->timeline.back().b, h->stream);
}

void tl_dump(dlsm_handle *h)
{
    if (!h->timeline_on || h->timeline.empty()) return;
    cudaDeviceSynchronize();
    for (auto &e : h->timeline) {
        float t0 = 0.f, t1 = 0.f;
        cudaEventElapsedTime(&t0, h->timeline.front().a, e.a);
        cudaEventElapsedTime(&t1, h->timeline.front().a, e.b);
        fprintf(stderr, "[dlsm timeline] %-18s %9.1f -> %9.1f us  (%7.1f)\n", e.name, t0 * 1e3, t1 * 1e3,
                (t1 - t0) * 1e3);
    }
    for (auto &e : h->timeline) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    h->timeline.clear();
}

void begin_phase(dlsm_handle *h, int cat)
{
    if (!h->timing) return;
    EventPair e;
    cudaEventCreate(&e.a);
    cudaEventCreate(&e.b);
    e.cat = cat;
    cudaEventRecord(e.a, h->stream);
    h->events.push_back(e);
}

void flush_events(dlsm_handle *h)
{
    if (h->events.empty()) return;
    cudaStreamSynchronize(h->stream);
    for (auto &e : h->events) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess) {
            if (e.cat == 0) h->ctr.latent_ms += ms; else h->ctr.other_ms += ms;
        }
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
    }
    h->events.clear();
}

void end_phase(dlsm_handle *h)
{
    if (!h->timing) return;
    cudaEventRecord(h->events.back().b, h->stream);
    if (h->events.size() > 2048) flush_events(h);
}

// Developer / test overrides.  Every one is a typed option of the handle (dlsm_set_option); the
// environment only supplies DEFAULTS, read once here -- nothing on the per-sweep path calls getenv.
void apply_sweep_mode(dlsm_handle *h)
{
    const int64_t m = h->opt[DLSM_OPT_SWEEP_MODE];
    h->sweep_mode = (m == DLSM_SWEEP_CHAIN || m == DLSM_SWEEP_CHAIN_DENSE) ? 1
                    : ((m == DLSM_SWEEP_SLICE || m == DLSM_SWEEP_SLICE_PLAIN) ? 2 : 0);
    h->dense_build = m == DLSM_SWEEP_CHAIN_DENSE; // the many-chains register budget, whatever C is
    h->no_pipeline = m == DLSM_SWEEP_SLICE_PLAIN;
}

void read_env_options(dlsm_handle *h)
{
    auto on = [](const char *name) { return getenv(name) != nullptr; };
    h->timeline_on = on("DLSM_TIMELINE");
    if (const char *m = getenv("DLSM_SWEEP_MODE")) {
        h->opt[DLSM_OPT_SWEEP_MODE] = !strcmp(m, "chain") ? DLSM_SWEEP_CHAIN
                                      : !strcmp(m, "chain-dense") ? DLSM_SWEEP_CHAIN_DENSE
                                      : !strcmp(m, "slice") ? DLSM_SWEEP_SLICE
                                      : !strcmp(m, "slice-plain") ? DLSM_SWEEP_SLICE_PLAIN : DLSM_SWEEP_AUTO;
    }
    if (const char *m = getenv("DLSM_FFBS"))
        h->opt[DLSM_OPT_FFBS_KERNEL] = !strcmp(m, "warp") ? DLSM_FFBS_WARP : DLSM_FFBS_THREAD;
    h->opt[DLSM_OPT_FFBS_SMEM_STAGE] = on("DLSM_FFBS_SMEM");
    if (const char *m = getenv("DLSM_FFBS_PER_SM")) h->opt[DLSM_OPT_FFBS_CTAS_PER_SM] = atoll(m);
    h->opt[DLSM_OPT_NO_GATHER_PACK] = on("DLSM_NO_GATHER_PACK");
    h->opt[DLSM_OPT_NO_TRACKED_LOGLIK] = on("DLSM_NO_LLCUR");
    h->opt[DLSM_OPT_CENTER_EXACT] = on("DLSM_CENTER_EXACT");
    h->opt[DLSM_OPT_HDP_SEGMENTED] = on("DLSM_HDP_SEGMENTED");
    h->opt[DLSM_OPT_NO_EARLY_X] = on("DLSM_NO_EARLY_X");
    h->opt[DLSM_OPT_NO_ROWSUM_CACHE] = on("DLSM_NO_ROWSUM");
    if (const char *m = getenv("DLSM_CHAIN_KERNEL"))
        h->opt[DLSM_OPT_CHAIN_KERNEL] = !strcmp(m, "block") ? DLSM_CHAIN_BLOCK
                                        : !strcmp(m, "block2") ? DLSM_CHAIN_BLOCK_PAIR
                                        : !strcmp(m, "node") ? DLSM_CHAIN_NODE
                                        : !strcmp(m, "rowsum") ? DLSM_CHAIN_NODE_ROWSUM : DLSM_CHAIN_AUTO;
    if (const char *m = getenv("DLSM_CC_KERNEL")) h->opt[DLSM_OPT_CC_KERNEL] = atoll(m);
    if (const char *m = getenv("DLSM_FULL_KERNEL")) h->opt[DLSM_OPT_FULL_KERNEL] = atoll(m);
    if (const char *m = getenv("DLSM_CCD_GROUP")) h->opt[DLSM_OPT_CCD_GROUP] = atoll(m);
    h->opt[DLSM_OPT_FFBS_NO_L2_WINDOW] = on("DLSM_FFBS_NO_L2_WINDOW");
    if (const char *m = getenv("DLSM_NO_CLUSTER")) h->opt[DLSM_OPT_NO_CLUSTER] = atoll(m) > 0 ? atoll(m) : 1;
    if (const char *m = getenv("DLSM_TRACE_CHUNK_BYTES")) h->opt[DLSM_OPT_TRACE_CHUNK_BYTES] = atoll(m);
    apply_sweep_mode(h);
}

NetView net_view(const dlsm_handle *h)
{
    NetView v;
    v.T = h->cfg.T; v.n = h->cfg.n; v.d = h->cfg.d; v.W = h->W;
    v.rowbits = h->rowbits; v.colbits = h->colbits;
    v.deg = h->deg; v.in_edges = h->in_edges; v.out_edges = h->out_edges;
    v.ctrl_in = h->ctrl_in; v.ctrl_out = h->ctrl_out;
    v.max_in = h->max_in; v.max_out = h->max_out; v.n_control = h->n_control;
    v.ctrl_per_chain = h->ctrl_sets > 1 ? 1 : 0;
    return v;
}

int need_inputs(dlsm_handle *h)
{
    if (h->lk == kCaseControl) {
        if (!h->have_edges || !h->have_ctrl)
            FAIL(h, DLSM_ERR_NOTSET, "case-control likelihood needs dlsm_set_edge_lists + dlsm_set_controls");
    } else if (!h->have_net) {
        FAIL(h, DLSM_ERR_NOTSET, "network not set (dlsm_set_network_dense)");
    }
    return DLSM_OK;
}

SweepParams sweep_params(dlsm_handle *h)
{
    SweepParams p;
    memset(&p, 0, sizeof(p));
    p.net = net_view(h);
    p.C = h->cfg.n_chains; p.K = h->cfg.K; p.prior = h->cfg.prior;
    p.tune = h->cfg.tune; p.tune_interval = h->cfg.tune_interval;
    p.X = F<double>(h, DLSM_F_X);
    p.intercept = F<double>(h, DLSM_F_INTERCEPT);
    p.rinv = h->rinv;
    p.z = F<int32_t>(h, DLSM_F_Z);
    p.mu = F<double>(h, DLSM_F_MU);
    p.sigma = F<double>(h, DLSM_F_SIGMA);
    p.lambda = F<double>(h, DLSM_F_LAMBDA);
    p.tau_sq = h->hy.tau_sq; p.sigma_sq = h->hy.sigma_sq;
    p.step = F<double>(h, DLSM_F_X_STEP);
    p.nacc = F<int32_t>(h, DLSM_F_X_NACC);
    p.nsteps = F<int32_t>(h, DLSM_F_X_NSTEPS);
    p.until = F<int32_t>(h, DLSM_F_X_UNTIL);
    p.seed = h->seed;
    p.sweep = h->sweep_idx[kRngLatent];
    p.chain_offset = (uint32_t)h->chain_offset;
    p.flags = h->d_flags;
    return p;
}

size_t sweep_smem(const dlsm_handle *h, bool xs)
{
    const size_t x = (size_t)h->cfg.T * sweep_rows_padded(h->lk == kUndirected, h->cfg.n) * h->cfg.d * sizeof(double);
    const int warps = h->cfg.T < 16 ? h->cfg.T : 16;
    return (xs ? x : 0) + warps * sweep_stage_doubles(h->cfg.d) * sizeof(double) +
           (size_t)h->cfg.T * sizeof(int) + 16;
}

constexpr size_t kMaxSmem = 227 * 1024;

template <int LK, int D, bool XS, int MAXT, int MINB>
int launch_sweep_t(dlsm_handle *h, const SweepParams &p, int warps)
{
    const size_t smem = sweep_smem(h, XS);
    if (LK != kCaseControl && p.rows) { // proposal-only evaluation on the row-sum cache
        auto kern = k_sweep<LK == kCaseControl ? kDirected : LK, D, XS, MAXT, MINB, true>;
        CU(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<h->cfg.n_chains, warps * 32, smem, h->stream>>>(p);
    } else {
        auto kern = k_sweep<LK, D, XS, MAXT, MINB>;
        CU(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<h->cfg.n_chains, warps * 32, smem, h->stream>>>(p);
    }
    CHECK_LAUNCH(h);
    return DLSM_OK;
}

#ifndef DLSM_SWEEP_MINB
#define DLSM_SWEEP_MINB 3
#endif
// Register budget follows the block size: chains with few time slices run 3 CTAs per SM.
template <int LK, int D, bool XS>
int launch_sweep_x(dlsm_handle *h, const SweepParams &p)
{
    const int warps = h->cfg.T < 16 ? h->cfg.T : 16;
    // At most one chain per SM (the reference's own single-chain use): nothing to share the
    // register file with, so the compiler may keep all four softplus chains of a trip in flight
    // (156 registers; 15 % lower sweep latency than the 72-register build, which serialises them
    // and relies on 27 resident warps to fill the gaps).
    if (h->cfg.n_chains <= h->sm_count && !h->dense_build) {
        if (warps <= 9) return launch_sweep_t<LK, D, XS, 288, 1>(h, p, warps);
        return launch_sweep_t<LK, D, XS, 512, 1>(h, p, warps);
    }
    // shared memory admits only two CTAs per SM (cfg 4: 98 KB per chain): give them the registers
    if (warps <= 10 && 3 * sweep_smem(h, XS) > kMaxSmem && !h->dense_build)
        return launch_sweep_t<LK, D, XS, 320, 2>(h, p, warps);
    if (warps <= 9) return launch_sweep_t<LK, D, XS, 288, DLSM_SWEEP_MINB>(h, p, warps);
    if (warps <= 10) return launch_sweep_t<LK, D, XS, 320, 3>(h, p, warps);
    return launch_sweep_t<LK, D, XS, 512, 2>(h, p, warps);
}

template <int LK>
int launch_sweep_lk(dlsm_handle *h, const SweepParams &p)
{
    const bool xs = sweep_smem(h, true) <= kMaxSmem;
    if (h->cfg.d == 2) return xs ? launch_sweep_x<LK, 2, true>(h, p) : launch_sweep_x<LK, 2, false>(h, p);
    return xs ? launch_sweep_x<LK, 0, true>(h, p) : launch_sweep_x<LK, 0, false>(h, p);
}

// ---- CTA-per-(chain, slice) variant ---------------------------------------------------------
int slice_warps(const dlsm_handle *h)
{
    int work = h->cfg.n;
    if (h->lk == kCaseControl) work = h->max_in + h->max_out + 2 * h->n_control;
    int nw = (work + 63) / 64;
    return nw < 1 ? 1 : (nw > 16 ? 16 : nw);
}

size_t slice_smem(const dlsm_handle *h, bool xs, int nw)
{
    const size_t x = (size_t)h->cfg.n * (h->cfg.d + (h->lk == kUndirected ? 0 : 1)) * sizeof(double);
    return (xs ? x : 0) + (sweep_stage_doubles(h->cfg.d) + 2 * (size_t)nw) * sizeof(double) + 16;
}

template <int LK, int D, bool XS>
int launch_slice_t(dlsm_handle *h, const SweepParams &p, int nw)
{
    const size_t CT = (size_t)h->cfg.n_chains * h->cfg.T;
    CU(h, cudaMemsetAsync(h->d_progress, 0, CT * sizeof(int), h->stream));
    CU(h, cudaMemsetAsync(h->d_ticket, 0, sizeof(unsigned int), h->stream));
    if (LK != kCaseControl && nw >= 3 && !h->no_pipeline) {
        // warp-specialised, pipelined variant: one control warp + nw compute warps
        const int warps = nw + 1;
        const size_t smem = slice_smem(h, XS, warps) + (2 * warps) * sizeof(double) + warps * sizeof(int);
        auto kern = k_sweep_slice_ws<LK == kCaseControl ? kDirected : LK, D, XS>;
        CU(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)CT, warps * 32, smem, h->stream>>>(p, h->d_progress, h->d_ticket);
    } else {
        const size_t smem = slice_smem(h, XS, nw);
        auto kern = k_sweep_slice<LK, D, XS>;
        CU(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)CT, nw * 32, smem, h->stream>>>(p, h->d_progress, h->d_ticket);
    }
    CHECK_LAUNCH(h);
    return DLSM_OK;
}

// ---- cluster-per-(chain, slice) variant: few slices, long rows (cfg 3) ---------------------------
size_t cluster_smem(const dlsm_handle *h)
{
    const dlsm_config &c = h->cfg;
    return (size_t)c.n * (c.d + (h->lk == kUndirected ? 0 : 1)) * sizeof(double) +
           (sweep_stage_doubles(c.d) + 2 * (size_t)kMaxTeam * 2) * sizeof(double) + 32;
}

template <int LK, int D>
int cluster_launch_t(dlsm_handle *h, const SweepParams *p, int CS, int ncomp, int *max_active)
{
    const size_t CT = (size_t)h->cfg.n_chains * h->cfg.T;
    const size_t smem = cluster_smem(h);
    auto kern = k_sweep_slice_cl<LK, D>;
    CU(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(CT * CS));
    cfg.blockDim = dim3(32 * (1 + ncomp));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = h->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    if (max_active) { // probe only: how many such clusters can be resident at once
        CU(h, cudaOccupancyMaxActiveClusters(max_active, kern, &cfg));
        return DLSM_OK;
    }
    CU(h, cudaMemsetAsync(h->d_progress, 0, CT * sizeof(int), h->stream));
    CU(h, cudaMemsetAsync(h->d_ticket, 0, sizeof(unsigned int), h->stream));
    CU(h, cudaLaunchKernelEx(&cfg, kern, *p, h->d_progress, h->d_ticket));
    CHECK_LAUNCH(h);
    return DLSM_OK;
}

int cluster_dispatch(dlsm_handle *h, const SweepParams *p, int CS, int ncomp, int *max_active)
{
    const bool d2 = h->cfg.d == 2;
    if (h->lk == kUndirected)
        return d2 ? cluster_launch_t<kUndirected, 2>(h, p, CS, ncomp, max_active)
                  : cluster_launch_t<kUndirected, 0>(h, p, CS, ncomp, max_active);
    return d2 ? cluster_launch_t<kDirected, 2>(h, p, CS, ncomp, max_active)
              : cluster_launch_t<kDirected, 0>(h, p, CS, ncomp, max_active);
}

// CTAs per cluster: the largest size whose C*T clusters can all be resident (they form a wavefront
// over the slices), so that one chain spreads over as many SMs as the part has; 0 = not applicable.
// DLSM_OPT_NO_CLUSTER: 0 block-speculative kernel with a two-block window (k_sweep_blkw), 1 no clusters,
// 2 per-node kernel (k_sweep_slice_cl), 3 block-speculative kernel without the window (k_sweep_blk).
int cluster_size(dlsm_handle *h)
{
    if (h->cluster_cs >= 0) return h->cluster_cs;
    h->cluster_cs = 0;
    const dlsm_config &c = h->cfg;
    const int CT = c.n_chains * c.T;
    const int64_t mode = h->opt[DLSM_OPT_NO_CLUSTER];
    if (h->lk == kCaseControl || h->no_pipeline || mode == 1 || CT * 2 > h->sm_count) return 0;
    h->cluster_blk = mode != 2;
    // the block-speculative kernel with the two-block window (k_sweep_blkw) unless DLSM_OPT_NO_CLUSTER = 3
    // asks for the plain one (k_sweep_blk)
    h->cluster_win = mode == 0 && blkw_smem_bytes(c.n, c.d, h->lk == kDirected, h->W) <= kMaxSmem;
    if (h->cluster_blk ? blk_smem_bytes(c.n, c.d, h->lk == kDirected, h->W) > kMaxSmem : cluster_smem(h) > kMaxSmem)
        return 0;
    const int chunks = (c.n + (h->lk == kUndirected ? 63 : 31)) / (h->lk == kUndirected ? 64 : 32);
    for (int CS = h->sm_count / CT < 8 ? h->sm_count / CT : 8; CS >= 2; CS--) {
        int ncomp, active = 0;
        if (h->cluster_blk) {
            ncomp = 16; // 16 warps per CTA: CS * 16 <= 128 column shares
            SweepParams p = sweep_params(h);
            if ((h->cluster_win ? blkw_launch(p, h->lk == kDirected, CS, nullptr, nullptr, nullptr, h->stream, &active)
                                : blk_launch(p, h->lk == kDirected, CS, ncomp, nullptr, nullptr, h->stream, &active)) != cudaSuccess) {
                cudaGetLastError();
                continue;
            }
        } else {
            ncomp = (chunks + CS - 1) / CS;
            ncomp = ncomp < 1 ? 1 : (ncomp > 15 ? 15 : ncomp);
            if (cluster_dispatch(h, nullptr, CS, ncomp, &active) != DLSM_OK) { cudaGetLastError(); continue; }
        }
        if (active >= CT) {
            h->cluster_cs = CS;
            h->cluster_ncomp = ncomp;
            break;
        }
    }
    return h->cluster_cs;
}

// case-control sweep, batch-parallel: a warp per node of a run of mutually independent nodes
int launch_cc_batch(dlsm_handle *h, const SweepParams &p)
{
    const dlsm_config &c = h->cfg;
    const size_t CT = (size_t)c.n_chains * c.T;
    // DLSM_OPT_CC_KERNEL: 0 auto (the dataflow kernel), 1 k_sweep_cc, 2 k_sweep_cc2, 3 k_sweep_cc3 (runs of
    // mutually independent nodes), 4 k_sweep_ccd (dataflow over nodes, double-buffered positions)
    const int64_t cck = h->opt[DLSM_OPT_CC_KERNEL];
    if (c.d == 2 && h->n_control <= 128 && (cck == 0 || cck == 4)) {
        const size_t cells = CT * c.n;
        if (!h->d_gather) CU(h, cudaMalloc((void **)&h->d_gather, cells * 4 * sizeof(double)));
        int launches = 0;
        CU(h, ccd_launch(p, h->d_gather, &h->ccd, h->sm_count, (int)h->opt[DLSM_OPT_CCD_GROUP], h->stream, &launches));
        h->ctr.kernel_launches += launches - 1;
        return DLSM_OK;
    }
    if (!h->cc_dep_valid) {
        const size_t cells = (size_t)h->ctrl_sets * c.T * c.n;
        cudaFree(h->d_cc_dep);
        h->d_cc_dep = nullptr;
        CU(h, cudaMalloc((void **)&h->d_cc_dep, cells * sizeof(int32_t)));
        k_cc_deps<<<(unsigned)((cells * 32 + 255) / 256), 256, 0, h->stream>>>(p.net, h->ctrl_sets, h->d_cc_dep);
        CHECK_LAUNCH(h);
        h->ctr.kernel_launches += 1;
        h->cc_dep_valid = true;
    }
    CU(h, cudaMemsetAsync(h->d_progress, 0, CT * sizeof(int), h->stream));
    CU(h, cudaMemsetAsync(h->d_ticket, 0, sizeof(unsigned int), h->stream));
    if (c.d == 2 && h->n_control <= 128 && cck == 3) {
        if (h->cc3_ok < 0) {
            int active = 0;
            h->cc3_ok = (cc3_launch(p, nullptr, nullptr, nullptr, h->stream, &active) == cudaSuccess &&
                         active >= (int)CT) ? 1 : 0;
            cudaGetLastError();
        }
        if (h->cc3_ok) {
            CU(h, cc3_launch(p, h->d_progress, h->d_ticket, h->d_cc_dep, h->stream, nullptr));
            return DLSM_OK;
        }
    }
    if (c.d == 2 && h->n_control <= 128 && cck == 2 &&
        cc2_smem_bytes(h->max_in, h->max_out, h->n_control) <= kMaxSmem) {
        // second generation: list indices staged per 32-node block, 256-bit gather records
        const size_t cells = (size_t)c.n_chains * c.T * c.n;
        if (!h->d_gather) CU(h, cudaMalloc((void **)&h->d_gather, cells * 4 * sizeof(double)));
        k_pack_gather<<<(unsigned)((cells + 255) / 256), 256, 0, h->stream>>>((const double *)p.X, (const double *)h->rinv,
                                                                              h->d_gather, c.n_chains, c.T, c.n);
        CHECK_LAUNCH(h);
        h->ctr.kernel_launches += 1;
        CU(h, cc2_launch(p, h->d_gather, h->d_progress, h->d_ticket, h->d_cc_dep, h->stream));
        return DLSM_OK;
    }
    const int nw = 16;
    const size_t smem = sweep_stage_doubles(c.d) * sizeof(double) + 64 * sizeof(int) + 16;
    if (c.d == 2) k_sweep_cc<2><<<(unsigned)CT, nw * 32, smem, h->stream>>>(p, h->d_progress, h->d_ticket, h->d_cc_dep);
    else k_sweep_cc<0><<<(unsigned)CT, nw * 32, smem, h->stream>>>(p, h->d_progress, h->d_ticket, h->d_cc_dep);
    CHECK_LAUNCH(h);
    return DLSM_OK;
}

static __global__ void k_sum_slices(int C, int T, const double *in, double *out)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double s = 0.0;
    for (int t = 0; t < T; t++) s += in[(size_t)c * T + t];
    out[c] = s;
}

bool use_slice_kernel(const dlsm_handle *h);

// the sweep kernel of this handle tracks the post-sweep log-likelihood although it is a slice kernel
static bool cluster_tracks_loglik(dlsm_handle *h)
{
    return use_slice_kernel(h) && h->lk != kCaseControl && cluster_size(h) >= 2 && h->cluster_win;
}

template <int LK>
int launch_slice_lk(dlsm_handle *h, const SweepParams &p)
{
    if (LK == kCaseControl && !h->no_pipeline) return launch_cc_batch(h, p);
    if (LK != kCaseControl && cluster_size(h) >= 2) {
        h->ctr.cluster_sweeps += 1;
        if (!h->cluster_blk) return cluster_dispatch(h, &p, h->cluster_cs, h->cluster_ncomp, nullptr);
        if (h->cluster_win) {
            // the kernel also hands over the full-network log-likelihood of the state it leaves behind
            // (per-slice sums, added in slice order)
            if (p.ll_cur && !h->d_ll_slices)
                CU(h, cudaMalloc((void **)&h->d_ll_slices, (size_t)h->cfg.n_chains * h->cfg.T * sizeof(double)));
            CU(h, blkw_launch(p, LK == kDirected, h->cluster_cs, h->d_progress, h->d_ticket,
                              p.ll_cur ? h->d_ll_slices : nullptr, h->stream, nullptr));
            if (p.ll_cur) {
                const int C = h->cfg.n_chains;
                k_sum_slices<<<(C + 127) / 128, 128, 0, h->stream>>>(C, h->cfg.T, (const double *)h->d_ll_slices, p.ll_cur);
                CHECK_LAUNCH(h);
                h->ctr.kernel_launches += 1;
            }
            return DLSM_OK;
        }
        CU(h, blk_launch(p, LK == kDirected, h->cluster_cs, h->cluster_ncomp, h->d_progress, h->d_ticket,
                         h->stream, nullptr));
        return DLSM_OK;
    }
    const int nw = slice_warps(h);
    const bool xs = slice_smem(h, true, nw + 1) + 4096 <= kMaxSmem / 2;
    if (h->cfg.d == 2) return xs ? launch_slice_t<LK, 2, true>(h, p, nw) : launch_slice_t<LK, 2, false>(h, p, nw);
    return xs ? launch_slice_t<LK, 0, true>(h, p, nw) : launch_slice_t<LK, 0, false>(h, p, nw);
}

// Which mapping serves a sweep (measured, profiles/r2b_ab_variants.json, r2c_*):
//   * few (chain, slice) pairs with long rows (C*T*2 <= SMs, n >= 256; cfg 3): a thread-block cluster per
//     pair, 32 nodes per cluster barrier (k_sweep_blk);
//   * otherwise one CTA per chain, one warp per slice: the block-speculative kernel (k_sweep_cb) up to
//     two chains per SM (its lanes carry independent rows: it needs no co-resident warps to hide
//     latency), the node-by-node kernel beyond that (with the row-sum cache for long rows);
//   * case-control lists: CTA per (chain, slice) kernels when a warp per slice cannot fill the part.
bool use_slice_kernel(const dlsm_handle *h)
{
    if (h->sweep_mode == 1) return false;
    if (h->sweep_mode == 2) return true;
    const size_t warps_chain_mode = (size_t)h->cfg.n_chains * (h->cfg.T < 16 ? h->cfg.T : 16);
    if (h->lk == kCaseControl) return h->cfg.n >= 256 && warps_chain_mode < (size_t)h->sm_count * 16;
    return h->cfg.n >= 256 && h->cfg.n_chains * h->cfg.T * 2 <= h->sm_count && h->opt[DLSM_OPT_NO_CLUSTER] != 1;
}

int chain_kernel(const dlsm_handle *h)
{
    const int64_t k = h->opt[DLSM_OPT_CHAIN_KERNEL];
    if (k != DLSM_CHAIN_AUTO) return (int)k;
    // (DLSM_SWEEP_CHAIN_DENSE = "as if there were many chains": the node-by-node kernels)
    if (h->cfg.n_chains <= 2 * h->sm_count && !h->dense_build) return DLSM_CHAIN_BLOCK;
    return h->cfg.n >= 256 ? DLSM_CHAIN_NODE_ROWSUM : DLSM_CHAIN_NODE;
}

// which kernel serves the one-CTA-per-chain mapping (exact likelihoods): DLSM_OPT_CHAIN_KERNEL
bool chain_blk(const dlsm_handle *h)
{
    return h->lk != kCaseControl && (chain_kernel(h) == DLSM_CHAIN_BLOCK || chain_kernel(h) == DLSM_CHAIN_BLOCK_PAIR);
}

int launch_sweep(dlsm_handle *h, const SweepParams &p)
{
    begin_phase(h, 0);
    int rc;
    if (use_slice_kernel(h)) {
        if (h->lk == kUndirected) rc = launch_slice_lk<kUndirected>(h, p);
        else if (h->lk == kDirected) rc = launch_slice_lk<kDirected>(h, p);
        else rc = launch_slice_lk<kCaseControl>(h, p);
    } else if (chain_blk(h)) {
        // block-speculative chain kernel (k_sweep_cb): 32 nodes per step, lanes = rows
        cudaError_t ce = cb_launch(p, h->lk == kDirected, h->stream, chain_kernel(h) == DLSM_CHAIN_BLOCK_PAIR);
        if (ce != cudaSuccess) { end_phase(h); CU(h, ce); }
        rc = DLSM_OK;
    } else {
        if (h->lk == kUndirected) rc = launch_sweep_lk<kUndirected>(h, p);
        else if (h->lk == kDirected) rc = launch_sweep_lk<kDirected>(h, p);
        else rc = launch_sweep_lk<kCaseControl>(h, p);
    }
    end_phase(h);
    if (rc == DLSM_OK) {
        h->ctr.kernel_launches += 1;
        h->ctr.sweeps += 1;
        h->ctr.node_updates += (uint64_t)h->cfg.n_chains * h->cfg.T * h->cfg.n;
    }
    return rc;
}

template <typename K, typename... A>
int launch_simple(dlsm_handle *h, K kern, dim3 grid, dim3 block, size_t smem, A... args)
{
    kern<<<grid, block, smem, h->stream>>>(args...);
    CHECK_LAUNCH(h);
    h->ctr.kernel_launches += 1;
    return DLSM_OK;
}

// full-network log-likelihood for two variants -> h->d_partial
int launch_full(dlsm_handle *h, const double *rinv0, const double *rinv1, int nv = 2)
{
    FullParams p;
    memset(&p, 0, sizeof(p));
    p.net = net_view(h);
    p.C = h->cfg.n_chains; p.tiles = h->full_tiles;
    p.X = F<double>(h, DLSM_F_X);
    p.bvar = h->d_bvar;
    p.rinv0 = rinv0; p.rinv1 = rinv1;
    p.partial = h->d_partial;
    p.flags = h->d_flags;
    p.same_r = (nv == 2 && rinv0 == rinv1) ? 1 : 0;
    if (h->lk == kCaseControl && h->cfg.d == 2 && (nv == 1 || p.same_r) && !h->opt[DLSM_OPT_NO_GATHER_PACK]) {
        // positions and reciprocal radii of this evaluation as 32-byte records: one 256-bit load
        // per gathered node instead of two loads (the kernel is bound by L1 gather wavefronts)
        const size_t cells = (size_t)h->cfg.n_chains * h->cfg.T * h->cfg.n;
        if (!h->d_gather) CU(h, cudaMalloc((void **)&h->d_gather, cells * 4 * sizeof(double)));
        int rc = launch_simple(h, k_pack_gather, dim3((unsigned)((cells + 255) / 256)), dim3(256), 0,
                               (const double *)p.X, rinv0, h->d_gather, h->cfg.n_chains, h->cfg.T, h->cfg.n);
        if (rc != DLSM_OK) return rc;
        p.gather = h->d_gather;
    }
    dim3 grid(h->cfg.T * h->full_tiles, h->cfg.n_chains);
    const size_t smem = (h->lk == kCaseControl) ? 0 : (size_t)h->cfg.n * h->cfg.d * sizeof(double);
    const bool d2 = h->cfg.d == 2;
    // one variant of an exact likelihood, d = 2, long rows: lanes = rows, broadcast columns (k_full_lr;
    // cfg 3: 0.346 -> 0.30 ms per pass; at n = 120 its 10 work items per slice do not fill the CTAs: +6 %)
    if (h->lk != kCaseControl && d2 && nv == 1 && h->opt[DLSM_OPT_FULL_KERNEL] != 1 &&
        (h->cfg.n >= 256 || h->opt[DLSM_OPT_FULL_KERNEL] == 2) &&
        fullr_smem_bytes(h->cfg.n, h->lk == kDirected) <= kMaxSmem) {
        CU(h, fullr_launch(p, h->lk == kDirected, grid, h->stream));
        h->ctr.kernel_launches += 1;
        return DLSM_OK;
    }
#define LAUNCH_FULL_NV(LK, D, NV)                                                                \
    do {                                                                                         \
        CU(h, cudaFuncSetAttribute(k_full<LK, D, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        k_full<LK, D, NV><<<grid, 256, smem, h->stream>>>(p);                                    \
    } while (0)
#define LAUNCH_FULL(LK)                                                                          \
    do {                                                                                         \
        if (d2 && nv == 1) LAUNCH_FULL_NV(LK, 2, 1);                                             \
        else if (d2) LAUNCH_FULL_NV(LK, 2, 2);                                                   \
        else if (nv == 1) LAUNCH_FULL_NV(LK, 0, 1);                                              \
        else LAUNCH_FULL_NV(LK, 0, 2);                                                           \
    } while (0)
    if (smem > kMaxSmem) FAIL(h, DLSM_ERR_UNSUPPORTED, "n too large for the exact full-network kernel");
    if (h->lk == kUndirected) LAUNCH_FULL(kUndirected);
    else if (h->lk == kDirected) LAUNCH_FULL(kDirected);
    else LAUNCH_FULL(kCaseControl);
#undef LAUNCH_FULL_NV
#undef LAUNCH_FULL
    CHECK_LAUNCH(h);
    h->ctr.kernel_launches += 1;
    return DLSM_OK;
}

// ---- row-sum cache ---------------------------------------------------------------------------
// used by the device loop (dlsm_run_sweeps / dlsm_run_traced) with the exact likelihoods and the
// chain kernel; the function-level entry points keep the two-variant evaluation
size_t rows_smem(const dlsm_handle *h)
{
    const dlsm_config &c = h->cfg;
    const int ns = 7 / (((c.n + 31) / 32 + 1) / 2 * (((c.n + 31) / 32 + 1 + 16) / 17)) + 2;
    return (size_t)(ns < c.T ? ns : c.T) * (((size_t)c.n * (c.d + (h->lk == kDirected ? 1 : 0)) + 1) & ~(size_t)1) * sizeof(double);
}

bool rows_enabled(const dlsm_handle *h)
{
    return h->lk != kCaseControl && !use_slice_kernel(h) && !h->opt[DLSM_OPT_NO_ROWSUM_CACHE] &&
           chain_kernel(h) == DLSM_CHAIN_NODE_ROWSUM &&
           !h->opt[DLSM_OPT_NO_TRACKED_LOGLIK] && rows_smem(h) <= kMaxSmem;
}

int ensure_rows(dlsm_handle *h)
{
    if (h->d_rows) return DLSM_OK;
    const dlsm_config &c = h->cfg;
    const size_t cells = (size_t)c.n_chains * c.T * c.n, slices = (size_t)c.n_chains * c.T;
    const int nb = (c.n + 31) / 32;
    h->rows_nb = nb;
    h->rows_half = (nb + 1) / 2;
    h->rows_R = (nb + 1 + 16) / 17;                    // runs of <= ~17 tiles
    h->rows_L = (nb + 1 + h->rows_R - 1) / h->rows_R;
    h->rows_ipc = h->rows_half * h->rows_R;
    h->rows_grid = (c.T * h->rows_ipc + 7) / 8;
    h->rows_ns = 7 / h->rows_ipc + 2;
    if (h->rows_ns > c.T) h->rows_ns = c.T;
    CU(h, cudaMalloc((void **)&h->d_rows, cells * 8));
    CU(h, cudaMalloc((void **)&h->d_scr, slices * sweep_rows_padded(true, c.n) * 8));
    CU(h, cudaMalloc((void **)&h->d_rows_own, slices * h->rows_half * h->rows_R * 2 * 32 * 8));
    CU(h, cudaMalloc((void **)&h->d_rows_part, (slices * ((size_t)nb * (nb - 1) / 2) * 32 + 32) * 8));
    CU(h, cudaMalloc((void **)&h->d_accflag, (size_t)c.n_chains * 4));
    const size_t need = (size_t)c.n_chains * h->rows_grid * 2 * 8;
    if (need > (size_t)c.n_chains * h->full_nblk * 2 * 8) { // k_rows has its own partial-sum layout
        CU(h, cudaStreamSynchronize(h->stream));
        cudaFree(h->d_partial);
        h->d_partial = nullptr;
        CU(h, cudaMalloc((void **)&h->d_partial, need));
    }
    h->rows_valid = false;
    return DLSM_OK;
}

// full-network log-likelihood of variant 0 (bvar, rinv0) -> d_partial, its row sums -> (own, part)
int launch_rows(dlsm_handle *h, const double *rinv0)
{
    const dlsm_config &c = h->cfg;
    RowsParams p;
    memset(&p, 0, sizeof(p));
    p.net = net_view(h);
    p.C = c.n_chains; p.nb = h->rows_nb; p.half = h->rows_half; p.R = h->rows_R; p.L = h->rows_L;
    p.ipc = h->rows_ipc; p.ns = h->rows_ns;
    p.X = F<double>(h, DLSM_F_X); p.bvar = h->d_bvar; p.rinv0 = rinv0;
    p.partial = h->d_partial; p.own = h->d_rows_own; p.part = h->d_rows_part;
    const dim3 grid(h->rows_grid, c.n_chains);
    const size_t smem = rows_smem(h);
#define LAUNCH_ROWS(LK, D)                                                                                \
    do {                                                                                                  \
        CU(h, cudaFuncSetAttribute(k_rows<LK, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        k_rows<LK, D><<<grid, 256, smem, h->stream>>>(p);                                                 \
    } while (0)
    if (h->lk == kUndirected) { if (c.d == 2) LAUNCH_ROWS(kUndirected, 2); else LAUNCH_ROWS(kUndirected, 0); }
    else { if (c.d == 2) LAUNCH_ROWS(kDirected, 2); else LAUNCH_ROWS(kDirected, 0); }
#undef LAUNCH_ROWS
    CHECK_LAUNCH(h);
    h->ctr.kernel_launches += 1;
    return DLSM_OK;
}

int rows_nblk(const dlsm_handle *h) { return h->rows_grid; }

int commit_rows(dlsm_handle *h, const int32_t *flag)
{
    const dlsm_config &c = h->cfg;
    const int warps = c.T * h->rows_nb;
    return launch_simple(h, k_rows_commit, dim3((warps + 7) / 8, c.n_chains), dim3(256), 0, flag, c.T, c.n,
                         h->rows_nb, h->rows_half, h->rows_R, (const double *)h->d_rows_own,
                         (const double *)h->d_rows_part, h->d_rows);
}

// rows <- the current state's row sums, DLSM_F_LOGLIK <- its log-likelihood
int refresh_rows(dlsm_handle *h)
{
    const int C = h->cfg.n_chains;
    int rc = launch_simple(h, k_bvar_current, dim3((C + 127) / 128), dim3(128), 0, C,
                           (const double *)F<double>(h, DLSM_F_INTERCEPT), h->d_bvar);
    if (rc == DLSM_OK) rc = launch_rows(h, h->rinv);
    if (rc == DLSM_OK)
        rc = launch_simple(h, k_sum_partials, dim3((C + 3) / 4), dim3(128), 0, C, rows_nblk(h),
                           (const double *)h->d_partial, h->d_ll2, F<double>(h, DLSM_F_LOGLIK));
    if (rc == DLSM_OK) rc = commit_rows(h, nullptr);
    if (rc == DLSM_OK) h->rows_valid = true;
    return rc;
}

int check_flags(dlsm_handle *h)
{
    unsigned int f = 0;
    CU(h, cudaMemcpyAsync(&f, h->d_flags, sizeof(f), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    if (f & 2u) h->ctr.ub_flags += 1;
    if (f) CU(h, cudaMemsetAsync(h->d_flags, 0, sizeof(unsigned int), h->stream));
    if (f & 1u) FAIL(h, DLSM_ERR_NONFINITE, "a log-likelihood ratio evaluated to NaN/inf");
    return DLSM_OK;
}

int upload(dlsm_handle *h, void *dst, const void *src, size_t bytes)
{
    CU(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream));
    return DLSM_OK;
}

int download(dlsm_handle *h, void *dst, const void *src, size_t bytes)
{
    CU(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return DLSM_OK;
}

int update_rinv(dlsm_handle *h)
{
    const size_t total = (size_t)h->cfg.n_chains * h->cfg.n;
    return launch_simple(h, k_rinv, dim3((unsigned)((total + 255) / 256)), dim3(256), 0,
                         (const double *)F<double>(h, DLSM_F_RADII), h->rinv, total);
}

void free_trace(dlsm_handle *h)
{
    for (auto &ch : h->chunk) {
        for (int f = 0; f < DLSM_F_COUNT_; f++) { cudaFree(ch.dev[f]); ch.dev[f] = nullptr; }
        cudaFree(ch.logp); ch.logp = nullptr;
        if (ch.filled) cudaEventDestroy(ch.filled);
        if (ch.drained) cudaEventDestroy(ch.drained);
        ch.filled = ch.drained = nullptr;
        ch.used = false;
    }
    h->trace_R = 0;
}

} // namespace

extern "C" {

int dlsm_abi_version(void) { return DLSM_ABI_VERSION; }

int dlsm_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char *dlsm_last_error(const dlsm_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int dlsm_create(const dlsm_config *cfg, dlsm_handle **out)
{
    if (!cfg || !out) { g_create_error = "null argument"; return DLSM_ERR_INVALID; }
    *out = nullptr;
    if (cfg->n_chains < 1 || cfg->T < 1 || cfg->n < 2 || cfg->d < 1 || cfg->d > kMaxD || cfg->K < 0) {
        g_create_error = "invalid shape (need n_chains>=1, T>=1, n>=2, 1<=d<=8, K>=0)";
        return DLSM_ERR_INVALID;
    }
    if (cfg->likelihood == DLSM_LIK_CASE_CONTROL && !cfg->is_directed) {
        // lsm.py:425-427
        g_create_error = "The case-control likelihood currently only supported for directed networks.";
        return DLSM_ERR_INVALID;
    }
    if (cfg->prior == DLSM_PRIOR_MIXTURE && cfg->K < 1) {
        g_create_error = "mixture prior needs K >= 1";
        return DLSM_ERR_INVALID;
    }
    if (dlsm_device_count() <= cfg->device) {
        g_create_error = "no CUDA device available (libdlsm has no CPU fallback)";
        return DLSM_ERR_CUDA;
    }
    dlsm_handle *h = new dlsm_handle();
    h->cfg = *cfg;
    memset(&h->ctr, 0, sizeof(h->ctr));
    h->hy.tau_sq = 2.0; h->hy.sigma_sq = 0.1;
    h->hy.intercept_prior[0] = h->hy.intercept_prior[1] = 0.0;
    h->hy.intercept_variance_prior = 2.0;
    h->lk = cfg->likelihood == DLSM_LIK_CASE_CONTROL ? kCaseControl
                                                      : (cfg->is_directed ? kDirected : kUndirected);
    h->W = ((cfg->n + 31) / 32 + 3) / 4 * 4;
    read_env_options(h);
    auto fail = [&](const char *what, cudaError_t e) {
        g_create_error = std::string(what) + ": " + cudaGetErrorString(e);
        dlsm_destroy(h);
        return DLSM_ERR_CUDA;
    };
    cudaError_t e;
    if ((e = cudaSetDevice(cfg->device)) != cudaSuccess) return fail("cudaSetDevice", e);
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, cfg->device);
    {
        int pmax = 0, wmax = 0;
        cudaDeviceGetAttribute(&pmax, cudaDevAttrMaxPersistingL2CacheSize, cfg->device);
        cudaDeviceGetAttribute(&wmax, cudaDevAttrMaxAccessPolicyWindowSize, cfg->device);
        h->l2_persist_max = (size_t)(pmax < wmax ? pmax : wmax);
    }
    if ((e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking)) != cudaSuccess)
        return fail("cudaStreamCreate", e);
    h->stream = h->own_stream;
    {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if ((e = cudaStreamCreateWithPriority(&h->side_stream, cudaStreamNonBlocking, hi)) != cudaSuccess)
            return fail("cudaStreamCreate(side)", e);
        cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming);
    }
    for (int f = 0; f < DLSM_F_COUNT_; f++) {
        const size_t bytes = field_elems(*cfg, f) * elem_size(f);
        h->field_bytes[f] = bytes;
        if (bytes == 0) continue;
        if ((e = cudaMalloc(&h->field[f], bytes)) != cudaSuccess) return fail("cudaMalloc(state)", e);
        cudaMemsetAsync(h->field[f], 0, bytes, h->stream);
    }
    const size_t C = cfg->n_chains, n = cfg->n, T = cfg->T;
    h->full_tiles = (int)((n + 63) / 64);
    h->full_nblk = cfg->T * h->full_tiles;
#define ALLOC(ptr, bytes)                                                              \
    if ((e = cudaMalloc((void **)&(ptr), (bytes))) != cudaSuccess) return fail("cudaMalloc", e)
    ALLOC(h->rinv, C * n * 8);
    ALLOC(h->d_partial, C * h->full_nblk * 2 * 8);
    ALLOC(h->d_bvar, C * 4 * 8);
    ALLOC(h->d_prop, C * 8);
    ALLOC(h->d_ll2, C * 2 * 8);
    ALLOC(h->d_small, C * 4 * 8);
    ALLOC(h->d_small_i, C * 4 * 4);
    ALLOC(h->d_flags, 4);
    ALLOC(h->d_progress, C * T * sizeof(int));
    ALLOC(h->d_ticket, 4);
    ALLOC(h->d_bad, 4);
    if (cfg->is_directed) {
        ALLOC(h->d_rprop, C * n * 8);
        ALLOC(h->d_rprop_inv, C * n * 8);
    }
#undef ALLOC
    cudaMemsetAsync(h->d_flags, 0, 4, h->stream);
    cudaMemsetAsync(h->rinv, 0, C * n * 8, h->stream);
    if ((e = cudaStreamSynchronize(h->stream)) != cudaSuccess) return fail("init", e);
    *out = h;
    return DLSM_OK;
}

void dlsm_destroy(dlsm_handle *h)
{
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->l2_persist_set) { // give the persisting-L2 set-aside of the label kernel's stage back to the device
        cudaDeviceSynchronize();
        cudaCtxResetPersistingL2Cache();
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);
    }
    for (auto &e : h->events) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    for (int f = 0; f < DLSM_F_COUNT_; f++) cudaFree(h->field[f]);
    void *ptrs[] = {h->rowbits, h->colbits, h->deg, h->in_edges, h->out_edges, h->ctrl_in,
                    h->ctrl_out, h->rinv, h->d_eps, h->d_logu, h->d_ratio, h->d_out, h->d_acc,
                    h->d_partial, h->d_bvar, h->d_prop, h->d_ll2, h->d_rprop, h->d_rprop_inv,
                    h->d_small, h->d_small_i, h->d_flags, h->d_bad, h->d_progress, h->d_ticket, h->d_ffbs_stage,
                    h->d_cc_dep, h->d_rows, h->d_scr, h->d_rows_own, h->d_rows_part, h->d_accflag};
    for (void *p : ptrs) cudaFree(p);
    if (h->x_stream) cudaStreamDestroy(h->x_stream);
    if (h->ev_x_ready) cudaEventDestroy(h->ev_x_ready);
    if (h->ev_x_copied) cudaEventDestroy(h->ev_x_copied);
    free_trace(h);
    cudaFree(h->d_logp);
    cudaFree(h->d_center);
    cudaFree(h->d_cooc);
    cudaFree(h->d_gather);
    ccd_free(h->ccd);
    cudaFree(h->d_ll_slices);
    cudaFree(h->d_radii_terms);
    cudaFree(h->d_proc_ref);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->side_stream) cudaStreamDestroy(h->side_stream);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
}

int dlsm_set_option(dlsm_handle *h, int option, int64_t value)
{
    if (!h) return DLSM_ERR_INVALID;
    if (option < 0 || option >= DLSM_OPT_COUNT_) FAIL(h, DLSM_ERR_INVALID, "unknown option %d", option);
    if (option == DLSM_OPT_SWEEP_MODE && (value < DLSM_SWEEP_AUTO || value > DLSM_SWEEP_SLICE_PLAIN))
        FAIL(h, DLSM_ERR_INVALID, "DLSM_OPT_SWEEP_MODE takes a dlsm_sweep_mode value");
    if (option == DLSM_OPT_FFBS_KERNEL && (value < DLSM_FFBS_AUTO || value > DLSM_FFBS_WARP))
        FAIL(h, DLSM_ERR_INVALID, "DLSM_OPT_FFBS_KERNEL takes a dlsm_ffbs_kernel value");
    if (value < 0) FAIL(h, DLSM_ERR_INVALID, "option values are non-negative");
    if (option == DLSM_OPT_NO_CLUSTER && value > 3) FAIL(h, DLSM_ERR_INVALID, "DLSM_OPT_NO_CLUSTER takes 0..3");
    if (option == DLSM_OPT_CC_KERNEL && value > 4) FAIL(h, DLSM_ERR_INVALID, "DLSM_OPT_CC_KERNEL takes 0..4");
    if (option == DLSM_OPT_CHAIN_KERNEL && value > DLSM_CHAIN_BLOCK_PAIR) FAIL(h, DLSM_ERR_INVALID, "DLSM_OPT_CHAIN_KERNEL takes a dlsm_chain_kernel value");
    h->opt[option] = value;
    h->rows_valid = false, h->sweeps_since_set = 0;
    h->cluster_cs = -1;
    h->cc3_ok = -1;
    if (option == DLSM_OPT_SWEEP_MODE) apply_sweep_mode(h);
    return DLSM_OK;
}

int dlsm_set_stream(dlsm_handle *h, void *s)
{
    if (!h) return DLSM_ERR_INVALID;
    CU(h, cudaStreamSynchronize(h->stream));
    h->stream = s ? static_cast<cudaStream_t>(s) : h->own_stream;
    return DLSM_OK;
}

int dlsm_synchronize(dlsm_handle *h)
{
    if (!h) return DLSM_ERR_INVALID;
    CU(h, cudaSetDevice(h->cfg.device));
    CU(h, cudaStreamSynchronize(h->stream));
    return DLSM_OK;
}

int dlsm_set_network_dense(dlsm_handle *h, const double *Y)
{
    if (!h || !Y) return DLSM_ERR_INVALID;
    CU(h, cudaSetDevice(h->cfg.device));
    const size_t T = h->cfg.T, n = h->cfg.n, W = h->W;
    if (!h->rowbits) CU(h, cudaMalloc((void **)&h->rowbits, T * n * W * 4));
    if (h->cfg.is_directed && !h->colbits) CU(h, cudaMalloc((void **)&h->colbits, T * n * W * 4));
    double *slice = nullptr;
    CU(h, cudaMalloc((void **)&slice, n * n * 8));
    CU(h, cudaMemsetAsync(h->d_bad, 0, 4, h->stream));
    for (size_t t = 0; t < T; t++) {
        CU(h, cudaMemcpyAsync(slice, Y + t * n * n, n * n * 8, cudaMemcpyHostToDevice, h->stream));
        const size_t warps = n * W;
        k_pack_rows<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, h->stream>>>(
            slice, (int)n, (int)W, h->rowbits + t * n * W, h->d_bad);
        if (h->cfg.is_directed)
            k_pack_cols<<<(unsigned)((n * W + 255) / 256), 256, 0, h->stream>>>(
                slice, (int)n, (int)W, h->colbits + t * n * W);
        h->ctr.kernel_launches += h->cfg.is_directed ? 2 : 1;
    }
    int bad = 0;
    cudaError_t e = cudaMemcpyAsync(&bad, h->d_bad, 4, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(slice);
    CU(h, e);
    CHECK_LAUNCH(h);
    if (bad) FAIL(h, DLSM_ERR_NONBINARY, "adjacency must be 0/1 (weighted or missing (-1) dyads are not supported on the device path)");
    h->have_net = true;
    h->rows_valid = false, h->sweeps_since_set = 0;
    return DLSM_OK;
}

int dlsm_set_edge_lists(dlsm_handle *h, const int32_t *degrees, const int32_t *in_edges,
                        int32_t max_in, const int32_t *out_edges, int32_t max_out)
{
    if (!h || !degrees || max_in < 0 || max_out < 0) return DLSM_ERR_INVALID;
    if ((max_in > 0 && !in_edges) || (max_out > 0 && !out_edges)) return DLSM_ERR_INVALID;
    CU(h, cudaSetDevice(h->cfg.device));
    const size_t TN = (size_t)h->cfg.T * h->cfg.n;
    cudaFree(h->deg); cudaFree(h->in_edges); cudaFree(h->out_edges);
    h->deg = h->in_edges = h->out_edges = nullptr;
    CU(h, cudaMalloc((void **)&h->deg, TN * 2 * 4));
    CU(h, cudaMalloc((void **)&h->in_edges, (TN * (size_t)max_in + 1) * 4));
    CU(h, cudaMalloc((void **)&h->out_edges, (TN * (size_t)max_out + 1) * 4));
    int rc = upload(h, h->deg, degrees, TN * 2 * 4);
    if (rc == DLSM_OK && max_in) rc = upload(h, h->in_edges, in_edges, TN * max_in * 4);
    if (rc == DLSM_OK && max_out) rc = upload(h, h->out_edges, out_edges, TN * max_out * 4);
    if (rc != DLSM_OK) return rc;
    CU(h, cudaStreamSynchronize(h->stream));
    h->max_in = max_in; h->max_out = max_out;
    h->have_edges = true;
    h->cc_dep_valid = false;
    return DLSM_OK;
}

int dlsm_set_network_edges(dlsm_handle *h, const int32_t *edges, size_t n_edges)
{
    if (!h || (n_edges > 0 && !edges)) return DLSM_ERR_INVALID;
    if (h->lk != kCaseControl) FAIL(h, DLSM_ERR_INVALID, "an edge-list network needs the case-control likelihood");
    CU(h, cudaSetDevice(h->cfg.device));
    int32_t *d_edges = nullptr;
    CU(h, cudaMalloc((void **)&d_edges, (n_edges * 3 + 1) * sizeof(int32_t)));
    cudaError_t e = cudaMemcpyAsync(d_edges, edges, n_edges * 3 * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream);
    int32_t *deg = nullptr, *in_e = nullptr, *out_e = nullptr;
    int max_in = 0, max_out = 0, status = 0;
    if (e == cudaSuccess)
        e = graph_build_edge_lists(d_edges, n_edges, h->cfg.T, h->cfg.n, &deg, &in_e, &max_in, &out_e, &max_out,
                                   &status, h->stream);
    cudaFree(d_edges);
    CU(h, e);
    if (status == 1) FAIL(h, DLSM_ERR_INVALID, "edge list: an index is out of range or a tie is a self loop");
    if (status == 2) FAIL(h, DLSM_ERR_INVALID, "edge list: a tie is listed twice");
    cudaFree(h->deg); cudaFree(h->in_edges); cudaFree(h->out_edges);
    h->deg = deg; h->in_edges = in_e; h->out_edges = out_e;
    h->max_in = max_in; h->max_out = max_out;
    h->have_edges = true;
    h->cc_dep_valid = false;
    h->ctr.kernel_launches += 4;
    return DLSM_OK;
}

int dlsm_edge_list_dims(dlsm_handle *h, int32_t *max_in, int32_t *max_out)
{
    if (!h || !max_in || !max_out) return DLSM_ERR_INVALID;
    if (!h->have_edges) FAIL(h, DLSM_ERR_NOTSET, "no edge lists on the device");
    *max_in = h->max_in; *max_out = h->max_out;
    return DLSM_OK;
}

int dlsm_get_edge_lists(dlsm_handle *h, int32_t *degrees, int32_t *in_edges, int32_t *out_edges)
{
    if (!h || !degrees) return DLSM_ERR_INVALID;
    if (!h->have_edges) FAIL(h, DLSM_ERR_NOTSET, "no edge lists on the device");
    CU(h, cudaSetDevice(h->cfg.device));
    const size_t TN = (size_t)h->cfg.T * h->cfg.n;
    int rc = download(h, degrees, h->deg, TN * 2 * 4);
    if (rc == DLSM_OK && in_edges && h->max_in) rc = download(h, in_edges, h->in_edges, TN * h->max_in * 4);
    if (rc == DLSM_OK && out_edges && h->max_out) rc = download(h, out_edges, h->out_edges, TN * h->max_out * 4);
    return rc;
}

int dlsm_set_controls(dlsm_handle *h, const int32_t *ctrl_in, const int32_t *ctrl_out,
                      int32_t n_control, int32_t n_sets)
{
    if (!h || !ctrl_in || !ctrl_out || n_control < 1) return DLSM_ERR_INVALID;
    if (n_sets != 1 && n_sets != h->cfg.n_chains)
        FAIL(h, DLSM_ERR_INVALID, "n_sets must be 1 or n_chains");
    CU(h, cudaSetDevice(h->cfg.device));
    const size_t bytes = (size_t)n_sets * h->cfg.T * h->cfg.n * n_control * 4;
    if (n_control != h->n_control || n_sets != h->ctrl_sets) {
        cudaFree(h->ctrl_in); cudaFree(h->ctrl_out);
        h->ctrl_in = h->ctrl_out = nullptr;
        CU(h, cudaMalloc((void **)&h->ctrl_in, bytes));
        CU(h, cudaMalloc((void **)&h->ctrl_out, bytes));
    }
    int rc = upload(h, h->ctrl_in, ctrl_in, bytes);
    if (rc == DLSM_OK) rc = upload(h, h->ctrl_out, ctrl_out, bytes);
    if (rc != DLSM_OK) return rc;
    CU(h, cudaStreamSynchronize(h->stream));
    h->n_control = n_control; h->ctrl_sets = n_sets;
    h->have_ctrl = true;
    h->cc_dep_valid = false;
    return DLSM_OK;
}

int dlsm_resample_controls(dlsm_handle *h, int32_t n_control, int32_t n_sets)
{
    if (!h || n_control < 1) return DLSM_ERR_INVALID;
    if (!h->have_edges) FAIL(h, DLSM_ERR_NOTSET, "dlsm_set_edge_lists has not been called");
    if (n_sets != 1 && n_sets != h->cfg.n_chains) FAIL(h, DLSM_ERR_INVALID, "n_sets must be 1 or n_chains");
    CU(h, cudaSetDevice(h->cfg.device));
    const dlsm_config &c = h->cfg;
    const size_t bytes = (size_t)n_sets * c.T * c.n * n_control * 4;
    if (n_control != h->n_control || n_sets != h->ctrl_sets || !h->ctrl_in) {
        CU(h, cudaStreamSynchronize(h->stream));
        cudaFree(h->ctrl_in); cudaFree(h->ctrl_out);
        h->ctrl_in = h->ctrl_out = nullptr;
        CU(h, cudaMalloc((void **)&h->ctrl_in, bytes));
        CU(h, cudaMalloc((void **)&h->ctrl_out, bytes));
        h->n_control = n_control; h->ctrl_sets = n_sets;
    }
    ControlParams p;
    memset(&p, 0, sizeof(p));
    p.sets = n_sets; p.T = c.T; p.n = c.n; p.n_control = n_control;
    p.max_in = h->max_in; p.max_out = h->max_out;
    p.deg = h->deg; p.in_edges = h->in_edges; p.out_edges = h->out_edges;
    p.ctrl_in = h->ctrl_in; p.ctrl_out = h->ctrl_out;
    p.seed = h->seed; p.sweep = h->control_draws; p.chain_offset = (uint32_t)h->chain_offset;
    const size_t warps = (size_t)n_sets * c.T * c.n * 2;
    int rc = launch_simple(h, k_resample_controls, dim3((unsigned)((warps + 3) / 4)), dim3(128), 0, p);
    if (rc != DLSM_OK) return rc;
    h->control_draws += 1;
    h->have_ctrl = true;
    h->cc_dep_valid = false;
    return DLSM_OK;
}

int dlsm_get_controls(dlsm_handle *h, int32_t *ctrl_in, int32_t *ctrl_out)
{
    if (!h || !ctrl_in || !ctrl_out) return DLSM_ERR_INVALID;
    if (!h->have_ctrl) FAIL(h, DLSM_ERR_NOTSET, "no control sets on the device");
    CU(h, cudaSetDevice(h->cfg.device));
    const size_t bytes = (size_t)h->ctrl_sets * h->cfg.T * h->cfg.n * h->n_control * 4;
    int rc = download(h, ctrl_in, h->ctrl_in, bytes);
    if (rc == DLSM_OK) rc = download(h, ctrl_out, h->ctrl_out, bytes);
    return rc;
}

int dlsm_set_state(dlsm_handle *h, int field, const void *host, size_t bytes)
{
    if (!h || !host || field < 0 || field >= DLSM_F_COUNT_) return DLSM_ERR_INVALID;
    if (field == DLSM_F_NCOUNT || field == DLSM_F_NK || field == DLSM_F_LOGLIK) FAIL(h, DLSM_ERR_INVALID, "read-only field");
    if (bytes != h->field_bytes[field] || bytes == 0)
        FAIL(h, DLSM_ERR_INVALID, "field %d expects %zu bytes, got %zu", field, h->field_bytes[field], bytes);
    CU(h, cudaSetDevice(h->cfg.device));
    int rc = upload(h, h->field[field], host, bytes);
    if (rc != DLSM_OK) return rc;
    if (field == DLSM_F_X || field == DLSM_F_INTERCEPT || field == DLSM_F_RADII) h->rows_valid = false, h->sweeps_since_set = 0;
    if (field == DLSM_F_RADII) rc = update_rinv(h);
    if (rc != DLSM_OK) return rc;
    CU(h, cudaStreamSynchronize(h->stream));
    return DLSM_OK;
}

int dlsm_get_state(dlsm_handle *h, int field, void *host, size_t bytes)
{
    if (!h || !host || field < 0 || field >= DLSM_F_COUNT_) return DLSM_ERR_INVALID;
    if (bytes != h->field_bytes[field] || bytes == 0)
        FAIL(h, DLSM_ERR_INVALID, "field %d expects %zu bytes, got %zu", field, h->field_bytes[field], bytes);
    CU(h, cudaSetDevice(h->cfg.device));
    return download(h, host, h->field[field], bytes);
}

int dlsm_set_hyper(dlsm_handle *h, const dlsm_hyper *hy)
{
    if (!h || !hy) return DLSM_ERR_INVALID;
    h->hy = *hy;
    return DLSM_OK;
}

int dlsm_set_rng(dlsm_handle *h, uint64_t seed, uint64_t chain_offset, uint64_t sweep_index)
{
    if (!h) return DLSM_ERR_INVALID;
    h->seed = seed;
    h->chain_offset = chain_offset;
    for (int k = 0; k < 5; k++) h->sweep_idx[k] = (uint32_t)sweep_index;
    return DLSM_OK;
}

static int ensure_replay_buffers(dlsm_handle *h)
{
    const size_t N = (size_t)h->cfg.n_chains * h->cfg.T * h->cfg.n;
    if (!h->d_eps) CU(h, cudaMalloc((void **)&h->d_eps, N * h->cfg.d * 8));
    if (!h->d_logu) CU(h, cudaMalloc((void **)&h->d_logu, N * 8));
    if (!h->d_ratio) CU(h, cudaMalloc((void **)&h->d_ratio, N * 8));
    if (!h->d_acc) CU(h, cudaMalloc((void **)&h->d_acc, N * 4));
    return DLSM_OK;
}

int dlsm_sweep_latent(dlsm_handle *h, const double *eps, const double *logu, int32_t *accepted,
                      double *ratio)
{
    if (!h) return DLSM_ERR_INVALID;
    if ((eps == nullptr) != (logu == nullptr)) FAIL(h, DLSM_ERR_INVALID, "eps and logu must both be given (replay) or both NULL (native)");
    CU(h, cudaSetDevice(h->cfg.device));
    int rc = need_inputs(h);
    if (rc != DLSM_OK) return rc;
    const size_t N = (size_t)h->cfg.n_chains * h->cfg.T * h->cfg.n;
    SweepParams p = sweep_params(h);
    h->rows_valid = false, h->sweeps_since_set = 0; // the function-level sweep evaluates both variants and keeps no row sums
    if (eps || accepted || ratio) {
        rc = ensure_replay_buffers(h);
        if (rc != DLSM_OK) return rc;
    }
    if (eps) {
        if ((rc = upload(h, h->d_eps, eps, N * h->cfg.d * 8)) != DLSM_OK) return rc;
        if ((rc = upload(h, h->d_logu, logu, N * 8)) != DLSM_OK) return rc;
        p.eps = h->d_eps; p.logu = h->d_logu;
    }
    if (accepted) p.accepted = h->d_acc;
    if (ratio) p.ratio = h->d_ratio;
    if ((rc = launch_sweep(h, p)) != DLSM_OK) return rc;
    if (!eps) h->sweep_idx[kRngLatent] += 1;
    if (accepted) CU(h, cudaMemcpyAsync(accepted, h->d_acc, N * 4, cudaMemcpyDeviceToHost, h->stream));
    if (ratio) CU(h, cudaMemcpyAsync(ratio, h->d_ratio, N * 8, cudaMemcpyDeviceToHost, h->stream));
    return check_flags(h);
}

// exact: numpy's summation order (dlsm_center, replay); otherwise a deterministic tree (device loop)
static int center_async(dlsm_handle *h, bool exact = true)
{
    const dlsm_config &c = h->cfg;
    const size_t rows = (size_t)c.T * c.n;
    double *X = F<double>(h, DLSM_F_X);
    begin_phase(h, 1);
    int rc;
    if (rows <= 16384) {
        rc = launch_simple(h, k_center, dim3(c.n_chains), dim3(256), 0, X, c.T, c.n, c.d);
    } else {
        const int B = (int)((rows + 8191) / 8192 < 128 ? (rows + 8191) / 8192 : 128);
        if (!h->d_center) CU(h, cudaMalloc((void **)&h->d_center, (size_t)c.n_chains * (128 + 1) * kMaxD * sizeof(double)));
        double *means = h->d_center, *partial = h->d_center + (size_t)c.n_chains * kMaxD;
        if (exact) {
            rc = launch_simple(h, k_center_mean_exact, dim3(c.n_chains), dim3(256), 0, (const double *)X,
                               c.T, c.n, c.d, means);
        } else {
            rc = launch_simple(h, k_center_partial, dim3(B, c.n_chains), dim3(256), 0, (const double *)X, c.T,
                               c.n, c.d, partial);
            if (rc == DLSM_OK)
                rc = launch_simple(h, k_center_total, dim3(c.n_chains), dim3(32), 0, (const double *)partial, B,
                                   c.d, (double)rows, means);
        }
        if (rc == DLSM_OK)
            rc = launch_simple(h, k_center_apply, dim3(B, c.n_chains), dim3(256), 0, X, c.T, c.n, c.d,
                               (const double *)means);
    }
    end_phase(h);
    return rc;
}

int dlsm_center(dlsm_handle *h)
{
    if (!h) return DLSM_ERR_INVALID;
    CU(h, cudaSetDevice(h->cfg.device));
    int rc = center_async(h);
    if (rc != DLSM_OK) return rc;
    CU(h, cudaStreamSynchronize(h->stream));
    return DLSM_OK;
}

// device-side part of sample_intercepts; d_eps/d_logu are device pointers or null (native)
// use_cur: the tracked log-likelihood (DLSM_F_LOGLIK) is current, evaluate proposals only.
// seed_cur: it is NOT current yet -- the first step evaluates proposal and current state in one pass
// (two variants per gathered record) and leaves the kept one in DLSM_F_LOGLIK for the steps after it.
static int intercepts_async(dlsm_handle *h, const double *d_eps, const double *d_logu,
                            int32_t *d_acc, double *d_ratio, bool use_cur = false, bool rows = false,
                            bool seed_cur = false)
{
    const int C = h->cfg.n_chains, m = h->cfg.is_directed ? 2 : 1;
    const dim3 g1((C + 127) / 128), b1(128);
    begin_phase(h, 1);
    for (int i = 0; i < m; i++) {
        ScalarMH p;
        memset(&p, 0, sizeof(p));
        p.C = C; p.which = i; p.nblk = rows ? rows_nblk(h) : h->full_nblk;
        p.accflag = rows ? h->d_accflag : nullptr;
        p.tune = h->cfg.tune; p.tune_interval = h->cfg.intercept_tune_interval[i];
        p.intercept = F<double>(h, DLSM_F_INTERCEPT);
        p.bvar = h->d_bvar; p.prop = h->d_prop; p.partial = h->d_partial;
        p.prior_mean = h->hy.intercept_prior[i]; p.prior_var = h->hy.intercept_variance_prior;
        p.step = F<double>(h, DLSM_F_B_STEP);
        p.nacc = F<int32_t>(h, DLSM_F_B_NACC);
        p.nsteps = F<int32_t>(h, DLSM_F_B_NSTEPS);
        p.until = F<int32_t>(h, DLSM_F_B_UNTIL);
        p.eps = d_eps; p.logu = d_logu; p.m = m;
        p.seed = h->seed; p.sweep = h->sweep_idx[kRngIntercept];
        p.chain_offset = (uint32_t)h->chain_offset;
        p.site = (uint32_t)((size_t)h->cfg.T * h->cfg.n);
        p.accepted = d_acc; p.ratio = d_ratio; p.flags = h->d_flags;
        const bool cur_now = use_cur || (seed_cur && i > 0);
        p.ll_cur = (cur_now || seed_cur) ? F<double>(h, DLSM_F_LOGLIK) : nullptr; p.use_cur = cur_now ? 1 : 0;
        int rc = launch_simple(h, k_intercept_propose, g1, b1, 0, p);
        if (rc != DLSM_OK) return rc;
        if (rows) rc = launch_rows(h, h->rinv); // the proposal's log-likelihood AND its row sums
        else rc = launch_full(h, h->rinv, h->rinv, cur_now ? 1 : 2);
        if (rc != DLSM_OK) return rc;
        if ((rc = launch_simple(h, k_intercept_finalize, dim3((C + 3) / 4), b1, 0, p)) != DLSM_OK) return rc; // warp per chain
        if (rows && (rc = commit_rows(h, h->d_accflag)) != DLSM_OK) return rc; // accepted chains adopt them
    }
    end_phase(h);
    if (!d_eps) h->sweep_idx[kRngIntercept] += 1;
    return DLSM_OK;
}

int dlsm_sample_intercepts(dlsm_handle *h, const double *eps, const double *logu,
                           int32_t *accepted, double *ratio)
{
    if (!h) return DLSM_ERR_INVALID;
    if ((eps == nullptr) != (logu == nullptr)) FAIL(h, DLSM_ERR_INVALID, "eps and logu must both be given or both NULL");
    CU(h, cudaSetDevice(h->cfg.device));
    int rc = need_inputs(h);
    if (rc != DLSM_OK) return rc;
    const size_t C = h->cfg.n_chains, m = h->cfg.is_directed ? 2 : 1;
    double *d_eps = nullptr, *d_logu = nullptr;
    h->rows_valid = false, h->sweeps_since_set = 0;
    if (eps) {
        d_eps = h->d_small; d_logu = h->d_small + C * 2;
        if ((rc = upload(h, d_eps, eps, C * m * 8)) != DLSM_OK) return rc;
        if ((rc = upload(h, d_logu, logu, C * m * 8)) != DLSM_OK) return rc;
    }
    if ((rc = intercepts_async(h, d_eps, d_logu, accepted ? h->d_small_i : nullptr,
                               ratio ? h->d_ll2 : nullptr)) != DLSM_OK)
        return rc;
    if (accepted) CU(h, cudaMemcpyAsync(accepted, h->d_small_i, C * m * 4, cudaMemcpyDeviceToHost, h->stream));
    if (ratio) CU(h, cudaMemcpyAsync(ratio, h->d_ll2, C * m * 8, cudaMemcpyDeviceToHost, h->stream));
    return check_flags(h);
}

static int radii_async(dlsm_handle *h, bool native, const double *d_logu, int32_t *d_acc,
                       double *d_ratio, bool use_cur = false, bool rows = false)
{
    const int C = h->cfg.n_chains, n = h->cfg.n;
    begin_phase(h, 1);
    int rc;
    const uint32_t site0 = (uint32_t)((size_t)h->cfg.T * n + 8);
    const int chunks = (n + kRadiiChunk - 1) / kRadiiChunk;
    if (!h->d_radii_terms) CU(h, cudaMalloc((void **)&h->d_radii_terms, (size_t)C * chunks * 6 * sizeof(double)));
    if (native) {
        rc = launch_simple(h, k_radii_gammas, dim3(chunks, C), dim3(256), 0, n,
                           (const double *)F<double>(h, DLSM_F_RADII),
                           (const double *)F<double>(h, DLSM_F_R_STEP), h->d_rprop, h->d_radii_terms,
                           h->seed, h->sweep_idx[kRngRadii], (uint32_t)h->chain_offset, site0);
        if (rc != DLSM_OK) return rc;
        rc = launch_simple(h, k_radii_propose, dim3(C), dim3(1024), 0, n, chunks, (const double *)h->d_radii_terms,
                           h->d_rprop, h->d_rprop_inv);
    } else {
        const size_t total = (size_t)C * n;
        rc = launch_simple(h, k_rinv, dim3((unsigned)((total + 255) / 256)), dim3(256), 0,
                           (const double *)h->d_rprop, h->d_rprop_inv, total);
    }
    if (rc != DLSM_OK) return rc;
    rc = launch_simple(h, k_bvar_current, dim3((C + 127) / 128), dim3(128), 0, C,
                       (const double *)F<double>(h, DLSM_F_INTERCEPT), h->d_bvar);
    if (rc != DLSM_OK) return rc;
    if (rows) rc = launch_rows(h, h->d_rprop_inv);
    else rc = launch_full(h, h->d_rprop_inv, h->rinv, use_cur ? 1 : 2);
    if (rc != DLSM_OK) return rc;
    RadiiMH p;
    memset(&p, 0, sizeof(p));
    p.C = C; p.n = n; p.nblk = rows ? rows_nblk(h) : h->full_nblk;
    p.accflag = rows ? h->d_accflag : nullptr;
    p.tune = h->cfg.radii_tune; p.tune_interval = h->cfg.radii_tune_interval;
    p.radii = F<double>(h, DLSM_F_RADII); p.rinv = h->rinv;
    p.prop = h->d_rprop; p.prop_rinv = h->d_rprop_inv; p.partial = h->d_partial;
    p.step = F<double>(h, DLSM_F_R_STEP);
    p.nacc = F<int32_t>(h, DLSM_F_R_NACC);
    p.nsteps = F<int32_t>(h, DLSM_F_R_NSTEPS);
    p.until = F<int32_t>(h, DLSM_F_R_UNTIL);
    p.logu = d_logu;
    p.seed = h->seed; p.sweep = h->sweep_idx[kRngRadii];
    p.chain_offset = (uint32_t)h->chain_offset; p.site = site0;
    p.accepted = d_acc; p.ratio = d_ratio; p.flags = h->d_flags;
    p.ll_cur = use_cur ? F<double>(h, DLSM_F_LOGLIK) : nullptr; p.use_cur = use_cur ? 1 : 0;
    rc = launch_simple(h, k_radii_terms, dim3(chunks, C), dim3(256), 0, p, h->d_radii_terms);
    if (rc != DLSM_OK) return rc;
    rc = launch_simple(h, k_radii_finalize, dim3(C), dim3(256), 0, p, (const double *)h->d_radii_terms, chunks);
    if (rc == DLSM_OK && rows) rc = commit_rows(h, h->d_accflag);
    end_phase(h);
    if (native) h->sweep_idx[kRngRadii] += 1;
    return rc;
}

int dlsm_sample_radii(dlsm_handle *h, const double *proposal, const double *logu,
                      int32_t *accepted, double *ratio)
{
    if (!h) return DLSM_ERR_INVALID;
    if (!h->cfg.is_directed) FAIL(h, DLSM_ERR_INVALID, "radii exist only for directed networks");
    if ((proposal == nullptr) != (logu == nullptr)) FAIL(h, DLSM_ERR_INVALID, "proposal and logu must both be given or both NULL");
    CU(h, cudaSetDevice(h->cfg.device));
    int rc = need_inputs(h);
    if (rc != DLSM_OK) return rc;
    const size_t C = h->cfg.n_chains, n = h->cfg.n;
    h->rows_valid = false, h->sweeps_since_set = 0;
    if (proposal) {
        if ((rc = upload(h, h->d_rprop, proposal, C * n * 8)) != DLSM_OK) return rc;
        if ((rc = upload(h, h->d_small, logu, C * 8)) != DLSM_OK) return rc;
    }
    rc = radii_async(h, proposal == nullptr, proposal ? h->d_small : nullptr,
                     accepted ? h->d_small_i : nullptr, ratio ? h->d_ll2 : nullptr);
    if (rc != DLSM_OK) return rc;
    if (accepted) CU(h, cudaMemcpyAsync(accepted, h->d_small_i, C * 4, cudaMemcpyDeviceToHost, h->stream));
    if (ratio) CU(h, cudaMemcpyAsync(ratio, h->d_ll2, C * 8, cudaMemcpyDeviceToHost, h->stream));
    return check_flags(h);
}

static int labels_async(dlsm_handle *h, const double *d_U, double *lik_out, int sample)
{
    const dlsm_config &c = h->cfg;
    LabelParams p;
    memset(&p, 0, sizeof(p));
    p.C = c.n_chains; p.T = c.T; p.n = c.n; p.d = c.d; p.K = c.K;
    p.X = F<double>(h, DLSM_F_X); p.mu = F<double>(h, DLSM_F_MU);
    p.sigma = F<double>(h, DLSM_F_SIGMA); p.lambda = F<double>(h, DLSM_F_LAMBDA);
    p.w = F<double>(h, DLSM_F_WEIGHTS);
    p.U = d_U;
    p.seed = h->seed; p.sweep = h->sweep_idx[kRngLabels]; p.chain_offset = (uint32_t)h->chain_offset;
    p.z = F<int32_t>(h, DLSM_F_Z);
    p.ncount = F<double>(h, DLSM_F_NCOUNT); p.nk = F<int32_t>(h, DLSM_F_NK);
    p.lik_out = lik_out; p.sample = sample;
    if (sample) {
        CU(h, cudaMemsetAsync(p.ncount, 0, h->field_bytes[DLSM_F_NCOUNT], h->stream));
        CU(h, cudaMemsetAsync(p.nk, 0, h->field_bytes[DLSM_F_NK], h->stream));
    }
    // thread-per-node kernel when its [T*K + 2K][TPB] shared-memory stage fits, else warp-per-node
    const size_t per_thread = ((size_t)c.T * c.K + 2 * c.K) * sizeof(double);
    const size_t extra = ((size_t)c.K * (c.d + 2) + (size_t)c.K * c.K) * sizeof(double);
    int rc;
    begin_phase(h, 1);
    // The thread-per-node kernel keeps (T*K + 2K) doubles per thread.  In shared memory that caps
    // the SM at 227 KB / footprint threads (8 warps at cfg 2) and blocks co-resident kernels; in an
    // L2-resident global stage ([entry][thread] per CTA, coalesced) the cap is the register file.
    // Measured at cfg 2: 408 us (shared) vs 277 us (global).  DLSM_FFBS_SMEM=1 forces the former.
    const size_t ctas = (size_t)((c.n + 63) / 64) * c.n_chains;
    const size_t need = ctas * per_thread * 64;
    // K <= 16: the register-resident kernel (persistent grid, stage indexed by CTA slot)
    const size_t tables = ((size_t)c.T * c.K * c.K + (size_t)c.K * (c.d + 2)) * sizeof(double);
    const int64_t ffbs_kernel = h->opt[DLSM_OPT_FFBS_KERNEL]; // thread / warp: force an older mapping (tests)
    const bool ffbs_warp = ffbs_kernel == DLSM_FFBS_WARP, ffbs_smem = h->opt[DLSM_OPT_FFBS_SMEM_STAGE] != 0;
    if (c.K <= 16 && tables <= 64 * 1024 && ffbs_kernel == DLSM_FFBS_AUTO && !ffbs_smem) {
        const int KC = (c.K + 3) / 4 * 4;
        const int tiles = (c.n + 63) / 64;
        const long items = (long)tiles * c.n_chains;
        auto launch = [&](auto kern) -> int {
            CU(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tables));
            int per_sm = 0;
            CU(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 64, tables));
            if (per_sm < 1) per_sm = 1;
            const int64_t cap = h->opt[DLSM_OPT_FFBS_CTAS_PER_SM];
            if (cap > 0 && cap < per_sm) per_sm = (int)cap;
            long grid = (long)h->sm_count * per_sm;
            if (grid > items) grid = items;
            const size_t stage = (size_t)grid * c.T * c.K * 64 * sizeof(double);
            if (stage > h->ffbs_stage_bytes) {
                cudaFree(h->d_ffbs_stage);
                h->d_ffbs_stage = nullptr;
                CU(h, cudaMalloc((void **)&h->d_ffbs_stage, stage));
                h->ffbs_stage_bytes = stage;
            }
            p.gstage = h->d_ffbs_stage;
            // The per-CTA stage is rewritten by every launch and only ever read back by the CTA that wrote
            // it: keep its lines in L2 (persisting access-policy window) so that they are not written
            // back to HBM between launches -- 139 MB of DRAM traffic per launch at cfg 2 otherwise.
            bool window = false;
            if (!h->opt[DLSM_OPT_FFBS_NO_L2_WINDOW] && h->l2_persist_max > 0 && stage <= (size_t)h->l2_persist_max) {
                if (h->l2_persist_set < stage) {
                    CU(h, cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, stage));
                    h->l2_persist_set = stage;
                }
                cudaStreamAttrValue av;
                memset(&av, 0, sizeof(av));
                av.accessPolicyWindow.base_ptr = h->d_ffbs_stage;
                av.accessPolicyWindow.num_bytes = stage;
                av.accessPolicyWindow.hitRatio = 1.0f;
                av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
                window = cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &av) == cudaSuccess;
                if (!window) cudaGetLastError();
            }
            const int lrc = launch_simple(h, kern, dim3((unsigned)grid), dim3(64), tables, p, tiles, (int)items);
            if (window) {
                cudaStreamAttrValue av;
                memset(&av, 0, sizeof(av));
                cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &av);
            }
            return lrc;
        };
        if (c.d == 2) {
            rc = KC == 4 ? launch(k_ffbs_r<4, 2>) : KC == 8 ? launch(k_ffbs_r<8, 2>)
                 : KC == 12 ? launch(k_ffbs_r<12, 2>) : launch(k_ffbs_r<16, 2>);
        } else {
            rc = KC == 4 ? launch(k_ffbs_r<4, 0>) : KC == 8 ? launch(k_ffbs_r<8, 0>)
                 : KC == 12 ? launch(k_ffbs_r<12, 0>) : launch(k_ffbs_r<16, 0>);
        }
    } else if (!ffbs_warp && !ffbs_smem && per_thread * 64 > 16 * 1024 && need <= ((size_t)4 << 30) &&
        extra <= kMaxSmem) {
        if (need > h->ffbs_stage_bytes) {
            cudaFree(h->d_ffbs_stage);
            h->d_ffbs_stage = nullptr;
            CU(h, cudaMalloc((void **)&h->d_ffbs_stage, need));
            h->ffbs_stage_bytes = need;
        }
        p.gstage = h->d_ffbs_stage;
        CU(h, cudaFuncSetAttribute(k_ffbs_t<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)extra));
        rc = launch_simple(h, k_ffbs_t<64>, dim3((c.n + 63) / 64, c.n_chains), dim3(64), extra, p);
    } else if (!ffbs_warp && per_thread * 64 + extra <= kMaxSmem / 2) {
        const size_t smem = per_thread * 64 + extra;
        CU(h, cudaFuncSetAttribute(k_ffbs_t<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        rc = launch_simple(h, k_ffbs_t<64>, dim3((c.n + 63) / 64, c.n_chains), dim3(64), smem, p);
    } else if (!ffbs_warp && per_thread * 32 + extra <= kMaxSmem) {
        const size_t smem = per_thread * 32 + extra;
        CU(h, cudaFuncSetAttribute(k_ffbs_t<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        rc = launch_simple(h, k_ffbs_t<32>, dim3((c.n + 31) / 32, c.n_chains), dim3(32), smem, p);
    } else {
        const int wpb = 4;
        const size_t per_warp = ((size_t)c.T * c.K + 3 * c.K) * sizeof(double);
        const size_t smem = per_warp * wpb;
        if (smem > kMaxSmem) FAIL(h, DLSM_ERR_UNSUPPORTED, "T*K too large for the label kernel");
        CU(h, cudaFuncSetAttribute(k_ffbs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const size_t warps = (size_t)c.n_chains * c.n;
        rc = launch_simple(h, k_ffbs, dim3((unsigned)((warps + wpb - 1) / wpb)), dim3(wpb * 32), smem, p);
    }
    end_phase(h);
    if (sample && !d_U) h->sweep_idx[kRngLabels] += 1;
    return rc;
}

int dlsm_sample_labels(dlsm_handle *h, const double *U)
{
    if (!h) return DLSM_ERR_INVALID;
    if (h->cfg.prior != DLSM_PRIOR_MIXTURE) FAIL(h, DLSM_ERR_INVALID, "labels exist only with the mixture prior");
    CU(h, cudaSetDevice(h->cfg.device));
    const size_t N = (size_t)h->cfg.n_chains * h->cfg.n * h->cfg.T;
    double *d_U = nullptr;
    if (U) {
        int rc = ensure_replay_buffers(h);
        if (rc != DLSM_OK) return rc;
        d_U = h->d_logu; // same element count (C*T*n)
        if ((rc = upload(h, d_U, U, N * 8)) != DLSM_OK) return rc;
    }
    int rc = labels_async(h, d_U, nullptr, 1);
    if (rc != DLSM_OK) return rc;
    CU(h, cudaStreamSynchronize(h->stream));
    return DLSM_OK;
}

// part: 0 = whole block, 1 = emission side (mu, sigma, lambda, tau^2, b), 2 = transition side
static int hdp_update_async(dlsm_handle *h, int part)
{
    const dlsm_config &c = h->cfg;
    HdpParams p;
    memset(&p, 0, sizeof(p));
    p.C = c.n_chains; p.T = c.T; p.n = c.n; p.d = c.d; p.K = c.K;
    p.X = F<double>(h, DLSM_F_X); p.z = F<int32_t>(h, DLSM_F_Z);
    p.ncount = F<double>(h, DLSM_F_NCOUNT); p.nk = F<int32_t>(h, DLSM_F_NK);
    p.mu = F<double>(h, DLSM_F_MU); p.sigma = F<double>(h, DLSM_F_SIGMA);
    p.lambda = F<double>(h, DLSM_F_LAMBDA); p.beta = F<double>(h, DLSM_F_BETA);
    p.weights = F<double>(h, DLSM_F_WEIGHTS); p.hyper = F<double>(h, DLSM_F_HYPER);
    p.pr = h->hdp_prior;
    p.seed = h->seed; p.sweep = h->sweep_idx[4]; p.chain_offset = (uint32_t)h->chain_offset;
    const size_t smem = hdp_smem_bytes(c.T, c.K, c.d);
    p.bin_rows = h->opt[DLSM_OPT_HDP_SEGMENTED] ? 0 : hdp_bin_rows(c.K, c.d);
    if (smem > kMaxSmem) FAIL(h, DLSM_ERR_UNSUPPORTED, "T*K*K too large for the HDP update kernel");
    begin_phase(h, 1);
    int rc;
    if (part == 1) {
        CU(h, cudaFuncSetAttribute(k_hdp_update<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        rc = launch_simple(h, k_hdp_update<1>, dim3(c.n_chains), dim3(128), smem, p);
    } else if (part == 2) {
        CU(h, cudaFuncSetAttribute(k_hdp_update<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        rc = launch_simple(h, k_hdp_update<2>, dim3(c.n_chains), dim3(128), smem, p);
    } else {
        CU(h, cudaFuncSetAttribute(k_hdp_update<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        rc = launch_simple(h, k_hdp_update<0>, dim3(c.n_chains), dim3(128), smem, p);
    }
    end_phase(h);
    if (part != 1) h->sweep_idx[4] += 1; // the emission side runs first when the block is split
    return rc;
}

int dlsm_set_hdp_prior(dlsm_handle *h, const dlsm_hdp_prior *pr)
{
    if (!h || !pr) return DLSM_ERR_INVALID;
    if (h->cfg.prior != DLSM_PRIOR_MIXTURE) FAIL(h, DLSM_ERR_INVALID, "needs the mixture prior");
    h->hdp_prior = *pr;
    h->have_hdp_prior = true;
    return DLSM_OK;
}

int dlsm_hdp_update(dlsm_handle *h)
{
    if (!h) return DLSM_ERR_INVALID;
    if (!h->have_hdp_prior) FAIL(h, DLSM_ERR_NOTSET, "dlsm_set_hdp_prior has not been called");
    CU(h, cudaSetDevice(h->cfg.device));
    int rc = hdp_update_async(h, 0);
    if (rc != DLSM_OK) return rc;
    CU(h, cudaStreamSynchronize(h->stream));
    return DLSM_OK;
}

static int procrustes_async(dlsm_handle *h);

// One sweep of the estimator loop body, everything asynchronous.  *tracked tells whether
// DLSM_F_LOGLIK holds the network log-likelihood of the state this sweep leaves behind.
static int one_sweep(dlsm_handle *h, uint32_t flags, bool *tracked)
{
    int rc;
    SweepParams p = sweep_params(h);
    const bool procrustes = h->have_proc_ref && !(flags & 1u);
    // the chain kernel with positions in shared memory centres them on the way out
    const bool fuse = !(flags & 1u) && !procrustes && !use_slice_kernel(h) &&
                      sweep_smem(h, true) <= kMaxSmem;
    p.fuse_center = fuse ? 1 : 0;
    // the chain kernel also hands over the full-network log-likelihood of the state it leaves
    // behind, so the intercept / radii MH below evaluates only its proposals
    bool use_cur = (!use_slice_kernel(h) || cluster_tracks_loglik(h)) && h->lk != kCaseControl &&
                   !h->opt[DLSM_OPT_NO_TRACKED_LOGLIK];
    p.ll_cur = use_cur ? F<double>(h, DLSM_F_LOGLIK) : nullptr;
    // ... and keeps the per-node row sums, so that a node-update evaluates its proposal only
    // (the first sweep after the state came in from outside -- dlsm_set_state, a function-level call --
    //  runs on the two-variant kernel: building the cache costs a k_rows pass, which only pays when
    //  further sweeps follow on the device)
    const bool rows = use_cur && rows_enabled(h) && (h->rows_valid || h->sweeps_since_set > 0);
    h->sweeps_since_set += 1;
    if (rows) {
        if ((rc = ensure_rows(h)) != DLSM_OK) return rc;
        if (!h->rows_valid) {
            tl_begin(h, "row sums");
            begin_phase(h, 1);
            rc = refresh_rows(h);
            end_phase(h);
            tl_end(h);
            if (rc != DLSM_OK) return rc;
        }
        p.rows = h->d_rows; p.scr = h->d_scr;
        h->ctr.rowsum_sweeps += 1;
    }
    if (h->x_copy_pending) { // the previous record of X is still on its way to the host
        CU(h, cudaStreamWaitEvent(h->stream, h->ev_x_copied, 0));
        h->x_copy_pending = false;
    }
    tl_begin(h, "sweep");
    if ((rc = launch_sweep(h, p)) != DLSM_OK) return rc;
    tl_end(h);
    h->sweep_idx[kRngLatent] += 1;
    if (procrustes && (rc = procrustes_async(h)) != DLSM_OK) return rc; // lsm.py:495-498
    if (!(flags & 1u) && !fuse && (rc = center_async(h, h->opt[DLSM_OPT_CENTER_EXACT] != 0)) != DLSM_OK) return rc;
    if (h->early_x_dst) { // X is final for this sweep: stream its record out now
        CU(h, cudaEventRecord(h->ev_x_ready, h->stream));
        CU(h, cudaStreamWaitEvent(h->x_stream, h->ev_x_ready, 0));
        CU(h, cudaMemcpyAsync(h->early_x_dst, h->field[DLSM_F_X], h->field_bytes[DLSM_F_X],
                              cudaMemcpyDeviceToHost, h->x_stream));
        CU(h, cudaEventRecord(h->ev_x_copied, h->x_stream));
        h->x_copy_pending = true;
        h->early_x_dst = nullptr;
    }
    // After centring, the label block (FFBS -> HDP update; latency-bound, few warps per SM)
    // and the intercept / radii MH (full-network kernel; issue-bound) are independent: run
    // the label block on a high-priority side stream so the two overlap.
    const bool labels = h->cfg.prior == DLSM_PRIOR_MIXTURE && !(flags & 8u);
    const bool hdp = labels && h->have_hdp_prior && !(flags & 16u);
    cudaStream_t main_stream = h->stream;
    if (labels) {
        // side stream: FFBS -> emission side of the HDP block (what the next latent sweep
        // needs) -> [event] -> transition side, which only the next label draw needs and which
        // therefore overlaps the next latent sweep as well
        CU(h, cudaEventRecord(h->ev_fork, main_stream));
        CU(h, cudaStreamWaitEvent(h->side_stream, h->ev_fork, 0));
        h->stream = h->side_stream;
        const bool timing = h->timing;
        h->timing = false; // phase events belong to the main stream
        tl_begin(h, "ffbs");
        rc = labels_async(h, nullptr, nullptr, 1);
        tl_end(h);
        tl_begin(h, "hdp emission");
        if (rc == DLSM_OK && hdp) rc = hdp_update_async(h, 1);
        tl_end(h);
        if (rc == DLSM_OK) {
            cudaError_t ce = cudaEventRecord(h->ev_join, h->side_stream);
            if (ce != cudaSuccess) rc = DLSM_ERR_CUDA;
        }
        tl_begin(h, "hdp transition");
        if (rc == DLSM_OK && hdp) rc = hdp_update_async(h, 2);
        tl_end(h);
        h->timing = timing;
        h->stream = main_stream;
        if (rc != DLSM_OK) return rc;
    }
    // Where the sweep kernel does not track the log-likelihood (case-control lists, CTA-per-slice
    // kernels) one evaluation of the current state serves all the MH steps of this sweep: 1 + 3
    // variant evaluations instead of 3 x 2.
    const bool any_mh = !(flags & 2u) || (h->cfg.is_directed && !(flags & 4u));
    // case-control lists: the full-network kernel is bound by its random 32-byte gathers, a second
    // variant on the same records is nearly free -- the first intercept step evaluates the current
    // state along with its proposal instead of a pass of its own
    const bool seed_cur = !use_cur && !(flags & 2u) && h->lk == kCaseControl && h->cfg.d == 2 &&
                          !h->opt[DLSM_OPT_NO_TRACKED_LOGLIK] && !h->opt[DLSM_OPT_NO_GATHER_PACK];
    if (!use_cur && any_mh && !seed_cur && !h->opt[DLSM_OPT_NO_TRACKED_LOGLIK]) {
        const int C = h->cfg.n_chains;
        tl_begin(h, "loglik(current)");
        rc = launch_simple(h, k_bvar_current, dim3((C + 127) / 128), dim3(128), 0, C,
                           (const double *)F<double>(h, DLSM_F_INTERCEPT), h->d_bvar);
        if (rc == DLSM_OK) rc = launch_full(h, h->rinv, h->rinv, 1);
        if (rc == DLSM_OK)
            rc = launch_simple(h, k_sum_partials, dim3((C + 3) / 4), dim3(128), 0, C, h->full_nblk,
                               (const double *)h->d_partial, h->d_ll2, F<double>(h, DLSM_F_LOGLIK));
        if (rc != DLSM_OK) return rc;
        tl_end(h);
        use_cur = true;
    }
    if (tracked) *tracked = use_cur;
    tl_begin(h, "intercepts");
    if (!(flags & 2u) && (rc = intercepts_async(h, nullptr, nullptr, nullptr, nullptr, use_cur, rows, seed_cur)) != DLSM_OK) return rc;
    tl_end(h);
    if (seed_cur) use_cur = true;
    if (tracked) *tracked = use_cur;
    tl_begin(h, "radii");
    if (h->cfg.is_directed && !(flags & 4u) &&
        (rc = radii_async(h, true, nullptr, nullptr, nullptr, use_cur, rows)) != DLSM_OK)
        return rc;
    tl_end(h);
    if (labels) CU(h, cudaStreamWaitEvent(main_stream, h->ev_join, 0));
    return DLSM_OK;
}

// the transition-side HDP update of the last sweep may still be in flight on the side stream
static int join_side_stream(dlsm_handle *h, uint32_t flags)
{
    if (h->cfg.prior == DLSM_PRIOR_MIXTURE && !(flags & 8u)) {
        CU(h, cudaEventRecord(h->ev_fork, h->side_stream));
        CU(h, cudaStreamWaitEvent(h->stream, h->ev_fork, 0));
    }
    return DLSM_OK;
}

int dlsm_run_sweeps(dlsm_handle *h, int32_t n_sweeps, uint32_t flags)
{
    if (!h || n_sweeps < 0) return DLSM_ERR_INVALID;
    CU(h, cudaSetDevice(h->cfg.device));
    int rc = need_inputs(h);
    if (rc != DLSM_OK) return rc;
    for (int s = 0; s < n_sweeps; s++)
        if ((rc = one_sweep(h, flags, nullptr)) != DLSM_OK) return rc;
    if ((rc = join_side_stream(h, flags)) != DLSM_OK) return rc;
    tl_dump(h);
    return check_flags(h);
}

// joint log-posterior of the current state -> out_dev[C]; ll_tracked: DLSM_F_LOGLIK is current
static int logp_async(dlsm_handle *h, double *out_dev, bool ll_tracked)
{
    const dlsm_config &c = h->cfg;
    int rc;
    LogpParams p;
    memset(&p, 0, sizeof(p));
    if (ll_tracked) {
        p.ll = F<double>(h, DLSM_F_LOGLIK); p.ll_stride = 1;
    } else {
        rc = launch_simple(h, k_bvar_current, dim3((c.n_chains + 127) / 128), dim3(128), 0, c.n_chains,
                           (const double *)F<double>(h, DLSM_F_INTERCEPT), h->d_bvar);
        if (rc != DLSM_OK) return rc;
        if ((rc = launch_full(h, h->rinv, h->rinv, 1)) != DLSM_OK) return rc;
        rc = launch_simple(h, k_sum_partials, dim3((c.n_chains + 3) / 4), dim3(128), 0, c.n_chains,
                           h->full_nblk, (const double *)h->d_partial, h->d_ll2, (double *)nullptr);
        if (rc != DLSM_OK) return rc;
        p.ll = h->d_ll2; p.ll_stride = 2;
    }
    p.C = c.n_chains; p.T = c.T; p.n = c.n; p.d = c.d; p.K = c.K; p.m = c.is_directed ? 2 : 1;
    p.mixture = c.prior == DLSM_PRIOR_MIXTURE; p.directed = c.is_directed;
    p.X = F<double>(h, DLSM_F_X); p.intercept = F<double>(h, DLSM_F_INTERCEPT);
    p.tau_sq = h->hy.tau_sq; p.sigma_sq = h->hy.sigma_sq;
    p.ic_prior0 = h->hy.intercept_prior[0]; p.ic_prior1 = h->hy.intercept_prior[1];
    p.ic_var = h->hy.intercept_variance_prior;
    if (p.mixture) {
        if (!h->have_hdp_prior) FAIL(h, DLSM_ERR_NOTSET, "dlsm_set_hdp_prior has not been called");
        p.mu = F<double>(h, DLSM_F_MU); p.sigma = F<double>(h, DLSM_F_SIGMA);
        p.lambda = F<double>(h, DLSM_F_LAMBDA); p.weights = F<double>(h, DLSM_F_WEIGHTS);
        p.beta = F<double>(h, DLSM_F_BETA); p.hyper = F<double>(h, DLSM_F_HYPER);
        p.z = F<int32_t>(h, DLSM_F_Z);
        p.pr = h->hdp_prior;
    }
    p.out = out_dev;
    return launch_simple(h, k_logp, dim3(c.n_chains), dim3(256), 0, p);
}

int dlsm_logp(dlsm_handle *h, double *out)
{
    if (!h || !out) return DLSM_ERR_INVALID;
    CU(h, cudaSetDevice(h->cfg.device));
    int rc = need_inputs(h);
    if (rc != DLSM_OK) return rc;
    const size_t C = h->cfg.n_chains;
    if (!h->d_logp) CU(h, cudaMalloc((void **)&h->d_logp, C * 8));
    if ((rc = logp_async(h, h->d_logp, false)) != DLSM_OK) return rc;
    if ((rc = download(h, out, h->d_logp, C * 8)) != DLSM_OK) return rc;
    return check_flags(h);
}

static int procrustes_async(dlsm_handle *h)
{
    const dlsm_config &c = h->cfg;
    double *X = F<double>(h, DLSM_F_X);
    begin_phase(h, 1);
    int rc;
    if (c.d == 2) rc = launch_simple(h, k_procrustes<2>, dim3(c.n_chains), dim3(256), 0, c.T, c.n, c.d, X, (const double *)h->d_proc_ref);
    else if (c.d == 3) rc = launch_simple(h, k_procrustes<3>, dim3(c.n_chains), dim3(256), 0, c.T, c.n, c.d, X, (const double *)h->d_proc_ref);
    else rc = launch_simple(h, k_procrustes<kMaxD>, dim3(c.n_chains), dim3(256), 0, c.T, c.n, c.d, X, (const double *)h->d_proc_ref);
    end_phase(h);
    return rc;
}

int dlsm_cooccurrence(dlsm_handle *h, uint32_t *out, uint64_t *n_samples, int32_t reset)
{
    if (!h || !n_samples) return DLSM_ERR_INVALID;
    CU(h, cudaSetDevice(h->cfg.device));
    const size_t cells = (size_t)h->cfg.T * h->cfg.n * h->cfg.n;
    *n_samples = h->d_cooc ? h->cooc_samples : 0;
    if (out) {
        if (h->d_cooc) {
            int rc = download(h, out, h->d_cooc, cells * sizeof(uint32_t));
            if (rc != DLSM_OK) return rc;
        } else {
            memset(out, 0, cells * sizeof(uint32_t));
        }
    }
    if (reset && h->d_cooc) {
        CU(h, cudaMemsetAsync(h->d_cooc, 0, cells * sizeof(uint32_t), h->stream));
        CU(h, cudaStreamSynchronize(h->stream));
        h->cooc_samples = 0;
    }
    return DLSM_OK;
}

int dlsm_edge_probas(dlsm_handle *h, int32_t chain, double *out)
{
    if (!h || !out) return DLSM_ERR_INVALID;
    const dlsm_config &c = h->cfg;
    if (chain < 0 || chain >= c.n_chains) FAIL(h, DLSM_ERR_INVALID, "chain out of range");
    CU(h, cudaSetDevice(c.device));
    const size_t cells = (size_t)c.T * c.n * c.n;
    double *d_out = nullptr;
    CU(h, cudaMalloc((void **)&d_out, cells * sizeof(double)));
    const double *X = F<double>(h, DLSM_F_X) + (size_t)chain * c.T * c.n * c.d;
    const double *ic = F<double>(h, DLSM_F_INTERCEPT) + (size_t)chain * 2;
    const double *ri = c.is_directed ? h->rinv + (size_t)chain * c.n : nullptr;
    int rc = launch_simple(h, k_edge_probas, dim3((unsigned)(((size_t)c.n * c.n + 255) / 256), c.T), dim3(256), 0,
                           X, ic, ri, c.n, c.d, (int)c.is_directed, d_out);
    if (rc == DLSM_OK) rc = download(h, out, d_out, cells * sizeof(double));
    cudaFree(d_out);
    return rc;
}

int dlsm_set_procrustes_ref(dlsm_handle *h, const double *Xref)
{
    if (!h) return DLSM_ERR_INVALID;
    CU(h, cudaSetDevice(h->cfg.device));
    if (!Xref) { h->have_proc_ref = false; return DLSM_OK; }
    const size_t bytes = h->field_bytes[DLSM_F_X];
    if (!h->d_proc_ref) CU(h, cudaMalloc((void **)&h->d_proc_ref, bytes));
    int rc = upload(h, h->d_proc_ref, Xref, bytes);
    if (rc != DLSM_OK) return rc;
    CU(h, cudaStreamSynchronize(h->stream));
    h->have_proc_ref = true;
    return DLSM_OK;
}

int dlsm_procrustes(dlsm_handle *h)
{
    if (!h) return DLSM_ERR_INVALID;
    if (!h->have_proc_ref) FAIL(h, DLSM_ERR_NOTSET, "dlsm_set_procrustes_ref has not been called");
    CU(h, cudaSetDevice(h->cfg.device));
    int rc = procrustes_async(h);
    if (rc != DLSM_OK) return rc;
    CU(h, cudaStreamSynchronize(h->stream));
    return DLSM_OK;
}

int dlsm_host_alloc(size_t bytes, void **out)
{
    if (!out || bytes == 0) return DLSM_ERR_INVALID;
    *out = nullptr;
    if (cudaHostAlloc(out, bytes, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        g_create_error = "cudaHostAlloc failed";
        return DLSM_ERR_CUDA;
    }
    return DLSM_OK;
}

int dlsm_host_free(void *p)
{
    if (!p) return DLSM_OK;
    return cudaFreeHost(p) == cudaSuccess ? DLSM_OK : DLSM_ERR_CUDA;
}

// (re)allocate the two device trace chunks for a trace specification
static int prepare_trace(dlsm_handle *h, const dlsm_trace_spec *sp, int n_records, bool early_x)
{
    const dlsm_config &c = h->cfg;
    size_t slot[DLSM_F_COUNT_] = {0};
    size_t per_record = sp->want_logp ? (size_t)c.n_chains * 8 : 0;
    int nseg = 0;
    for (int f = 0; f < DLSM_F_COUNT_; f++) {
        const bool all = (sp->fields_all >> f) & 1u, first = (sp->fields_first >> f) & 1u;
        if (!all && !first) continue;
        if (h->field_bytes[f] == 0) FAIL(h, DLSM_ERR_INVALID, "traced field %d does not exist in this configuration", f);
        slot[f] = all ? h->field_bytes[f] : h->field_bytes[f] / c.n_chains;
        if (early_x && f == DLSM_F_X) continue; // copied straight from the live state, no ring slot
        per_record += slot[f];
        nseg++;
    }
    if (nseg > kMaxSnapSeg) FAIL(h, DLSM_ERR_INVALID, "at most %d traced fields", kMaxSnapSeg);
    if (per_record == 0 && !early_x) { free_trace(h); return DLSM_OK; }
    if (per_record == 0) per_record = 1; // only the early-copied positions: events, no ring buffers
    const bool same_spec = h->trace_R > 0 && h->trace_all == sp->fields_all &&
                           h->trace_first == sp->fields_first && h->trace_logp == (sp->want_logp ? 1 : 0) &&
                           h->trace_early_x == early_x;
    if (same_spec && h->trace_R >= n_records) return DLSM_OK; // the common case: no driver query
    size_t free_b = 0, total_b = 0;
    CU(h, cudaMemGetInfo(&free_b, &total_b));
    size_t budget = (size_t)512 << 20; // per chunk
    if (budget > free_b / 8) budget = free_b / 8;
    if (h->opt[DLSM_OPT_TRACE_CHUNK_BYTES] > 0) budget = (size_t)h->opt[DLSM_OPT_TRACE_CHUNK_BYTES];
    long R = (long)(budget / per_record);
    if (R < 1) R = 1;
    if (R > 1024) R = 1024;
    if (R > n_records) R = n_records > 0 ? n_records : 1;
    if (same_spec && h->trace_R >= R) return DLSM_OK;
    CU(h, cudaStreamSynchronize(h->stream));
    if (h->copy_stream) CU(h, cudaStreamSynchronize(h->copy_stream));
    free_trace(h);
    if (!h->copy_stream) CU(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (auto &ch : h->chunk) {
        for (int f = 0; f < DLSM_F_COUNT_; f++)
            if (slot[f] && !(early_x && f == DLSM_F_X)) CU(h, cudaMalloc(&ch.dev[f], slot[f] * (size_t)R));
        if (sp->want_logp) CU(h, cudaMalloc((void **)&ch.logp, (size_t)c.n_chains * 8 * (size_t)R));
        CU(h, cudaEventCreateWithFlags(&ch.filled, cudaEventDisableTiming));
        CU(h, cudaEventCreateWithFlags(&ch.drained, cudaEventDisableTiming));
    }
    memcpy(h->trace_slot, slot, sizeof(slot));
    h->trace_all = sp->fields_all; h->trace_first = sp->fields_first;
    h->trace_logp = sp->want_logp ? 1 : 0;
    h->trace_early_x = early_x;
    h->trace_R = (int)R;
    return DLSM_OK;
}

// device chunk -> host destination records [first, first + count), on the copy stream
static int drain_chunk(dlsm_handle *h, int which, size_t first, int count, void *const *dst, double *logp_dst)
{
    auto &ch = h->chunk[which];
    CU(h, cudaStreamWaitEvent(h->copy_stream, ch.filled, 0));
    for (int f = 0; f < DLSM_F_COUNT_; f++) {
        if (!h->trace_slot[f] || (h->early_x_active && f == DLSM_F_X)) continue;
        char *to = static_cast<char *>(dst[f]) + first * h->trace_slot[f];
        CU(h, cudaMemcpyAsync(to, ch.dev[f], h->trace_slot[f] * (size_t)count, cudaMemcpyDeviceToHost, h->copy_stream));
    }
    if (h->trace_logp) {
        const size_t rb = (size_t)h->cfg.n_chains * 8;
        CU(h, cudaMemcpyAsync(reinterpret_cast<char *>(logp_dst) + first * rb, ch.logp, rb * (size_t)count,
                              cudaMemcpyDeviceToHost, h->copy_stream));
    }
    CU(h, cudaEventRecord(ch.drained, h->copy_stream));
    return DLSM_OK;
}

static int run_traced_body(dlsm_handle *h, int32_t n_sweeps, uint32_t flags, const dlsm_trace_spec *sp,
                           void *const *dst, double *logp_dst);

int dlsm_run_traced(dlsm_handle *h, int32_t n_sweeps, uint32_t flags, const dlsm_trace_spec *sp,
                    void *const *dst, double *logp_dst)
{
    if (!h || !sp || n_sweeps < 0) return DLSM_ERR_INVALID;
    const int rc = run_traced_body(h, n_sweeps, flags, sp, dst, logp_dst);
    if (rc != DLSM_OK) {
        // common error exit: no copy into the caller's (possibly pinned, soon to be freed) buffers
        // may still be in flight when the error is reported, and the early-copy state is cleared
        const std::string keep = h->err;
        if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
        if (h->x_stream) cudaStreamSynchronize(h->x_stream);
        cudaStreamSynchronize(h->side_stream);
        cudaStreamSynchronize(h->stream);
        cudaGetLastError();
        h->early_x_active = false;
        h->early_x_dst = nullptr;
        h->x_copy_pending = false;
        h->err = keep;
    }
    return rc;
}

static int run_traced_body(dlsm_handle *h, int32_t n_sweeps, uint32_t flags, const dlsm_trace_spec *sp,
                           void *const *dst, double *logp_dst)
{
    if (sp->thin < 1) FAIL(h, DLSM_ERR_INVALID, "thin must be >= 1");
    if (sp->fields_all & sp->fields_first) FAIL(h, DLSM_ERR_INVALID, "a field is traced either for all chains or for the first");
    CU(h, cudaSetDevice(h->cfg.device));
    int rc = need_inputs(h);
    if (rc != DLSM_OK) return rc;
    const int n_records = n_sweeps / sp->thin;
    for (int f = 0; f < DLSM_F_COUNT_; f++)
        if ((((sp->fields_all | sp->fields_first) >> f) & 1u) && n_records > 0 && (!dst || !dst[f]))
            FAIL(h, DLSM_ERR_INVALID, "no destination for traced field %d", f);
    if (sp->want_logp && n_records > 0 && !logp_dst) FAIL(h, DLSM_ERR_INVALID, "no destination for the log-posterior trace");
    // Positions traced for every chain into page-locked memory leave early (see one_sweep); a
    // pageable destination would block the host inside the sweep, so it goes through the ring.
    bool early_x = false;
    if (n_records > 0 && !(flags & 1u) && ((sp->fields_all >> DLSM_F_X) & 1u) &&
        h->field_bytes[DLSM_F_X] >= ((size_t)1 << 20) && !h->opt[DLSM_OPT_NO_EARLY_X]) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, dst[DLSM_F_X]) == cudaSuccess && at.type == cudaMemoryTypeHost) early_x = true;
        cudaGetLastError();
    }
    if ((rc = prepare_trace(h, sp, n_records, early_x)) != DLSM_OK) return rc;
    const bool tracing = (h->trace_R > 0 || early_x) && n_records > 0;
    // records per chunk: the allocated capacity, but at least ~4 chunks per call so that the copy
    // of one chunk overlaps the sweeps filling the next even in short, heavy runs
    int R = h->trace_R;
    if (R > (n_records + 3) / 4) R = (n_records + 3) / 4;
    if (R < 1) R = 1;
    if (early_x && !h->ev_x_ready) {
        CU(h, cudaEventCreateWithFlags(&h->ev_x_ready, cudaEventDisableTiming));
        CU(h, cudaEventCreateWithFlags(&h->ev_x_copied, cudaEventDisableTiming));
        CU(h, cudaStreamCreateWithFlags(&h->x_stream, cudaStreamNonBlocking));
    }
    h->early_x_active = early_x;
    int cur = 0, fill = 0;          // chunk being filled, records in it
    size_t done = 0;                // records handed to earlier chunks
    int pending = -1, pending_n = 0; // filled chunk not yet drained (one chunk of lookahead keeps
    size_t pending_first = 0;        // the launch queue busy while a pageable copy blocks the host)
    for (auto &ch : h->chunk) ch.used = false;
    for (int s = 0; s < n_sweeps; s++) {
        bool tracked = false;
        const bool record = tracing && (s + 1) % sp->thin == 0;
        h->early_x_dst = (early_x && record)
                             ? static_cast<char *>(dst[DLSM_F_X]) + (done + fill) * h->trace_slot[DLSM_F_X]
                             : nullptr;
        if ((rc = one_sweep(h, flags, &tracked)) != DLSM_OK) return rc;
        if (!record) continue;
        if ((rc = join_side_stream(h, flags)) != DLSM_OK) return rc;
        if (sp->cooc_mode && h->cfg.prior == DLSM_PRIOR_MIXTURE && (int64_t)(done + fill) >= sp->cooc_from) {
            const dlsm_config &c = h->cfg;
            const size_t cells = (size_t)c.T * c.n * c.n;
            if (!h->d_cooc) {
                CU(h, cudaMalloc((void **)&h->d_cooc, cells * sizeof(uint32_t)));
                CU(h, cudaMemsetAsync(h->d_cooc, 0, cells * sizeof(uint32_t), h->stream));
                h->cooc_samples = 0;
            }
            const int C_use = sp->cooc_mode == 2 ? c.n_chains : 1;
            rc = launch_simple(h, k_cooc_accumulate, dim3((unsigned)(((size_t)c.n * c.n + 255) / 256), c.T), dim3(256),
                               0, (const int32_t *)F<int32_t>(h, DLSM_F_Z), C_use, c.T, c.n, h->d_cooc);
            if (rc != DLSM_OK) return rc;
            h->cooc_samples += (uint64_t)C_use;
        }
        auto &ch = h->chunk[cur];
        if (fill == 0 && ch.used) CU(h, cudaStreamWaitEvent(h->stream, ch.drained, 0));
        if (h->trace_logp &&
            (rc = logp_async(h, ch.logp + (size_t)fill * h->cfg.n_chains, tracked)) != DLSM_OK)
            return rc;
        SnapParams sn;
        memset(&sn, 0, sizeof(sn));
        size_t max_words = 0;
        for (int f = 0; f < DLSM_F_COUNT_; f++) {
            if (!h->trace_slot[f] || (early_x && f == DLSM_F_X)) continue;
            sn.src[sn.nseg] = static_cast<const uint32_t *>(h->field[f]);
            sn.dst[sn.nseg] = reinterpret_cast<uint32_t *>(static_cast<char *>(ch.dev[f]) + (size_t)fill * h->trace_slot[f]);
            sn.words[sn.nseg] = h->trace_slot[f] / 4;
            if (sn.words[sn.nseg] > max_words) max_words = sn.words[sn.nseg];
            sn.nseg++;
        }
        if (sn.nseg) {
            size_t blocks = (max_words / 4 + 255) / 256;
            if (blocks < 1) blocks = 1;
            if (blocks > 592) blocks = 592; // 4 CTAs per SM
            if ((rc = launch_simple(h, k_snapshot, dim3((unsigned)blocks, sn.nseg), dim3(256), 0, sn)) != DLSM_OK) return rc;
        }
        fill++;
        if (fill == R) {
            CU(h, cudaEventRecord(ch.filled, h->stream));
            ch.used = true;
            if (pending >= 0 && (rc = drain_chunk(h, pending, pending_first, pending_n, dst, logp_dst)) != DLSM_OK) return rc;
            pending = cur; pending_n = fill; pending_first = done;
            done += fill; fill = 0; cur ^= 1;
        }
    }
    if (tracing) {
        if (pending >= 0 && (rc = drain_chunk(h, pending, pending_first, pending_n, dst, logp_dst)) != DLSM_OK) return rc;
        if (fill > 0) {
            CU(h, cudaEventRecord(h->chunk[cur].filled, h->stream));
            h->chunk[cur].used = true;
            if ((rc = drain_chunk(h, cur, done, fill, dst, logp_dst)) != DLSM_OK) return rc;
        }
        CU(h, cudaStreamSynchronize(h->copy_stream));
        if (early_x) CU(h, cudaStreamSynchronize(h->x_stream));
    }
    h->early_x_active = false;
    h->x_copy_pending = false; // the copy stream has been drained
    if ((rc = join_side_stream(h, flags)) != DLSM_OK) return rc;
    return check_flags(h);
}

int dlsm_loglik_partial(dlsm_handle *h, double *out)
{
    if (!h || !out) return DLSM_ERR_INVALID;
    CU(h, cudaSetDevice(h->cfg.device));
    int rc = need_inputs(h);
    if (rc != DLSM_OK) return rc;
    if ((rc = ensure_replay_buffers(h)) != DLSM_OK) return rc;
    const size_t N = (size_t)h->cfg.n_chains * h->cfg.T * h->cfg.n;
    SweepParams p = sweep_params(h);
    const dim3 grid((unsigned)((N * 32 + 255) / 256)), block(256);
    const bool d2 = h->cfg.d == 2;
    if (h->lk == kUndirected) rc = d2 ? launch_simple(h, k_partial<kUndirected, 2>, grid, block, 0, p, h->d_ratio) : launch_simple(h, k_partial<kUndirected, 0>, grid, block, 0, p, h->d_ratio);
    else if (h->lk == kDirected) rc = d2 ? launch_simple(h, k_partial<kDirected, 2>, grid, block, 0, p, h->d_ratio) : launch_simple(h, k_partial<kDirected, 0>, grid, block, 0, p, h->d_ratio);
    else rc = d2 ? launch_simple(h, k_partial<kCaseControl, 2>, grid, block, 0, p, h->d_ratio) : launch_simple(h, k_partial<kCaseControl, 0>, grid, block, 0, p, h->d_ratio);
    if (rc != DLSM_OK) return rc;
    return download(h, out, h->d_ratio, N * 8);
}

int dlsm_loglik_full(dlsm_handle *h, double *out)
{
    if (!h || !out) return DLSM_ERR_INVALID;
    CU(h, cudaSetDevice(h->cfg.device));
    int rc = need_inputs(h);
    if (rc != DLSM_OK) return rc;
    const int C = h->cfg.n_chains;
    rc = launch_simple(h, k_bvar_current, dim3((C + 127) / 128), dim3(128), 0, C,
                       (const double *)F<double>(h, DLSM_F_INTERCEPT), h->d_bvar);
    if (rc != DLSM_OK) return rc;
    if ((rc = launch_full(h, h->rinv, h->rinv)) != DLSM_OK) return rc;
    rc = launch_simple(h, k_sum_partials, dim3((C + 3) / 4), dim3(128), 0, C, h->full_nblk,
                       (const double *)h->d_partial, h->d_ll2, (double *)nullptr);
    if (rc != DLSM_OK) return rc;
    std::vector<double> tmp((size_t)C * 2);
    if ((rc = download(h, tmp.data(), h->d_ll2, (size_t)C * 16)) != DLSM_OK) return rc;
    for (int c = 0; c < C; c++) out[c] = tmp[(size_t)c * 2 + 1];
    return DLSM_OK;
}

int dlsm_gaussian_likelihood(dlsm_handle *h, double *out)
{
    if (!h || !out) return DLSM_ERR_INVALID;
    if (h->cfg.prior != DLSM_PRIOR_MIXTURE) FAIL(h, DLSM_ERR_INVALID, "needs the mixture prior");
    CU(h, cudaSetDevice(h->cfg.device));
    const size_t N = (size_t)h->cfg.n_chains * h->cfg.n * h->cfg.T * h->cfg.K;
    double *d = nullptr;
    CU(h, cudaMalloc((void **)&d, N * 8));
    int rc = labels_async(h, nullptr, d, 0);
    if (rc == DLSM_OK) rc = download(h, out, d, N * 8);
    cudaFree(d);
    return rc;
}

int dlsm_debug_set_counts(dlsm_handle *h, int field, const void *host, size_t bytes)
{
    if (!h || !host || (field != DLSM_F_NCOUNT && field != DLSM_F_NK)) return DLSM_ERR_INVALID;
    if (bytes != h->field_bytes[field] || bytes == 0) FAIL(h, DLSM_ERR_INVALID, "bad size");
    CU(h, cudaSetDevice(h->cfg.device));
    int rc = upload(h, h->field[field], host, bytes);
    if (rc != DLSM_OK) return rc;
    CU(h, cudaStreamSynchronize(h->stream));
    return DLSM_OK;
}

int dlsm_debug_rowsums(dlsm_handle *h, double *out)
{
    if (!h || !out) return DLSM_ERR_INVALID;
    CU(h, cudaSetDevice(h->cfg.device));
    int rc = need_inputs(h);
    if (rc != DLSM_OK) return rc;
    if (!rows_enabled(h)) FAIL(h, DLSM_ERR_UNSUPPORTED, "no row-sum cache in this configuration");
    if ((rc = ensure_rows(h)) != DLSM_OK) return rc;
    if (!h->rows_valid && (rc = refresh_rows(h)) != DLSM_OK) return rc;
    return download(h, out, h->d_rows, (size_t)h->cfg.n_chains * h->cfg.T * h->cfg.n * 8);
}

int dlsm_debug_draws(dlsm_handle *h, double *eps, double *logu)
{
    if (!h || !eps || !logu) return DLSM_ERR_INVALID;
    CU(h, cudaSetDevice(h->cfg.device));
    int rc = ensure_replay_buffers(h);
    if (rc != DLSM_OK) return rc;
    const size_t N = (size_t)h->cfg.n_chains * h->cfg.T * h->cfg.n;
    SweepParams p = sweep_params(h);
    const dim3 grid((unsigned)((N + 255) / 256)), block(256);
    rc = (h->cfg.d == 2) ? launch_simple(h, k_debug_draws<2>, grid, block, 0, p, h->d_eps, h->d_logu)
                         : launch_simple(h, k_debug_draws<0>, grid, block, 0, p, h->d_eps, h->d_logu);
    if (rc != DLSM_OK) return rc;
    if ((rc = download(h, eps, h->d_eps, N * h->cfg.d * 8)) != DLSM_OK) return rc;
    return download(h, logu, h->d_logu, N * 8);
}

int dlsm_enable_timing(dlsm_handle *h, int on)
{
    if (!h) return DLSM_ERR_INVALID;
    flush_events(h);
    h->timing = on != 0;
    return DLSM_OK;
}

int dlsm_get_counters(dlsm_handle *h, dlsm_counters *out)
{
    if (!h || !out) return DLSM_ERR_INVALID;
    CU(h, cudaSetDevice(h->cfg.device));
    flush_events(h);
    *out = h->ctr;
    return DLSM_OK;
}

} // extern "C"
