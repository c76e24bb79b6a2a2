/*
 * dlsm.h -- C-ABI of libdlsm.so: the B200 (sm_100a) implementation of dynetlsm's blocked
 * Metropolis-Hastings-within-Gibbs hot path.
 *
 * This is the drop-in boundary.  The reference has no FFI of its own for
 * this path -- its "native boundary" is Python -> Cython `def` functions on typed memoryviews --
 * so each entry point below names the reference interface it replaces (paths relative to the
 * reference root).  The binding a maintainer would add to the reference is a ctypes stub; see
 * INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; every array is C-contiguous; fp64 unless noted; label and
 *     index arrays are int32 on this side (the reference uses int64; the host layer converts).
 *   - host pointers are borrowed for the duration of the call; the handle owns all device memory.
 *   - every call returns 0 on success or a negative dlsm_status; dlsm_last_error() gives text.
 *   - there is NO CPU fallback: every compute entry point fails with DLSM_ERR_CUDA when no CUDA
 *     device / sm_100 image is available.
 *   - calls on one handle must be serialised by the caller; one CUDA stream per handle.
 *   - all per-chain arrays carry a leading chain axis C = cfg.n_chains.
 */
#ifndef DLSM_H
#define DLSM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DLSM_ABI_VERSION 2

typedef struct dlsm_handle dlsm_handle;

typedef enum {
    DLSM_OK = 0,
    DLSM_ERR_INVALID = -1,   /* bad argument / shape / state */
    DLSM_ERR_CUDA = -2,      /* CUDA runtime error or no device */
    DLSM_ERR_NONBINARY = -3, /* adjacency holds a value other than 0/1 (weights, -1 missing) */
    DLSM_ERR_NOTSET = -4,    /* a required input (network, case-control lists, ...) is missing */
    DLSM_ERR_NONFINITE = -5, /* a log-likelihood evaluated to NaN/inf */
    DLSM_ERR_UNSUPPORTED = -6
} dlsm_status;

/* likelihood kinds */
#define DLSM_LIK_EXACT 0        /* static_network_fast.pyx:17 / directed_likelihoods_fast.pyx:46 */
#define DLSM_LIK_CASE_CONTROL 1 /* directed_likelihoods_fast.pyx:83 (directed only, lsm.py:425) */
/* latent-position priors */
#define DLSM_PRIOR_LSM 0     /* sample_latent_positions.py:131-140 */
#define DLSM_PRIOR_MIXTURE 1 /* sample_latent_positions.py:187-199 */

typedef struct {
    int32_t n_chains;     /* independent chains sharing one network */
    int32_t T, n, d;      /* time steps, nodes, latent dimension (d <= 8) */
    int32_t K;            /* mixture components (0 for the plain LSM) */
    int32_t is_directed;  /* lsm.py:236 */
    int32_t likelihood;   /* DLSM_LIK_* */
    int32_t prior;        /* DLSM_PRIOR_* */
    int32_t device;       /* CUDA device ordinal */
    /* Metropolis tuner wiring (metropolis.py:85-94).  tune < 0 means tune=None. */
    int32_t tune;                 /* latent + intercept samplers */
    int32_t tune_interval;        /* latent samplers */
    int32_t intercept_tune_interval[2];
    int32_t radii_tune;           /* lsm.py:470 passes None, hdp_lpcm.py:745 passes tune */
    int32_t radii_tune_interval;
    int32_t reserved[4];
} dlsm_config;

/* state fields for dlsm_set_state / dlsm_get_state; shapes with leading chain axis C */
typedef enum {
    DLSM_F_X = 0,          /* f64 (C,T,n,d)  latent positions */
    DLSM_F_INTERCEPT = 1,  /* f64 (C,2)      [beta] or [beta_in, beta_out] */
    DLSM_F_RADII = 2,      /* f64 (C,n) */
    DLSM_F_Z = 3,          /* i32 (C,T,n)    labels */
    DLSM_F_MU = 4,         /* f64 (C,K,d) */
    DLSM_F_SIGMA = 5,      /* f64 (C,K)      variances */
    DLSM_F_LAMBDA = 6,     /* f64 (C,) */
    DLSM_F_WEIGHTS = 7,    /* f64 (C,T,K,K)  w[0,0,:] = initial distribution (hdp_lpcm.py:123) */
    DLSM_F_X_STEP = 8,     /* f64 (C,T,n)    Metropolis.step_size of each latent sampler */
    DLSM_F_X_NACC = 9,     /* i32 (C,T,n)    n_accepted */
    DLSM_F_X_NSTEPS = 10,  /* i32 (C,T,n)    n_steps */
    DLSM_F_X_UNTIL = 11,   /* i32 (C,T,n)    steps_until_tune */
    DLSM_F_B_STEP = 12,    /* f64 (C,2)      intercept samplers */
    DLSM_F_B_NACC = 13,    /* i32 (C,2) */
    DLSM_F_B_NSTEPS = 14,  /* i32 (C,2) */
    DLSM_F_B_UNTIL = 15,   /* i32 (C,2) */
    DLSM_F_R_STEP = 16,    /* f64 (C,)       radii sampler (Dirichlet concentration scale) */
    DLSM_F_R_NACC = 17,    /* i32 (C,) */
    DLSM_F_R_NSTEPS = 18,  /* i32 (C,) */
    DLSM_F_R_UNTIL = 19,   /* i32 (C,) */
    DLSM_F_NCOUNT = 20,    /* f64 (C,T,K,K)  transition counts of the last label draw (read-only) */
    DLSM_F_NK = 21,        /* i32 (C,T,K)    occupancy counts of the last label draw (read-only) */
    DLSM_F_BETA = 22,      /* f64 (C,K)      global HDP weights beta */
    DLSM_F_HYPER = 23,     /* f64 (C,8)      [gamma, alpha_init, alpha, kappa, mean_variance_prior, b, 0, 0] */
    DLSM_F_LOGLIK = 24,    /* f64 (C,)       network log-likelihood of the state dlsm_run_sweeps left behind,
                              tracked by the device loop (read-only; dense likelihoods, chain kernel) */
    DLSM_F_COUNT_
} dlsm_field;

/* hyper-parameters of the MH targets */
typedef struct {
    double tau_sq;            /* lsm.py:241 */
    double sigma_sq;          /* lsm.py:242 */
    double intercept_prior[2];        /* prior means (lsm.py:239 'auto' resolved by the host) */
    double intercept_variance_prior;  /* lsm.py:240 */
} dlsm_hyper;

/* fixed hyper-hyper-parameters of the sticky HDP-HMM mixture (hdp_lpcm.py:385-455, :750-793) */
typedef struct {
    double a;                     /* sigma_k ~ InvGamma(a/2, b/2) */
    double a0, b0;                /* mean_variance_prior ~ InvGamma(a0/2, b0/2)  (if resample_mvp) */
    double c0, d0;                /* b ~ Gamma(c0/2, 2/d0)                        (if resample_b)   */
    double lambda_prior, lambda_variance_prior;
    double gamma_prior_shape, gamma_prior_rate;
    double alpha_init_shape, alpha_init_rate;
    double alpha_kappa_shape, alpha_kappa_rate;
    int32_t resample_mvp, resample_b;
} dlsm_hdp_prior;

/* ---- lifecycle ------------------------------------------------------------------------- */
int dlsm_abi_version(void);
int dlsm_device_count(void); /* 0 when no usable CUDA device */
int dlsm_create(const dlsm_config *cfg, dlsm_handle **out);
void dlsm_destroy(dlsm_handle *h);
const char *dlsm_last_error(const dlsm_handle *h); /* h may be NULL: last creation error */
/* run on a caller-provided cudaStream_t (e.g. the framework's current stream); NULL restores
 * the handle's own stream */
int dlsm_set_stream(dlsm_handle *h, void *cuda_stream);
int dlsm_synchronize(dlsm_handle *h);

/* Developer / test options: which kernel variant serves a step.  None changes results beyond the
 * documented rounding (the parity tests run every variant).  Each option also has an environment
 * variable that only supplies its DEFAULT, read once by dlsm_create (never on the sweep path):
 * DLSM_SWEEP_MODE=chain|chain-dense|slice|slice-plain, DLSM_FFBS=thread|warp, DLSM_FFBS_SMEM,
 * DLSM_FFBS_PER_SM=<n>, DLSM_NO_GATHER_PACK, DLSM_NO_LLCUR, DLSM_CENTER_EXACT, DLSM_HDP_SEGMENTED,
 * DLSM_NO_EARLY_X, DLSM_TRACE_CHUNK_BYTES=<bytes>, DLSM_NO_ROWSUM, DLSM_NO_CLUSTER=1|2,
 * DLSM_CHAIN_KERNEL=rowsum|node|block|block2, DLSM_CC_KERNEL=1|2|3|4, DLSM_FULL_KERNEL=1|2, DLSM_CCD_GROUP=<chains>,
 * DLSM_FFBS_NO_L2_WINDOW. */
typedef enum {
    DLSM_OPT_SWEEP_MODE = 0,        /* dlsm_sweep_mode: which latent-sweep kernel (default: heuristic) */
    DLSM_OPT_FFBS_KERNEL = 1,       /* dlsm_ffbs_kernel: label kernel mapping */
    DLSM_OPT_FFBS_SMEM_STAGE = 2,   /* 1: thread-per-node label kernel keeps its stage in shared memory */
    DLSM_OPT_FFBS_CTAS_PER_SM = 3,  /* cap of the persistent label grid (0 = occupancy) */
    DLSM_OPT_NO_GATHER_PACK = 4,    /* 1: case-control full-network kernel gathers X and 1/r separately */
    DLSM_OPT_NO_TRACKED_LOGLIK = 5, /* 1: intercept / radii MH evaluate current AND proposal every time */
    DLSM_OPT_CENTER_EXACT = 6,      /* 1: long-chain centring inside the device loop uses numpy's serial order */
    DLSM_OPT_HDP_SEGMENTED = 7,     /* 1: segmented-butterfly sufficient statistics in the HDP update */
    DLSM_OPT_NO_EARLY_X = 8,        /* 1: position records always go through the trace ring */
    DLSM_OPT_TRACE_CHUNK_BYTES = 9, /* device bytes of one trace-ring chunk (0 = 512 MB or free/8) */
    DLSM_OPT_NO_ROWSUM_CACHE = 10,  /* 1: the device loop evaluates proposal AND current position of every
                                       node-update afresh instead of keeping per-node row sums */
    DLSM_OPT_NO_CLUSTER = 11,       /* few (chain, slice) pairs, long rows: 0 = block-speculative sweep on a
                                       thread-block cluster per pair with a two-block window (k_sweep_blkw),
                                       1 = no clusters (CTA per pair), 2 = per-node cluster kernel
                                       (k_sweep_slice_cl), 3 = block-speculative sweep without the window
                                       (k_sweep_blk) */
    DLSM_OPT_CHAIN_KERNEL = 12,     /* dlsm_chain_kernel: the one-CTA-per-chain sweep kernel (exact likelihoods) */
    DLSM_OPT_CC_KERNEL = 13,        /* case-control sweep of (chain, slice) pairs: 0 = auto (k_sweep_ccd: dataflow
                                       over nodes, positions double-buffered for the sweep, d = 2 and <= 128
                                       controls; else k_sweep_cc), 1 = k_sweep_cc (runs of mutually independent
                                       nodes), 2 = k_sweep_cc2 (per-block list staging, 256-bit gather records),
                                       3 = k_sweep_cc3 (2-CTA cluster per pair), 4 = k_sweep_ccd */
    DLSM_OPT_FULL_KERNEL = 14,      /* full-network log-likelihood of one variant (exact likelihoods, d = 2): 0 = auto
                                       (k_full_lr: lanes = rows, broadcast columns, for n >= 256), 1 = k_full (folded
                                       rows, lanes = columns), 2 = k_full_lr at any n */
    DLSM_OPT_CCD_GROUP = 15,        /* chains per launch of the dataflow case-control sweep (0 = heuristic: about 64
                                       warps per (chain, slice) pair) */
    DLSM_OPT_FFBS_NO_L2_WINDOW = 16, /* 1: no persisting-L2 access-policy window over the label kernel's global stage */
    DLSM_OPT_COUNT_
} dlsm_option;
typedef enum {
    DLSM_SWEEP_AUTO = 0, DLSM_SWEEP_CHAIN = 1, DLSM_SWEEP_CHAIN_DENSE = 2, DLSM_SWEEP_SLICE = 3,
    DLSM_SWEEP_SLICE_PLAIN = 4
} dlsm_sweep_mode;
typedef enum { DLSM_FFBS_AUTO = 0, DLSM_FFBS_THREAD = 1, DLSM_FFBS_WARP = 2 } dlsm_ffbs_kernel;
typedef enum {
    DLSM_CHAIN_AUTO = 0,        /* block kernel up to two chains per SM, else row-sum cache for n >= 256, else node */
    DLSM_CHAIN_NODE_ROWSUM = 1, /* node by node; the device loop evaluates proposals only (row-sum cache) */
    DLSM_CHAIN_NODE = 2,        /* node by node, proposal and current position evaluated afresh (k_sweep) */
    DLSM_CHAIN_BLOCK = 3,       /* 32 nodes per step, lanes = rows (k_sweep_cb) */
    DLSM_CHAIN_BLOCK_PAIR = 4   /* the same with two warps per slice where a chain has an SM to itself
                                   (k_sweep_cbp; C <= 148, d = 2, T <= 15; measured: -2 % at n = 500, +40 % at
                                   n = 120 -- an A/B switch, not a default) */
} dlsm_chain_kernel;
int dlsm_set_option(dlsm_handle *h, int option, int64_t value);

/* ---- network -------------------------------------------------------------------------- */
/* Y (T,n,n) fp64 0/1 as passed to DynamicNetworkLSM.fit (lsm.py:319-343).  Bit-packed on the
 * device: row-major always, transposed too when is_directed. */
int dlsm_set_network_dense(dlsm_handle *h, const double *Y);
/* Sparse input for networks whose dense Y cannot exist (cfg 5).  Replaces
 * DirectedCaseControlSampler.init (case_control_likelihood.py:37-73): degrees (T,n,2) [in,out],
 * in_edges (T,n,max_in), out_edges (T,n,max_out), zero-padded. */
int dlsm_set_edge_lists(dlsm_handle *h, const int32_t *degrees, const int32_t *in_edges,
                        int32_t max_in, const int32_t *out_edges, int32_t max_out);
/* The same bookkeeping built ON THE DEVICE from the ties themselves: edges (n_edges, 3) int32 rows
 * (t, sender, receiver), any order, every tie once, no self ties.  Produces what the reference derives
 * from the dense tensor -- degrees, in/out lists in ascending order, zero-padded to the largest degree
 * (case_control_likelihood.py:44-70) -- in O(n_edges) work.  dlsm_edge_list_dims / dlsm_get_edge_lists
 * read the lists back (degrees (T,n,2), in_edges (T,n,max_in), out_edges (T,n,max_out); list pointers
 * may be NULL). */
int dlsm_set_network_edges(dlsm_handle *h, const int32_t *edges, size_t n_edges);
int dlsm_edge_list_dims(dlsm_handle *h, int32_t *max_in, int32_t *max_out);
int dlsm_get_edge_lists(dlsm_handle *h, int32_t *degrees, int32_t *in_edges, int32_t *out_edges);
/* control sets (case_control_likelihood.py:75-112): (S,T,n,n_control), -1 padded, where S = 1
 * (shared by all chains) or S = n_chains */
int dlsm_set_controls(dlsm_handle *h, const int32_t *ctrl_in, const int32_t *ctrl_out,
                      int32_t n_control, int32_t n_sets);

/* Device-side redraw of the control sets (DirectedCaseControlSampler.sample,
 * case_control_likelihood.py:75-112) on the Philox streams: for every node and direction
 * min(n - degree - 1, n_control) distinct non-neighbours, uniformly, -1 padded; n_sets = 1 (shared)
 * or n_chains (one set per chain).  Needs dlsm_set_edge_lists.  dlsm_get_controls copies the
 * current sets (n_sets,T,n,n_control) to the host. */
int dlsm_resample_controls(dlsm_handle *h, int32_t n_control, int32_t n_sets);
int dlsm_get_controls(dlsm_handle *h, int32_t *ctrl_in, int32_t *ctrl_out);

/* ---- chain state ----------------------------------------------------------------------- */
int dlsm_set_state(dlsm_handle *h, int field, const void *host, size_t bytes);
int dlsm_get_state(dlsm_handle *h, int field, void *host, size_t bytes);
int dlsm_set_hyper(dlsm_handle *h, const dlsm_hyper *hy);
/* device Philox4x32-10 streams: key = seed, counters = (site, sweep, chain+chain_offset, draw) */
int dlsm_set_rng(dlsm_handle *h, uint64_t seed, uint64_t chain_offset, uint64_t sweep_index);

/* ---- hot path -------------------------------------------------------------------------- */
/* One latent-position sweep for every chain: for t asc, j asc a random-walk MH update of X[t,j]
 * (sample_latent_positions.py:92-146 / :149-206 + metropolis.py:40-54,96-136), executed as a
 * bit-exact time-slice wavefront.
 *   replay:  eps (C,T,n,d) recorded standard normals, logu (C,T,n) recorded log-uniforms.
 *   native:  eps == logu == NULL -> device Philox streams (dlsm_set_rng).
 * Optional outputs (host, may be NULL): accepted (C,T,n) i32, ratio (C,T,n) f64. */
int dlsm_sweep_latent(dlsm_handle *h, const double *eps, const double *logu, int32_t *accepted,
                      double *ratio);
/* X -= mean(X, axis=(time, node))  (lsm.py:501, hdp_lpcm.py:852), numpy summation order */
int dlsm_center(dlsm_handle *h);
/* sample_intercepts (sample_coefficients.py:12-88): m = 1 (undirected) or 2 (directed) MH steps
 * on the full-network log-likelihood.  replay: eps (C,m), logu (C,m); native: NULL. */
int dlsm_sample_intercepts(dlsm_handle *h, const double *eps, const double *logu,
                           int32_t *accepted, double *ratio);
/* sample_radii (sample_coefficients.py:91-121 + metropolis.py:57-82).  replay: proposal (C,n) =
 * the recorded Dirichlet draw after the zero guard, logu (C,); native: NULL (device gamma
 * variates). */
int dlsm_sample_radii(dlsm_handle *h, const double *proposal, const double *logu,
                      int32_t *accepted, double *ratio);
/* sample_labels_block (sample_labels.py:134-190 + gaussian_likelihood_fast.pyx:30-54): HDP-HMM
 * forward-filter/backward-sample per node.  replay: U (C,n,T) raw uniforms, node-major;
 * native: NULL.  Updates DLSM_F_Z, DLSM_F_NCOUNT, DLSM_F_NK. */
int dlsm_sample_labels(dlsm_handle *h, const double *U);
/* The conjugate / auxiliary-variable block that follows the label draw in one HDP-LPCM sweep
 * (hdp_lpcm.py:881-1023 with sample_auxillary.py:6-50, sample_concentration.py:6-21): table
 * counts, override variables, beta, w0, w[t,k], mu_k, sigma_k, lambda, the tau^2 / b hyper-priors
 * and the concentration parameters, one CTA per chain on device Philox streams.  Reads
 * DLSM_F_NCOUNT / F_NK / F_Z / F_X, updates F_MU, F_SIGMA, F_LAMBDA, F_BETA, F_WEIGHTS, F_HYPER.
 * Native RNG only: in replay mode the host performs this block with the numpy RandomState. */
int dlsm_set_hdp_prior(dlsm_handle *h, const dlsm_hdp_prior *pr);
int dlsm_hdp_update(dlsm_handle *h);
/* n_sweeps x [latent -> center -> intercepts -> (radii) -> (labels -> (hdp update))] with device
 * RNG, no host round trip (the loop bodies lsm.py:483-523 / hdp_lpcm.py:840-1023).
 * flags: bit0 skip center, bit1 skip intercepts, bit2 skip radii, bit3 skip labels, bit4 skip the
 * HDP update (it also needs dlsm_set_hdp_prior to have been called). */
int dlsm_run_sweeps(dlsm_handle *h, int32_t n_sweeps, uint32_t flags);

/* ---- the rest of the estimator loop, device-resident ---------------- */
/* Joint log-posterior of the current state, out (C,): LSM lsm.py:576-625; HDP-LPCM
 * hdp_lpcm.py:1188-1280 (needs dlsm_set_hdp_prior and the F_BETA / F_HYPER fields). */
int dlsm_logp(dlsm_handle *h, double *out);
/* In-loop longitudinal Procrustes (lsm.py:495-498 -> procrustes.py:28-35): while a reference
 * configuration Xref (C,T,n,d) is set, every sweep of dlsm_run_sweeps / dlsm_run_traced rotates
 * X onto it (one rotation for all time steps) between the latent sweep and the centring.
 * NULL clears it.  dlsm_procrustes applies the rotation once (parity probe). */
int dlsm_set_procrustes_ref(dlsm_handle *h, const double *Xref);
int dlsm_procrustes(dlsm_handle *h);
/* What a stored sample consists of (the per-iteration trace writes lsm.py:475-477, :526-566 and
 * hdp_lpcm.py:823-837, :1025-1069). */
typedef struct {
    uint32_t fields_all;   /* bit f: store DLSM_F_<f> of every chain with each record */
    uint32_t fields_first; /* bit f: store DLSM_F_<f> of chain 0 only */
    int32_t thin;          /* one record every `thin` sweeps (>= 1) */
    int32_t want_logp;     /* also store the joint log-posterior of every chain */
    int32_t cooc_mode;     /* co-clustering counts (label_utils.py:40-62): 0 off, 1 chain 0, 2 all chains pooled */
    int32_t cooc_from;     /* first record of this call that is counted (burn-in) */
    int32_t reserved[2];
} dlsm_trace_spec;
/* dlsm_run_sweeps that also records the chain: after every `thin`-th sweep the traced fields (and
 * the log-posterior) are gathered into a device ring and streamed to the host on a copy stream
 * while the following sweeps run.  dst[f] (f < DLSM_F_COUNT_) receives field f as
 * (n_sweeps / thin, C or 1, <field shape per chain>), logp_dst (n_sweeps / thin, C); pageable or
 * pinned (dlsm_host_alloc) memory, valid until the call returns.  Position records of >= 1 MB going
 * to pinned memory skip the ring: they are copied from the live state right after the centring, on
 * their own stream, while the rest of the sweep still runs. */
int dlsm_run_traced(dlsm_handle *h, int32_t n_sweeps, uint32_t flags, const dlsm_trace_spec *spec,
                    void *const *dst, double *logp_dst);
/* co-clustering counts accumulated on the device by dlsm_run_traced (cooc_mode): out (T,n,n) u32 (may
 * be NULL), *n_samples = label configurations counted; reset != 0 clears the accumulator afterwards */
int dlsm_cooccurrence(dlsm_handle *h, uint32_t *out, uint64_t *n_samples, int32_t reset);
/* page-locked host memory for trace destinations */
int dlsm_host_alloc(size_t bytes, void **out);
int dlsm_host_free(void *p);

/* Edge probabilities of one chain at its current state, out (T,n,n), zero diagonal: directed
 * directed_network_probas (directed_likelihoods_fast.pyx:273-294, the estimators' probas_,
 * hdp_lpcm.py:480-492), undirected expit(beta - dist) (lsm.py:296-305). */
int dlsm_edge_probas(dlsm_handle *h, int32_t chain, double *out);

/* ---- parity probes --------------------------------------------------------------------- */
/* per-node log-likelihood at the current state, out (C,T,n):
 * partial_loglikelihood / directed_partial_loglikelihood / approx_directed_partial_loglikelihood */
int dlsm_loglik_partial(dlsm_handle *h, double *out);
/* full-network log-likelihood at the current state, out (C,):
 * network_likelihoods.py:16-33 / directed_likelihoods_fast.pyx:185-270 */
int dlsm_loglik_full(dlsm_handle *h, double *out);
/* emission densities of compute_gaussian_likelihood(normalize=False), out (C,n,T,K) */
int dlsm_gaussian_likelihood(dlsm_handle *h, double *out);
/* overwrite the read-only label-count fields (DLSM_F_NCOUNT / DLSM_F_NK) -- test probe for
 * dlsm_hdp_update, which normally consumes what dlsm_sample_labels just produced */
int dlsm_debug_set_counts(dlsm_handle *h, int field, const void *host, size_t bytes);
/* the device loop's row-sum cache, out (C,T,n): rows[c,t,j] = per-node log-likelihood of node j at the
 * current state as the loop tracks it (computed by k_rows if the cache is not current);
 * DLSM_ERR_UNSUPPORTED where the loop keeps no cache (case-control lists, CTA-per-slice kernels) */
int dlsm_debug_rowsums(dlsm_handle *h, double *out);
/* the raw draws the NEXT native latent sweep will consume: eps (C,T,n,d), logu (C,T,n) */
int dlsm_debug_draws(dlsm_handle *h, double *eps, double *logu);

/* ---- counters -------------------------------------------------------------------------- */
typedef struct {
    uint64_t kernel_launches;  /* kernels of this library launched on the handle so far */
    uint64_t node_updates;     /* latent MH node-updates executed */
    uint64_t sweeps;           /* latent sweeps executed (per chain) */
    double latent_ms;          /* device time in the latent sweep kernel (when timing enabled) */
    double other_ms;           /* device time in the other hot-path kernels */
    uint64_t ub_flags;         /* case-control lists that hit the reference's out-of-bounds quirk */
    uint64_t cluster_sweeps;   /* latent sweeps served by the thread-block-cluster kernel */
    uint64_t rowsum_sweeps;    /* latent sweeps served from the row-sum cache (proposal-only evaluation) */
} dlsm_counters;
int dlsm_enable_timing(dlsm_handle *h, int on); /* CUDA events around every phase */
int dlsm_get_counters(dlsm_handle *h, dlsm_counters *out);

#ifdef __cplusplus
}
#endif
#endif /* DLSM_H */
