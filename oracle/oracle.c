/*
 * oracle.c -- CPU restatement (plain C, fp64) of dynetlsm's blocked Metropolis-Hastings-within-Gibbs
 * hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the *checker* for the CUDA product path in
 * dynetlsm_b200/csrc/.  It is imported/linked only by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs.  The product package never calls it and has
 * no CPU fallback.
 *
 * Parity status: PINNED.  The reference's own tests assert shapes only, so
 * this restatement is pinned against outputs of the reference itself, run in the build container
 * from /root/reference with fixed seeds (oracle/make_golden.py -> tests/golden/ *.npz; checked by
 * tests/test_oracle_golden.py) and, where oracle/_ref holds the compiled reference Cython kernels,
 * directly against those (tests/test_oracle_golden.py, tests/test_oracle_golden_big.py, tests/test_ref_driver.py).
 *
 * Every function cites the reference file:line it follows (paths relative to the reference root).
 * Arithmetic is IEEE fp64 with NO fused multiply-add contraction (build with -ffp-contract=off):
 * the reference's Cython is compiled for baseline x86-64 (no FMA) and its numpy expressions round
 * every elementary operation separately.
 *
 * Build:  make -C oracle        (gcc -O2 -ffp-contract=off -fPIC -shared)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * numpy's pairwise summation (numpy/_core/src/umath/loops_utils.h.src, `pairwise_sum`), used by
 * every `np.sum` over a contiguous 1-D fp64 array on the reference's path
 * (network_likelihoods.py:33, sample_labels.py:169, sample_latent_positions.py:133-140).
 * Third-party arithmetic (numpy, pinned 1.18.1 in the reference CI, 2.3.5 here; the algorithm is
 * unchanged between them): blocks of <=128 summed with 8 interleaved accumulators, larger inputs
 * split recursively at a multiple of 8.
 * ------------------------------------------------------------------------------------------ */
static double np_pairwise_sum(const double *a, int64_t n)
{
    if (n < 8) {
        double res = -0.0;
        for (int64_t i = 0; i < n; i++) res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8];
        int64_t i;
        for (int k = 0; k < 8; k++) r[k] = a[k];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int k = 0; k < 8; k++) r[k] += a[i + k];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    } else {
        int64_t n2 = n / 2;
        n2 -= n2 % 8;
        return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
    }
}

ORC_API double orc_np_sum(const double *a, int64_t n) { return np_pairwise_sum(a, n); }

/* 0.5 * np.sum(v*v) / s   -- the Gaussian prior kernels of sample_latent_positions.py:131-140 */
static double half_sumsq_over(const double *v, int d, double s)
{
    double sq[64];
    for (int k = 0; k < d; k++) sq[k] = v[k] * v[k];
    return 0.5 * np_pairwise_sum(sq, d) / s;
}

/* ------------------------------------------------------------------------------------------
 * K1  static_network_fast.pyx:17-44  partial_loglikelihood
 * ------------------------------------------------------------------------------------------ */
ORC_API double orc_partial_loglikelihood(const double *Y, const double *X, int n, int d,
                                         double intercept, int node, int squared)
{
    double loglik = 0;
    for (int i = 0; i < n; i++) {
        if (i == node) continue;
        double dist = 0;
        for (int k = 0; k < d; k++) {
            double df = X[(size_t)i * d + k] - X[(size_t)node * d + k];
            dist += df * df;
        }
        double eta = squared ? intercept - dist : intercept - sqrt(dist);
        loglik += Y[(size_t)node * n + i] * eta;
        loglik -= log(1 + exp(eta));
    }
    return loglik;
}

/* ------------------------------------------------------------------------------------------
 * K2  directed_likelihoods_fast.pyx:46-80  directed_partial_loglikelihood
 * ------------------------------------------------------------------------------------------ */
ORC_API double orc_directed_partial_loglikelihood(const double *Y, const double *X,
                                                  const double *radii, int n, int d, double b_in,
                                                  double b_out, int node, int squared)
{
    double loglik = 0;
    for (int j = 0; j < n; j++) {
        if (j == node) continue;
        double dist = 0;
        for (int k = 0; k < d; k++) {
            double df = X[(size_t)j * d + k] - X[(size_t)node * d + k];
            dist += df * df;
        }
        if (!squared) dist = sqrt(dist);
        double eta = b_in * (1 - dist / radii[j]);
        eta += b_out * (1 - dist / radii[node]);
        loglik += Y[(size_t)node * n + j] * eta - log(1 + exp(eta));
        eta = b_in * (1 - dist / radii[node]);
        eta += b_out * (1 - dist / radii[j]);
        loglik += Y[(size_t)j * n + node] * eta - log(1 + exp(eta));
    }
    return loglik;
}

/* ------------------------------------------------------------------------------------------
 * K3  directed_likelihoods_fast.pyx:83-182  approx_directed_partial_loglikelihood
 * (case-control estimator).  Quirks kept on purpose:
 *   - the SECOND control loop tests the `in` sentinel (:161) but indexes `control_nodes_out` (:167);
 *   - control_adj divides by the number of non-sentinel controls (0 -> inf*0 = NaN, as in C).
 * A sentinel (-1) reached through the out list in the second loop is undefined behaviour in the
 * reference (negative index with wraparound=False); here it terminates the loop and sets *ub=1.
 * ------------------------------------------------------------------------------------------ */
static double pair_dist(const double *X, int d, int64_t a, int64_t b, int squared)
{
    double dist = 0;
    for (int k = 0; k < d; k++) {
        double df = X[(size_t)a * d + k] - X[(size_t)b * d + k];
        dist += df * df;
    }
    return squared ? dist : sqrt(dist);
}

ORC_API double orc_approx_directed_partial_loglikelihood(
    const double *X, const double *radii, const int64_t *in_edges, int max_in,
    const int64_t *out_edges, int max_out, const int64_t *degree, const int64_t *ctrl_in,
    const int64_t *ctrl_out, int n_control, int n, int d, double b_in, double b_out, int node,
    int squared, int *ub)
{
    int in_degree = (int)degree[(size_t)node * 2 + 0];
    int out_degree = (int)degree[(size_t)node * 2 + 1];
    double loglik = 0, control = 0, n_control_edges = 0, control_adj, eta, dist;
    if (ub) *ub = 0;

    for (int j = 0; j < in_degree; j++) { /* :108-119 */
        int64_t k = in_edges[(size_t)node * max_in + j];
        dist = pair_dist(X, d, k, node, squared);
        eta = b_in * (1 - dist / radii[node]);
        eta += b_out * (1 - dist / radii[k]);
        loglik += eta - log(1 + exp(eta));
    }
    for (int j = 0; j < out_degree; j++) { /* :122-133 */
        int64_t k = out_edges[(size_t)node * max_out + j];
        dist = pair_dist(X, d, k, node, squared);
        eta = b_in * (1 - dist / radii[k]);
        eta += b_out * (1 - dist / radii[node]);
        loglik += eta - log(1 + exp(eta));
    }
    for (int j = 0; j < n_control; j++) { /* :136-152 */
        int64_t k = ctrl_in[(size_t)node * n_control + j];
        if (k == -1) break;
        dist = pair_dist(X, d, k, node, squared);
        eta = b_in * (1 - dist / radii[node]);
        eta += b_out * (1 - dist / radii[k]);
        control += log(1 + exp(eta));
        n_control_edges += 1;
    }
    control_adj = (double)(n - in_degree - 1) / n_control_edges; /* :155 */
    loglik -= control_adj * control;

    control = 0;
    n_control_edges = 0;
    for (int j = 0; j < n_control; j++) { /* :160-176 */
        if (ctrl_in[(size_t)node * n_control + j] == -1) break; /* sic: tests the IN list */
        int64_t k = ctrl_out[(size_t)node * n_control + j];
        if (k < 0) { if (ub) *ub = 1; break; }
        dist = pair_dist(X, d, k, node, squared);
        eta = b_in * (1 - dist / radii[k]);
        eta += b_out * (1 - dist / radii[node]);
        control += log(1 + exp(eta));
        n_control_edges += 1;
    }
    control_adj = (double)(n - out_degree - 1) / n_control_edges; /* :179 */
    loglik -= control_adj * control;
    return loglik;
}

/* ------------------------------------------------------------------------------------------
 * latent_space.py:19-33 calculate_distances -> sklearn.metrics.euclidean_distances (third-party,
 * scikit-learn unpinned in requirements.txt:4; 1.9.0 here).  Published algorithm, fp64 input:
 *   XX = einsum('ij,ij->i'); D = -2 X X^T (BLAS dgemm); D += XX[:,None]; D += XX[None,:];
 *   D = max(D, 0); fill_diagonal(D, 0); sqrt(D).
 * The BLAS dot's rounding (FMA use / order) cannot be restated exactly; parity for everything
 * downstream of `dist` is therefore by tolerance (1e-10 relative), not bitwise.
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_calculate_distances(const double *X, int T, int n, int d, int squared,
                                     double *dist)
{
    double *xx = (double *)malloc(sizeof(double) * (size_t)n);
    for (int t = 0; t < T; t++) {
        const double *Xt = X + (size_t)t * n * d;
        double *D = dist + (size_t)t * n * n;
        for (int i = 0; i < n; i++) {
            double s = 0;
            for (int k = 0; k < d; k++) s += Xt[(size_t)i * d + k] * Xt[(size_t)i * d + k];
            xx[i] = s;
        }
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) {
                double dot = 0;
                for (int k = 0; k < d; k++) dot += Xt[(size_t)i * d + k] * Xt[(size_t)j * d + k];
                double v = -2 * dot;
                v += xx[i];
                v += xx[j];
                if (v < 0) v = 0;
                if (i == j) v = 0;
                D[(size_t)i * n + j] = squared ? v : sqrt(v);
            }
    }
    free(xx);
}

/* ------------------------------------------------------------------------------------------
 * K4  directed_likelihoods_fast.pyx:185-205  directed_network_loglikelihood_fast
 * ------------------------------------------------------------------------------------------ */
ORC_API double orc_directed_network_loglikelihood(const double *Y, const double *dist,
                                                  const double *radii, int T, int n, double b_in,
                                                  double b_out)
{
    double loglik = 0;
    for (int t = 0; t < T; t++)
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) {
                if (i == j) continue;
                size_t o = ((size_t)t * n + i) * n + j;
                double d_in = (1 - dist[o] / radii[j]);
                double d_out = (1 - dist[o] / radii[i]);
                double eta = b_in * d_in + b_out * d_out;
                loglik += Y[o] * eta - log(1 + exp(eta));
            }
    return loglik;
}

/* ------------------------------------------------------------------------------------------
 * K5  network_likelihoods.py:26-33  dynamic_network_loglikelihood_undirected
 *     + array_utils.py:4-8 triu_indices_from_3d(k=1): terms in (t, i, j>i) C order, reduced by
 *     np.sum (pairwise).
 * ------------------------------------------------------------------------------------------ */
ORC_API double orc_undirected_network_loglikelihood(const double *Y, const double *dist, int T,
                                                    int n, double intercept)
{
    size_t m = (size_t)T * n * (n - 1) / 2, c = 0;
    double *terms = (double *)malloc(sizeof(double) * (m ? m : 1));
    for (int t = 0; t < T; t++)
        for (int i = 0; i < n; i++)
            for (int j = i + 1; j < n; j++) {
                size_t o = ((size_t)t * n + i) * n + j;
                double eta = intercept - dist[o];
                terms[c++] = Y[o] * eta - log(1 + exp(eta));
            }
    double s = np_pairwise_sum(terms, (int64_t)m);
    free(terms);
    return s;
}

/* ------------------------------------------------------------------------------------------
 * K6  directed_likelihoods_fast.pyx:208-270  approx_directed_network_loglikelihood
 * (out-edges and out-controls only; distances recomputed from X).
 * ------------------------------------------------------------------------------------------ */
ORC_API double orc_approx_directed_network_loglikelihood(
    const double *X, const double *radii, const int64_t *out_edges, int max_out,
    const int64_t *degree, const int64_t *ctrl, int n_control, int T, int n, int d, double b_in,
    double b_out, int squared)
{
    double loglik = 0;
    for (int t = 0; t < T; t++) {
        const double *Xt = X + (size_t)t * n * d;
        for (int i = 0; i < n; i++) {
            int out_degree = (int)degree[((size_t)t * n + i) * 2 + 1];
            for (int j = 0; j < out_degree; j++) {
                int64_t k = out_edges[((size_t)t * n + i) * max_out + j];
                double dist = pair_dist(Xt, d, k, i, squared);
                double eta = b_in * (1 - dist / radii[k]);
                eta += b_out * (1 - dist / radii[i]);
                loglik += eta - log(1 + exp(eta));
            }
            double control = 0, n_control_edges = 0;
            for (int j = 0; j < n_control; j++) {
                int64_t k = ctrl[((size_t)t * n + i) * n_control + j];
                if (k == -1) break;
                double dist = pair_dist(Xt, d, k, i, squared);
                double eta = b_in * (1 - dist / radii[k]);
                eta += b_out * (1 - dist / radii[i]);
                control += log(1 + exp(eta));
                n_control_edges += 1;
            }
            double control_adj = (double)(n - out_degree - 1) / n_control_edges;
            loglik -= control_adj * control;
        }
    }
    return loglik;
}

/* ------------------------------------------------------------------------------------------
 * K7  gaussian_likelihood_fast.pyx:17-54  spherical_normal_log_pdf / compute_gaussian_likelihood
 * X is one node's trajectory (T, d); out is (T, K) = exp(loglik) (normalize: subtract row max).
 * ------------------------------------------------------------------------------------------ */
static double spherical_normal_log_pdf(const double *x, const double *mean, double var, int d)
{
    double sum_sq = 0.0;
    for (int k = 0; k < d; k++) {
        double df = x[k] - mean[k];
        sum_sq += df * df;
    }
    sum_sq *= 0.5 * (1. / var);
    return -0.5 * d * log(2 * M_PI * var) - sum_sq;
}

ORC_API void orc_compute_gaussian_likelihood(const double *X, const double *mu,
                                             const double *sigma, double lmbda, int T, int K,
                                             int d, int normalize, double *out)
{
    double muk[64];
    for (int t = 0; t < T; t++) {
        for (int k = 0; k < K; k++) {
            if (t == 0) {
                out[(size_t)t * K + k] = spherical_normal_log_pdf(X + (size_t)t * d,
                                                                  mu + (size_t)k * d, sigma[k], d);
            } else {
                for (int j = 0; j < d; j++)
                    muk[j] = lmbda * mu[(size_t)k * d + j] + (1 - lmbda) * X[(size_t)(t - 1) * d + j];
                out[(size_t)t * K + k] = spherical_normal_log_pdf(X + (size_t)t * d, muk, sigma[k], d);
            }
        }
        if (normalize) {
            double mx = out[(size_t)t * K];
            for (int k = 1; k < K; k++) if (out[(size_t)t * K + k] > mx) mx = out[(size_t)t * K + k];
            for (int k = 0; k < K; k++) out[(size_t)t * K + k] -= mx;
        }
    }
    for (size_t i = 0; i < (size_t)T * K; i++) out[i] = exp(out[i]);
}

/* ------------------------------------------------------------------------------------------
 * metropolis.py:5-37 tuning tables, :85-136 Metropolis state machine.
 * Per-sampler state: step_size, n_accepted, n_steps, steps_until_tune.  tune < 0 means None.
 * ------------------------------------------------------------------------------------------ */
static double tune_random_walk(double s, double r)
{
    if (r < 0.001) s *= 0.1;
    else if (r < 0.05) s *= 0.5;
    else if (r < 0.25) s *= 0.9;
    else if (r > 0.95) s *= 10.0;
    else if (r > 0.75) s *= 2.0;
    else if (r > 0.4) s *= 1.1;
    return s;
}

static double tune_dirichlet(double s, double r)
{
    if (r < 0.001) s *= 10.0;
    else if (r < 0.05) s *= 2;
    else if (r < 0.25) s *= 1.1;
    else if (r > 0.95) s *= 0.1;
    else if (r > 0.75) s *= 0.5;
    else if (r > 0.4) s *= 0.9;
    return s;
}

ORC_API void orc_metropolis_bookkeep(double *step, int32_t *n_accepted, int32_t *n_steps,
                                     int32_t *until, int tune, int tune_interval, int accepted,
                                     int is_dirichlet)
{
    *n_accepted += accepted; /* :113 */
    *n_steps += 1;           /* :114 */
    if (tune < 0) return;    /* :117 tune is None */
    if (*n_steps < tune && *until == 0) { /* :123 */
        double rate = (double)*n_accepted / (double)tune_interval;
        *step = is_dirichlet ? tune_dirichlet(*step, rate) : tune_random_walk(*step, rate);
        *n_accepted = 0;
        *until = tune_interval;
    } else {
        *until -= 1;
    }
}

/* ------------------------------------------------------------------------------------------
 * a1/a2  sample_latent_positions.py:92-146 (LSM prior) and :149-206 (mixture prior), with
 * metropolis.py:40-54 random_walk_metropolis driven by RECORDED raw draws (replay):
 * eps (T,n,d) standard normals, logu (T,n) = np.log(rand()).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int32_t T, n, d, K;
    int32_t is_directed, use_cc, prior_kind, squared; /* prior_kind 0 = LSM, 1 = mixture */
    int32_t max_in, max_out, n_control, tune, tune_interval, pad0;
    double tau_sq, sigma_sq, lmbda;
    const double *Y;         /* (T,n,n) or NULL when use_cc */
    double *X;               /* (T,n,d) in/out */
    const double *intercept; /* (1,) or (2,) */
    const double *radii;     /* (n,) or NULL */
    const int64_t *in_edges, *out_edges, *degrees, *ctrl_in, *ctrl_out;
    const double *mu;        /* (K,d) */
    const double *sigma;     /* (K,) variances */
    const int64_t *z;        /* (T,n) */
    double *step;            /* (T,n) tuner state, in/out */
    int32_t *n_accepted, *n_steps, *until;
    const double *eps;       /* (T,n,d) */
    const double *logu;      /* (T,n) */
    int32_t *accepted;       /* (T,n) out */
    double *ratio;           /* (T,n) out: logp(x) - logp(x0) */
    double *logp_new;        /* (T,n) out (may be NULL) */
    double *logp_old;        /* (T,n) out (may be NULL) */
} orc_sweep_args;

static double node_loglik(const orc_sweep_args *a, int t, int j)
{
    const int n = a->n, d = a->d;
    const double *Xt = a->X + (size_t)t * n * d;
    if (a->is_directed) {
        if (a->use_cc) {
            size_t o = (size_t)t * n;
            return orc_approx_directed_partial_loglikelihood(
                Xt, a->radii, a->in_edges + o * a->max_in, a->max_in,
                a->out_edges + o * a->max_out, a->max_out, a->degrees + o * 2,
                a->ctrl_in + o * a->n_control, a->ctrl_out + o * a->n_control, a->n_control, n, d,
                a->intercept[0], a->intercept[1], j, a->squared, NULL);
        }
        return orc_directed_partial_loglikelihood(a->Y + (size_t)t * n * n, Xt, a->radii, n, d,
                                                  a->intercept[0], a->intercept[1], j,
                                                  a->squared);
    }
    return orc_partial_loglikelihood(a->Y + (size_t)t * n * n, Xt, n, d, a->intercept[0], j,
                                     a->squared);
}

/* closure `logp` of sample_latent_positions.py:100-142 / :158-201 evaluated at the value
 * currently stored in X[t,j] */
static double node_logp(const orc_sweep_args *a, int t, int j)
{
    const int n = a->n, d = a->d, T = a->T;
    const double *x = a->X + ((size_t)t * n + j) * d;
    double loglik = node_loglik(a, t, j);
    double diff[64];
    if (a->prior_kind == 0) {
        if (t == 0) { /* :131-132 */
            loglik -= half_sumsq_over(x, d, a->tau_sq);
        } else { /* :134-135 */
            const double *xp = a->X + ((size_t)(t - 1) * n + j) * d;
            for (int k = 0; k < d; k++) diff[k] = x[k] - xp[k];
            loglik -= half_sumsq_over(diff, d, a->sigma_sq);
        }
        if (t < T - 1) { /* :137-139 */
            const double *xn = a->X + ((size_t)(t + 1) * n + j) * d;
            for (int k = 0; k < d; k++) diff[k] = xn[k] - x[k];
            loglik -= half_sumsq_over(diff, d, a->sigma_sq);
        }
    } else {
        const double lm = a->lmbda, oml = 1 - lm;
        int64_t zc = a->z[(size_t)t * n + j];
        const double *m = a->mu + (size_t)zc * d;
        if (t == 0) { /* :188-190 */
            for (int k = 0; k < d; k++) diff[k] = x[k] - m[k];
        } else { /* :192-193 */
            const double *xp = a->X + ((size_t)(t - 1) * n + j) * d;
            for (int k = 0; k < d; k++) diff[k] = (x[k] - oml * xp[k]) - lm * m[k];
        }
        loglik -= half_sumsq_over(diff, d, a->sigma[zc]);
        if (t < T - 1) { /* :196-199 */
            int64_t zn = a->z[(size_t)(t + 1) * n + j];
            const double *mn = a->mu + (size_t)zn * d;
            const double *xn = a->X + ((size_t)(t + 1) * n + j) * d;
            for (int k = 0; k < d; k++) diff[k] = (xn[k] - oml * x[k]) - lm * mn[k];
            loglik -= half_sumsq_over(diff, d, a->sigma[zn]);
        }
    }
    return loglik;
}

ORC_API int orc_sweep_latent(orc_sweep_args *a)
{
    const int T = a->T, n = a->n, d = a->d;
    if (d > 64) return -1;
    double x0[64], x[64];
    for (int t = 0; t < T; t++)
        for (int j = 0; j < n; j++) {
            size_t s = (size_t)t * n + j;
            double *xs = a->X + s * d;
            for (int k = 0; k < d; k++) {
                x0[k] = xs[k];
                x[k] = x0[k] + a->step[s] * a->eps[s * d + k]; /* metropolis.py:44 */
            }
            memcpy(xs, x, sizeof(double) * d);
            double lp_new = node_logp(a, t, j); /* :47 logp(x) first ... */
            memcpy(xs, x0, sizeof(double) * d);
            double lp_old = node_logp(a, t, j); /* ... then logp(x0) */
            double r = lp_new - lp_old;
            int acc = 1;
            if (a->logu[s] >= r) acc = 0; /* :50 */
            if (acc) memcpy(xs, x, sizeof(double) * d);
            a->accepted[s] = acc;
            a->ratio[s] = r;
            if (a->logp_new) a->logp_new[s] = lp_new;
            if (a->logp_old) a->logp_old[s] = lp_old;
            orc_metropolis_bookkeep(&a->step[s], &a->n_accepted[s], &a->n_steps[s], &a->until[s],
                                    a->tune, a->tune_interval, acc, 0);
        }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * lsm.py:501 / hdp_lpcm.py:852   X -= np.mean(X, axis=(0, 1))
 * numpy reduces the two leading axes of a C-contiguous (T,n,d) array by a plain sequential
 * accumulation per column (pairwise summation only applies along the contiguous inner axis),
 * then divides by T*n.
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_center(double *X, int T, int n, int d)
{
    for (int k = 0; k < d; k++) {
        double s = 0.0;
        for (size_t r = 0; r < (size_t)T * n; r++) s += X[r * d + k];
        double m = s / (double)((size_t)T * n);
        for (size_t r = 0; r < (size_t)T * n; r++) X[r * d + k] -= m;
    }
}

/* ------------------------------------------------------------------------------------------
 * a10  sample_coefficients.py:12-88 sample_intercepts  (random-walk MH on the full-network
 * log-likelihood with a Gaussian prior), replay-driven: eps[i], logu[i] per intercept.
 * `dist` is the (T,n,n) distance cache (lsm.py:504-505).  With use_cc the case-control
 * full-network estimator K6 is used and dist is ignored (lsm.py:504 passes dist=None).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int32_t T, n, d, is_directed, use_cc, max_out, n_control, tune;
    int32_t tune_interval[2];
    double prior_mean[2];
    double prior_var;
    const double *Y, *X, *dist, *radii;
    const int64_t *out_edges, *degrees, *ctrl_out;
    double *intercept;   /* (1,) or (2,) in/out */
    double *step;        /* (2,) tuner state */
    int32_t *n_accepted, *n_steps, *until;
    const double *eps;   /* (1,) or (2,) */
    const double *logu;  /* (1,) or (2,) */
    int32_t *accepted;   /* out */
    double *ratio;       /* out */
    double *loglik_new, *loglik_old; /* out: network log-likelihood part of logp(x), logp(x0) */
} orc_intercept_args;

static double full_loglik(const orc_intercept_args *a, const double *radii, double b0, double b1)
{
    if (a->is_directed) {
        if (a->use_cc)
            return orc_approx_directed_network_loglikelihood(a->X, radii, a->out_edges, a->max_out,
                                                             a->degrees, a->ctrl_out, a->n_control,
                                                             a->T, a->n, a->d, b0, b1, 0);
        return orc_directed_network_loglikelihood(a->Y, a->dist, radii, a->T, a->n, b0, b1);
    }
    return orc_undirected_network_loglikelihood(a->Y, a->dist, a->T, a->n, b0);
}

ORC_API int orc_sample_intercepts(orc_intercept_args *a)
{
    int m = a->is_directed ? 2 : 1;
    for (int i = 0; i < m; i++) {
        double x0 = a->intercept[i];
        double x = x0 + a->step[i] * a->eps[i];
        double b[2] = {a->intercept[0], m == 2 ? a->intercept[1] : 0.0};
        b[i] = x;
        double ll_new = full_loglik(a, a->radii, b[0], b[1]);
        double df = x - a->prior_mean[i];
        double lp_new = ll_new - (df * df) / (2 * a->prior_var);
        b[i] = x0;
        double ll_old = full_loglik(a, a->radii, b[0], b[1]);
        df = x0 - a->prior_mean[i];
        double lp_old = ll_old - (df * df) / (2 * a->prior_var);
        double r = lp_new - lp_old;
        int acc = 1;
        if (a->logu[i] >= r) acc = 0;
        if (acc) a->intercept[i] = x;
        a->accepted[i] = acc;
        a->ratio[i] = r;
        if (a->loglik_new) a->loglik_new[i] = ll_new;
        if (a->loglik_old) a->loglik_old[i] = ll_old;
        orc_metropolis_bookkeep(&a->step[i], &a->n_accepted[i], &a->n_steps[i], &a->until[i],
                                a->tune, a->tune_interval[i], acc, 0);
    }
    return 0;
}

/* scipy.stats.dirichlet.logpdf (third-party, scipy 1.4.1 pinned in CI / 1.18.1 here):
 *   -[sum(gammaln(alpha)) - gammaln(sum(alpha))] + sum(xlogy(alpha - 1, x)) */
ORC_API double orc_dirichlet_logpdf(const double *x, const double *alpha, int n)
{
    double sa = 0, sl = 0, sx = 0;
    for (int i = 0; i < n; i++) {
        sa += alpha[i];
        sl += lgamma(alpha[i]);
        double c = alpha[i] - 1;
        sx += (c == 0.0) ? 0.0 : c * log(x[i]);
    }
    return -(sl - lgamma(sa)) + sx;
}

/* ------------------------------------------------------------------------------------------
 * a10  sample_coefficients.py:91-121 sample_radii + metropolis.py:57-82 dirichlet_metropolis,
 * replay-driven: `proposal` is the recorded Dirichlet draw AFTER the zero-guard (:65-67).
 * ------------------------------------------------------------------------------------------ */
ORC_API int orc_sample_radii(orc_intercept_args *a, double *radii /* in/out (n,) */,
                             const double *proposal, double logu, double *step_size,
                             int32_t *n_accepted, int32_t *n_steps, int32_t *until, int tune,
                             int tune_interval, int32_t *accepted, double *ratio)
{
    int n = a->n;
    double *al = (double *)malloc(sizeof(double) * n * 2);
    double s = *step_size;
    double r = full_loglik(a, proposal, a->intercept[0], a->intercept[1]) -
               full_loglik(a, radii, a->intercept[0], a->intercept[1]);
    for (int i = 0; i < n; i++) {
        al[i] = s * proposal[i];
        al[n + i] = s * radii[i];
    }
    r += orc_dirichlet_logpdf(radii, al, n) - orc_dirichlet_logpdf(proposal, al + n, n);
    int acc = 1;
    if (logu >= r) acc = 0;
    if (acc) memcpy(radii, proposal, sizeof(double) * n);
    *accepted = acc;
    *ratio = r;
    orc_metropolis_bookkeep(step_size, n_accepted, n_steps, until, tune, tune_interval, acc, 1);
    free(al);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * a12  sample_labels.py:134-190 sample_labels_block (+ :16-19 sample_categorical), replay-driven:
 * U (n,T) holds the raw uniform of each categorical draw in node-major order;
 * u = cdf[K-1] * U reproduces legacy RandomState.uniform(0, cdf[-1]) bitwise.
 * Kept quirk: bwds_msg is allocated once outside the node loop, so bwds_msg[T-1] == 1.
 * np.dot(w[t], pm) is BLAS (order unspecified): restated as an ascending-k dot product.
 * Outputs: z (T,n) int64, ncount (T,K,K) fp64 (ncount[0,0,k] = initial-state counts),
 * nk (T,K) int64.  (resp is the one-hot of z and is rebuilt by the caller.)
 * ------------------------------------------------------------------------------------------ */
ORC_API int orc_sample_labels_block(const double *X, const double *mu, const double *sigma,
                                    double lmbda, const double *w, const double *U, int T, int n,
                                    int d, int K, int64_t *z, double *ncount, int64_t *nk,
                                    double *probas_out /* (n,T,K) or NULL */)
{
    double *lik = (double *)malloc(sizeof(double) * T * K);
    double *bwd = (double *)malloc(sizeof(double) * T * K);
    double *pm = (double *)malloc(sizeof(double) * T * K);
    double *xi = (double *)malloc(sizeof(double) * T * d);
    double *p = (double *)malloc(sizeof(double) * K);
    for (int i = 0; i < T * K; i++) { bwd[i] = 1.0; pm[i] = 0.0; }
    memset(ncount, 0, sizeof(double) * T * K * K);
    memset(nk, 0, sizeof(int64_t) * T * K);
    for (int i = 0; i < n; i++) {
        for (int t = 0; t < T; t++)
            memcpy(xi + (size_t)t * d, X + ((size_t)t * n + i) * d, sizeof(double) * d);
        orc_compute_gaussian_likelihood(xi, mu, sigma, lmbda, T, K, d, 0, lik); /* :159-160 */
        for (int t = T - 1; t > 0; t--) { /* :164-169 */
            for (int k = 0; k < K; k++) pm[t * K + k] = lik[t * K + k] * bwd[t * K + k];
            for (int j = 0; j < K; j++) {
                double s = 0;
                for (int k = 0; k < K; k++) s += w[((size_t)t * K + j) * K + k] * pm[t * K + k];
                bwd[(t - 1) * K + j] = s;
            }
            double tot = np_pairwise_sum(bwd + (t - 1) * K, K);
            for (int j = 0; j < K; j++) bwd[(t - 1) * K + j] /= tot;
        }
        for (int k = 0; k < K; k++) pm[k] = lik[k] * bwd[k]; /* :170 */
        int64_t zp = 0;
        for (int t = 0; t < T; t++) { /* :173-188 */
            const double *wr = (t == 0) ? w : w + ((size_t)t * K + zp) * K;
            double c = 0;
            for (int k = 0; k < K; k++) {
                p[k] = wr[k] * pm[t * K + k];
                if (probas_out) probas_out[((size_t)i * T + t) * K + k] = p[k];
                c += p[k];
                p[k] = c; /* cumsum */
            }
            double u = 0.0 + (p[K - 1] - 0.0) * U[(size_t)i * T + t];
            int64_t zz = 0;
            for (int k = 0; k < K; k++) zz += (u > p[k]);
            if (zz >= K) zz = K - 1; /* unreachable for finite cdf; guards the count arrays */
            z[(size_t)t * n + i] = zz;
            if (t == 0) ncount[zz] += 1;
            else ncount[((size_t)t * K + zp) * K + zz] += 1;
            nk[(size_t)t * K + zz] += 1;
            zp = zz;
        }
    }
    free(lik); free(bwd); free(pm); free(xi); free(p);
    return 0;
}
