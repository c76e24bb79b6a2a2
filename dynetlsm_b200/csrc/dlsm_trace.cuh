// dlsm_trace.cuh -- what the estimator loops do around the hot path once per stored sample, moved
// to the device so that a whole fit() runs without a host round trip per sweep:
//   k_logp        joint log-posterior of the current state, one value per chain
//                 (lsm.py:576-625 for the LSM, hdp_lpcm.py:1188-1280 for the HDP-LPCM); the network
//                 term comes from the tracked log-likelihood or from k_full
//   k_procrustes  in-loop longitudinal Procrustes rotation onto a reference configuration
//                 (lsm.py:495-498 -> procrustes.py:20-35 -> scipy.linalg.orthogonal_procrustes)
//   k_snapshot    gathers the traced state fields of one sample into a slot of the device trace ring
#pragma once
#include "dlsm_kernels.cuh"

namespace dlsm {

struct LogpParams {
    int C, T, n, d, K, m;      // m = number of intercepts (1 undirected, 2 directed)
    int mixture, directed;
    const double *X;           // [C][T][n][d]
    const double *intercept;   // [C][2]
    const double *ll;          // network log-likelihood of chain c at ll[c * ll_stride]
    int ll_stride;
    double tau_sq, sigma_sq, ic_prior0, ic_prior1, ic_var;
    // mixture prior only
    const double *mu, *sigma, *lambda, *weights, *beta, *hyper;
    const int32_t *z;
    dlsm_hdp_prior pr;
    double *out;               // [C]
};

__device__ __forceinline__ double clip_tiny(double v) { return v > 2.2250738585072014e-308 ? v : 2.2250738585072014e-308; }

// scipy.stats.dirichlet.logpdf(x, alpha) with the reference's clipping of non-positive entries
// (distributions.py:72-102 -> hdp_lpcm.py:1193-1203); alpha_k = scale * base[k] + (k == hot ? extra : 0)
__device__ inline double dirichlet_logpdf_row(const double *x, const double *base, double scale,
                                              int hot, double extra, double flat, int K)
{
    // the reference clips a vector only when one of its entries is <= 0, and then ALL of them
    // (np.clip on the whole array): a subnormal weight in a row without zeros keeps its value
    bool a_nonpos = false, x_nonpos = false;
    for (int k = 0; k < K; k++) {
        double a = base ? scale * base[k] : flat;
        if (k == hot) a += extra;
        a_nonpos |= !(a > 0.0);
        x_nonpos |= !(x[k] > 0.0);
    }
    double sa = 0.0, sl = 0.0, sx = 0.0;
    for (int k = 0; k < K; k++) {
        double a = base ? scale * base[k] : flat;
        if (k == hot) a += extra;
        if (a_nonpos) a = clip_tiny(a);
        const double xv = x_nonpos ? clip_tiny(x[k]) : x[k];
        sa += a;
        sl += lgamma(a);
        const double am1 = a - 1.0;
        sx += (am1 == 0.0) ? 0.0 : am1 * log(xv);
    }
    return -(sl - lgamma(sa)) + sx;
}

// log of the standard normal mass on (a, b), tail-aware
__device__ inline double log_gauss_mass(double a, double b)
{
    const double r = 0.70710678118654752440;
    double mass;
    if (a > 0.0) mass = 0.5 * (erfc(a * r) - erfc(b * r));
    else if (b < 0.0) mass = 0.5 * (erfc(-b * r) - erfc(-a * r));
    else mass = 1.0 - 0.5 * erfc(-a * r) - 0.5 * erfc(b * r);
    return log(mass);
}

__global__ void __launch_bounds__(256) k_logp(const LogpParams p)
{
    __shared__ double sh[8];
    const int c = blockIdx.x, T = p.T, n = p.n, d = p.d, K = p.K;
    const double *X = p.X + (size_t)c * T * n * d;
    double acc = 0.0;
    if (!p.mixture) {
        // lsm.py:607-614
        const double it = 0.5 / p.tau_sq, is = 0.5 / p.sigma_sq;
        for (int r = threadIdx.x; r < T * n; r += blockDim.x) {
            const double *x = X + (size_t)r * d;
            double q = 0.0;
            if (r < n) {
                for (int k = 0; k < d; k++) q += x[k] * x[k];
                acc -= q * it;
            } else {
                const double *xp = x - (size_t)n * d;
                for (int k = 0; k < d; k++) { const double df = x[k] - xp[k]; q += df * df; }
                acc -= q * is;
            }
        }
    } else {
        const double *mu = p.mu + (size_t)c * K * d, *sg = p.sigma + (size_t)c * K;
        const double *w = p.weights + (size_t)c * T * K * K, *beta = p.beta + (size_t)c * K;
        const double *hy = p.hyper + (size_t)c * 8;
        const int32_t *z = p.z + (size_t)c * T * n;
        const double lam = p.lambda[c];
        const double gam = hy[0], a_init = hy[1], alpha = hy[2], kappa = hy[3], bb = hy[5];
        // Dirichlet terms (hdp_lpcm.py:1193-1203): beta, w[0,0], w[t,k] for t >= 1
        const int rows = 2 + (T - 1) * K;
        for (int r = threadIdx.x; r < rows; r += blockDim.x) {
            if (r == 0) acc += dirichlet_logpdf_row(beta, nullptr, 0.0, -1, 0.0, gam / K, K);
            else if (r == 1) acc += dirichlet_logpdf_row(w, beta, a_init, -1, 0.0, 0.0, K);
            else {
                const int t = 1 + (r - 2) / K, k = (r - 2) % K;
                acc += dirichlet_logpdf_row(w + ((size_t)t * K + k) * K, beta, alpha, k, kappa, 0.0, K);
            }
        }
        // label chains, AR(1) mixture density of the positions and the sigma[z] prior term
        // (hdp_lpcm.py:1208-1212, :1227-1243, :1251-1253; the last one runs over sigma[z], i.e.
        // once per (t, i))
        const double ha1 = 0.5 * p.pr.a + 1.0, hb = 0.5 * bb;
        for (int r = threadIdx.x; r < T * n; r += blockDim.x) {
            const int t = r / n;
            const int zz = z[r];
            const double *x = X + (size_t)r * d, *m = mu + (size_t)zz * d;
            const double s = sg[zz], ls = log(s);
            double q = 0.0;
            if (t == 0) {
                for (int k = 0; k < d; k++) { const double df = x[k] - m[k]; q += df * df; }
                acc += log(w[zz]);
            } else {
                const double *xp = x - (size_t)n * d;
                for (int k = 0; k < d; k++) {
                    const double df = x[k] - (1.0 - lam) * xp[k] - lam * m[k];
                    q += df * df;
                }
                acc += log(w[((size_t)t * K + z[r - n]) * K + zz]);
            }
            acc += -0.5 * ls - 0.5 * q / s;
            acc += -ha1 * ls - hb / s;
        }
        if (threadIdx.x == 0) {
            const double mvp = hy[4];
            double q = 0.0;
            for (int e = 0; e < K * d; e++) q += mu[e] * mu[e];
            acc -= 0.5 * q / mvp;                                   // :1246-1248
            const double sd = sqrt(p.pr.lambda_variance_prior);     // :1256-1262 truncnorm on (0, 1)
            const double zl = (lam - p.pr.lambda_prior) / sd;
            acc += -0.5 * zl * zl - 0.91893853320467274178 - log(sd) -
                   log_gauss_mass((0.0 - p.pr.lambda_prior) / sd, (1.0 - p.pr.lambda_prior) / sd);
            if (p.directed) acc += lgamma((double)n);               // Dirichlet(1,...,1) on the radii
            if (p.pr.resample_mvp) acc += -(0.5 * p.pr.a0 + 1.0) * log(mvp) - 0.5 * p.pr.b0 / mvp;
            if (p.pr.resample_b) acc += (p.pr.c0 - 1.0) * log(bb) - p.pr.d0 * bb;
        }
    }
    if (threadIdx.x == 0) {
        // intercept prior (lsm.py:617-623, hdp_lpcm.py:1215-1224)
        const double d0 = p.intercept[c * 2] - p.ic_prior0;
        acc -= 0.5 * d0 * d0 / p.ic_var;
        if (p.m == 2) {
            const double d1 = p.intercept[c * 2 + 1] - p.ic_prior1;
            acc -= 0.5 * d1 * d1 / p.ic_var;
        }
    }
    const double tot = block_sum(acc, sh);
    if (threadIdx.x == 0) p.out[c] = p.ll[(size_t)c * p.ll_stride] + tot;
}

// ---------------------------------------------------------------------------------------------
// In-loop Procrustes: R = argmin over orthogonal R of ||X R - X_ref||_F with X, X_ref flattened to
// (T n, d); R = U V^T from the SVD of M = X^T X_ref (scipy.linalg.orthogonal_procrustes(X, X_ref)).
// One CTA per chain: M by block reductions, a one-sided Jacobi SVD of the d x d matrix on one
// thread (d <= 8), then X <- X R in place.
// ---------------------------------------------------------------------------------------------
template <int DM>
__global__ void __launch_bounds__(256) k_procrustes(int T, int n, int d_rt, double *Xall,
                                                    const double *ref_all)
{
    __shared__ double sh[8];
    __shared__ double M[DM * DM], R[DM * DM];
    const int d = (DM == kMaxD) ? d_rt : DM;
    const int c = blockIdx.x;
    const size_t rows = (size_t)T * n;
    double *X = Xall + (size_t)c * rows * d;
    const double *ref = ref_all + (size_t)c * rows * d;
    double m[DM * DM];
#pragma unroll
    for (int e = 0; e < DM * DM; e++) m[e] = 0.0;
    for (size_t r = threadIdx.x; r < rows; r += blockDim.x) {
#pragma unroll
        for (int a = 0; a < DM; a++)
#pragma unroll
            for (int b = 0; b < DM; b++)
                if (a < d && b < d) m[a * DM + b] = fma(X[r * d + a], ref[r * d + b], m[a * DM + b]);
    }
#pragma unroll
    for (int e = 0; e < DM * DM; e++) {
        const double s = block_sum(m[e], sh);
        if (threadIdx.x == 0) M[e] = s;
    }
    if (threadIdx.x == 0) {
        double A[DM * DM], V[DM * DM];
        for (int a = 0; a < DM; a++)
            for (int b = 0; b < DM; b++) {
                A[a * DM + b] = (a < d && b < d) ? M[a * DM + b] : 0.0;
                V[a * DM + b] = (a == b) ? 1.0 : 0.0;
            }
        for (int sweep = 0; sweep < 40; sweep++) {
            double off = 0.0;
            for (int pcol = 0; pcol < d - 1; pcol++)
                for (int q = pcol + 1; q < d; q++) {
                    double al = 0.0, be = 0.0, ga = 0.0;
                    for (int i = 0; i < d; i++) {
                        al += A[i * DM + pcol] * A[i * DM + pcol];
                        be += A[i * DM + q] * A[i * DM + q];
                        ga += A[i * DM + pcol] * A[i * DM + q];
                    }
                    if (fabs(ga) <= 1e-300 || fabs(ga) <= 1e-17 * sqrt(al * be)) continue;
                    off = fmax(off, fabs(ga) / sqrt(al * be));
                    const double zeta = (be - al) / (2.0 * ga);
                    const double tt = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                    const double cs = 1.0 / sqrt(1.0 + tt * tt), sn = cs * tt;
                    for (int i = 0; i < d; i++) {
                        const double ap = A[i * DM + pcol], aq = A[i * DM + q];
                        A[i * DM + pcol] = cs * ap - sn * aq;
                        A[i * DM + q] = sn * ap + cs * aq;
                        const double vp = V[i * DM + pcol], vq = V[i * DM + q];
                        V[i * DM + pcol] = cs * vp - sn * vq;
                        V[i * DM + q] = sn * vp + cs * vq;
                    }
                }
            if (off < 1e-15) break;
        }
        // A = U S: normalise the columns (a zero column leaves that direction unrotated)
        for (int k = 0; k < d; k++) {
            double nr = 0.0;
            for (int i = 0; i < d; i++) nr += A[i * DM + k] * A[i * DM + k];
            nr = sqrt(nr);
            for (int i = 0; i < d; i++) A[i * DM + k] = nr > 0.0 ? A[i * DM + k] / nr : V[i * DM + k];
        }
        for (int i = 0; i < d; i++)
            for (int j = 0; j < d; j++) {
                double s = 0.0;
                for (int k = 0; k < d; k++) s += A[i * DM + k] * V[j * DM + k];
                R[i * DM + j] = s;
            }
    }
    __syncthreads();
    for (size_t r = threadIdx.x; r < rows; r += blockDim.x) {
        double x[DM], y[DM];
#pragma unroll
        for (int a = 0; a < DM; a++) x[a] = (a < d) ? X[r * d + a] : 0.0;
#pragma unroll
        for (int b = 0; b < DM; b++) {
            double s = 0.0;
#pragma unroll
            for (int a = 0; a < DM; a++)
                if (a < d && b < d) s = fma(x[a], R[a * DM + b], s);
            y[b] = s;
        }
#pragma unroll
        for (int b = 0; b < DM; b++)
            if (b < d) X[r * d + b] = y[b];
    }
}

// ---------------------------------------------------------------------------------------------
// One stored sample: copy every traced field into its slot of the device trace ring.
// grid = (blocks, nseg); 4-byte words (every field is fp64 or int32).
// ---------------------------------------------------------------------------------------------
constexpr int kMaxSnapSeg = 16;
struct SnapParams {
    const uint32_t *src[kMaxSnapSeg];
    uint32_t *dst[kMaxSnapSeg];
    unsigned long long words[kMaxSnapSeg];
    int nseg;
};

__global__ void __launch_bounds__(256) k_snapshot(const SnapParams p)
{
    const int s = blockIdx.y;
    const uint32_t *src = p.src[s];
    uint32_t *dst = p.dst[s];
    const unsigned long long words = p.words[s];
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (((((uintptr_t)src) | ((uintptr_t)dst)) & 15) == 0) {
        const unsigned long long quads = words >> 2;
        const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
        uint4 *d4 = reinterpret_cast<uint4 *>(dst);
        for (unsigned long long q = i; q < quads; q += stride) d4[q] = s4[q];
        for (unsigned long long q = (quads << 2) + i; q < words; q += stride) dst[q] = src[q];
    } else {
        for (; i < words; i += stride) dst[i] = src[i];
    }
}


// ---------------------------------------------------------------------------------------------
// Posterior co-clustering counts (label_utils.py:40-62 calculate_posterior_cooccurrence):
// cooc[t][i][j] += #{chains c < C_use : z[c,t,i] == z[c,t,j]} for the current labels.
// grid = (ceil(n*n/256), T), block = 256
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_cooc_accumulate(const int32_t *z /* [C][T][n] */, int C_use, int T,
                                                         int n, uint32_t *cooc /* [T][n][n] */)
{
    const int t = blockIdx.y;
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)n * n) return;
    const int i = (int)(e / n), j = (int)(e % n);
    uint32_t cnt = 0;
    for (int c = 0; c < C_use; c++) {
        const int32_t *zc = z + ((size_t)c * T + t) * n;
        cnt += (zc[i] == zc[j]) ? 1u : 0u;
    }
    cooc[(size_t)t * n * n + e] += cnt;
}

// ---------------------------------------------------------------------------------------------
// Edge probabilities at a chain's current state (K9 directed_network_probas,
// directed_likelihoods_fast.pyx:273-294; undirected: expit(beta - dist), lsm.py:296-305), zero diagonal.
// out (T, n, n);  grid = (ceil(n*n/256), T), block = 256
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_edge_probas(const double *X /* [T][n][d] */, const double *intercept,
                                                     const double *rinv /* [n] or null */, int n, int d,
                                                     int directed, double *out)
{
    const int t = blockIdx.y;
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)n * n) return;
    const int i = (int)(e / n), j = (int)(e % n);
    double pr = 0.0;
    if (i != j) {
        const double *xi = X + ((size_t)t * n + i) * d, *xj = X + ((size_t)t * n + j) * d;
        double s = 0.0;
        for (int k = 0; k < d; k++) { const double df = xi[k] - xj[k]; s += df * df; }
        const double dist = sqrt(s);
        const double eta = directed ? intercept[0] * (1.0 - dist * rinv[j]) + intercept[1] * (1.0 - dist * rinv[i])
                                    : intercept[0] - dist;
        pr = 1.0 / (1.0 + exp(-eta));
    }
    out[(size_t)t * n * n + e] = pr;
}

// ---------------------------------------------------------------------------------------------
// Case-control sets on the device (DirectedCaseControlSampler.sample,
// case_control_likelihood.py:75-112): for every (set, t, i) and both directions, min(n - deg - 1,
// n_control) DISTINCT nodes drawn uniformly from the non-neighbours of i (i itself excluded), the
// rest of the row padded with -1.  One warp per (set, t, i, direction): rounds of 32 Philox
// candidates, each accepted in lane order unless it is i, a neighbour, already in the list or a
// duplicate of a lower lane -- sequential rejection sampling, hence a uniform random subset in
// uniform random order.  The reference consumes numpy's rng.choice; this is the device-RNG
// counterpart (same distribution, not the same draws).
// grid = ceil(sets * T * n * 2 / warps per CTA), block = 128
// ---------------------------------------------------------------------------------------------
constexpr uint32_t kRngControls = 5;

struct ControlParams {
    int sets, T, n, n_control, max_in, max_out;
    const int32_t *deg;        // [T][n][2] in, out
    const int32_t *in_edges;   // [T][n][max_in]
    const int32_t *out_edges;  // [T][n][max_out]
    int32_t *ctrl_in, *ctrl_out; // [sets][T][n][n_control]
    uint64_t seed;
    uint32_t sweep, chain_offset;
};

__global__ void __launch_bounds__(128) k_resample_controls(const ControlParams p)
{
    const int lane = threadIdx.x & 31;
    const size_t wid = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t total = (size_t)p.sets * p.T * p.n * 2;
    if (wid >= total) return;
    const int dir = (int)(wid & 1);                 // 0: in-controls, 1: out-controls
    const size_t row = wid >> 1;                    // (set, t, i)
    const int i = (int)(row % p.n);
    const size_t ti = row % ((size_t)p.T * p.n);
    const int set = (int)(row / ((size_t)p.T * p.n));
    const int deg = p.deg[ti * 2 + dir];
    const int32_t *nb = dir == 0 ? p.in_edges + ti * p.max_in : p.out_edges + ti * p.max_out;
    int32_t *out = (dir == 0 ? p.ctrl_in : p.ctrl_out) + row * p.n_control;
    int want = p.n - deg - 1;
    want = want < p.n_control ? want : p.n_control;
    want = want > 0 ? want : 0;
    int have = 0;
    for (uint32_t round = 0; have < want && round < (1u << 20); round++) {
        const U2 u = philox_u2(p.seed, (uint32_t)(ti * 2 + dir), p.sweep, (uint32_t)set + p.chain_offset,
                               kRngControls, round * 16 + (lane >> 1));
        int c = (int)(((lane & 1) ? u.b : u.a) * p.n);
        c = c < p.n ? c : p.n - 1;
        bool ok = c != i;
        for (int q = 0; ok && q < deg; q++) ok = nb[q] != c;          // neighbour lists are short
        for (int q = 0; ok && q < have; q++) ok = out[q] != c;        // already drawn
        const unsigned same = __match_any_sync(kFull, ok ? c : -1 - lane); // duplicates in the batch
        ok = ok && lane == __ffs(same) - 1;
        const unsigned acc = __ballot_sync(kFull, ok);
        const int pos = have + __popc(acc & ((1u << lane) - 1u));
        if (ok && pos < want) out[pos] = c;
        have += __popc(acc);
        have = have < want ? have : want;
        __syncwarp();
    }
    for (int q = want + lane; q < p.n_control; q += 32) out[q] = -1;
}

} // namespace dlsm
