// dlsm_hdp.cuh -- the conjugate / auxiliary-variable block of one HDP-LPCM sweep on the device
// (reference: hdp_lpcm.py:881-1023, sample_auxillary.py:6-50,
// sample_concentration.py:6-21, distributions.py:72-102).  One CTA per chain, everything small
// lives in shared memory; all draws come from per-(chain, sweep, site) Philox streams, so the
// block is only distributionally -- not draw-for-draw -- equal to the host version
// (dynetlsm_b200/hdp_updates.py), which stays the path of the bit-exact replay mode.
#pragma once
#include "dlsm_device.cuh"
#include "../../include/dlsm.h"

#ifndef DLSM_HDP_BIN_BYTES
#define DLSM_HDP_BIN_BYTES (24 * 1024)
#endif

namespace dlsm {

constexpr uint32_t kRngHdp = 4;

// sequential stream of uniforms / normals / gammas / betas on one Philox (site) lane
struct Stream {
    uint64_t seed;
    uint32_t site, sweep, chain, blk;
    double spare;
    bool has_spare;
    __device__ Stream(uint64_t s, uint32_t site_, uint32_t sweep_, uint32_t chain_)
        : seed(s), site(site_), sweep(sweep_), chain(chain_), blk(0), spare(0.0), has_spare(false) {}
    __device__ double uniform()
    {
        if (has_spare) { has_spare = false; return spare; }
        const U2 u = philox_u2(seed, site, sweep, chain, kRngHdp, blk++);
        spare = u.b;
        has_spare = true;
        return u.a;
    }
    double nspare = 0.0;
    bool has_nspare = false;
    __device__ double normal()
    {
        if (has_nspare) { has_nspare = false; return nspare; }
        double z0;
        box_muller(philox_u2(seed, site, sweep, chain, kRngHdp, blk++), z0, nspare);
        has_nspare = true;
        return z0;
    }
    // Marsaglia & Tsang (2000); shape < 1 boosted by U^(1/shape)
    __device__ double gamma(double shape)
    {
        if (!(shape > 0.0)) return 0.0;
        double boost = 1.0;
        if (shape < 1.0) {
            boost = exp(log(uniform()) / shape); // U^(1/shape) without the cost of pow()
            shape += 1.0;
        }
        const double dd = shape - 1.0 / 3.0, cc = 1.0 / sqrt(9.0 * dd);
        for (int it = 0; it < 256; it++) {
            const double z = normal();
            double v = 1.0 + cc * z;
            if (v <= 0.0) continue;
            v = v * v * v;
            const double u = uniform(), z2 = z * z;
            if (u < 1.0 - 0.0331 * z2 * z2) return boost * dd * v; // squeeze: no logarithms
            if (log(u) < 0.5 * z2 + dd - dd * v + dd * log(v)) return boost * dd * v;
        }
        return boost * dd;
    }
    __device__ double beta(double a, double b)
    {
        const double x = gamma(a), y = gamma(b);
        return x / (x + y);
    }
    // Bernoulli trials draw 32-bit uniforms, four per Philox block (blocks tagged in the high
    // half of the counter range so they never collide with the 52-bit draws of this stream)
    uint32_t w4[4];
    int nw = 0;
    uint32_t blk32 = 0;
    __device__ int bernoulli(double p)
    {
        if (nw == 0) {
            const W4 o = philox_w4(seed, site, sweep, chain, kRngHdp, 0x800000u | blk32++);
            w4[0] = o.x; w4[1] = o.y; w4[2] = o.z; w4[3] = o.w;
            nw = 4;
        }
        const uint32_t r = w4[--nw];
        return ((double)r + 0.5) * 2.3283064365386963e-10 < p ? 1 : 0;
    }
    __device__ int binomial(int n, double p)
    {
        int s = 0;
        for (int i = 0; i < n; i++) s += bernoulli(p);
        return s;
    }
};

struct HdpParams {
    int C, T, n, d, K;
    const double *X;        // [C][T][n][d]
    const int32_t *z;       // [C][T][n]
    const double *ncount;   // [C][T][K][K]
    const int32_t *nk;      // [C][T][K]
    double *mu, *sigma, *lambda, *beta, *weights, *hyper;
    dlsm_hdp_prior pr;
    uint64_t seed;
    uint32_t sweep, chain_offset;
    int bin_rows;           // private rows of the per-cluster binning (0: per-warp segmented reduction)
};

// Escobar & West (1995) auxiliary-variable update of a DP concentration parameter
__device__ inline double concentration(Stream &g, double alpha, double n_clusters, double n_samples,
                                       double shape, double rate)
{
    const double eta = g.beta(alpha + 1.0, n_samples);
    double m_shape = shape + n_clusters - 1.0;
    const double m_scale = rate - log(eta);
    const double odds = (m_shape / m_scale) * (1.0 / n_samples);
    if (g.bernoulli(odds / (1.0 + odds))) m_shape += 1.0;
    return g.gamma(m_shape) / m_scale;
}

// The block factorises into two groups that are conditionally independent given the labels and
// the positions, so they are separate launches (PART) that may run in either order or concurrently:
//   PART 1 "emission side":  mu_k, sigma_k, lambda, tau^2 (mvp), b        -- what the NEXT latent
//                            sweep needs; reads X, z, nk
//   PART 2 "transition side": tables m, overrides, beta, w0, w[t,k], gamma, alpha_init, alpha, kappa
//                            -- only the next label draw needs them; reads the counts
//   PART 0 = both.
// dynamic smem: ints m[T*K*K], wover[T*K], gstart[T*K*K]; doubles mbar[K], newbeta[K], scal[16], rowlog[T*K],
// bins[rows][2*K*d + K], totals[2*K*d + K]
// DLSM_HDP_EMIS_MINB: CTAs per SM the emission-side build is compiled for (9 lets 9 x 148 = 1 332 chains
// run as one wave; the default build needs 64 registers = 8 CTAs)
#ifndef DLSM_HDP_EMIS_MINB
#define DLSM_HDP_EMIS_MINB 1
#endif
template <int PART>
__global__ void __launch_bounds__(128, PART == 1 ? DLSM_HDP_EMIS_MINB : 1) k_hdp_update(const HdpParams p)
{
    constexpr bool kEmis = PART != 2, kTrans = PART != 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int T = p.T, n = p.n, d = p.d, K = p.K, KK = K * K;
    const int c = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    int *m = reinterpret_cast<int *>(smem_raw);          // [T][K][K]
    int *wov = m + T * KK;                               // [T][K] override counts (t >= 1)
    int *gstart = wov + T * K;                           // [T][K][K] first Philox group of a cell
    double *mbar = reinterpret_cast<double *>(gstart + T * KK + ((2 * T * KK + T * K) & 1));
    double *nbeta = mbar + K, *scal = nbeta + K;
    double *rowlog = scal + 16;                          // [T*K] log-beta terms of the alpha+kappa update
    const int nbin = 2 * K * d + K;                      // S0 | S1 | R
    const int brows = p.bin_rows > 0 ? p.bin_rows : 4;
    double *bins = rowlog + T * K;                       // [brows][nbin] private partial sums
    double *S0 = bins + brows * nbin, *S1 = S0 + K * d, *R = S1 + K * d; // totals
    const double *X = p.X + (size_t)c * T * n * d;
    const int32_t *z = p.z + (size_t)c * T * n;
    const double *cnt = p.ncount + (size_t)c * T * KK;
    const int32_t *nk = p.nk + (size_t)c * T * K;
    double *mu = p.mu + (size_t)c * K * d, *sigma = p.sigma + (size_t)c * K;
    double *beta = p.beta + (size_t)c * K, *w = p.weights + (size_t)c * T * KK;
    double *hy = p.hyper + (size_t)c * 8;
    const uint32_t chain = (uint32_t)c + p.chain_offset;
    const double gamma0 = hy[0], alpha_init = hy[1], alpha = hy[2], kappa = hy[3];
    const double mvp = hy[4], bpar = hy[5];
    const double lm = p.lambda[c];
    uint32_t site = 0; // families of streams get disjoint site ranges

    if (kTrans) {
    // ---- 1. table counts m (sample_auxillary.py:6-28) ----
    // m[cell] = sum_{i < count} Bernoulli(pr / (pr + i)).  The draws of all cells are flattened
    // into groups of four customers (one Philox block each) and dealt to the threads, so the few
    // heavy cells (the self-transitions) do not serialise on one lane: draw (cell, i) always uses
    // word 3 - i % 4 of block i / 4 of the cell's stream, whoever computes it.
    {
        const int ncell = T * KK, per = (ncell + nt - 1) / nt;
        const int lo = min(tid * per, ncell), hi = min(lo + per, ncell);
        int loc = 0;
        for (int cell = lo; cell < hi; cell++) {
            const int t = cell / KK, j = (cell / K) % K;
            const int groups = (t > 0 || j == 0) ? ((int)cnt[cell] + 3) >> 2 : 0;
            gstart[cell] = loc;
            loc += groups;
            m[cell] = 0;
        }
        int incl = loc;
        for (int o = 1; o < 32; o <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, incl, o);
            if ((tid & 31) >= o) incl += up;
        }
        int *wtot = reinterpret_cast<int *>(scal); // scal is not in use yet
        if ((tid & 31) == 31) wtot[tid >> 5] = incl;
        __syncthreads();
        int base = 0, total = 0;
        for (int wq = 0; wq < (nt >> 5); wq++) {
            if (wq < (tid >> 5)) base += wtot[wq];
            total += wtot[wq];
        }
        const int off = base + incl - loc;
        for (int cell = lo; cell < hi; cell++) gstart[cell] += off;
        __syncthreads();
        for (int g = tid; g < total; g += nt) {
            int a = 0, b = ncell; // the last cell whose first group is <= g owns g
            while (b - a > 1) {
                const int mid = (a + b) >> 1;
                if (gstart[mid] <= g) a = mid; else b = mid;
            }
            const int cell = a, gi = g - gstart[cell];
            const int t = cell / KK, j = (cell / K) % K, k = cell % K;
            const double pr = (t == 0) ? alpha_init * beta[k] : alpha * beta[k] + (j == k ? kappa : 0.0);
            const int count = (int)cnt[cell];
            const W4 o = philox_w4(p.seed, site + cell, p.sweep, chain, kRngHdp, 0x800000u | (uint32_t)gi);
            const uint32_t wd[4] = {o.w, o.z, o.y, o.x};
            int mm = 0;
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const int i = 4 * gi + r;
                if (i < count && ((double)wd[r] + 0.5) * 2.3283064365386963e-10 < pr / (pr + i)) mm++;
            }
            if (mm) atomicAdd(&m[cell], mm);
        }
    }
    site += T * KK;
    __syncthreads();
    // ---- 2. override variables and m_bar (sample_auxillary.py:31-50) ----
    const double rho0 = kappa / (alpha + kappa);
    for (int e = tid; e < T * K; e += nt) {
        const int t = e / K, j = e % K;
        int wv = 0;
        if (t > 0) {
            Stream g(p.seed, site + e, p.sweep, chain);
            wv = g.binomial(m[t * KK + j * K + j], rho0 / (rho0 + beta[j] * (1.0 - rho0)));
        }
        wov[e] = wv;
    }
    site += T * K;
    __syncthreads();
    for (int k = tid; k < K; k += nt) {
        double s = m[k]; // m[0,0,k]
        for (int t = 1; t < T; t++)
            for (int j = 0; j < K; j++) s += m[t * KK + j * K + k] - (j == k ? wov[t * K + j] : 0);
        mbar[k] = s;
    }
    __syncthreads();
    // ---- 3. beta ~ Dir(gamma/K + m_bar) (hdp_lpcm.py:887) ----
    for (int k = tid; k < K; k += nt) {
        Stream g(p.seed, site + k, p.sweep, chain);
        nbeta[k] = g.gamma(gamma0 / K + mbar[k]);
    }
    site += K;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int k = 0; k < K; k++) s += nbeta[k];
        for (int k = 0; k < K; k++) nbeta[k] /= s;
    }
    __syncthreads();
    // ---- 4. w0 and w[t,k] rows (hdp_lpcm.py:890-898), clipped Dirichlet parameters ----
    for (int cell = tid; cell < T * KK; cell += nt) {
        const int t = cell / KK, j = (cell / K) % K, k = cell % K;
        if (t == 0 && j != 0) continue;
        double a = (t == 0) ? alpha_init * nbeta[k] + nk[k]
                            : alpha * nbeta[k] + (j == k ? kappa : 0.0) + cnt[cell];
        if (a <= 0.0) a = 2.2250738585072014e-308;
        Stream g(p.seed, site + cell, p.sweep, chain);
        w[cell] = g.gamma(a);
    }
    site += T * KK;
    __syncthreads();
    for (int row = tid; row < T * K; row += nt) {
        const int t = row / K, j = row % K;
        if (t == 0 && j != 0) continue;
        double s = 0.0;
        for (int k = 0; k < K; k++) s += w[row * K + k];
        for (int k = 0; k < K; k++) w[row * K + k] /= s;
    }
    // alpha + kappa auxiliary variables, one (t >= 1, j) row per thread (hdp_lpcm.py:998-1012)
    if (tid < 16) scal[tid] = 0.0;
    __syncthreads();
    const double ak_old = alpha + kappa;
    for (int row = K + tid; row < T * K; row += nt) {
        double nd = 0.0, mrow = 0.0;
        for (int k = 0; k < K; k++) { nd += cnt[row * K + k]; mrow += m[row * K + k]; }
        atomicAdd(&scal[5], mrow);                  // sum of m[1:]
        atomicAdd(&scal[6], (double)wov[row]);      // override successes
        double lb = 0.0;
        if (nd > 0.0) {
            Stream g(p.seed, site + row, p.sweep, chain);
            atomicAdd(&scal[2], (double)g.bernoulli(nd / (nd + ak_old)));
            lb = log(g.beta(ak_old + 1.0, nd));
            atomicAdd(&scal[4], mrow);
        }
        rowlog[row] = lb; // summed in row order below: the integer-valued sums above are exact in
                          // any order, this one is not
    }
    site += T * K;
    __syncthreads();
    // one scalar task per thread, each on its own Philox stream (no long serial gamma chains)
    if (tid == 3) { // gamma (:977-983)
        Stream g(p.seed, site + 3, p.sweep, chain);
        double ncl = 0.0, nsm = 0.0;
        for (int k = 0; k < K; k++) { ncl += mbar[k] > 0.0; nsm += mbar[k]; }
        hy[0] = concentration(g, gamma0, ncl, nsm, p.pr.gamma_prior_shape, p.pr.gamma_prior_rate);
    } else if (tid == 4) { // alpha_init (:989-995)
        Stream g(p.seed, site + 4, p.sweep, chain);
        double m00 = 0.0;
        for (int k = 0; k < K; k++) m00 += m[k];
        hy[1] = concentration(g, alpha_init, m00, (double)n, p.pr.alpha_init_shape, p.pr.alpha_init_rate);
    } else if (tid == 5) { // alpha + kappa and rho (:1010-1023)
        Stream g(p.seed, site + 5, p.sweep, chain);
        double slog = 0.0;
        for (int row = K; row < T * K; row++) slog += rowlog[row];
        const double ak = g.gamma(p.pr.alpha_kappa_shape + scal[4] - scal[2]) /
                          (p.pr.alpha_kappa_rate - slog);
        const double rho = g.beta(8.0 + scal[6], scal[5] - scal[6] + 2.0);
        hy[3] = ak * rho;
        hy[2] = ak - ak * rho;
    }
    for (int k = tid; k < K; k += nt) beta[k] = nbeta[k];
    } // kTrans
    site = 2 * T * KK + 2 * T * K + K + 8; // fixed layout of the Philox site ranges, whatever PART runs
    if (kEmis) {
    // ---- 5. cluster means (hdp_lpcm.py:901-919) ----
    // per-cluster sums, reproducible: a warp reduces its 32 entries cluster by cluster with the
    // fixed butterfly and its leader lane adds into the warp's private bins; the four warps' bins
    // are then added in warp order (fp64 atomics would make the chain depend on the arrival order)
    const int lane = tid & 31, warp = tid >> 5;
    double *mybins = bins + (p.bin_rows > 0 ? tid : warp) * nbin;
    for (int e = tid; e < brows * nbin; e += nt) bins[e] = 0.0;
    __syncthreads();
    if (p.bin_rows > 0) {
        // thread r < bin_rows owns row r of the bins and the entries e = r (mod bin_rows), in order
        if (tid < p.bin_rows)
            for (int e = tid; e < T * n; e += p.bin_rows) {
                const int t = e / n, i = e % n;
                double *dst = mybins + (z[e] + (t > 0 ? K : 0)) * d; // S0 bins, then S1 bins
                for (int q = 0; q < d; q++) {
                    double v = X[(size_t)e * d + q];
                    if (t > 0) v -= (1.0 - lm) * X[((size_t)(t - 1) * n + i) * d + q];
                    dst[q] += v;
                }
            }
    } else {
        // large K * d: a warp reduces its 32 entries cluster by cluster with the fixed butterfly
        for (int e0 = warp * 32; e0 < T * n; e0 += nt) {
            const int e = e0 + lane;
            const bool valid = e < T * n;
            const int t = valid ? e / n : 0, i = valid ? e % n : 0;
            const int key = valid ? z[e] + (t > 0 ? K : 0) : -1;
            unsigned todo = __ballot_sync(0xffffffffu, valid);
            while (todo) {
                const int kk = __shfl_sync(0xffffffffu, key, __ffs(todo) - 1);
                const bool mine = valid && key == kk;
                for (int q = 0; q < d; q++) {
                    double v = 0.0;
                    if (mine) {
                        v = X[(size_t)e * d + q];
                        if (t > 0) v -= (1.0 - lm) * X[((size_t)(t - 1) * n + i) * d + q];
                    }
                    v = warp_sum(v);
                    if (lane == 0) mybins[kk * d + q] += v;
                }
                todo &= ~__ballot_sync(0xffffffffu, mine);
            }
        }
    }
    __syncthreads();
    for (int e = tid; e < 2 * K * d; e += nt) { // rows added in row order: reproducible
        double sacc = 0.0;
        for (int r = 0; r < brows; r++) sacc += bins[r * nbin + e];
        S0[e] = sacc;
    }
    __syncthreads();
    for (int k = tid; k < K; k += nt) {
        double prec = 1.0 / mvp + nk[k] / sigma[k];
        double nrest = 0.0;
        for (int t = 1; t < T; t++) nrest += nk[t * K + k];
        prec += (lm * lm / sigma[k]) * nrest;
        const double var = 1.0 / prec, sd = sqrt(var);
        Stream g(p.seed, site + k, p.sweep, chain);
        for (int q = 0; q < d; q++) {
            const double mean = ((1.0 / sigma[k]) * S0[k * d + q] + (lm / sigma[k]) * S1[k * d + q]) * var;
            mu[k * d + q] = mean + sd * g.normal();
        }
    }
    site += K;
    __syncthreads();
    // ---- 6. cluster variances (hdp_lpcm.py:922-937) ----
    if (p.bin_rows > 0) {
        if (tid < p.bin_rows)
            for (int e = tid; e < T * n; e += p.bin_rows) {
                const int t = e / n, i = e % n, k = z[e];
                double r2 = 0.0;
                for (int q = 0; q < d; q++) {
                    const double df = X[(size_t)e * d + q] -
                                      ((t == 0) ? mu[k * d + q]
                                                : (1.0 - lm) * X[((size_t)(t - 1) * n + i) * d + q] + lm * mu[k * d + q]);
                    r2 += df * df;
                }
                mybins[2 * K * d + k] += r2;
            }
    } else {
        for (int e0 = warp * 32; e0 < T * n; e0 += nt) {
            const int e = e0 + lane;
            const bool valid = e < T * n;
            const int t = valid ? e / n : 0, i = valid ? e % n : 0;
            const int k = valid ? z[e] : -1;
            double r2 = 0.0;
            if (valid)
                for (int q = 0; q < d; q++) {
                    const double df = X[(size_t)e * d + q] -
                                      ((t == 0) ? mu[k * d + q]
                                                : (1.0 - lm) * X[((size_t)(t - 1) * n + i) * d + q] + lm * mu[k * d + q]);
                    r2 += df * df;
                }
            unsigned todo = __ballot_sync(0xffffffffu, valid);
            while (todo) {
                const int kk = __shfl_sync(0xffffffffu, k, __ffs(todo) - 1);
                const bool mine = valid && k == kk;
                const double v = warp_sum(mine ? r2 : 0.0);
                if (lane == 0) mybins[2 * K * d + kk] += v;
                todo &= ~__ballot_sync(0xffffffffu, mine);
            }
        }
    }
    __syncthreads();
    for (int k = tid; k < K; k += nt) {
        const int e = 2 * K * d + k;
        double sacc = 0.0;
        for (int r = 0; r < brows; r++) sacc += bins[r * nbin + e];
        R[k] = sacc;
    }
    __syncthreads();
    for (int k = tid; k < K; k += nt) {
        double tot = 0.0;
        for (int t = 0; t < T; t++) tot += nk[t * K + k];
        const double shape = 0.5 * (tot * d + p.pr.a), rate = 0.5 * bpar + 0.5 * R[k];
        Stream g(p.seed, site + k, p.sweep, chain);
        sigma[k] = rate / g.gamma(shape);
    }
    site += K;
    __syncthreads();
    // ---- 7. lambda ~ truncated normal on (0,1); 8. the tau^2 and b hyper-priors ----
    __syncthreads();
    if (tid < 2) scal[tid] = 0.0;
    __syncthreads();
    {
        double ml = 0.0, sl = 0.0;
        for (int e = n + tid; e < T * n; e += nt) { // t >= 1
            const int t = e / n, i = e % n, k = z[e];
            const double *x = X + (size_t)e * d, *xp = X + ((size_t)(t - 1) * n + i) * d;
            for (int q = 0; q < d; q++) {
                const double dm = mu[k * d + q] - xp[q];
                ml += (dm / sigma[k]) * (x[q] - xp[q]);
                sl += dm * dm / sigma[k];
            }
        }
        ml = warp_sum(ml);
        sl = warp_sum(sl);
        if ((tid & 31) == 0) { scal[8 + (tid >> 5)] = ml; scal[12 + (tid >> 5)] = sl; }
    }
    __syncthreads();
    if (tid == 0) { // (hdp_lpcm.py:940-954), inverse-cdf draw
        Stream g(p.seed, site + 0, p.sweep, chain);
        const double sum_ml = ((scal[8] + scal[9]) + scal[10]) + scal[11];
        const double sum_sl = ((scal[12] + scal[13]) + scal[14]) + scal[15];
        const double var = 1.0 / (1.0 / p.pr.lambda_variance_prior + sum_sl);
        const double mean = (sum_ml + p.pr.lambda_prior / p.pr.lambda_variance_prior) * var;
        const double sd = sqrt(var);
        const double lo = normcdf((0.0 - mean) / sd), hi = normcdf((1.0 - mean) / sd);
        double u = lo + (hi - lo) * g.uniform();
        u = fmin(fmax(u, 1e-300), 1.0 - 1e-16);
        p.lambda[c] = fmin(fmax(mean + sd * normcdfinv(u), 0.0), 1.0);
    } else if (tid == 1) { // tau^2 hyper-prior (hdp_lpcm.py:957-962)
        if (p.pr.resample_mvp) {
            Stream g(p.seed, site + 1, p.sweep, chain);
            double bb = 0.5 * p.pr.b0;
            for (int e = 0; e < K * d; e++) bb += 0.5 * mu[e] * mu[e];
            hy[4] = bb / g.gamma(0.5 * (p.pr.a0 + K));
        }
    } else if (tid == 2) { // b hyper-prior (:965-972)
        if (p.pr.resample_b) {
            Stream g(p.seed, site + 2, p.sweep, chain);
            double sc = 0.5 * p.pr.d0;
            for (int k = 0; k < K; k++) sc += 0.5 * (1.0 / sigma[k]);
            hy[5] = g.gamma(0.5 * (p.pr.c0 + K * p.pr.a)) / sc;
        }
    }
    } // kEmis
}

// private binning rows: as many of the 128 threads as fit a 24 KB bin area (multiples of 32)
inline int hdp_bin_rows(int K, int d)
{
    const size_t nbin = (size_t)2 * K * d + K;
    const size_t rows = ((size_t)DLSM_HDP_BIN_BYTES) / (nbin * sizeof(double));
    return rows >= 128 ? 128 : (rows >= 64 ? 64 : (rows >= 32 ? 32 : 0));
}

inline size_t hdp_smem_bytes(int T, int K, int d)
{
    const size_t ints = (size_t)2 * T * K * K + (size_t)T * K;
    const size_t nbin = (size_t)2 * K * d + K;
    const int rows = hdp_bin_rows(K, d);
    return (ints + (ints & 1)) * sizeof(int) +
           ((size_t)2 * K + 16 + (size_t)T * K + ((rows > 0 ? rows : 4) + 1) * nbin) * sizeof(double) + 16;
}

} // namespace dlsm
