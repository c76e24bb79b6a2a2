// dlsm_device.cuh -- device-side building blocks shared by every kernel of libdlsm.so.
//
// fp64 throughout (the reference's arithmetic type).  Rounding-sensitive steps
// (proposal, squared distance, priors) use the explicit-rounding intrinsics so that nvcc cannot
// contract them into FMAs: the reference (Cython built for baseline x86-64, numpy elementwise
// ops) rounds every multiply and add separately, and accepted states must be bit-identical.
#pragma once
#include <cmath>
#include <cstdint>
#include <cuda_runtime.h>
#include "dlsm_tables.cuh"

namespace dlsm {

constexpr int kMaxD = 8;
constexpr unsigned kFull = 0xffffffffu;

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11).  Counter = (site, sweep, chain, (kind<<24)|block),
// key = 64-bit seed.  Each call yields two 52-bit uniforms in (0,1).
// ---------------------------------------------------------------------------------------------
enum RngKind : uint32_t { kRngLatent = 0, kRngIntercept = 1, kRngRadii = 2, kRngLabels = 3 };

__host__ __device__ inline void philox_round(uint32_t &c0, uint32_t &c1, uint32_t &c2,
                                             uint32_t &c3, uint32_t k0, uint32_t k1)
{
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
}

struct U2 { double a, b; };
struct W4 { uint32_t x, y, z, w; };

__host__ __device__ inline U2 philox_u2(uint64_t seed, uint32_t site, uint32_t sweep,
                                        uint32_t chain, uint32_t kind, uint32_t block)
{
    uint32_t c0 = site, c1 = sweep, c2 = chain, c3 = (kind << 24) | (block & 0xffffffu);
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; r++) {
        philox_round(c0, c1, c2, c3, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    const double s = 2.220446049250313e-16; // 2^-52
    U2 u;
    u.a = ((double)(((uint64_t)c0 << 20) | (c1 >> 12)) + 0.5) * s;
    u.b = ((double)(((uint64_t)c2 << 20) | (c3 >> 12)) + 0.5) * s;
    return u;
}

// raw 4 x 32 bits of one Philox block (cheap Bernoulli trials: 32-bit resolution is plenty)
__host__ __device__ inline W4 philox_w4(uint64_t seed, uint32_t site, uint32_t sweep, uint32_t chain,
                                        uint32_t kind, uint32_t block)
{
    uint32_t c0 = site, c1 = sweep, c2 = chain, c3 = (kind << 24) | (block & 0xffffffu);
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; r++) {
        philox_round(c0, c1, c2, c3, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    W4 o;
    o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

// Box-Muller pair from two uniforms in (0,1)
__device__ inline void box_muller(const U2 u, double &z0, double &z1)
{
    const double r = sqrt(-2.0 * log(u.a));
    double s, c;
    sincospi(2.0 * u.b, &s, &c);
    z0 = r * c;
    z1 = r * s;
}

// the draws of one latent node-update: eps[0..d) and log(u)
template <int DM>
__device__ inline void latent_draws(uint64_t seed, uint32_t site, uint32_t sweep, uint32_t chain,
                                    int d, double (&eps)[DM], double &logu)
{
    const U2 ua = philox_u2(seed, site, sweep, chain, kRngLatent, 0);
    logu = log(ua.a);
#pragma unroll
    for (int p = 0; p < (DM + 1) / 2; p++) {
        if (2 * p < d) {
            double z0, z1;
            box_muller(philox_u2(seed, site, sweep, chain, kRngLatent, 1 + p), z0, z1);
            eps[2 * p] = z0;
            if (2 * p + 1 < DM) eps[2 * p + 1] = z1;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// log(1 + exp(eta)) exactly as the reference writes it (static_network_fast.pyx:42,
// directed_likelihoods_fast.pyx:73): no softplus guard, overflow to +inf above eta ~ 709.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double log1pexp_naive(double eta) { return log(1.0 + exp(eta)); }

// ---------------------------------------------------------------------------------------------
// Fused fp64 softplus, |abs error| < 1e-15 on the whole real line (tests/test_host_numerics.py):
//   log(1+e^x) = max(x,0) + log1p(t),  t = e^{-|x|} in (0,1]
//   t:      |x| = n ln2/32 - r, |r| <= ln2/64;  t = 2^{-(n>>5)} * 2^{-(n&31)/32} * e^{r}  (deg-6)
//   log1p:  y = 1+t (rounding error c recovered), y = (1+z)/R[i] with i = top 6 mantissa bits,
//           log y = -log R[i] + (z - z^2/2 + ... + z^7/7),  |z| <= 1/128;  + c*R[i]
// 26 fp64 instructions + 3 table loads (1.3 KB of tables, L1-resident) against ~87 fp64 + ~190
// integer/branch instructions for CUDA's exp() followed by log() (ncu, profiles/r1_*).  The
// reference formula is the naive one; the two agree to ~1e-16 absolute, far inside the 1e-10
// relative parity bound on the summed log-likelihoods.
// ---------------------------------------------------------------------------------------------
#define DLSM_X(v) v,
static __device__ const double d_exp_tab[64] = {DLSM_EXP_TAB(DLSM_X)};
static __device__ const double d_expp_tab[64] = {DLSM_EXPP_TAB(DLSM_X)};
#define DLSM_X2(r, l) {r, l},
static __device__ __align__(16) const double2 d_rl_tab[129] = {DLSM_RL_TAB(DLSM_X2)}; // {R[i], -log R[i]}
static __device__ const double d_exp32_tab[32] = {DLSM_EXP32_TAB(DLSM_X)};                     // 2^(-j/32)
static __device__ __align__(16) const double2 d_rl32_tab[33] = {DLSM_RL32_TAB(DLSM_X2)};         // {R[i], -log R[i]}, R[i] ~ 1/(1+i/32)
#undef DLSM_X2
static const double h_exp32_tab[32] = {DLSM_EXP32_TAB(DLSM_X)};
static const double h_rcp32_tab[33] = {DLSM_RCP32_TAB(DLSM_X)};
static const double h_log32_tab[33] = {DLSM_LOG32_TAB(DLSM_X)};
static const double h_exp_tab[64] = {DLSM_EXP_TAB(DLSM_X)};
static const double h_expp_tab[64] = {DLSM_EXPP_TAB(DLSM_X)};
static const double h_rcp_tab[129] = {DLSM_RCP_TAB(DLSM_X)};
static const double h_log_tab[129] = {DLSM_LOG_TAB(DLSM_X)};
#undef DLSM_X

// scalar constants live in the constant bank so that DFMA takes them as c[][] operands instead of
// burning registers / UMOV+IMAD pairs on 64-bit immediates
#define DLSM_SP_CONSTS                                                                          \
    {DLSM_32_OVER_LN2, 6755399441055744.0 /* 1.5*2^52 */, DLSM_LN2_32_HI, DLSM_LN2_32_LO,        \
     1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5, 1.0,                                  \
     1.0 / 7.0, -1.0 / 6.0, 0.2, -0.25, 1.0 / 3.0, -0.5}
static __constant__ double d_spc[16] = DLSM_SP_CONSTS;
static const double h_spc[16] = DLSM_SP_CONSTS;
// constants of the small-table softplus (fast_log1pexp_neg below)
#define DLSM_SP2_CONSTS                                                                         \
    {DLSM_NE2_OVER_LN2, 6755399441055744.0 /* 1.5*2^52 */, DLSM_LN2_OVER_NE2, 1.0,               \
     1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5,                                       \
     -1.0 / 8.0, 1.0 / 7.0, -1.0 / 6.0, 0.2, -0.25, 1.0 / 3.0, -0.5}
static __constant__ double d_sp2[16] = DLSM_SP2_CONSTS;
static const double h_sp2[16] = DLSM_SP2_CONSTS;

// log1p(exp(-|eta|)): the part of softplus that is left after splitting off max(eta, 0).
// Branch-free and select-free on purpose (the sweep kernels are issue- and L1-bound, ncu
// profiles/r2a): independent evaluations in one thread interleave freely.  23 fp64 instructions and
// two table loads whose 32 lanes touch at most 2 + 5 cache lines (256-byte and 528-byte tables; the
// first generation, kept as _v1 below, used 512 bytes + 2 KB = 4 + 16 lines per warp and was the
// kernel's L1 bottleneck).
//   t = e^{-a}:  a = n ln2/32 - r, |r| <= ln2/64;  t = 2^{-(n>>5)} 2^{-(n&31)/32} e^r   (degree 6;
//                one-step reduction: the rounding of ln2/32 costs < 4e-17 ABSOLUTE in t)
//   log1p(t):    y = 1 + t;  i = round(32 (y-1)), R = fl(1/(1 + i/32));  z = y R - 1, |z| <= 1/64;
//                log y = -log R + (z - z^2/2 + ... - z^8/8)
// The rounding of y = 1 + t (<= 1.1e-16 absolute) is NOT compensated: the result is accurate to
// ~2e-16 ABSOLUTE, which is what sums of O(1) terms need (relative accuracy in the far tail,
// where the term is < 1e-9, is ~1e-7); tests/test_host_numerics.py.
// n is clamped (unsigned) so that a > ~690 -- where t < 1e-300 -- cannot wrap the exponent.
__host__ __device__ __forceinline__ double fast_log1pexp_neg(double a /* = |eta| */)
{
#ifdef __CUDA_ARCH__
    const double *ET = d_exp32_tab, *K = d_sp2;
#else
    const double *ET = h_exp32_tab, *RT = h_rcp32_tab, *LT = h_log32_tab, *K = h_sp2;
#endif
    const double kf = fma(a, K[0], K[1]); // round(a * 32/ln2) lands in the low mantissa bits
    union { double f; long long i; unsigned u[2]; } cv;
    cv.f = kf;
    const unsigned n = cv.u[0];
    const double nf = kf - K[1];
    const double r = fma(nf, K[2], -a);
    double p = fma(r, K[4], K[5]);
    p = fma(r, p, K[6]);
    p = fma(r, p, K[7]);
    p = fma(r, p, K[8]);
    p = fma(r, p, K[3]);
    p = fma(r, p, K[3]);
    cv.f = p * ET[n & 31u];
    const unsigned sh = n >> 5;
    cv.u[1] -= (sh < 1000u ? sh : 1000u) << 20; // * 2^-(n>>5)
    const double y = K[3] + cv.f;
    cv.f = y;
    unsigned i = (cv.u[1] - 0x3ff00000u + 0x4000u) >> 15; // 0..32
    i = i < 32u ? i : 32u; // NaN / garbage never indexes out of the table
#ifdef __CUDA_ARCH__
    const double2 rl = __ldg(&d_rl32_tab[i]); // reciprocal and log in one 128-bit load
    const double Ri = rl.x, Li = rl.y;
#else
    const double Ri = RT[i], Li = LT[i];
#endif
    const double z = fma(y, Ri, -K[3]);
    double q = fma(z, K[9], K[10]);
    q = fma(z, q, K[11]);
    q = fma(z, q, K[12]);
    q = fma(z, q, K[13]);
    q = fma(z, q, K[14]);
    q = fma(z, q, K[15]);
    q = fma(z, q, K[3]);
    return fma(z, q, Li);
}

// first generation (64- and 129-entry tables, degrees 5 and 6), kept for A/B measurements
__host__ __device__ __forceinline__ double fast_log1pexp_neg_v1(double a /* = |eta| */)
{
#ifdef __CUDA_ARCH__
    const double *ET = d_exp_tab, *K = d_spc;
#else
    const double *ET = h_exp_tab, *RT = h_rcp_tab, *LT = h_log_tab, *K = h_spc;
#endif
    const double kf = fma(a, K[0], K[1]); // round(a * 64/ln2) lands in the low mantissa bits
    union { double f; long long i; unsigned u[2]; } cv;
    cv.f = kf;
    const unsigned n = cv.u[0];
    const double nf = kf - K[1];
    double r = fma(nf, K[2], -a);
    r = fma(nf, K[3], r);
    double p = fma(r, K[5], K[6]);
    p = fma(r, p, K[7]);
    p = fma(r, p, K[8]);
    p = fma(r, p, K[9]);
    p = fma(r, p, K[9]);
    cv.f = p * ET[n & 63u];
    const unsigned sh = n >> 6;
    cv.u[1] -= (sh < 1000u ? sh : 1000u) << 20; // * 2^-(n>>6)
    const double y = K[9] + cv.f;
    cv.f = y;
    unsigned i = (cv.u[1] - 0x3ff00000u + 0x1000u) >> 13; // 0..128
    i = i < 128u ? i : 128u; // NaN / garbage never indexes out of the table
#ifdef __CUDA_ARCH__
    const double2 rl = __ldg(&d_rl_tab[i]); // reciprocal and log in one 128-bit load
    const double Ri = rl.x, Li = rl.y;
#else
    const double Ri = RT[i], Li = LT[i];
#endif
    const double z = fma(y, Ri, -K[9]);
    double q = fma(z, K[11], K[12]);
    q = fma(z, q, K[13]);
    q = fma(z, q, K[14]);
    q = fma(z, q, K[15]);
    q = fma(z, q, K[9]);
    return fma(z, q, Li);
}

// log(1 + e^eta) = max(eta, 0) + log1p(e^{-|eta|}); max(eta,0) = (eta + |eta|)/2 keeps NaN alive.
// |abs error| < 1e-15 everywhere, relative error < 1e-15 in the negative tail
// (tests/test_host_numerics.py).
#ifdef DLSM_SOFTPLUS_V1 // A/B builds: the first-generation (large-table) evaluation everywhere
#define DLSM_L1PEN fast_log1pexp_neg_v1
#else
#define DLSM_L1PEN fast_log1pexp_neg
#endif
__host__ __device__ __forceinline__ double fast_log1pexp(double eta)
{
    const double a = fabs(eta);
    return fma(0.5, a, 0.5 * eta) + DLSM_L1PEN(a);
}

// One Bernoulli-logit term  y*eta - log(1 + e^eta)  with y in {0,1} passed as ym = y - 1/2:
//     y eta - max(eta,0) - L = (y - 1/2) eta - |eta|/2 - L,   L = log1p(e^{-|eta|})
// three fp64 instructions around L and no selects.
__host__ __device__ __forceinline__ double logit_term(double ym, double eta)
{
    const double a = fabs(eta);
    return fma(ym, eta, fma(-0.5, a, -DLSM_L1PEN(a)));
}

// Branch-free fp64 exp(x) on the same tables (relative error < 1e-15 for |x| < 700; flushes to 0
// below x = -708 where the reference's libm returns denormals, +inf above 709.7).
__host__ __device__ __forceinline__ double fast_exp(double x)
{
#ifdef __CUDA_ARCH__
    const double *PT = d_expp_tab, *K = d_spc;
#else
    const double *PT = h_expp_tab, *K = h_spc;
#endif
    const double xc = fmin(fmax(x, -1000.0), 1000.0); // keeps the integer part in range; NaN -> handled below
    const double kf = fma(xc, K[0], K[1]);
    union { double f; long long i; unsigned u[2]; int s[2]; } cv;
    cv.f = kf;
    const int n = cv.s[0];
    const double nf = kf - K[1];
    double r = fma(nf, -K[2], xc);
    r = fma(nf, -K[3], r);
    double p = fma(r, K[5], K[6]);
    p = fma(r, p, K[7]);
    p = fma(r, p, K[8]);
    p = fma(r, p, K[9]);
    p = fma(r, p, K[9]);
    cv.f = p * PT[n & 63];
    int e = n >> 6;
    e = e < -1022 ? -1022 : (e > 1023 ? 1023 : e);
    cv.s[1] += e << 20;
    double v = cv.f;
    v = (x < -708.0) ? 0.0 : v;
    v = (x > 709.7) ? (double)INFINITY : v;
    return (x != x) ? x : v;
}

// Branch-free fp64 sqrt: MUFU.RSQ64H seed (~22 bits), one Goldschmidt step (~44 bits) and a final
// fused correction g += (s - g*g) * h (h only needs the seed's accuracy there), which squares the
// error again: exactly rounded on 4M random inputs (tools/ubench/sqrt_check.cu).  The argument
// must be > 0 and normal: callers add 1e-300 so that coincident points give 1e-150 instead of
// 0 * inf (absorbed exactly for every s > 1e-284).  NaN propagates.
__device__ __forceinline__ double fast_sqrt_pos(double s)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
    double g = s * y;
    const double h = 0.5 * y;
    g = fma(g, fma(-h, g, 0.5), g);
    return fma(fma(-g, g, s), h, g);
}

__device__ __forceinline__ double fast_sqrt(double s) { return fast_sqrt_pos(s + 1e-300); }

// Euclidean distance for the likelihood terms: squared differences accumulated by FMA on top of
// the 1e-300 guard (one rounding fewer than the reference's separate multiply and add -- inside
// the 1e-10 parity bound of the log-likelihoods, and irrelevant to the bit-exactness of the
// states, which only depends on decisions)
template <int DM>
__device__ __forceinline__ double fast_dist(const double (&a)[DM], const double (&b)[DM], int d)
{
    double s = 1e-300;
#pragma unroll
    for (int k = 0; k < DM; k++)
        if (k < d) {
            const double df = a[k] - b[k];
            s = fma(df, df, s);
        }
    return fast_sqrt_pos(s);
}

#ifndef DLSM_NAIVE_SOFTPLUS
__device__ __forceinline__ double log1pexp(double eta) { return fast_log1pexp(eta); }
#else
__device__ __forceinline__ double log1pexp(double eta) { return log1pexp_naive(eta); }
#endif

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// Reduce two per-lane partial sums over the warp with 6 double shuffles instead of 10: the first
// exchange folds (a, b) so that lanes 0-15 carry a and lanes 16-31 carry b, four butterfly steps
// finish both, one more exchange hands every lane both totals.
__device__ __forceinline__ void warp_sum2(double &a, double &b, int lane)
{
    const bool up = lane & 16;
    double keep = up ? b : a, give = up ? a : b;
    keep += __shfl_xor_sync(kFull, give, 16);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) keep += __shfl_xor_sync(kFull, keep, o);
    const double other = __shfl_xor_sync(kFull, keep, 16);
    a = up ? other : keep;
    b = up ? keep : other;
}

// numpy pairwise_sum for n <= 8 (d <= 8 reductions in the priors): serial below 8, the
// 8-accumulator tree at exactly 8.
template <int DM>
__device__ __forceinline__ double np_sum_small(const double (&a)[DM], int d)
{
    if (DM == 8 && d == 8) {
        return __dadd_rn(__dadd_rn(__dadd_rn(a[0], a[1]), __dadd_rn(a[2], a[3])),
                         __dadd_rn(__dadd_rn(a[4 % DM], a[5 % DM]), __dadd_rn(a[6 % DM], a[7 % DM])));
    }
    double s = -0.0;
#pragma unroll
    for (int k = 0; k < DM; k++)
        if (k < d) s = __dadd_rn(s, a[k]);
    return s;
}

// 0.5 * np.sum(v*v) / s
template <int DM>
__device__ __forceinline__ double half_sumsq_over(const double (&v)[DM], int d, double s)
{
    double sq[DM];
#pragma unroll
    for (int k = 0; k < DM; k++) sq[k] = (k < d) ? __dmul_rn(v[k], v[k]) : 0.0;
    return __ddiv_rn(__dmul_rn(0.5, np_sum_small<DM>(sq, d)), s);
}

// squared Euclidean distance with the reference's rounding: dist += (a - b) ** 2, serial in k
template <int DM>
__device__ __forceinline__ double sqdist(const double (&a)[DM], const double (&b)[DM], int d)
{
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < DM; k++)
        if (k < d) {
            const double df = __dsub_rn(a[k], b[k]);
            s = __dadd_rn(s, __dmul_rn(df, df));
        }
    return s;
}

template <int DM>
__device__ __forceinline__ void load_pos(const double *p, int d, double (&x)[DM])
{
    if (DM == 2) {
        const double2 v = *reinterpret_cast<const double2 *>(p);
        x[0] = v.x;
        x[1 % DM] = v.y;
    } else {
#pragma unroll
        for (int k = 0; k < DM; k++) x[k] = (k < d) ? p[k] : 0.0;
    }
}

// Metropolis tuner (metropolis.py:5-37, :113-136); tune < 0 means tune=None
__device__ inline double tune_random_walk(double s, double r)
{
    if (r < 0.001) s *= 0.1;
    else if (r < 0.05) s *= 0.5;
    else if (r < 0.25) s *= 0.9;
    else if (r > 0.95) s *= 10.0;
    else if (r > 0.75) s *= 2.0;
    else if (r > 0.4) s *= 1.1;
    return s;
}

__device__ inline double tune_dirichlet(double s, double r)
{
    if (r < 0.001) s *= 10.0;
    else if (r < 0.05) s *= 2.0;
    else if (r < 0.25) s *= 1.1;
    else if (r > 0.95) s *= 0.1;
    else if (r > 0.75) s *= 0.5;
    else if (r > 0.4) s *= 0.9;
    return s;
}

__device__ inline void metropolis_bookkeep(double &step, int &n_accepted, int &n_steps, int &until,
                                           int tune, int tune_interval, int accepted,
                                           bool dirichlet)
{
    n_accepted += accepted;
    n_steps += 1;
    if (tune < 0) return;
    if (n_steps < tune && until == 0) {
        const double rate = (double)n_accepted / (double)tune_interval;
        step = dirichlet ? tune_dirichlet(step, rate) : tune_random_walk(step, rate);
        n_accepted = 0;
        until = tune_interval;
    } else {
        until -= 1;
    }
}

} // namespace dlsm
