// dlsm_cbp.cu -- k_sweep_cbp: the block-speculative chain kernel (k_sweep_cb, dlsm_blk.cu) with TWO warps
// per time slice, for chains that have an SM to themselves (at most one chain per SM: a single-chain
// fit, cfg 4 sharded over 8 GPUs = 128 chains per GPU).
//
// k_sweep_cb at one CTA per SM runs 10 warps (cfg 4): 2.75 ms per sweep at 128 registers, 4.27 ms at 96 --
// while two co-resident CTAs (20 warps at 96 registers) finish twice the work in 5.14 ms: with one
// chain per SM the kernel is bound by the latency of its dependent fp64 chains, not by the pipe.  Here
// the 32 rows of a block keep their lanes, but the columns of the slice are dealt to two warps in
// alternating groups of 128; the second warp hands its four per-lane partial sums (columns below /
// above the block, proposal / current position) to the first through shared memory, which then
// resolves the block exactly as k_sweep_cb does.  Two named barriers (bar.sync 1 + t, 64) per block
// and slice: partial sums ready, commits visible.
// Same decisions as the sequential sweep (exact; sums differ in order only).
// MEASURED (B200): cfg 4, 128 chains 2.69 ms vs 2.75 ms (k_sweep_cb at 128 registers); cfg 2, 148 chains
// 0.48 vs 0.34 ms (rows of 120 columns are one group: the second warp only adds barriers).  The
// one-chain-per-SM regime is not latency-bound after all -- kept as DLSM_CHAIN_BLOCK_PAIR, not a default.
// grid = C (<= SMs), block = 64 * T (T <= 15); d = 2, positions in shared memory
#include "dlsm_kernels.cuh"
#include "dlsm_dyad.cuh"
#include "dlsm_blk.h"

namespace dlsm {

constexpr int kCbpMaxT = 15; // named barriers 1..T

static size_t cbp_smem_bytes(int T, int n, int d)
{
    const size_t x = (((size_t)T * n * d + 1) & ~(size_t)1) * sizeof(double);
    return x + (size_t)T * 32 * d * sizeof(double) + (size_t)T * 32 * 4 * sizeof(double) + (size_t)T * sizeof(int) + 64;
}

bool cbp_applicable(int T, int n, int d)
{
    return d == 2 && T <= kCbpMaxT && cbp_smem_bytes(T, n, d) <= (size_t)227 * 1024;
}

// MAXT: 640 threads (T <= 10: up to 102 registers) or 960 (T <= 15: 68 registers)
template <int LK, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) k_sweep_cbp(const SweepParams p)
{
    constexpr int DM = 2;
    constexpr bool kDir = LK != kUndirected;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int T = p.net.T, n = p.net.n, d = 2, W = p.net.W;
    const int c = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = warp >> 1, half = warp & 1; // the slice's first warp resolves, both walk columns
    const size_t chain_elems = (size_t)T * n * d;
    double *Xg = p.X + (size_t)c * chain_elems;
    double *Xc = reinterpret_cast<double *>(smem_raw);
    double *stage_base = Xc + ((chain_elems + 1) & ~(size_t)1);
    double *st_prop = stage_base + (size_t)t * 32 * d;                  // [T][32][d]
    double *xpart = stage_base + (size_t)T * 32 * d + (size_t)t * 32 * 4; // [T][32][4]
    volatile int *progress = reinterpret_cast<volatile int *>(stage_base + (size_t)T * 32 * d + (size_t)T * 32 * 4);
    for (size_t e = threadIdx.x; e < chain_elems; e += blockDim.x) Xc[e] = Xg[e];
    for (int q = threadIdx.x; q < T; q += blockDim.x) progress[q] = 0;
    __syncthreads();

    const double b0 = p.intercept[c * 2 + 0], b1 = p.intercept[c * 2 + 1];
    const double *rinv = kDir ? p.rinv + (size_t)c * n : nullptr;
    const uint32_t chain_id = (uint32_t)c + p.chain_offset;
    bool nonfinite = false;
    double full_acc = 0.0;

    {
        double *Xt = Xc + (size_t)t * n * d;
        for (int jb = 0; jb < n; jb += 32) {
            const int jend = (n - jb) < 32 ? (n - jb) : 32;
            const bool vrow = lane < jend;
            const int jl = vrow ? jb + lane : jb;
            const size_t gs = ((size_t)c * T + t) * n + jl;
            // ---- lane-parallel preparation (both warps: each needs the row's two positions) ----
            double xn[DM], xo[DM], logu = 0.0, inv = 0.0, nn = 0.0, no = 0.0;
            double my_step = p.step[gs];
            int my_nacc = 0, my_nsteps = 0, my_until = 0, zc = 0;
            load_pos<DM>(Xt + (size_t)jl * d, d, xo);
            {
                double eps[DM];
                if (p.eps) {
                    eps[0] = p.eps[gs * d]; eps[1] = p.eps[gs * d + 1];
                    logu = p.logu[gs];
                } else {
                    latent_draws<DM>(p.seed, (uint32_t)(t * n + jl), p.sweep, chain_id, d, eps, logu);
                }
                xn[0] = __dadd_rn(xo[0], __dmul_rn(my_step, eps[0]));
                xn[1] = __dadd_rn(xo[1], __dmul_rn(my_step, eps[1]));
                if (half == 0) {
                    st_prop[lane * d] = xn[0]; st_prop[lane * d + 1] = xn[1];
                    my_nacc = p.nacc[gs]; my_nsteps = p.nsteps[gs]; my_until = p.until[gs];
                    inv = (t == 0) ? 1.0 / p.tau_sq : 1.0 / p.sigma_sq;
                    if (p.prior != 0) {
                        zc = p.z[((size_t)c * T + t) * n + jl];
                        inv = 1.0 / p.sigma[(size_t)c * p.K + zc];
                    }
                    if (t < T - 1) { // X[t+1, j] is still last sweep's value: slice t+1 trails this one
                        double xnx[DM];
                        const volatile double *q = Xc + ((size_t)(t + 1) * n + jl) * d;
                        xnx[0] = q[0]; xnx[1] = q[1];
                        nn = prior_next<DM>(p, c, t, jl, xn, xnx);
                        no = prior_next<DM>(p, c, t, jl, xo, xnx);
                    }
                }
            }
            __syncwarp();
            // ---- this warp's columns: alternating groups of 128; sums over i < jb (lo) and i >= jb + 32 (hi) ----
            const double rj = kDir ? rinv[jl] : 0.0;
            const uint32_t *rowb = p.net.rowbits + ((size_t)t * n + jl) * W;
            const uint32_t *colb = kDir ? p.net.colbits + ((size_t)t * n + jl) * W : nullptr;
            double lo_n = 0.0, lo_o = 0.0, hi_n = 0.0, hi_o = 0.0;
            auto column = [&](int i, uint32_t wr, uint32_t wc, double &an, double &ao) {
                double xi[DM];
                load_pos<DM>(Xt + (size_t)i * d, d, xi);
                const double ri = kDir ? __ldg(rinv + i) : 0.0;
                const double yr = ymask(wr, i & 31), yc = ymask(wc, i & 31);
                an += dyad<LK, DM>(xi, ri, xn, rj, yr, yc, b0, b1, d);
                ao += dyad<LK, DM>(xi, ri, xo, rj, yr, yc, b0, b1, d);
            };
            for (int w4 = 4 * half; w4 * 32 < n; w4 += 8) {
                const uint4 r4 = __ldg(reinterpret_cast<const uint4 *>(rowb + w4));
                uint4 c4 = make_uint4(0u, 0u, 0u, 0u);
                if (kDir) c4 = __ldg(reinterpret_cast<const uint4 *>(colb + w4));
                const uint32_t rw[4] = {r4.x, r4.y, r4.z, r4.w}, cw[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int base = (w4 + k) * 32;
                    if (base >= n) break;
                    if (base == jb) continue; // the block's own columns: below
                    const int cnt = (n - base) < 32 ? (n - base) : 32;
                    double a_n = 0.0, a_o = 0.0, b_n = 0.0, b_o = 0.0;
                    int ii = 0;
                    for (; ii + 1 < cnt; ii += 2) { // two columns per trip: four independent chains
                        column(base + ii, rw[k], cw[k], a_n, a_o);
                        column(base + ii + 1, rw[k], cw[k], b_n, b_o);
                    }
                    if (ii < cnt) column(base + ii, rw[k], cw[k], a_n, a_o);
                    if (base < jb) { lo_n += a_n + b_n; lo_o += a_o + b_o; }
                    else { hi_n += a_n + b_n; hi_o += a_o + b_o; }
                }
            }
            // the block's own columns at their old positions, dealt by parity: below the row -> lo, above -> hi
            const uint32_t wrb = __ldg(rowb + (jb >> 5)), wcb = kDir ? __ldg(colb + (jb >> 5)) : 0u;
            for (int ic = half; ic < jend; ic += 2) {
                double tn = 0.0, to = 0.0;
                column(jb + ic, wrb, wcb, tn, to);
                if (ic < lane) { lo_n += tn; lo_o += to; }
                else if (ic > lane) { hi_n += tn; hi_o += to; }
            }
            if (half == 1) {
                double *xp4 = xpart + lane * 4;
                xp4[0] = lo_n; xp4[1] = lo_o; xp4[2] = hi_n; xp4[3] = hi_o;
            }
            asm volatile("bar.sync %0, 64;" ::"r"(1 + t) : "memory"); // the second warp's partial sums are in xpart
            unsigned mask = 0u;
            if (half == 0) {
                const double *xp4 = xpart + lane * 4;
                lo_n += xp4[0]; lo_o += xp4[1]; hi_n += xp4[2]; hi_o += xp4[3];
                // ---- wavefront: the whole block of slice t-1 must be final ----
                double xp[DM] = {0.0, 0.0};
                if (t > 0) {
                    while (progress[t - 1] < jb + jend) { __nanosleep(DLSM_SPIN_NS); }
                    __threadfence_block();
                    const volatile double *q = Xc + ((size_t)(t - 1) * n + jl) * d;
                    xp[0] = q[0]; xp[1] = q[1];
                }
                const double pr_n = prior_prev<DM>(p, c, t, zc, inv, xn, xp);
                const double pr_o = prior_prev<DM>(p, c, t, zc, inv, xo, xp);
                // ---- resolve the block's nodes in order ----
                double my_ratio = 0.0;
                for (int jj = 0; jj < jend; jj++) {
                    int acc = 0;
                    if (lane == jj) {
                        double lp_new = __dsub_rn(lo_n + hi_n, pr_n), lp_old = __dsub_rn(lo_o + hi_o, pr_o);
                        if (t < T - 1) {
                            lp_new = __dsub_rn(lp_new, nn);
                            lp_old = __dsub_rn(lp_old, no);
                        }
                        my_ratio = __dsub_rn(lp_new, lp_old);
                        acc = (logu >= my_ratio) ? 0 : 1; // metropolis.py:50 (NaN accepts)
                        nonfinite |= !(my_ratio == my_ratio) || my_ratio - my_ratio != 0.0;
                        full_acc += acc ? lo_n : lo_o; // dyads {i < j} at the kept state
                    }
                    acc = __shfl_sync(kFull, acc, jj);
                    if (acc) { // the rows behind trade node jj's old terms for its new ones (it is below them)
                        mask |= 1u << jj;
                        double xin[DM], xio[DM];
                        load_pos<DM>(st_prop + jj * d, d, xin);
                        load_pos<DM>(Xt + (size_t)(jb + jj) * d, d, xio);
                        const double ri = kDir ? __ldg(rinv + jb + jj) : 0.0;
                        const double yr = ymask(wrb, jj), yc = ymask(wcb, jj);
                        const double dn = dyad<LK, DM>(xin, ri, xn, rj, yr, yc, b0, b1, d) -
                                          dyad<LK, DM>(xio, ri, xn, rj, yr, yc, b0, b1, d);
                        const double dd = dyad<LK, DM>(xin, ri, xo, rj, yr, yc, b0, b1, d) -
                                          dyad<LK, DM>(xio, ri, xo, rj, yr, yc, b0, b1, d);
                        if (lane > jj) { lo_n += dn; lo_o += dd; }
                    }
                }
                // ---- commit, bookkeeping ----
                const int my_acc = (mask >> lane) & 1u;
                if (vrow) {
                    if (my_acc) { Xt[(size_t)jl * d] = xn[0]; Xt[(size_t)jl * d + 1] = xn[1]; }
                    if (p.ratio) p.ratio[gs] = my_ratio;
                    if (p.accepted) p.accepted[gs] = my_acc;
                    metropolis_bookkeep(my_step, my_nacc, my_nsteps, my_until, p.tune, p.tune_interval, my_acc, false);
                    p.step[gs] = my_step; p.nacc[gs] = my_nacc; p.nsteps[gs] = my_nsteps; p.until[gs] = my_until;
                }
                __threadfence_block();
                __syncwarp();
                if (lane == 0) progress[t] = jb + jend;
            }
            asm volatile("bar.sync %0, 64;" ::"r"(1 + t) : "memory"); // the commits are visible to the second warp
        }
    }
    if (nonfinite) atomicOr(p.flags, 1u);
    __syncthreads();
    if (p.ll_cur) { // full-network log-likelihood of the post-sweep state, summed in slice order
        full_acc = warp_sum(full_acc);
        double *wsum = stage_base;
        if (lane == 0 && half == 0) wsum[t] = full_acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double sacc = 0.0;
            for (int w = 0; w < T; w++) sacc += wsum[w];
            p.ll_cur[c] = sacc;
        }
        __syncthreads();
    }
    if (p.fuse_center) { // X -= mean(X, axis=(0,1)), numpy's serial order (bit-identical to k_center)
        double *mean = stage_base;
        if ((int)threadIdx.x < d) {
            const size_t rows = (size_t)T * n;
            double sacc = 0.0;
            for (size_t r = 0; r < rows; r++) sacc = __dadd_rn(sacc, Xc[r * d + threadIdx.x]);
            mean[threadIdx.x] = __ddiv_rn(sacc, (double)rows);
        }
        __syncthreads();
        for (size_t e = threadIdx.x; e < chain_elems; e += blockDim.x) Xg[e] = __dsub_rn(Xc[e], mean[e % d]);
    } else {
        for (size_t e = threadIdx.x; e < chain_elems; e += blockDim.x) Xg[e] = Xc[e];
    }
}

template <int LK, int MAXT>
static cudaError_t cbp_launch_t(const SweepParams &p, cudaStream_t stream)
{
    const size_t smem = cbp_smem_bytes(p.net.T, p.net.n, p.net.d);
    auto kern = k_sweep_cbp<LK, MAXT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<p.C, 64 * p.net.T, smem, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t cbp_launch(const SweepParams &p, bool directed, cudaStream_t stream)
{
    if (p.net.T <= 10)
        return directed ? cbp_launch_t<kDirected, 640>(p, stream) : cbp_launch_t<kUndirected, 640>(p, stream);
    return directed ? cbp_launch_t<kDirected, 960>(p, stream) : cbp_launch_t<kUndirected, 960>(p, stream);
}

} // namespace dlsm
