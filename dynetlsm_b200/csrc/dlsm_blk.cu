// dlsm_blk.cu -- k_sweep_blk: the BLOCK-SPECULATIVE latent-position sweep of one (chain, slice)
// on a thread-block cluster (its own translation unit: it builds in seconds).
//
// The reference updates the nodes of a slice one after the other (sample_latent_positions.py:98-99):
// node j's MH ratio needs l_j(x') - l_j(x) with every earlier node at its NEW position.  The
// per-node kernels (k_sweep, k_sweep_slice_ws, k_sweep_slice_cl) therefore pay a reduction, a
// decision and -- across warps or CTAs -- a synchronisation PER NODE (2 000 x 20 per sweep at cfg 3).
// Here 32 consecutive nodes are resolved per synchronisation, with exactly the sequential result:
//
//   1. every CTA of the cluster stages the block's 32 proposals (same inputs, same arithmetic);
//   2. PARALLEL: lane l of every warp is row node j = jb + l; the warp walks ITS share of the
//      columns i outside the block and accumulates, per lane, sum_i D(x_i, x'_j) and sum_i D(x_i, x_j)
//      (D = the dyad's log-likelihood terms).  No shuffles, no masks: x_i / 1/r_i are broadcast
//      shared-memory loads, the adjacency bits of row j come from a TMA-staged (cp.async.bulk) copy
//      of the block's 32 bit-rows.  The 32 x 32 dyads INSIDE the block are evaluated for the four
//      combinations (column old / new) x (row old / new) and kept individually;
//   3. partial sums and the in-block table travel to the leader CTA through distributed shared
//      memory (st.shared::cluster), one cluster barrier;
//   4. SERIAL, one warp of the leader: node after node, lane j holds l_j(x'), l_j(x) with every
//      in-block column at its old position; an accepted node i pushes its (new - old) terms to the
//      lanes j > i.  ~100 cycles per node instead of a cross-CTA round trip;
//   5. the leader broadcasts the 32 decisions as ONE word; every CTA commits the accepted proposals
//      from its own stage.  Second cluster barrier.
//
// Exact: the decisions are those of the sequential sweep (sums differ in order only, ~1e-13 relative).
#include "dlsm_kernels.cuh"
#include "dlsm_dyad.cuh"
#include "dlsm_blk.h"

#include <cstdio>
#include <cstring>

namespace dlsm {

constexpr int kBlkMaxTeam = 128; // warps of a cluster

// phase timing of the leader CTA's thread 0 (developer build: -DDLSM_BLK_TIMING, printed per cluster)
#ifdef DLSM_BLK_TIMING
#define BLK_T(i) do { if (threadIdx.x == 0) { const long long now_ = clock64(); tacc[i] += now_ - tprev; tprev = now_; } } while (0)
#else
#define BLK_T(i) do { } while (0)
#endif

__device__ __forceinline__ void mbar_expect_tx(uint32_t addr, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}
// bulk asynchronous copy global -> shared memory (TMA engine, UBLKCP), completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_cta(uint32_t addr, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t}" ::"r"(addr), "r"(parity) : "memory");
}
__device__ __forceinline__ void cl_st_s32(uint32_t addr, int v)
{
    asm volatile("st.shared::cluster.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// shared-memory carve-up (doubles unless noted); every CTA uses the same layout
struct BlkLayout {
    size_t x, rinv, prop, x0, logu, nn, no, inv, zc, slots, red, ctab, bits, total;
};
__host__ __device__ inline BlkLayout blk_layout(int n, int d, bool directed, int W)
{
    BlkLayout L;
    size_t o = 0;
    L.x = o; o += ((size_t)n * d + 1) & ~(size_t)1;
    L.rinv = o; o += directed ? (((size_t)n + 1) & ~(size_t)1) : 0;
    L.prop = o; o += 32 * (size_t)d;
    L.x0 = o; o += 32 * (size_t)d;
    L.logu = o; o += 32;
    L.nn = o; o += 32;
    L.no = o; o += 32;
    L.inv = o; o += 32;
    L.zc = o; o += 16;                                   // 32 ints
    L.slots = o; o += (size_t)kBlkMaxTeam * 32 * 2;      // leader: [team][32][2]
    L.red = o; o += 16 * 32 * 2;                         // leader: [warp][32][2] second-stage partial sums
    L.ctab = o; o += 32 * 32 * 4;                        // leader: [column][row][4]
    L.bits = o; o += (size_t)2 * 2 * 32 * W / 2;         // [buffer][row/col][32][W] uint32
    L.total = o;
    return L;
}

template <int LK, int D>
__global__ void __launch_bounds__(512, 1) k_sweep_blk(const SweepParams p, int *progress_g, unsigned int *ticket)
{
    constexpr int DM = (D == 0) ? kMaxD : D;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long s_bar[2]; // adjacency bits of block b land on s_bar[b & 1]
    __shared__ int s_ticket, s_mask, s_ready; // s_ready: blocks whose decisions have reached this CTA
    const int T = p.net.T, n = p.net.n, d = (D == 0) ? p.net.d : D, W = p.net.W;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int rank = (int)cl_rank(), CS = (int)cl_size();
    const int team = CS * nwarps, gw = rank * nwarps + warp;
    const bool leader = rank == 0;
    constexpr bool kDir = LK != kUndirected;
    const BlkLayout L = blk_layout(n, d, kDir, W);
    double *sm = reinterpret_cast<double *>(smem_raw);
    double *Xt = sm + L.x, *s_rinv = sm + L.rinv, *st_prop = sm + L.prop, *st_x0 = sm + L.x0;
    double *st_logu = sm + L.logu, *st_nn = sm + L.nn, *st_no = sm + L.no, *st_inv = sm + L.inv;
    int *st_zc = reinterpret_cast<int *>(sm + L.zc);
    double *slots = sm + L.slots, *ctab = sm + L.ctab, *red = sm + L.red;
    uint32_t *bits = reinterpret_cast<uint32_t *>(sm + L.bits); // [2][2][32][W]

    if (threadIdx.x == 0) {
        if (leader) s_ticket = (int)atomicAdd(ticket, 1u);
        s_mask = 0;
        s_ready = 0;
        mbar_init(smem_addr(&s_bar[0]), 1);
        mbar_init(smem_addr(&s_bar[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cl_sync();
    const int tk = cl_ld_s32(cl_map(smem_addr(&s_ticket), 0));
    const int c = tk / T, t = tk % T;
    double *Xchain = p.X + (size_t)c * T * n * d;
    double *Xg = Xchain + (size_t)t * n * d;
    for (int e = threadIdx.x; e < n * d; e += blockDim.x) Xt[e] = Xg[e];
    if (kDir) {
        const double *rg = p.rinv + (size_t)c * n;
        for (int e = threadIdx.x; e < n; e += blockDim.x) s_rinv[e] = rg[e];
    }
    int *prog = progress_g + (size_t)c * T;
    const double b0 = p.intercept[c * 2 + 0], b1 = p.intercept[c * 2 + 1];
    const uint32_t chain_id = (uint32_t)c + p.chain_offset;
    // this warp's share of the columns
    const int cpw = (n + team - 1) / team;
    const int lo = gw * cpw < n ? gw * cpw : n, hi = (lo + cpw < n) ? lo + cpw : n;
    const uint32_t a_slots0 = cl_map(smem_addr(slots), 0), a_ctab0 = cl_map(smem_addr(ctab), 0);
    bool nonfinite = false;

    // TMA staging of the block's adjacency bit-rows (32 rows x W words, row- and column-major copies)
    auto stage_bits = [&](int blk) {
        const int jb2 = blk * 32;
        const int rows = (n - jb2) < 32 ? (n - jb2) : 32;
        const uint32_t bytes = (uint32_t)rows * W * 4;
        const uint32_t bar = smem_addr(&s_bar[blk & 1]);
        uint32_t *dst = bits + (size_t)(blk & 1) * 2 * 32 * W;
        mbar_expect_tx(bar, kDir ? 2 * bytes : bytes);
        bulk_g2s(smem_addr(dst), p.net.rowbits + ((size_t)t * n + jb2) * W, bytes, bar);
        if (kDir) bulk_g2s(smem_addr(dst + 32 * W), p.net.colbits + ((size_t)t * n + jb2) * W, bytes, bar);
    };
    if (threadIdx.x == 0) stage_bits(0);
    __syncthreads();

    const int nblk = (n + 31) / 32;
#ifdef DLSM_BLK_TIMING
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = clock64();
#endif
    for (int blk = 0; blk < nblk; blk++) {
        const int jb = blk * 32;
        const int jend = (n - jb) < 32 ? (n - jb) : 32;
        // ---- 1. stage the block: every CTA the proposals, the leader also what the decisions need ----
        const int jl = jb + lane;
        const bool mine = (warp == 0) && (lane < jend);
        const size_t gs = ((size_t)c * T + t) * n + (lane < jend ? jl : jb);
        double my_step = 0.0;
        int my_nacc = 0, my_nsteps = 0, my_until = 0;
        if (warp == 0) {
            double eps[DM], x0[DM], x[DM], logu = 0.0;
#pragma unroll
            for (int k = 0; k < DM; k++) { x0[k] = 0.0; x[k] = 0.0; eps[k] = 0.0; }
            if (mine) {
                load_pos<DM>(Xt + (size_t)jl * d, d, x0);
                my_step = p.step[gs];
                if (p.eps) {
#pragma unroll
                    for (int k = 0; k < DM; k++) eps[k] = (k < d) ? p.eps[gs * d + k] : 0.0;
                    logu = p.logu[gs];
                } else {
                    latent_draws<DM>(p.seed, (uint32_t)(t * n + jl), p.sweep, chain_id, d, eps, logu);
                }
#pragma unroll
                for (int k = 0; k < DM; k++) x[k] = (k < d) ? __dadd_rn(x0[k], __dmul_rn(my_step, eps[k])) : 0.0;
            }
#pragma unroll
            for (int k = 0; k < DM; k++)
                if (k < d) { st_prop[lane * d + k] = x[k]; st_x0[lane * d + k] = x0[k]; }
            if (leader && mine) {
                my_nacc = p.nacc[gs]; my_nsteps = p.nsteps[gs]; my_until = p.until[gs];
                st_logu[lane] = logu;
                double inv = (t == 0) ? 1.0 / p.tau_sq : 1.0 / p.sigma_sq;
                int zc = 0;
                if (p.prior != 0) {
                    zc = p.z[((size_t)c * T + t) * n + jl];
                    inv = 1.0 / p.sigma[(size_t)c * p.K + zc];
                }
                st_inv[lane] = inv;
                st_zc[lane] = zc;
                double nn = 0.0, no = 0.0;
                if (t < T - 1) { // slice t+1 (another cluster) cannot have touched nodes >= jb yet
                    double xnx[DM];
                    const volatile double *q = Xchain + ((size_t)(t + 1) * n + jl) * d;
#pragma unroll
                    for (int k = 0; k < DM; k++) xnx[k] = (k < d) ? q[k] : 0.0;
                    nn = prior_next<DM>(p, c, t, jl, x, xnx);
                    no = prior_next<DM>(p, c, t, jl, x0, xnx);
                }
                st_nn[lane] = nn;
                st_no[lane] = no;
            }
        }
        if (threadIdx.x == 0 && blk + 1 < nblk) stage_bits(blk + 1); // the other buffer is free since barrier B
        __syncthreads();
        BLK_T(0); // staging
        mbar_wait_cta(smem_addr(&s_bar[blk & 1]), (uint32_t)((blk >> 1) & 1));
        BLK_T(1); // adjacency bits (TMA)

        // ---- 2. parallel phase: lane = row node jb + lane, this warp's columns ----
        const bool vrow = lane < jend;
        // leader, warp 0: if slice t-1 has already resolved this block (the usual case: it runs ahead),
        // fetch its positions now -- the L2 round trips hide under the parallel phase
        double xp[DM];
#pragma unroll
        for (int k = 0; k < DM; k++) xp[k] = 0.0;
        bool have_xp = (t == 0);
        if (leader && warp == 0 && t > 0 && ld_acquire_gpu(prog + t - 1) >= jb + jend) {
            have_xp = true;
            if (vrow) {
                const volatile double *q = Xchain + ((size_t)(t - 1) * n + jl) * d;
#pragma unroll
                for (int k = 0; k < DM; k++) if (k < d) xp[k] = q[k];
            }
        }
        double xn[DM], xo[DM];
        load_pos<DM>(st_prop + lane * d, d, xn);
        load_pos<DM>(st_x0 + lane * d, d, xo);
        const double rj = kDir ? s_rinv[vrow ? jl : jb] : 0.0;
        const uint32_t *rowb = bits + (size_t)(blk & 1) * 2 * 32 * W + (size_t)lane * W;
        const uint32_t *colb = rowb + 32 * W;
        double acc_n = 0.0, acc_o = 0.0;
        {
            int wi = -1;
            uint32_t wr = 0, wc = 0;
            auto column = [&](int i, double &an, double &ao) {
                if ((i >> 5) != wi) {
                    wi = i >> 5;
                    wr = rowb[wi];
                    wc = kDir ? colb[wi] : 0u;
                }
                double xi[DM];
                load_pos<DM>(Xt + (size_t)i * d, d, xi);
                const double ri = kDir ? s_rinv[i] : 0.0;
                const double yr = ymask(wr, i & 31), yc = ymask(wc, i & 31);
                an += dyad<LK, DM>(xi, ri, xn, rj, yr, yc, b0, b1, d);
                ao += dyad<LK, DM>(xi, ri, xo, rj, yr, yc, b0, b1, d);
            };
            // columns below and above the block (two accumulator pairs for instruction-level parallelism)
            double an2 = 0.0, ao2 = 0.0;
            int i = lo;
            for (; i + 1 < hi; i += 2) {
                if (i + 1 >= jb && i < jb + 32) { // this pair touches the block: one by one
                    if (i < jb || i >= jb + 32) column(i, acc_n, acc_o);
                    if (i + 1 < jb || i + 1 >= jb + 32) column(i + 1, an2, ao2);
                    continue;
                }
                column(i, acc_n, acc_o);
                column(i + 1, an2, ao2);
            }
            if (i < hi && (i < jb || i >= jb + 32)) column(i, acc_n, acc_o);
            acc_n += an2;
            acc_o += ao2;
        }
        {
            const uint32_t s = a_slots0 + (uint32_t)((gw * 32 + lane) * 2 * sizeof(double));
            cl_st_f64(s, vrow ? acc_n : 0.0);
            cl_st_f64(s + 8, vrow ? acc_o : 0.0);
        }
        // the dyads inside the block, column by column: (column old | new) x (row new | old)
        {
            const uint32_t wr = rowb[jb >> 5];
            const uint32_t wc = kDir ? colb[jb >> 5] : 0u;
            for (int ic = gw; ic < jend; ic += team) {
                double xio[DM], xin[DM];
                load_pos<DM>(st_x0 + ic * d, d, xio);
                load_pos<DM>(st_prop + ic * d, d, xin);
                const double ri = kDir ? s_rinv[jb + ic] : 0.0;
                const double yr = ymask(wr, ic), yc = ymask(wc, ic);
                const bool ok = vrow && lane != ic;
                const double on = dyad<LK, DM>(xio, ri, xn, rj, yr, yc, b0, b1, d);
                const double oo = dyad<LK, DM>(xio, ri, xo, rj, yr, yc, b0, b1, d);
                const double nw = dyad<LK, DM>(xin, ri, xn, rj, yr, yc, b0, b1, d);
                const double no = dyad<LK, DM>(xin, ri, xo, rj, yr, yc, b0, b1, d);
                const uint32_t s = a_ctab0 + (uint32_t)((ic * 32 + lane) * 4 * sizeof(double));
                cl_st_f64(s, ok ? on : 0.0);
                cl_st_f64(s + 8, ok ? oo : 0.0);
                cl_st_f64(s + 16, ok ? nw : 0.0);
                cl_st_f64(s + 24, ok ? no : 0.0);
            }
        }
        BLK_T(2); // parallel phase (this warp)
        cl_sync(); // ---- 3. A: partial sums and the in-block table are in the leader's shared memory ----
        BLK_T(3); // cluster barrier A (incl. waiting for the slowest warp of the cluster)

        // ---- 4. serial phase: the leader adds the partial sums with all its warps, one warp decides ----
        if (leader) {
            double r_n = 0.0, r_o = 0.0;
            for (int w = warp; w < team; w += nwarps) { // team order within a warp's share
                r_n += slots[(w * 32 + lane) * 2];
                r_o += slots[(w * 32 + lane) * 2 + 1];
            }
            red[(warp * 32 + lane) * 2] = r_n;
            red[(warp * 32 + lane) * 2 + 1] = r_o;
            __syncthreads();
            BLK_T(4); // second-stage sums
        }
        if (leader && warp == 0) {
            if (!have_xp) { // the whole block of slice t-1 must be final (wavefront at block granularity)
                while (ld_acquire_gpu(prog + t - 1) < jb + jend) { __nanosleep(DLSM_SPIN_NS); }
                if (vrow) {
                    const volatile double *q = Xchain + ((size_t)(t - 1) * n + jl) * d;
#pragma unroll
                    for (int k = 0; k < DM; k++) if (k < d) xp[k] = q[k];
                }
            }
            double A_n = 0.0, A_o = 0.0, pr_n = 0.0, pr_o = 0.0, nn = 0.0, no = 0.0, logu = 0.0, my_ratio = 0.0;
            if (vrow) {
                const double inv = st_inv[lane];
                const int zc = st_zc[lane];
                pr_n = prior_prev<DM>(p, c, t, zc, inv, xn, xp);
                pr_o = prior_prev<DM>(p, c, t, zc, inv, xo, xp);
                nn = st_nn[lane]; no = st_no[lane]; logu = st_logu[lane];
                for (int w = 0; w < nwarps; w++) { // second stage, warp order
                    A_n += red[(w * 32 + lane) * 2];
                    A_o += red[(w * 32 + lane) * 2 + 1];
                }
                double c_n = 0.0, c_o = 0.0, e_n = 0.0, e_o = 0.0;
                for (int ic = 0; ic + 1 < jend; ic += 2) { // every in-block column at its OLD position
                    c_n += ctab[(ic * 32 + lane) * 4];
                    c_o += ctab[(ic * 32 + lane) * 4 + 1];
                    e_n += ctab[((ic + 1) * 32 + lane) * 4];
                    e_o += ctab[((ic + 1) * 32 + lane) * 4 + 1];
                }
                if (jend & 1) {
                    c_n += ctab[((jend - 1) * 32 + lane) * 4];
                    c_o += ctab[((jend - 1) * 32 + lane) * 4 + 1];
                }
                A_n += c_n + e_n;
                A_o += c_o + e_o;
            }
            unsigned mask = 0u;
            for (int jj = 0; jj < jend; jj++) {
                int acc = 0;
                if (lane == jj) {
                    double lp_new = __dsub_rn(A_n, pr_n), lp_old = __dsub_rn(A_o, pr_o);
                    if (t < T - 1) {
                        lp_new = __dsub_rn(lp_new, nn);
                        lp_old = __dsub_rn(lp_old, no);
                    }
                    my_ratio = __dsub_rn(lp_new, lp_old);
                    acc = (logu >= my_ratio) ? 0 : 1; // metropolis.py:50 (NaN accepts)
                    nonfinite |= !(my_ratio == my_ratio) || my_ratio - my_ratio != 0.0;
                }
                acc = __shfl_sync(kFull, acc, jj);
                if (acc) { // the later rows trade column jj's old terms for its new ones
                    mask |= 1u << jj;
                    if (lane > jj && vrow) {
                        const double *e = ctab + (jj * 32 + lane) * 4;
                        A_n += e[2] - e[0];
                        A_o += e[3] - e[1];
                    }
                }
            }
            // the decisions leave first (one word per CTA, then its ready counter with release
            // semantics): the other CTAs start the next block while the bookkeeping below still runs
            if (lane < CS) {
                cl_st_s32(cl_map(smem_addr(&s_mask), (uint32_t)lane), (int)mask);
                cl_st_release_s32(cl_map(smem_addr(&s_ready), (uint32_t)lane), blk + 1);
            }
            BLK_T(5); // wavefront wait + priors + serial resolve + broadcast
            const int my_acc = (mask >> lane) & 1u;
            if (vrow) {
                if (my_acc) {
#pragma unroll
                    for (int k = 0; k < DM; k++) if (k < d) Xg[(size_t)jl * d + k] = xn[k];
                }
                if (p.ratio) p.ratio[gs] = my_ratio;
                if (p.accepted) p.accepted[gs] = my_acc;
                metropolis_bookkeep(my_step, my_nacc, my_nsteps, my_until, p.tune, p.tune_interval, my_acc, false);
                p.step[gs] = my_step; p.nacc[gs] = my_nacc; p.nsteps[gs] = my_nsteps; p.until[gs] = my_until;
            }
            __syncwarp();
            if (lane == 0) st_release_gpu(prog + t, jb + jend); // slice t+1 may resolve this block
            BLK_T(6); // bookkeeping + release
        }
        // ---- 5. every CTA waits for the decisions of this block (local poll) and commits them ----
        while (cl_ld_acquire_local_s32(smem_addr(&s_ready)) < blk + 1) { __nanosleep(DLSM_SPIN_NS); }
        if (warp == 0 && vrow && ((unsigned)s_mask >> lane) & 1u) {
#pragma unroll
            for (int k = 0; k < DM; k++) if (k < d) Xt[(size_t)jl * d + k] = xn[k];
        }
        __syncthreads();
        BLK_T(7); // commit + CTA barrier
    }
#ifdef DLSM_BLK_TIMING
    if (threadIdx.x == 0 && (rank == 0 || rank == 1) && (t == 0 || t == T / 2))
        printf("blk timing t=%d rank=%d cycles: stage %lld bits %lld parallel %lld barrierA %lld sums %lld resolve %lld book %lld commit %lld\n",
               t, rank, tacc[0], tacc[1], tacc[2], tacc[3], tacc[4], tacc[5], tacc[6], tacc[7]);
#endif
    if (nonfinite) atomicOr(p.flags, 1u);
    cl_sync(); // no CTA may exit while a peer can still address its shared memory
}

// ---------------------------------------------------------------------------------------------
// k_sweep_blkw: k_sweep_blk with a TWO-BLOCK WINDOW.
// Phase timing of k_sweep_blk at cfg 3 (clock64 probes, profiles/r2d_blk_phase_timing.txt): the
// parallel phase is 58 % of a block's 32 us; staging (7 %), the leader's sums + serial resolve (19-25 %,
// incl. the wait for slice t-1), bookkeeping and the decision round trip keep 95 of the cluster's 96
// warps idle for the rest.  Here the leader's control warp resolves block b WHILE the work warps of the
// whole cluster run the parallel phase of block b+1:
//   * block b+1's rows meet the 32 columns of block b in BOTH states (old / proposed), like the
//     columns of their own block: a second 32 x 32 x 4 table ("previous-block table", kept in the
//     shared memory of CTA 1 and read by the leader through DSMEM); once block b is decided the leader
//     picks, per row, the entries of the kept column states;
//   * warp 0 of every CTA is a control warp: it stages the proposals two blocks ahead (three stage
//     buffers), the leader's also resolves; warps 1-15 are work warps and reduce their partial sums
//     inside the CTA first, so one 32 x 2 vector per CTA travels to the leader (double-buffered);
//   * one cluster barrier per block: it publishes block b+1's partial sums and tables, block b's
//     commits and the staging of block b+2.
// Same decisions as the sequential sweep (exact; the sums differ in order only).
// grid = C*T clusters of CS >= 2 CTAs, block = 512; dynamic smem = blkw_layout().total doubles
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double cl_ld_f64(uint32_t addr)
{
    double v;
    asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}

constexpr int kBlkwMaxCS = 8;
struct BlkwLayout {
    size_t x, rinv, stage, cpart, part, tab, bits, total;
    int stage_doubles;
};
__host__ __device__ inline BlkwLayout blkw_layout(int n, int d, bool directed, int W)
{
    BlkwLayout L;
    size_t o = 0;
    L.x = o; o += ((size_t)n * d + 1) & ~(size_t)1;
    L.rinv = o; o += directed ? (((size_t)n + 1) & ~(size_t)1) : 0;
    // one stage: prop[32 d] x0[32 d] logu[32] nn[32] no[32] inv[32] step[32] | zc nacc nsteps until (4 x 32 ints)
    L.stage_doubles = 64 * d + 5 * 32 + 64;
    L.stage = o; o += (size_t)3 * L.stage_doubles;
    L.cpart = o; o += 15 * 32 * 4;                       // this CTA's work warps: [warp][32][4] {all new, all old, below new, below old}
    L.part = o; o += (size_t)2 * kBlkwMaxCS * 32 * 4;    // leader: [buffer][CTA][32][4]
    L.tab = o; o += (size_t)2 * 32 * 32 * 4;             // CTA 0: own-block tables, CTA 1: previous-block tables
    L.bits = o; o += (size_t)2 * 2 * 32 * W / 2;         // [buffer][row/col][32][W] uint32
    L.total = o;
    return L;
}

template <int LK, int D>
__global__ void __launch_bounds__(512, 1) k_sweep_blkw(const SweepParams p, int *progress_g, unsigned int *ticket,
                                                      double *ll_slices)
{
    constexpr int DM = (D == 0) ? kMaxD : D;
    constexpr bool kDir = LK != kUndirected;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long s_bar[2];
    __shared__ int s_ticket, s_mask, s_ready;
    const int T = p.net.T, n = p.net.n, d = (D == 0) ? p.net.d : D, W = p.net.W;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int rank = (int)cl_rank(), CS = (int)cl_size();
    const bool leader = rank == 0;
    const int nwork = nwarps - 1, teamW = CS * nwork, ww = rank * nwork + (warp - 1); // work warps of the cluster
    const BlkwLayout L = blkw_layout(n, d, kDir, W);
    double *sm = reinterpret_cast<double *>(smem_raw);
    double *Xt = sm + L.x, *s_rinv = sm + L.rinv, *cpart = sm + L.cpart, *part = sm + L.part, *tab = sm + L.tab;
    uint32_t *bits = reinterpret_cast<uint32_t *>(sm + L.bits); // [2][2][32][W]
    auto st_base = [&](int k) { return sm + L.stage + (size_t)(k % 3) * L.stage_doubles; };
    auto st_ints = [&](int k) { return reinterpret_cast<int *>(st_base(k) + 64 * d + 160); };

    if (threadIdx.x == 0) {
        if (leader) s_ticket = (int)atomicAdd(ticket, 1u);
        s_mask = 0;
        s_ready = 0;
        mbar_init(smem_addr(&s_bar[0]), 1);
        mbar_init(smem_addr(&s_bar[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cl_sync();
    const int tk = cl_ld_s32(cl_map(smem_addr(&s_ticket), 0));
    const int c = tk / T, t = tk % T;
    double *Xchain = p.X + (size_t)c * T * n * d;
    double *Xg = Xchain + (size_t)t * n * d;
    for (int e = threadIdx.x; e < n * d; e += blockDim.x) Xt[e] = Xg[e];
    if (kDir) {
        const double *rg = p.rinv + (size_t)c * n;
        for (int e = threadIdx.x; e < n; e += blockDim.x) s_rinv[e] = rg[e];
    }
    int *prog = progress_g + (size_t)c * T;
    const double b0 = p.intercept[c * 2 + 0], b1 = p.intercept[c * 2 + 1];
    const uint32_t chain_id = (uint32_t)c + p.chain_offset;
    const int cpw = (n + teamW - 1) / teamW; // this work warp's share of the columns
    const int lo = (warp > 0 && ww * cpw < n) ? ww * cpw : n, hi = (lo + cpw < n) ? lo + cpw : n;
    const uint32_t a_part0 = cl_map(smem_addr(part), 0);
    const uint32_t a_tab_own = cl_map(smem_addr(tab), 0), a_tab_prev = cl_map(smem_addr(tab), 1);
    const int nblk = (n + 31) / 32;
    bool nonfinite = false;
    unsigned mask_prev = 0u; // leader's control warp: decisions of the block before the one being resolved
    double slice_ll = 0.0;   // ... and the dyads {i < j} of the slice at the states the sweep leaves them in

    // TMA staging of a block's adjacency bit-rows (32 rows x W words, row- and column-major copies)
    auto stage_bits = [&](int blk) {
        const int jb2 = blk * 32;
        const int rows = (n - jb2) < 32 ? (n - jb2) : 32;
        const uint32_t bytes = (uint32_t)rows * W * 4;
        const uint32_t bar = smem_addr(&s_bar[blk & 1]);
        uint32_t *dst = bits + (size_t)(blk & 1) * 2 * 32 * W;
        mbar_expect_tx(bar, kDir ? 2 * bytes : bytes);
        bulk_g2s(smem_addr(dst), p.net.rowbits + ((size_t)t * n + jb2) * W, bytes, bar);
        if (kDir) bulk_g2s(smem_addr(dst + 32 * W), p.net.colbits + ((size_t)t * n + jb2) * W, bytes, bar);
    };

    // control warp: proposals (every CTA) and what the decisions need (leader) of block k
    auto stage_block = [&](int k) {
        const int jb = k * 32, jend = (n - jb) < 32 ? (n - jb) : 32, jl = jb + lane;
        const bool mine = lane < jend;
        const size_t gs = ((size_t)c * T + t) * n + (mine ? jl : jb);
        double *sb = st_base(k);
        int *si = st_ints(k);
        double eps[DM], x0[DM], x[DM], logu = 0.0, my_step = 0.0;
#pragma unroll
        for (int q = 0; q < DM; q++) { x0[q] = 0.0; x[q] = 0.0; eps[q] = 0.0; }
        if (mine) {
            load_pos<DM>(Xt + (size_t)jl * d, d, x0);
            my_step = p.step[gs];
            if (p.eps) {
#pragma unroll
                for (int q = 0; q < DM; q++) eps[q] = (q < d) ? p.eps[gs * d + q] : 0.0;
                logu = p.logu[gs];
            } else {
                latent_draws<DM>(p.seed, (uint32_t)(t * n + jl), p.sweep, chain_id, d, eps, logu);
            }
#pragma unroll
            for (int q = 0; q < DM; q++) x[q] = (q < d) ? __dadd_rn(x0[q], __dmul_rn(my_step, eps[q])) : 0.0;
        }
#pragma unroll
        for (int q = 0; q < DM; q++)
            if (q < d) { sb[lane * d + q] = x[q]; sb[32 * d + lane * d + q] = x0[q]; }
        if (leader && mine) {
            sb[64 * d + lane] = logu;
            double inv = (t == 0) ? 1.0 / p.tau_sq : 1.0 / p.sigma_sq;
            int zc = 0;
            if (p.prior != 0) {
                zc = p.z[((size_t)c * T + t) * n + jl];
                inv = 1.0 / p.sigma[(size_t)c * p.K + zc];
            }
            double nn = 0.0, no = 0.0;
            if (t < T - 1) { // slice t+1 (another cluster) cannot have touched nodes >= jb yet
                double xnx[DM];
                const volatile double *q2 = Xchain + ((size_t)(t + 1) * n + jl) * d;
#pragma unroll
                for (int q = 0; q < DM; q++) xnx[q] = (q < d) ? q2[q] : 0.0;
                nn = prior_next<DM>(p, c, t, jl, x, xnx);
                no = prior_next<DM>(p, c, t, jl, x0, xnx);
            }
            sb[64 * d + 32 + lane] = nn;
            sb[64 * d + 64 + lane] = no;
            sb[64 * d + 96 + lane] = inv;
            sb[64 * d + 128 + lane] = my_step;
            si[lane] = zc; si[32 + lane] = p.nacc[gs]; si[64 + lane] = p.nsteps[gs]; si[96 + lane] = p.until[gs];
        }
        __syncwarp();
    };

    // work warps: lane = row node of block k; this warp's columns outside blocks k-1 and k, its share of
    // the two tables, the CTA's partial sums
    auto parallel = [&](int k) {
        const int jb = k * 32, jend = (n - jb) < 32 ? (n - jb) : 32, jl = jb + lane;
        const bool vrow = lane < jend;
        const double *sb = st_base(k);
        double xn[DM], xo[DM];
        load_pos<DM>(sb + lane * d, d, xn);
        load_pos<DM>(sb + 32 * d + lane * d, d, xo);
        const double rj = kDir ? s_rinv[vrow ? jl : jb] : 0.0;
        mbar_wait_cta(smem_addr(&s_bar[k & 1]), (uint32_t)((k >> 1) & 1));
        const uint32_t *rowb = bits + (size_t)(k & 1) * 2 * 32 * W + (size_t)lane * W;
        const uint32_t *colb = rowb + 32 * W;
        const int ex_lo = k > 0 ? jb - 32 : jb, ex_hi = jb + 32; // columns that enter through the tables
        double acc_n = 0.0, acc_o = 0.0, low_n = 0.0, low_o = 0.0;
        {
            int wi = -1;
            uint32_t wr = 0, wc = 0;
            auto column = [&](int i, double &an, double &ao) {
                if ((i >> 5) != wi) {
                    wi = i >> 5;
                    wr = rowb[wi];
                    wc = kDir ? colb[wi] : 0u;
                }
                double xi[DM];
                load_pos<DM>(Xt + (size_t)i * d, d, xi);
                const double ri = kDir ? s_rinv[i] : 0.0;
                const double yr = ymask(wr, i & 31), yc = ymask(wc, i & 31);
                an += dyad<LK, DM>(xi, ri, xn, rj, yr, yc, b0, b1, d);
                ao += dyad<LK, DM>(xi, ri, xo, rj, yr, yc, b0, b1, d);
            };
            auto range = [&](int a, int e) { // two columns per trip: two accumulator pairs
                double an2 = 0.0, ao2 = 0.0;
                int i = a;
                for (; i + 1 < e; i += 2) {
                    column(i, acc_n, acc_o);
                    column(i + 1, an2, ao2);
                }
                if (i < e) column(i, acc_n, acc_o);
                acc_n += an2;
                acc_o += ao2;
            };
            range(lo, hi < ex_lo ? hi : ex_lo);   // columns below the window: below every row of the block
            low_n = acc_n; low_o = acc_o;
            range(lo > ex_hi ? lo : ex_hi, hi);   // columns above it
        }
        {
            double *cp = cpart + ((warp - 1) * 32 + lane) * 4;
            cp[0] = vrow ? acc_n : 0.0; cp[1] = vrow ? acc_o : 0.0;
            cp[2] = vrow ? low_n : 0.0; cp[3] = vrow ? low_o : 0.0;
        }
        // the 32 columns of the previous block (slots 0-31) and of this block (32-63), both states each:
        // (column old | new) x (row new | old)
        for (int cc = ww; cc < 64; cc += teamW) {
            const bool prev = cc < 32;
            const int ic = prev ? cc : cc - 32;
            if (prev ? (k == 0) : (ic >= jend)) continue;
            const double *sc = prev ? st_base(k - 1) : sb;
            const int node = (prev ? jb - 32 : jb) + ic;
            double xio[DM], xin[DM];
            load_pos<DM>(sc + 32 * d + ic * d, d, xio);
            load_pos<DM>(sc + ic * d, d, xin);
            const double ri = kDir ? s_rinv[node] : 0.0;
            const uint32_t wr = rowb[node >> 5];
            const uint32_t wc = kDir ? colb[node >> 5] : 0u;
            const double yr = ymask(wr, ic), yc = ymask(wc, ic);
            const bool ok = vrow && (prev || lane != ic);
            const double on = dyad<LK, DM>(xio, ri, xn, rj, yr, yc, b0, b1, d);
            const double oo = dyad<LK, DM>(xio, ri, xo, rj, yr, yc, b0, b1, d);
            const double nw = dyad<LK, DM>(xin, ri, xn, rj, yr, yc, b0, b1, d);
            const double no = dyad<LK, DM>(xin, ri, xo, rj, yr, yc, b0, b1, d);
            const uint32_t s = (prev ? a_tab_prev : a_tab_own) +
                               (uint32_t)((((k & 1) * 32 + ic) * 32 + lane) * 4 * sizeof(double));
            cl_st_f64(s, ok ? on : 0.0);
            cl_st_f64(s + 8, ok ? oo : 0.0);
            cl_st_f64(s + 16, ok ? nw : 0.0);
            cl_st_f64(s + 24, ok ? no : 0.0);
        }
        asm volatile("bar.sync 1, %0;" ::"r"(nwork * 32) : "memory"); // the work warps' partial sums are in cpart
        const int tw = threadIdx.x - 32;
        if (tw < 128) { // one 32 x 4 vector per CTA travels to the leader
            const int l = tw & 31, v = tw >> 5;
            double sacc = 0.0;
            for (int w = 0; w < nwork; w++) sacc += cpart[(w * 32 + l) * 4 + v];
            cl_st_f64(a_part0 + (uint32_t)(((((k & 1) * kBlkwMaxCS + rank) * 32 + l) * 4 + v) * sizeof(double)), sacc);
        }
    };

    // leader's control warp: block b's sums, the serial resolve, the decisions, the bookkeeping
    auto resolve = [&](int b) {
        const int jb = b * 32, jend = (n - jb) < 32 ? (n - jb) : 32, jl = jb + lane;
        const bool vrow = lane < jend;
        const size_t gs = ((size_t)c * T + t) * n + (vrow ? jl : jb);
        const double *sb = st_base(b);
        const int *si = st_ints(b);
        double xn[DM], xo[DM], xp[DM];
        load_pos<DM>(sb + lane * d, d, xn);
        load_pos<DM>(sb + 32 * d + lane * d, d, xo);
#pragma unroll
        for (int q = 0; q < DM; q++) xp[q] = 0.0;
        if (t > 0) { // the whole block of slice t-1 must be final (wavefront at block granularity)
            while (ld_acquire_gpu(prog + t - 1) < jb + jend) { __nanosleep(DLSM_SPIN_NS); }
            if (vrow) {
                const volatile double *q2 = Xchain + ((size_t)(t - 1) * n + jl) * d;
#pragma unroll
                for (int q = 0; q < DM; q++) if (q < d) xp[q] = q2[q];
            }
        }
        double A_n = 0.0, A_o = 0.0, pr_n = 0.0, pr_o = 0.0, nn = 0.0, no = 0.0, logu = 0.0, my_ratio = 0.0;
        double B_n = 0.0, B_o = 0.0, my_ll = 0.0; // B: the part of A that comes from columns below the row
        const double *otab = tab + (size_t)(b & 1) * 32 * 32 * 4; // own-block table (this CTA's shared memory)
        if (vrow) {
            const double inv = sb[64 * d + 96 + lane];
            const int zc = si[lane];
            pr_n = prior_prev<DM>(p, c, t, zc, inv, xn, xp);
            pr_o = prior_prev<DM>(p, c, t, zc, inv, xo, xp);
            nn = sb[64 * d + 32 + lane]; no = sb[64 * d + 64 + lane]; logu = sb[64 * d + lane];
            for (int r = 0; r < CS; r++) { // CTA order
                const double *pp = part + (((b & 1) * kBlkwMaxCS + r) * 32 + lane) * 4;
                A_n += pp[0]; A_o += pp[1]; B_n += pp[2]; B_o += pp[3];
            }
            double c_n = 0.0, c_o = 0.0, e_n = 0.0, e_o = 0.0;
            for (int ic = 0; ic < jend; ic++) { // every own-block column at its OLD position
                const double on = otab[(ic * 32 + lane) * 4], oo = otab[(ic * 32 + lane) * 4 + 1];
                c_n += on; c_o += oo;
                if (ic < lane) { e_n += on; e_o += oo; }
            }
            A_n += c_n; A_o += c_o;
            B_n += e_n; B_o += e_o;
            if (b > 0) { // the previous block's columns at the states block b-1 was left in (table in CTA 1)
                double q_n = 0.0, q_o = 0.0;
                const uint32_t base = a_tab_prev + (uint32_t)((size_t)(b & 1) * 32 * 32 * 4 * sizeof(double));
                for (int pc = 0; pc < 32; pc++) {
                    const uint32_t e = base + (uint32_t)(((pc * 32 + lane) * 4 + (((mask_prev >> pc) & 1u) ? 2 : 0)) * sizeof(double));
                    q_n += cl_ld_f64(e);
                    q_o += cl_ld_f64(e + 8);
                }
                A_n += q_n; A_o += q_o;
                B_n += q_n; B_o += q_o;
            }
        }
        unsigned mask = 0u;
        for (int jj = 0; jj < jend; jj++) {
            int acc = 0;
            if (lane == jj) {
                double lp_new = __dsub_rn(A_n, pr_n), lp_old = __dsub_rn(A_o, pr_o);
                if (t < T - 1) {
                    lp_new = __dsub_rn(lp_new, nn);
                    lp_old = __dsub_rn(lp_old, no);
                }
                my_ratio = __dsub_rn(lp_new, lp_old);
                acc = (logu >= my_ratio) ? 0 : 1; // metropolis.py:50 (NaN accepts)
                nonfinite |= !(my_ratio == my_ratio) || my_ratio - my_ratio != 0.0;
                my_ll = acc ? B_n : B_o; // dyads {i < j} at the kept state, every i < j final
            }
            acc = __shfl_sync(kFull, acc, jj);
            if (acc) { // the later rows trade column jj's old terms for its new ones
                mask |= 1u << jj;
                if (lane > jj && vrow) {
                    const double *e = otab + (jj * 32 + lane) * 4;
                    A_n += e[2] - e[0]; A_o += e[3] - e[1];
                    B_n += e[2] - e[0]; B_o += e[3] - e[1];
                }
            }
        }
        // the decisions leave first (one word per CTA, then its ready counter with release semantics)
        if (lane < CS) {
            cl_st_s32(cl_map(smem_addr(&s_mask), (uint32_t)lane), (int)mask);
            cl_st_release_s32(cl_map(smem_addr(&s_ready), (uint32_t)lane), b + 1);
        }
        const int my_acc = (mask >> lane) & 1u;
        if (vrow) {
            if (my_acc) {
#pragma unroll
                for (int q = 0; q < DM; q++) if (q < d) Xg[(size_t)jl * d + q] = xn[q];
            }
            if (p.ratio) p.ratio[gs] = my_ratio;
            if (p.accepted) p.accepted[gs] = my_acc;
            double my_step = sb[64 * d + 128 + lane];
            int my_nacc = si[32 + lane], my_nsteps = si[64 + lane], my_until = si[96 + lane];
            metropolis_bookkeep(my_step, my_nacc, my_nsteps, my_until, p.tune, p.tune_interval, my_acc, false);
            p.step[gs] = my_step; p.nacc[gs] = my_nacc; p.nsteps[gs] = my_nsteps; p.until[gs] = my_until;
        }
        __syncwarp();
        if (lane == 0) st_release_gpu(prog + t, jb + jend); // slice t+1 may resolve this block
        mask_prev = mask;
        slice_ll += warp_sum(vrow ? my_ll : 0.0);
    };

    // ---- prologue: bits and proposals of blocks 0 and 1, parallel phase of block 0 ----
    if (threadIdx.x == 0) {
        stage_bits(0);
        if (nblk > 1) stage_bits(1);
    }
    if (warp == 0) {
        stage_block(0);
        if (nblk > 1) stage_block(1);
    }
    __syncthreads();
    if (warp > 0) parallel(0);
    cl_sync();

    for (int b = 0; b < nblk; b++) {
        if (warp == 0) {
            // (bit buffer b & 1 was last read by the parallel phase of block b, which ended before the
            //  barrier that closed the previous iteration)
            if (lane == 0 && b + 2 < nblk) stage_bits(b + 2);
            if (leader) resolve(b);
            if (b + 2 < nblk) stage_block(b + 2);
        } else if (b + 1 < nblk) {
            parallel(b + 1);
        }
        // block b's decisions (local poll), commit into this CTA's copy of the slice
        while (cl_ld_acquire_local_s32(smem_addr(&s_ready)) < b + 1) { __nanosleep(DLSM_SPIN_NS); }
        if (warp == 1) {
            const int jb = b * 32, jend = (n - jb) < 32 ? (n - jb) : 32;
            if (lane < jend && (((unsigned)s_mask >> lane) & 1u)) {
                const double *sb = st_base(b);
#pragma unroll
                for (int q = 0; q < DM; q++) if (q < d) Xt[(size_t)(jb + lane) * d + q] = sb[lane * d + q];
            }
        }
        // publishes: block b+1's partial sums and tables, block b's commits, block b+2's proposals
        cl_sync();
    }
    if (nonfinite) atomicOr(p.flags, 1u);
    // the full-network log-likelihood of the state the sweep leaves behind = the sum over the slices
    // (every dyad {i < j} was last evaluated when j was decided, with i already final)
    if (ll_slices && leader && threadIdx.x == 0) ll_slices[(size_t)c * T + t] = slice_ll;
    cl_sync(); // no CTA may exit while a peer can still address its shared memory
}

size_t blkw_smem_bytes(int n, int d, bool directed, int W)
{
    return blkw_layout(n, d, directed, W).total * sizeof(double) + 16;
}

template <int LK, int D>
static cudaError_t blkw_launch_t(const SweepParams &p, int CS, int *progress, unsigned int *ticket,
                                 double *ll_slices, cudaStream_t stream, int *max_active)
{
    const size_t CT = (size_t)p.C * p.net.T;
    const size_t smem = blkw_smem_bytes(p.net.n, p.net.d, LK != kUndirected, p.net.W);
    auto kern = k_sweep_blkw<LK, D>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(CT * CS));
    cfg.blockDim = dim3(512);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    if (max_active) return cudaOccupancyMaxActiveClusters(max_active, kern, &cfg);
    if ((e = cudaMemsetAsync(progress, 0, CT * sizeof(int), stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(ticket, 0, sizeof(unsigned int), stream)) != cudaSuccess) return e;
    return cudaLaunchKernelEx(&cfg, kern, p, progress, ticket, ll_slices);
}

cudaError_t blkw_launch(const SweepParams &p, bool directed, int CS, int *progress, unsigned int *ticket,
                        double *ll_slices, cudaStream_t stream, int *max_active)
{
    const bool d2 = p.net.d == 2;
    if (CS < 2 || CS > kBlkwMaxCS) return cudaErrorInvalidValue;
    if (!directed)
        return d2 ? blkw_launch_t<kUndirected, 2>(p, CS, progress, ticket, ll_slices, stream, max_active)
                  : blkw_launch_t<kUndirected, 0>(p, CS, progress, ticket, ll_slices, stream, max_active);
    return d2 ? blkw_launch_t<kDirected, 2>(p, CS, progress, ticket, ll_slices, stream, max_active)
              : blkw_launch_t<kDirected, 0>(p, CS, progress, ticket, ll_slices, stream, max_active);
}

// ---------------------------------------------------------------------------------------------
// k_sweep_cb: the block-speculative sweep for MANY chains -- one CTA per chain, one warp per time
// slice (the mapping of k_sweep), 32 nodes per step instead of one.
// Lane l of the slice's warp is row node j = jb + l.  The warp walks ALL columns i of the slice
// once per block: x_i is a broadcast shared-memory load, the adjacency bit comes from the lane's
// own bit-row (one 128-bit load per 128 columns), both MH evaluations D(x_i, x'_j), D(x_i, x_j)
// accumulate in the lane -- no warp reductions, no masks, no clamped indices.  The block's own 32
// columns enter at their OLD positions; the nodes are then resolved in order by the same warp:
// lane j decides from its sums, and an accepted node pushes (new - old) terms to the lanes behind
// it (4 dyads per lane per accepted node).  Same decisions as the node-by-node sweep (sums differ
// in order only); the wavefront over the slices advances a block at a time.
// The full-network log-likelihood of the state the sweep leaves behind is tracked as in k_sweep:
// the sums over the columns i < j of the kept variant (lo_*), added up over all node-updates.
// grid = C, block = 32 * min(T, 16); dynamic smem = [T*n*d doubles] + nwarps * 32*d doubles + T ints
// ---------------------------------------------------------------------------------------------
template <int LK, int D, bool XS, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) k_sweep_cb(const SweepParams p)
{
    constexpr int DM = (D == 0) ? kMaxD : D;
    constexpr bool kDir = LK != kUndirected;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int T = p.net.T, n = p.net.n, d = (D == 0) ? p.net.d : D, W = p.net.W;
    const int c = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const size_t chain_elems = (size_t)T * n * d;
    double *Xg = p.X + (size_t)c * chain_elems;
    double *Xc, *stage_base;
    if (XS) {
        Xc = reinterpret_cast<double *>(smem_raw);
        stage_base = Xc + ((chain_elems + 1) & ~(size_t)1);
        for (size_t e = threadIdx.x; e < chain_elems; e += blockDim.x) Xc[e] = Xg[e];
    } else {
        Xc = Xg;
        stage_base = reinterpret_cast<double *>(smem_raw);
    }
    double *st_prop = stage_base + (size_t)warp * 32 * d;
    volatile int *progress = reinterpret_cast<volatile int *>(stage_base + (size_t)nwarps * 32 * d);
    for (int t = threadIdx.x; t < T; t += blockDim.x) progress[t] = 0;
    __syncthreads();

    const double b0 = p.intercept[c * 2 + 0], b1 = p.intercept[c * 2 + 1];
    const double *rinv = kDir ? p.rinv + (size_t)c * n : nullptr;
    const uint32_t chain_id = (uint32_t)c + p.chain_offset;
    bool nonfinite = false;
    double full_acc = 0.0;

    for (int t = warp; t < T; t += nwarps) {
        double *Xt = Xc + (size_t)t * n * d;
        for (int jb = 0; jb < n; jb += 32) {
            const int jend = (n - jb) < 32 ? (n - jb) : 32;
            const bool vrow = lane < jend;
            const int jl = vrow ? jb + lane : jb;
            const size_t gs = ((size_t)c * T + t) * n + jl;
            // ---- lane-parallel preparation ----
            double xn[DM], xo[DM], logu = 0.0, inv = 0.0, nn = 0.0, no = 0.0;
            double my_step = p.step[gs];
            int my_nacc = p.nacc[gs], my_nsteps = p.nsteps[gs], my_until = p.until[gs], zc = 0;
            load_pos<DM>(Xt + (size_t)jl * d, d, xo);
            {
                double eps[DM];
                if (p.eps) {
#pragma unroll
                    for (int k = 0; k < DM; k++) eps[k] = (k < d) ? p.eps[gs * d + k] : 0.0;
                    logu = p.logu[gs];
                } else {
                    latent_draws<DM>(p.seed, (uint32_t)(t * n + jl), p.sweep, chain_id, d, eps, logu);
                }
#pragma unroll
                for (int k = 0; k < DM; k++) {
                    xn[k] = (k < d) ? __dadd_rn(xo[k], __dmul_rn(my_step, eps[k])) : 0.0;
                    if (k < d) st_prop[lane * d + k] = xn[k];
                }
                inv = (t == 0) ? 1.0 / p.tau_sq : 1.0 / p.sigma_sq;
                if (p.prior != 0) {
                    zc = p.z[((size_t)c * T + t) * n + jl];
                    inv = 1.0 / p.sigma[(size_t)c * p.K + zc];
                }
                if (t < T - 1) { // X[t+1, j] is still last sweep's value: slice t+1 trails this one
                    double xnx[DM];
                    const volatile double *q = Xc + ((size_t)(t + 1) * n + jl) * d;
#pragma unroll
                    for (int k = 0; k < DM; k++) xnx[k] = (k < d) ? q[k] : 0.0;
                    nn = prior_next<DM>(p, c, t, jl, xn, xnx);
                    no = prior_next<DM>(p, c, t, jl, xo, xnx);
                }
            }
            __syncwarp();
            // ---- all columns of the slice: sums over i < jb (lo) and i >= jb + 32 (hi) ----
            const double rj = kDir ? rinv[jl] : 0.0;
            const uint32_t *rowb = p.net.rowbits + ((size_t)t * n + jl) * W;
            const uint32_t *colb = kDir ? p.net.colbits + ((size_t)t * n + jl) * W : nullptr;
            double lo_n = 0.0, lo_o = 0.0, hi_n = 0.0, hi_o = 0.0;
            uint32_t wrb = 0, wcb = 0; // the block's own word
            auto column = [&](int i, uint32_t wr, uint32_t wc, double &an, double &ao) {
                double xi[DM];
                load_pos<DM>(Xt + (size_t)i * d, d, xi);
                const double ri = kDir ? __ldg(rinv + i) : 0.0;
                const double yr = ymask(wr, i & 31), yc = ymask(wc, i & 31);
                an += dyad<LK, DM>(xi, ri, xn, rj, yr, yc, b0, b1, d);
                ao += dyad<LK, DM>(xi, ri, xo, rj, yr, yc, b0, b1, d);
            };
            for (int w4 = 0; w4 * 32 < n; w4 += 4) {
                const uint4 r4 = __ldg(reinterpret_cast<const uint4 *>(rowb + w4));
                uint4 c4 = make_uint4(0u, 0u, 0u, 0u);
                if (kDir) c4 = __ldg(reinterpret_cast<const uint4 *>(colb + w4));
                const uint32_t rw[4] = {r4.x, r4.y, r4.z, r4.w}, cw[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int base = (w4 + k) * 32;
                    if (base >= n) break;
                    if (base == jb) { wrb = rw[k]; wcb = cw[k]; continue; }
                    const int cnt = (n - base) < 32 ? (n - base) : 32;
                    double a_n = 0.0, a_o = 0.0, b_n = 0.0, b_o = 0.0;
                    int ii = 0;
                    // (four columns per trip / eight chains per lane on a 204-register build, tried for chains that
                    //  have an SM to themselves: 15.1 ms instead of 2.75 ms at cfg 4 with 128 chains -- removed)
                    for (; ii + 1 < cnt; ii += 2) { // two columns per trip: four independent chains
                        column(base + ii, rw[k], cw[k], a_n, a_o);
                        column(base + ii + 1, rw[k], cw[k], b_n, b_o);
                    }
                    if (ii < cnt) column(base + ii, rw[k], cw[k], a_n, a_o);
                    if (base < jb) { lo_n += a_n + b_n; lo_o += a_o + b_o; }
                    else { hi_n += a_n + b_n; hi_o += a_o + b_o; }
                }
            }
            // the block's own columns at their old positions: below the row -> lo, above -> hi
            for (int ic = 0; ic < jend; ic++) {
                double tn = 0.0, to = 0.0;
                column(jb + ic, wrb, wcb, tn, to);
                if (ic < lane) { lo_n += tn; lo_o += to; }
                else if (ic > lane) { hi_n += tn; hi_o += to; }
            }
            // ---- wavefront: the whole block of slice t-1 must be final ----
            double xp[DM];
#pragma unroll
            for (int k = 0; k < DM; k++) xp[k] = 0.0;
            if (t > 0) {
                while (progress[t - 1] < jb + jend) { __nanosleep(DLSM_SPIN_NS); }
                __threadfence_block();
                const volatile double *q = Xc + ((size_t)(t - 1) * n + jl) * d;
#pragma unroll
                for (int k = 0; k < DM; k++) if (k < d) xp[k] = q[k];
            }
            const double pr_n = prior_prev<DM>(p, c, t, zc, inv, xn, xp);
            const double pr_o = prior_prev<DM>(p, c, t, zc, inv, xo, xp);
            // ---- resolve the block's nodes in order ----
            unsigned mask = 0u;
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
                if (my_acc) {
#pragma unroll
                    for (int k = 0; k < DM; k++) if (k < d) Xt[(size_t)jl * d + k] = xn[k];
                }
                if (p.ratio) p.ratio[gs] = my_ratio;
                if (p.accepted) p.accepted[gs] = my_acc;
                metropolis_bookkeep(my_step, my_nacc, my_nsteps, my_until, p.tune, p.tune_interval, my_acc, false);
                p.step[gs] = my_step; p.nacc[gs] = my_nacc; p.nsteps[gs] = my_nsteps; p.until[gs] = my_until;
            }
            __threadfence_block();
            __syncwarp();
            if (lane == 0) progress[t] = jb + jend;
            __syncwarp();
        }
    }
    if (nonfinite) atomicOr(p.flags, 1u);
    if (p.ll_cur) { // full-network log-likelihood of the post-sweep state, summed in warp order
        full_acc = warp_sum(full_acc);
        __syncthreads();
        double *wsum = stage_base;
        if (lane == 0) wsum[warp] = full_acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double sacc = 0.0;
            for (int w = 0; w < nwarps; w++) sacc += wsum[w];
            p.ll_cur[c] = sacc;
        }
    }
    if (XS) {
        __syncthreads();
        if (p.fuse_center) { // X -= mean(X, axis=(0,1)), numpy's serial order (bit-identical to k_center)
            double *mean = stage_base;
            __syncthreads();
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
}

size_t cb_smem_bytes(int T, int n, int d, bool xs)
{
    const int warps = T < 16 ? T : 16;
    const size_t x = (((size_t)T * n * d + 1) & ~(size_t)1) * sizeof(double);
    return (xs ? x : 0) + (size_t)warps * 32 * d * sizeof(double) + (size_t)T * sizeof(int) + 64;
}

template <int LK, int D, bool XS, int MAXT, int MINB>
static cudaError_t cb_launch_t(const SweepParams &p, int warps, cudaStream_t stream)
{
    const size_t smem = cb_smem_bytes(p.net.T, p.net.n, p.net.d, XS);
    auto kern = k_sweep_cb<LK, D, XS, MAXT, MINB>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<p.C, warps * 32, smem, stream>>>(p);
    return cudaGetLastError();
}

template <int LK, int D, bool XS>
static cudaError_t cb_launch_x(const SweepParams &p, int ctas_per_sm_by_smem, cudaStream_t stream)
{
    const int warps = p.net.T < 16 ? p.net.T : 16;
    if (warps <= 9 && ctas_per_sm_by_smem >= 3) return cb_launch_t<LK, D, XS, 288, 3>(p, warps, stream);
    if (warps <= 10 && ctas_per_sm_by_smem >= 2) return cb_launch_t<LK, D, XS, 320, 2>(p, warps, stream);
    return cb_launch_t<LK, D, XS, 512, 1>(p, warps, stream);
}

cudaError_t cb_launch(const SweepParams &p, bool directed, cudaStream_t stream, bool allow_pair)
{
    const size_t max_smem = 227 * 1024;
    const bool xs = cb_smem_bytes(p.net.T, p.net.n, p.net.d, true) <= max_smem;
    // a chain with an SM to itself: two warps per slice (k_sweep_cbp, dlsm_cbp.cu)
    if (allow_pair && p.C <= 148 && p.net.d == 2 && cbp_applicable(p.net.T, p.net.n, p.net.d))
        return cbp_launch(p, directed, stream);
    int per_sm = (int)(max_smem / (cb_smem_bytes(p.net.T, p.net.n, p.net.d, xs) + 1024));
    if (p.C <= 148) per_sm = 1; // at most one chain per SM: all the registers
    const bool d2 = p.net.d == 2;
#define CB(LK)                                                                                     \
    (d2 ? (xs ? cb_launch_x<LK, 2, true>(p, per_sm, stream) : cb_launch_x<LK, 2, false>(p, per_sm, stream)) \
        : (xs ? cb_launch_x<LK, 0, true>(p, per_sm, stream) : cb_launch_x<LK, 0, false>(p, per_sm, stream)))
    return directed ? CB(kDirected) : CB(kUndirected);
#undef CB
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
size_t blk_smem_bytes(int n, int d, bool directed, int W)
{
    return blk_layout(n, d, directed, W).total * sizeof(double) + 16;
}

template <int LK, int D>
static cudaError_t blk_launch_t(const SweepParams &p, int CS, int nwarps, int *progress, unsigned int *ticket,
                                cudaStream_t stream, int *max_active)
{
    const size_t CT = (size_t)p.C * p.net.T;
    const size_t smem = blk_smem_bytes(p.net.n, p.net.d, LK != kUndirected, p.net.W);
    auto kern = k_sweep_blk<LK, D>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(CT * CS));
    cfg.blockDim = dim3(32 * nwarps);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    if (max_active) return cudaOccupancyMaxActiveClusters(max_active, kern, &cfg);
    if ((e = cudaMemsetAsync(progress, 0, CT * sizeof(int), stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(ticket, 0, sizeof(unsigned int), stream)) != cudaSuccess) return e;
    return cudaLaunchKernelEx(&cfg, kern, p, progress, ticket);
}

cudaError_t blk_launch(const SweepParams &p, bool directed, int CS, int nwarps, int *progress,
                       unsigned int *ticket, cudaStream_t stream, int *max_active)
{
    const bool d2 = p.net.d == 2;
    if (!directed)
        return d2 ? blk_launch_t<kUndirected, 2>(p, CS, nwarps, progress, ticket, stream, max_active)
                  : blk_launch_t<kUndirected, 0>(p, CS, nwarps, progress, ticket, stream, max_active);
    return d2 ? blk_launch_t<kDirected, 2>(p, CS, nwarps, progress, ticket, stream, max_active)
              : blk_launch_t<kDirected, 0>(p, CS, nwarps, progress, ticket, stream, max_active);
}

} // namespace dlsm
