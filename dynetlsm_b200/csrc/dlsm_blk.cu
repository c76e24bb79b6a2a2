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
#include "dlsm_blk.h"

#include <cstring>

namespace dlsm {

constexpr int kBlkMaxTeam = 128; // warps of a cluster

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

// the dyad {row node j, column node i}: both directions for the directed model
//   yr: bit Y[j, i] (j sends), yc: bit Y[i, j] (i sends), as y - 1/2
template <int LK, int DM>
__device__ __forceinline__ double dyad(const double (&xi)[DM], double ri, const double (&xj)[DM], double rj,
                                       double yr, double yc, double b0, double b1, int d)
{
    const double dist = fast_dist<DM>(xi, xj, d);
    if (LK == kUndirected) return logit_term(yr, b0 - dist);
    return logit_term(yr, eta_directed(b0, b1, dist, ri, rj)) + logit_term(yc, eta_directed(b0, b1, dist, rj, ri));
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
        mbar_wait_cta(smem_addr(&s_bar[blk & 1]), (uint32_t)((blk >> 1) & 1));

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
        cl_sync(); // ---- 3. A: partial sums and the in-block table are in the leader's shared memory ----

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
        }
        // ---- 5. every CTA waits for the decisions of this block (local poll) and commits them ----
        while (cl_ld_acquire_local_s32(smem_addr(&s_ready)) < blk + 1) { __nanosleep(DLSM_SPIN_NS); }
        if (warp == 0 && vrow && ((unsigned)s_mask >> lane) & 1u) {
#pragma unroll
            for (int k = 0; k < DM; k++) if (k < d) Xt[(size_t)jl * d + k] = xn[k];
        }
        __syncthreads();
    }
    if (nonfinite) atomicOr(p.flags, 1u);
    cl_sync(); // no CTA may exit while a peer can still address its shared memory
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
                    if (MINB == 1) { // a whole SM per chain: four columns per trip, eight independent chains
                        double c_n = 0.0, c_o = 0.0, e_n = 0.0, e_o = 0.0;
                        for (; ii + 3 < cnt; ii += 4) {
                            column(base + ii, rw[k], cw[k], a_n, a_o);
                            column(base + ii + 1, rw[k], cw[k], b_n, b_o);
                            column(base + ii + 2, rw[k], cw[k], c_n, c_o);
                            column(base + ii + 3, rw[k], cw[k], e_n, e_o);
                        }
                        a_n += c_n; a_o += c_o; b_n += e_n; b_o += e_o;
                    }
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
    if (warps <= 10) return cb_launch_t<LK, D, XS, 320, 1>(p, warps, stream); // up to 204 registers
    return cb_launch_t<LK, D, XS, 512, 1>(p, warps, stream);
}

cudaError_t cb_launch(const SweepParams &p, bool directed, cudaStream_t stream)
{
    const size_t max_smem = 227 * 1024;
    const bool xs = cb_smem_bytes(p.net.T, p.net.n, p.net.d, true) <= max_smem;
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
