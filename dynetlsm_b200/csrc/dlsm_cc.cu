// dlsm_cc.cu -- k_sweep_cc2: the batch-parallel case-control sweep of large sparse networks
// (cfg 5: n = 50 000, ~240 list entries per node), second generation.
//
// The first kernel (k_sweep_cc, dlsm_kernels.cuh) spends a run of ~11 mutually independent nodes in
// ~10 us, almost all of it L2 latency in a CHAIN of dependent accesses per node: wavefront flag ->
// neighbour slice's position -> degrees -> list indices -> gathered positions -> radii.  Here every
// access that does not depend on the sweep itself is hoisted to once per 32-node block and made
// coalesced:
//   * the block's list indices (in/out edges, in/out controls, degrees: 32 contiguous rows each)
//     are copied into shared memory by the whole CTA while warp 0 stages the proposals;
//   * the wavefront is checked once per block (slice t-1 must have finished the block), so the
//     "previous slice" prior terms of all 32 nodes are computed lane-parallel during staging;
//   * positions and reciprocal radii are gathered as ONE 256-bit record {x, y, 1/r, 0} per list
//     entry (k_pack_gather before the sweep; accepted moves update the record).
// A node evaluation is then: shared-memory indices -> one round of gathers -> arithmetic.  The runs
// of independent nodes, their order of accumulation and hence the decisions are those of k_sweep_cc
// (bit-identical; tests/test_gpu_edge_cases.py::test_case_control_sweep_all_mappings).
#include "dlsm_kernels.cuh"
#include "dlsm_cc.h"

#include <cstring>

namespace dlsm {

// 256-bit load through the coherent path: the records are rewritten by this CTA during the sweep, so
// the read-only (.nc) variant used by k_full is not allowed here
__device__ __forceinline__ void ld256c(const double *p, double &a, double &b, double &c, double &d)
{
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p) : "memory");
}

struct CcLayout {
    size_t prop, logu, nn, no, prn, pro, acc, dep, deg, ie, oe, ci, co, total; // in 4-byte words from the base
};
__host__ __device__ inline CcLayout cc_layout(int max_in, int max_out, int nc)
{
    CcLayout L;
    size_t o = 0;
    L.prop = o; o += 32 * 2 * 2;        // 32 x double2
    L.logu = o; o += 64; L.nn = o; o += 64; L.no = o; o += 64; L.prn = o; o += 64; L.pro = o; o += 64;
    L.acc = o; o += 32; L.dep = o; o += 32; L.deg = o; o += 64;
    L.ie = o; o += (size_t)32 * max_in; L.oe = o; o += (size_t)32 * max_out;
    L.ci = o; o += (size_t)32 * nc; L.co = o; o += (size_t)32 * nc;
    L.total = (o + 3) & ~(size_t)3;
    return L;
}

// grid = C*T (atomic ticket), block = 512; d = 2
__global__ void __launch_bounds__(512, 1) k_sweep_cc2(const SweepParams p, double *G, int *progress_g,
                                                      unsigned int *ticket, const int32_t *dep_all)
{
    constexpr int DM = 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_ticket;
    const int T = p.net.T, n = p.net.n, d = 2, nc = p.net.n_control;
    const int max_in = p.net.max_in, max_out = p.net.max_out;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    if (threadIdx.x == 0) s_ticket = (int)atomicAdd(ticket, 1u);
    __syncthreads();
    const int c = s_ticket / T, t = s_ticket % T;
    const CcLayout L = cc_layout(max_in, max_out, nc);
    int32_t *sw = reinterpret_cast<int32_t *>(smem_raw);
    double *st_prop = reinterpret_cast<double *>(sw + L.prop);
    double *st_logu = reinterpret_cast<double *>(sw + L.logu), *st_nn = reinterpret_cast<double *>(sw + L.nn);
    double *st_no = reinterpret_cast<double *>(sw + L.no), *st_prn = reinterpret_cast<double *>(sw + L.prn);
    double *st_pro = reinterpret_cast<double *>(sw + L.pro);
    int *st_acc = sw + L.acc, *st_dep = sw + L.dep, *s_deg = sw + L.deg;
    int *s_ie = sw + L.ie, *s_oe = sw + L.oe, *s_ci = sw + L.ci, *s_co = sw + L.co;

    double *Xchain = p.X + (size_t)c * T * n * d;
    double *Xt = Xchain + (size_t)t * n * d;
    double *Gt = G + ((size_t)c * T + t) * n * 4;
    int *prog = progress_g + (size_t)c * T;
    const int32_t *dep = dep_all + ((size_t)(p.net.ctrl_per_chain ? c : 0) * T + t) * n;
    const size_t ctrl_slice = ((size_t)(p.net.ctrl_per_chain ? c : 0) * T + t) * n;
    const double b0 = p.intercept[c * 2 + 0], b1 = p.intercept[c * 2 + 1];
    const uint32_t chain_id = (uint32_t)c + p.chain_offset;
    bool nonfinite = false;

    for (int jb = 0; jb < n; jb += 32) {
        const int jend = (n - jb) < 32 ? (n - jb) : 32;
        const int jl = jb + lane;
        const bool mine = (warp == 0) && (lane < jend);
        const size_t gs = ((size_t)c * T + t) * n + (lane < jend ? jl : jb);
        double my_step = 0.0;
        int my_nacc = 0, my_nsteps = 0, my_until = 0;
        if (warp == 0) {
            // ---- staging: proposals, uniforms, all prior terms (lane = node) ----
            if (t > 0) { // slice t-1 must have finished this block (one poll per 32 nodes)
                if (lane == 0) while (ld_acquire_gpu(prog + t - 1) < jb + jend) { __nanosleep(DLSM_SPIN_NS); }
                __syncwarp();
            }
            if (mine) {
                double eps[DM], x0[DM], x[DM], logu;
                load_pos<DM>(Xt + (size_t)jl * d, d, x0);
                my_step = p.step[gs]; my_nacc = p.nacc[gs]; my_nsteps = p.nsteps[gs]; my_until = p.until[gs];
                if (p.eps) {
                    eps[0] = p.eps[gs * d]; eps[1] = p.eps[gs * d + 1];
                    logu = p.logu[gs];
                } else {
                    latent_draws<DM>(p.seed, (uint32_t)(t * n + jl), p.sweep, chain_id, d, eps, logu);
                }
                x[0] = __dadd_rn(x0[0], __dmul_rn(my_step, eps[0]));
                x[1] = __dadd_rn(x0[1], __dmul_rn(my_step, eps[1]));
                st_prop[lane * 2] = x[0]; st_prop[lane * 2 + 1] = x[1];
                st_logu[lane] = logu;
                double inv = (t == 0) ? 1.0 / p.tau_sq : 1.0 / p.sigma_sq;
                int zc = 0;
                if (p.prior != 0) {
                    zc = p.z[((size_t)c * T + t) * n + jl];
                    inv = 1.0 / p.sigma[(size_t)c * p.K + zc];
                }
                double nn = 0.0, no = 0.0;
                if (t < T - 1) { // slice t+1 (another CTA) cannot have touched nodes >= jb yet
                    double xnx[DM];
                    const volatile double *q = Xchain + ((size_t)(t + 1) * n + jl) * d;
                    xnx[0] = q[0]; xnx[1] = q[1];
                    nn = prior_next<DM>(p, c, t, jl, x, xnx);
                    no = prior_next<DM>(p, c, t, jl, x0, xnx);
                }
                st_nn[lane] = nn; st_no[lane] = no;
                double xp[DM] = {0.0, 0.0};
                if (t > 0) {
                    const volatile double *q = Xchain + ((size_t)(t - 1) * n + jl) * d;
                    xp[0] = q[0]; xp[1] = q[1];
                }
                st_prn[lane] = prior_prev<DM>(p, c, t, zc, inv, x, xp);
                st_pro[lane] = prior_prev<DM>(p, c, t, zc, inv, x0, xp);
                st_dep[lane] = dep[jl];
            }
        } else {
            // ---- the block's list rows are contiguous in global memory: coalesced copies ----
            const size_t r0 = (size_t)t * n + jb;
            const int tid = threadIdx.x - 32, nth = blockDim.x - 32;
            for (int e = tid; e < jend * 2; e += nth) s_deg[e] = p.net.deg[r0 * 2 + e];
            for (int e = tid; e < jend * max_in; e += nth) s_ie[e] = p.net.in_edges[r0 * max_in + e];
            for (int e = tid; e < jend * max_out; e += nth) s_oe[e] = p.net.out_edges[r0 * max_out + e];
            const size_t c0 = (ctrl_slice + jb) * nc;
            for (int e = tid; e < jend * nc; e += nth) { s_ci[e] = p.net.ctrl_in[c0 + e]; s_co[e] = p.net.ctrl_out[c0 + e]; }
        }
        __syncthreads();
        int b = 0;
        while (b < jend) {
            // run [b, b+len): the longest run whose members do not read an earlier member
            const int cand = b + lane;
            const bool fits = lane < nwarps && cand < jend && st_dep[cand] <= jb + b;
            const unsigned run = __ballot_sync(kFull, fits);
            const int len = (run == kFull) ? 32 : __ffs(~run) - 1; // >= 1: dep[j] <= j always
            int acc = 0;
            const int jj = b + warp, j = jb + jj;
            if (warp < len) {
                double x[DM], x0[DM];
                load_pos<DM>(st_prop + jj * 2, d, x);
                double rj, pad;
                ld256c(Gt + (size_t)j * 4, x0[0], x0[1], rj, pad);
                const int indeg = s_deg[jj * 2], outdeg = s_deg[jj * 2 + 1];
                const int *ie = s_ie + jj * max_in, *oe = s_oe + jj * max_out;
                const int *ci = s_ci + jj * nc, *co = s_co + jj * nc;
                double e_n = 0.0, e_o = 0.0, ci_n = 0.0, ci_o = 0.0, co_n = 0.0, co_o = 0.0;
                auto eta_pair = [&](int k, bool k_sends, double &vn, double &vo) {
                    double xk[DM], rk, pd;
                    ld256c(Gt + (size_t)k * 4, xk[0], xk[1], rk, pd);
                    const double dn = fast_dist<DM>(xk, x, d);
                    const double dd = fast_dist<DM>(xk, x0, d);
                    const double r_recv = k_sends ? rj : rk, r_send = k_sends ? rk : rj;
                    vn = eta_directed(b0, b1, dn, r_recv, r_send);
                    vo = eta_directed(b0, b1, dd, r_recv, r_send);
                };
                auto edge_list = [&](const int *lst, int len2, bool k_sends) { // order of k_sweep_cc
                    for (int q = lane; q < len2; q += 64) {
                        const int qb = q + 32;
                        const bool two = qb < len2;
                        const int ka = lst[q], kb = lst[two ? qb : q];
                        double an, ao, bn, bo;
                        eta_pair(ka, k_sends, an, ao);
                        eta_pair(kb, k_sends, bn, bo);
                        e_n += logit_term(0.5, an);
                        e_o += logit_term(0.5, ao);
                        if (two) {
                            e_n += logit_term(0.5, bn);
                            e_o += logit_term(0.5, bo);
                        }
                    }
                };
                edge_list(ie, indeg, true);
                edge_list(oe, outdeg, false);
                int m = nc, m_out;
                int cin[4], cout[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int q = u * 32 + lane;
                    cin[u] = q < nc ? ci[q] : 0;
                    cout[u] = q < nc ? co[q] : 0;
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const unsigned bal = __ballot_sync(kFull, u * 32 + lane < nc && cin[u] == -1);
                    if (bal && m == nc) m = u * 32 + __ffs(bal) - 1;
                }
                m_out = m;
#pragma unroll
                for (int u = 0; u < 4; u++) { // the reference reads X[-1] here: flag + stop
                    const unsigned bal = __ballot_sync(kFull, u * 32 + lane < m && cout[u] < 0);
                    if (bal && m_out == m) {
                        m_out = u * 32 + __ffs(bal) - 1;
                        if (lane == 0) atomicOr(p.flags, 2u);
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int q = u * 32 + lane;
                    const bool vi = q < m, vo_ = q < m_out;
                    double an, ao, bn, bo;
                    eta_pair(vi ? cin[u] : j, true, an, ao);
                    eta_pair(vo_ ? cout[u] : j, false, bn, bo);
                    const double la = log1pexp(an), lb = log1pexp(ao), lc = log1pexp(bn), ld = log1pexp(bo);
                    if (vi) { ci_n += la; ci_o += lb; }
                    if (vo_) { co_n += lc; co_o += ld; }
                }
                e_n = warp_sum(e_n); e_o = warp_sum(e_o);
                ci_n = warp_sum(ci_n); ci_o = warp_sum(ci_o);
                co_n = warp_sum(co_n); co_o = warp_sum(co_o);
                const double adj_in = (double)(n - indeg - 1) / (double)m;
                const double adj_out = (double)(n - outdeg - 1) / (double)m_out;
                const double ll_new = (e_n - adj_in * ci_n) - adj_out * co_n;
                const double ll_old = (e_o - adj_in * ci_o) - adj_out * co_o;
                double lp_new = __dsub_rn(ll_new, st_prn[jj]), lp_old = __dsub_rn(ll_old, st_pro[jj]);
                if (t < T - 1) {
                    lp_new = __dsub_rn(lp_new, st_nn[jj]);
                    lp_old = __dsub_rn(lp_old, st_no[jj]);
                }
                const double ratio = __dsub_rn(lp_new, lp_old);
                acc = (st_logu[jj] >= ratio) ? 0 : 1;
                if (lane == 0) {
                    st_acc[jj] = acc;
                    nonfinite |= !(ratio == ratio) || ratio - ratio != 0.0;
                    if (p.ratio) p.ratio[((size_t)c * T + t) * n + j] = ratio;
                }
            }
            __syncthreads(); // every member has read what it needs: commits may start
            if (warp < len && acc && lane < 2) {
                const double v = st_prop[jj * 2 + lane];
                Xt[(size_t)j * 2 + lane] = v;
                Gt[(size_t)j * 4 + lane] = v;
            }
            __syncthreads();
            b += len;
        }
        if (mine) {
            metropolis_bookkeep(my_step, my_nacc, my_nsteps, my_until, p.tune, p.tune_interval, st_acc[lane], false);
            p.step[gs] = my_step; p.nacc[gs] = my_nacc; p.nsteps[gs] = my_nsteps; p.until[gs] = my_until;
            if (p.accepted) p.accepted[gs] = st_acc[lane];
        }
        __syncthreads(); // (also orders the block's commits before the flag below)
        if (threadIdx.x == 0) st_release_gpu(prog + t, jb + jend); // slice t+1 may take this block
    }
    if (nonfinite) atomicOr(p.flags, 1u);
}

// ---------------------------------------------------------------------------------------------
// k_sweep_cc3: the batch-parallel case-control sweep on a 2-CTA CLUSTER per (chain, slice).
// ncu on k_sweep_cc (profiles/r2_baseline_k_sweep_cc_cfg5_ncu_details.txt): 80 CTAs on 148 SMs, a run
// of ~11 independent nodes costs ~21 000 issue slots on its one SM (1 900 instructions per
// node-update) -- ~40 % of the run's 10.7 us is issue time, the rest gather latency and barriers.
// Here the two CTAs of a cluster share every node: CTA 0 walks the node's IN lists (in-edges,
// in-controls), CTA 1 its OUT lists, each on its own SM, and CTA 1 hands its four partial sums (and
// the usable-control count) to CTA 0 through distributed shared memory; CTA 0 decides and commits.
// Two cluster barriers per run take the place of the two CTA barriers.  160 CTAs of 12 warps at
// <= 85 registers: two CTAs fit an SM, so all (chain, slice) clusters of cfg 5 are co-resident.
// Same runs (capped at 12 nodes), same decisions as the sequential sweep; the edge sums are added
// as (in-list total) + (out-list total) instead of one running sum.
// grid = 2*C*T, cluster (2,1,1), block = 384; d = 2, n_control <= 128
// ---------------------------------------------------------------------------------------------
template <bool OUT_SIDE>
__device__ __forceinline__ void cc_side_eval(const SweepParams &p, const double *Xt, const double *rinv, size_t r,
                                             size_t coff, int j, const double (&x)[2], const double (&x0)[2],
                                             double b0, double b1, int lane, double &e_n, double &e_o,
                                             double &c_n, double &c_o, int &m_used, int &deg_side)
{
    const NetView &net = p.net;
    const int nc = net.n_control, d = 2;
    const double rj = __ldg(rinv + j);
    deg_side = net.deg[r * 2 + (OUT_SIDE ? 1 : 0)];
    const int32_t *lst = OUT_SIDE ? net.out_edges + r * net.max_out : net.in_edges + r * net.max_in;
    const int32_t *ci = net.ctrl_in + coff, *co = net.ctrl_out + coff;
    e_n = e_o = c_n = c_o = 0.0;
    auto eta_pair = [&](int k, double &vn, double &vo) {
        double xk[2];
        const double2 v = __ldcg(reinterpret_cast<const double2 *>(Xt + (size_t)k * d));
        xk[0] = v.x; xk[1] = v.y;
        const double rk = __ldg(rinv + k);
        const double dn = fast_dist<2>(xk, x, d), dd = fast_dist<2>(xk, x0, d);
        const double r_recv = OUT_SIDE ? rk : rj, r_send = OUT_SIDE ? rj : rk;
        vn = eta_directed(b0, b1, dn, r_recv, r_send);
        vo = eta_directed(b0, b1, dd, r_recv, r_send);
    };
    for (int q = lane; q < deg_side; q += 32) { // edge list of this side
        double vn, vo;
        eta_pair(lst[q], vn, vo);
        e_n += logit_term(0.5, vn);
        e_o += logit_term(0.5, vo);
    }
    // usable controls = prefix of ctrl_in before its first -1 (:137; :161 tests the IN list while
    // walking the OUT list, so the out side needs it too)
    int m = nc;
    int cin[4], cmy[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const int q = u * 32 + lane;
        cin[u] = q < nc ? ci[q] : 0;
        cmy[u] = OUT_SIDE ? (q < nc ? co[q] : 0) : cin[u];
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const unsigned bal = __ballot_sync(kFull, u * 32 + lane < nc && cin[u] == -1);
        if (bal && m == nc) m = u * 32 + __ffs(bal) - 1;
    }
    int mm = m;
    if (OUT_SIDE) {
#pragma unroll
        for (int u = 0; u < 4; u++) { // the reference reads X[-1] here: flag + stop
            const unsigned bal = __ballot_sync(kFull, u * 32 + lane < m && cmy[u] < 0);
            if (bal && mm == m) {
                mm = u * 32 + __ffs(bal) - 1;
                if (lane == 0) atomicOr(p.flags, 2u);
            }
        }
    }
    m_used = mm;
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const bool ok = u * 32 + lane < mm;
        double vn, vo;
        eta_pair(ok ? cmy[u] : j, vn, vo);
        const double la = log1pexp(vn), lb = log1pexp(vo);
        if (ok) { c_n += la; c_o += lb; }
    }
    e_n = warp_sum(e_n); e_o = warp_sum(e_o);
    c_n = warp_sum(c_n); c_o = warp_sum(c_o);
}

__global__ void __launch_bounds__(384, 2) k_sweep_cc3(const SweepParams p, int *progress_g, unsigned int *ticket,
                                                      const int32_t *dep_all)
{
    constexpr int DM = 2;
    __shared__ __align__(16) double st_prop[64], st_logu[32], st_nn[32], st_no[32], st_prn[32], st_pro[32];
    __shared__ __align__(16) double s_slot[32 * 6]; // CTA 0: the out side's {e_n, e_o, c_n, c_o, m_out, outdeg} per node
    __shared__ int st_acc[32], st_dep[32], s_ticket;
    const int T = p.net.T, n = p.net.n, d = 2, nc = p.net.n_control;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int rank = (int)cl_rank();
    const bool leader = rank == 0;
    if (threadIdx.x == 0 && leader) s_ticket = (int)atomicAdd(ticket, 1u);
    cl_sync();
    const int tk = cl_ld_s32(cl_map(smem_addr(&s_ticket), 0));
    const int c = tk / T, t = tk % T;
    double *Xchain = p.X + (size_t)c * T * n * d;
    double *Xt = Xchain + (size_t)t * n * d;
    int *prog = progress_g + (size_t)c * T;
    const double *rinv = p.rinv + (size_t)c * n;
    const int32_t *dep = dep_all + ((size_t)(p.net.ctrl_per_chain ? c : 0) * T + t) * n;
    const double b0 = p.intercept[c * 2 + 0], b1 = p.intercept[c * 2 + 1];
    const uint32_t chain_id = (uint32_t)c + p.chain_offset;
    const uint32_t a_slot0 = cl_map(smem_addr(s_slot), 0);
    bool nonfinite = false;

    for (int jb = 0; jb < n; jb += 32) {
        const int jend = (n - jb) < 32 ? (n - jb) : 32;
        const int jl = jb + lane;
        const bool mine = (warp == 0) && (lane < jend);
        const size_t gs = ((size_t)c * T + t) * n + (lane < jend ? jl : jb);
        double my_step = 0.0;
        int my_nacc = 0, my_nsteps = 0, my_until = 0;
        if (warp == 0) {
            if (leader && t > 0) { // slice t-1 must have finished this block (one poll per 32 nodes)
                if (lane == 0) while (ld_acquire_gpu(prog + t - 1) < jb + jend) { __nanosleep(DLSM_SPIN_NS); }
                __syncwarp();
            }
            if (mine) { // both CTAs stage the proposals (same inputs, same arithmetic)
                double eps[DM], x0[DM], x[DM], logu;
                const double2 v0 = __ldcg(reinterpret_cast<const double2 *>(Xt + (size_t)jl * d));
                x0[0] = v0.x; x0[1] = v0.y;
                my_step = p.step[gs];
                if (p.eps) {
                    eps[0] = p.eps[gs * d]; eps[1] = p.eps[gs * d + 1];
                    logu = p.logu[gs];
                } else {
                    latent_draws<DM>(p.seed, (uint32_t)(t * n + jl), p.sweep, chain_id, d, eps, logu);
                }
                x[0] = __dadd_rn(x0[0], __dmul_rn(my_step, eps[0]));
                x[1] = __dadd_rn(x0[1], __dmul_rn(my_step, eps[1]));
                st_prop[lane * 2] = x[0]; st_prop[lane * 2 + 1] = x[1];
                st_dep[lane] = dep[jl];
                if (leader) {
                    my_nacc = p.nacc[gs]; my_nsteps = p.nsteps[gs]; my_until = p.until[gs];
                    st_logu[lane] = logu;
                    double inv = (t == 0) ? 1.0 / p.tau_sq : 1.0 / p.sigma_sq;
                    int zc = 0;
                    if (p.prior != 0) {
                        zc = p.z[((size_t)c * T + t) * n + jl];
                        inv = 1.0 / p.sigma[(size_t)c * p.K + zc];
                    }
                    double nn = 0.0, no = 0.0;
                    if (t < T - 1) {
                        double xnx[DM];
                        const volatile double *q = Xchain + ((size_t)(t + 1) * n + jl) * d;
                        xnx[0] = q[0]; xnx[1] = q[1];
                        nn = prior_next<DM>(p, c, t, jl, x, xnx);
                        no = prior_next<DM>(p, c, t, jl, x0, xnx);
                    }
                    st_nn[lane] = nn; st_no[lane] = no;
                    double xp[DM] = {0.0, 0.0};
                    if (t > 0) {
                        const volatile double *q = Xchain + ((size_t)(t - 1) * n + jl) * d;
                        xp[0] = q[0]; xp[1] = q[1];
                    }
                    st_prn[lane] = prior_prev<DM>(p, c, t, zc, inv, x, xp);
                    st_pro[lane] = prior_prev<DM>(p, c, t, zc, inv, x0, xp);
                }
            }
        }
        __syncthreads();
        int b = 0;
        while (b < jend) {
            const int cand = b + lane;
            const bool fits = lane < nwarps && cand < jend && st_dep[cand] <= jb + b;
            const unsigned run = __ballot_sync(kFull, fits);
            const int len = (run == kFull) ? 32 : __ffs(~run) - 1;
            const int jj = b + warp, j = jb + jj;
            double e_n = 0.0, e_o = 0.0, c_n = 0.0, c_o = 0.0;
            int m_used = nc, deg_side = 0;
            double x[DM] = {0.0, 0.0}, x0[DM] = {0.0, 0.0};
            if (warp < len) {
                x[0] = st_prop[jj * 2]; x[1] = st_prop[jj * 2 + 1];
                const double2 v0 = __ldcg(reinterpret_cast<const double2 *>(Xt + (size_t)j * d));
                x0[0] = v0.x; x0[1] = v0.y;
                const size_t r = (size_t)t * n + j;
                const size_t coff = ((size_t)(p.net.ctrl_per_chain ? c : 0) * T * n + r) * nc;
                if (leader) cc_side_eval<false>(p, Xt, rinv, r, coff, j, x, x0, b0, b1, lane, e_n, e_o, c_n, c_o, m_used, deg_side);
                else cc_side_eval<true>(p, Xt, rinv, r, coff, j, x, x0, b0, b1, lane, e_n, e_o, c_n, c_o, m_used, deg_side);
                if (!leader && lane == 0) { // the out side's sums travel to CTA 0
                    const uint32_t s = a_slot0 + (uint32_t)(jj * 6 * sizeof(double));
                    cl_st_f64(s, e_n); cl_st_f64(s + 8, e_o); cl_st_f64(s + 16, c_n); cl_st_f64(s + 24, c_o);
                    cl_st_f64(s + 32, (double)m_used); cl_st_f64(s + 40, (double)deg_side);
                }
            }
            cl_sync(); // both sides have read what they need; the out side's sums are in CTA 0
            int acc = 0;
            if (leader && warp < len) {
                const double *sl = s_slot + jj * 6;
                const double adj_in = (double)(n - deg_side - 1) / (double)m_used;
                const double adj_out = (double)(n - (int)sl[5] - 1) / sl[4];
                const double ll_new = ((e_n + sl[0]) - adj_in * c_n) - adj_out * sl[2];
                const double ll_old = ((e_o + sl[1]) - adj_in * c_o) - adj_out * sl[3];
                double lp_new = __dsub_rn(ll_new, st_prn[jj]), lp_old = __dsub_rn(ll_old, st_pro[jj]);
                if (t < T - 1) {
                    lp_new = __dsub_rn(lp_new, st_nn[jj]);
                    lp_old = __dsub_rn(lp_old, st_no[jj]);
                }
                const double ratio = __dsub_rn(lp_new, lp_old);
                acc = (st_logu[jj] >= ratio) ? 0 : 1;
                if (lane == 0) {
                    st_acc[jj] = acc;
                    nonfinite |= !(ratio == ratio) || ratio - ratio != 0.0;
                    if (p.ratio) p.ratio[((size_t)c * T + t) * n + j] = ratio;
                }
                if (acc && lane < 2) Xt[(size_t)j * 2 + lane] = lane ? x[1] : x[0];
            }
            cl_sync(); // the commits are visible to both CTAs
            b += len;
        }
        if (mine && leader) {
            metropolis_bookkeep(my_step, my_nacc, my_nsteps, my_until, p.tune, p.tune_interval, st_acc[lane], false);
            p.step[gs] = my_step; p.nacc[gs] = my_nacc; p.nsteps[gs] = my_nsteps; p.until[gs] = my_until;
            if (p.accepted) p.accepted[gs] = st_acc[lane];
        }
        __syncthreads();
        if (leader && threadIdx.x == 0) st_release_gpu(prog + t, jb + jend);
    }
    if (nonfinite) atomicOr(p.flags, 1u);
    cl_sync();
}

cudaError_t cc3_launch(const SweepParams &p, int *progress, unsigned int *ticket, const int32_t *dep,
                       cudaStream_t stream, int *max_active)
{
    const size_t CT = (size_t)p.C * p.net.T;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(CT * 2));
    cfg.blockDim = dim3(384);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    if (max_active) return cudaOccupancyMaxActiveClusters(max_active, k_sweep_cc3, &cfg);
    return cudaLaunchKernelEx(&cfg, k_sweep_cc3, p, progress, ticket, dep);
}

size_t cc2_smem_bytes(int max_in, int max_out, int n_control)
{
    return cc_layout(max_in, max_out, n_control).total * 4 + 16;
}

cudaError_t cc2_launch(const SweepParams &p, double *G, int *progress, unsigned int *ticket, const int32_t *dep,
                       cudaStream_t stream)
{
    const size_t CT = (size_t)p.C * p.net.T;
    const size_t smem = cc2_smem_bytes(p.net.max_in, p.net.max_out, p.net.n_control);
    cudaError_t e = cudaFuncSetAttribute(k_sweep_cc2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_sweep_cc2<<<(unsigned)CT, 512, smem, stream>>>(p, G, progress, ticket, dep);
    return cudaGetLastError();
}

} // namespace dlsm
