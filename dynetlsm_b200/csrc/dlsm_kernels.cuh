// dlsm_kernels.cuh -- the sm_100a kernels of the MH-within-Gibbs hot path.
//
// Mapping (DESIGN.md has the full rationale):
//   k_sweep      one CTA per chain, one warp per time slice, slices run as a bit-exact wavefront:
//                slice t updates node j once slice t-1 has finished node j (then X[t-1,j] is new
//                and X[t+1,j] is still old, exactly the reference's lexicographic (t,j) order,
//                sample_latent_positions.py:98-99).  Positions live in shared memory when the
//                chain fits (T*n*d*8 B), otherwise they stay in global/L2.  The j-reduction of the
//                pairwise logistic log-likelihood is 32 lanes wide, fp64, warp-shuffle tree; both
//                MH evaluations (proposal and current) share one pass over the row.
//   k_full_*     O(T n^2) full-network log-likelihood for two parameter variants in one pass
//                (intercept / radii MH evaluate logp(x) and logp(x0) on the same distances).
//   k_ffbs       one warp per (chain, node): emission densities, backward messages, forward draws.
#pragma once
#include "dlsm_device.cuh"

#ifndef DLSM_SPIN_NS
#define DLSM_SPIN_NS 20
#endif

namespace dlsm {

enum Lik : int { kUndirected = 0, kDirected = 1, kCaseControl = 2 };

struct NetView {
    int T, n, d, W;             // W = 32-bit words per adjacency row (multiple of 4)
    const uint32_t *rowbits;    // [T][n][W]  bit i of row j = Y[t, j, i]
    const uint32_t *colbits;    // [T][n][W]  bit i of row j = Y[t, i, j]   (directed only)
    // case-control lists (int32; -1 sentinels in the control sets)
    const int32_t *deg;         // [T][n][2]  in, out
    const int32_t *in_edges;    // [T][n][max_in]
    const int32_t *out_edges;   // [T][n][max_out]
    const int32_t *ctrl_in;     // [S][T][n][n_control]
    const int32_t *ctrl_out;
    int max_in, max_out, n_control, ctrl_per_chain;
};

struct SweepParams {
    NetView net;
    int C, K, prior, tune, tune_interval;
    double *X;                 // [C][T][n][d]
    const double *intercept;   // [C][2]
    const double *rinv;        // [C][n]   1 / radii
    const int32_t *z;          // [C][T][n]
    const double *mu;          // [C][K][d]
    const double *sigma;       // [C][K]
    const double *lambda;      // [C]
    double tau_sq, sigma_sq;
    double *step;              // [C][T][n]
    int32_t *nacc, *nsteps, *until;
    const double *eps;         // replay: [C][T][n][d]  (nullptr -> Philox)
    const double *logu;        // replay: [C][T][n]
    uint64_t seed;
    uint32_t sweep, chain_offset;
    int32_t *accepted;         // optional [C][T][n]
    double *ratio;             // optional [C][T][n]
    unsigned int *flags;       // bit0 non-finite ratio, bit1 case-control out-of-bounds quirk
    int fuse_center;           // k_sweep (positions in shared memory): centre before the write-back
    double *ll_cur;            // optional [C]: full-network log-likelihood of the post-sweep state
    // row-sum cache (exact likelihoods, device loop): rows[c][t][j] = sum_i term(i, j) at the current
    // state, so that a node-update evaluates its PROPOSAL only; scr[c][t][i] parks the proposal's
    // per-pair terms until the decision (see node_loglik1 / rows_apply)
    double *rows;              // [C][T][n] or nullptr (two-variant evaluation)
    double *scr;               // [C][T][n]
};

// latent dimension: a compile-time constant in the specialised (D == 2) instantiations
template <int DM>
__device__ __forceinline__ int latent_dim(const NetView &net) { return DM == kMaxD ? net.d : DM; }

// eta of a directed dyad given dist and the two reciprocal radii
//   sender s -> receiver r :  b_in * (1 - dist / r_r) + b_out * (1 - dist / r_s)
__device__ __forceinline__ double eta_directed(double b_in, double b_out, double dist,
                                               double rinv_recv, double rinv_send)
{
    return b_in * (1.0 - dist * rinv_recv) + b_out * (1.0 - dist * rinv_send);
}

// y in {0,1} -> y - 1/2 as a double, built from the adjacency bit with integer ops only
__device__ __forceinline__ double ymask(uint32_t word, int lane)
{
    const uint32_t sign = ((~word >> lane) & 1u) << 31; // bit clear -> -0.5
    return __hiloint2double((int)(0x3fe00000u | sign), 0);
}

__device__ __forceinline__ double vmask(bool v) { return __hiloint2double(v ? 0x3ff00000 : 0, 0); }

// ---------------------------------------------------------------------------------------------
// Per-node pairwise sums for the proposal (xn) and the current position (xo) of node j in slice
// t, executed by one warp.  Returns the two log-likelihoods (warp-uniform).
// ---------------------------------------------------------------------------------------------
// LO (needs nteam == 1, skip < 0): also return, per lane, the part of both sums that comes from
// i < j (lo_new / lo_old, NOT reduced over the warp).  Summed over the nodes of a slice with the accepted variant picked each
// time, these give the full-network log-likelihood of the post-sweep state for free: the dyad
// {i, j}, i < j, is last evaluated when j is updated, with x_i already final.
// INTR: give interior trips (64 existing pairs, none of them the self pair) their own unmasked body;
// pays off for long rows (cfg 4: +2.7 %), costs 1.5 % at n = 120 where every trip is a boundary trip
// or nearly, so only the long-row instantiation of k_sweep turns it on.
template <int LK, int DM, bool LO = false, bool INTR = false>
__device__ __forceinline__ void node_loglik2(const NetView &net, const double *Xt /* [n][d] */,
                                             const double *rinv /* [n] */, int chain, int t, int j,
                                             const double (&xn)[DM], const double (&xo)[DM],
                                             double b0, double b1, int lane, double &ll_new,
                                             double &ll_old, unsigned int *flags, int wteam = 0,
                                             int nteam = 1, int skip = -1, double *lo_new = nullptr,
                                             double *lo_old = nullptr)
{
    // skip: a second index left out of the sums (the pipelined slice kernel adds that pair later)
    // wteam / nteam: this warp's rank in, and the size of, the team of warps that shares the row.
    // The outputs are this warp's PARTIAL sums (the whole sums when nteam == 1); they are linear in
    // the per-pair terms, so the caller adds the partials of the team in a fixed order.
    const int n = net.n, d = latent_dim<DM>(net);
    if (LK == kUndirected) {
        // K1 static_network_fast.pyx:17-44.  Two 32-node chunks per trip, branch- and select-free
        // (out-of-range lanes and the self pair are evaluated on a clamped index and multiplied
        // by a 0/1 mask): four independent sqrt/softplus chains per lane keep the fp64 pipe busy.
        const uint32_t *row = net.rowbits + ((size_t)t * n + j) * net.W;
        double an = 0.0, ao = 0.0, an2 = 0.0, ao2 = 0.0;
        for (int base = wteam * 64; base < n; base += 64 * nteam) {
            const int i0 = base + lane, i1 = i0 + 32;
            const uint2 w = __ldg(reinterpret_cast<const uint2 *>(row + (base >> 5)));
            if (INTR && base + 64 <= n && (unsigned)(j - base) >= 64u &&
                (skip < 0 || (unsigned)(skip - base) >= 64u)) {
                // interior trip (warp-uniform): all 64 pairs exist and none is the self pair, so no
                // masks, no clamped indices and -- the pairs being all below or all above j -- no
                // lower-triangle bookkeeping either
                const double y0 = ymask(w.x, lane), y1 = ymask(w.y, lane);
                double xa[DM], xb[DM];
                load_pos<DM>(Xt + (size_t)i0 * d, d, xa);
                load_pos<DM>(Xt + (size_t)i1 * d, d, xb);
                an += logit_term(y0, b0 - fast_dist<DM>(xa, xn, d));
                ao += logit_term(y0, b0 - fast_dist<DM>(xa, xo, d));
                an2 += logit_term(y1, b0 - fast_dist<DM>(xb, xn, d));
                ao2 += logit_term(y1, b0 - fast_dist<DM>(xb, xo, d));
                continue;
            }
            const double v0 = vmask((i0 < n) && (i0 != j) && (i0 != skip));
            const double v1 = vmask((i1 < n) && (i1 != j) && (i1 != skip));
            const double y0 = ymask(w.x, lane), y1 = ymask(w.y, lane);
            double xa[DM], xb[DM];
            load_pos<DM>(Xt + (size_t)(i0 < n ? i0 : n - 1) * d, d, xa);
            load_pos<DM>(Xt + (size_t)(i1 < n ? i1 : n - 1) * d, d, xb);
            const double en0 = b0 - fast_dist<DM>(xa, xn, d);
            const double eo0 = b0 - fast_dist<DM>(xa, xo, d);
            const double en1 = b0 - fast_dist<DM>(xb, xn, d);
            const double eo1 = b0 - fast_dist<DM>(xb, xo, d);
            const double tn0 = logit_term(y0, en0), to0 = logit_term(y0, eo0);
            const double tn1 = logit_term(y1, en1), to1 = logit_term(y1, eo1);
            if (LO && (unsigned)(j - base) < 64u) {
                // the one trip that straddles j: everything accumulated so far is i < j
                const double l0 = vmask(i0 < j), l1 = vmask(i1 < j);
                *lo_new = fma(l0, tn0, fma(l1, tn1, an + an2));
                *lo_old = fma(l0, to0, fma(l1, to1, ao + ao2));
            }
            an = fma(v0, tn0, an);
            ao = fma(v0, to0, ao);
            an2 = fma(v1, tn1, an2);
            ao2 = fma(v1, to1, ao2);
        }
        ll_new = an + an2;
        ll_old = ao + ao2;
        warp_sum2(ll_new, ll_old, lane);
    } else if (LK == kDirected) {
        // K2 directed_likelihoods_fast.pyx:46-80 (branch- and select-free, four chains per lane)
        const uint32_t *row = net.rowbits + ((size_t)t * n + j) * net.W;
        const uint32_t *col = net.colbits + ((size_t)t * n + j) * net.W;
        const double rj = rinv[j];
        double an = 0.0, ao = 0.0;
        for (int base = wteam * 32; base < n; base += 32 * nteam) {
            const int i = base + lane;
            const double y_ji = ymask(__ldg(row + (base >> 5)), lane); // Y[node, i]: node sends
            const double y_ij = ymask(__ldg(col + (base >> 5)), lane); // Y[i, node]: i sends
            const double v = vmask((i < n) && (i != j) && (i != skip));
            const int ic = i < n ? i : n - 1;
            double xi[DM];
            load_pos<DM>(Xt + (size_t)ic * d, d, xi);
            const double ri = rinv[ic];
            const double dn = fast_dist<DM>(xi, xn, d);
            const double dd = fast_dist<DM>(xi, xo, d);
            const double tn = logit_term(y_ji, eta_directed(b0, b1, dn, ri, rj)) +
                              logit_term(y_ij, eta_directed(b0, b1, dn, rj, ri));
            const double to = logit_term(y_ji, eta_directed(b0, b1, dd, ri, rj)) +
                              logit_term(y_ij, eta_directed(b0, b1, dd, rj, ri));
            if (LO && (unsigned)(j - base) < 32u) {
                const double l = vmask(i < j);
                *lo_new = fma(l, tn, an);
                *lo_old = fma(l, to, ao);
            }
            an = fma(v, tn, an);
            ao = fma(v, to, ao);
        }
        ll_new = an;
        ll_old = ao;
        warp_sum2(ll_new, ll_old, lane);
    } else {
        // K3 directed_likelihoods_fast.pyx:83-182 (case-control estimator)
        const size_t r = (size_t)t * n + j;
        const int indeg = net.deg[r * 2 + 0], outdeg = net.deg[r * 2 + 1];
        const int32_t *ie = net.in_edges + r * net.max_in;
        const int32_t *oe = net.out_edges + r * net.max_out;
        const size_t coff =
            ((size_t)(net.ctrl_per_chain ? chain : 0) * net.T * n + r) * net.n_control;
        const int32_t *ci = net.ctrl_in + coff;
        const int32_t *co = net.ctrl_out + coff;
        const double rj = rinv[j];
        double e_n = 0.0, e_o = 0.0;     // edge terms
        double ci_n = 0.0, ci_o = 0.0;   // control sums over the in lists
        double co_n = 0.0, co_o = 0.0;   // control sums over the out lists
        auto eta_pair = [&](int k, bool k_sends, double &vn, double &vo) {
            double xk[DM];
            load_pos<DM>(Xt + (size_t)k * d, d, xk);
            const double rk = rinv[k];
            const double dn = fast_dist<DM>(xk, xn, d);
            const double dd = fast_dist<DM>(xk, xo, d);
            const double r_recv = k_sends ? rj : rk, r_send = k_sends ? rk : rj;
            vn = eta_directed(b0, b1, dn, r_recv, r_send);
            vo = eta_directed(b0, b1, dd, r_recv, r_send);
        };
        const int q0 = wteam * 32 + lane, qs = 32 * nteam;
        // Edge lists, two trips per pass: both index loads and then both position gathers are
        // independent, so their L2 round trips overlap (a list walk is a chain of dependent loads).
        // Per-lane accumulation order is unchanged (q, q + qs, q + 2 qs, ...).
        auto edge_list = [&](const int32_t *lst, int len, bool k_sends) {
            for (int q = q0; q < len; q += 2 * qs) {
                const int qb = q + qs;
                const bool two = qb < len;
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
        if (nteam == 1 && indeg <= 16 && outdeg <= 16) {
            // short lists (the sparse networks this likelihood is for): both in one trip, the in-list on
            // lanes 0-15 and the out-list on lanes 16-31 -- one pair evaluation per lane instead of four
            const bool out_side = lane >= 16;
            const int q = lane & 15;
            const bool live = q < (out_side ? outdeg : indeg);
            const int k = live ? (out_side ? oe[q] : ie[q]) : j;
            double xk[DM];
            load_pos<DM>(Xt + (size_t)k * d, d, xk);
            const double rk = rinv[k];
            const double dn = fast_dist<DM>(xk, xn, d), dd = fast_dist<DM>(xk, xo, d);
            const double r_recv = out_side ? rk : rj, r_send = out_side ? rj : rk; // in-list: k sends to node
            const double tn = logit_term(0.5, eta_directed(b0, b1, dn, r_recv, r_send));
            const double to = logit_term(0.5, eta_directed(b0, b1, dd, r_recv, r_send));
            if (live) { e_n += tn; e_o += to; }
        } else {
            edge_list(ie, indeg, true);  // :108-119
            edge_list(oe, outdeg, false); // :122-133
        }
        int m = net.n_control, m_out;
        if (nteam == 1 && net.n_control <= 128) {
            // all control indices of the node in registers with one round trip (4 + 4 loads)
            int cin[4], cout[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int q = u * 32 + lane;
                cin[u] = q < net.n_control ? ci[q] : 0;
                cout[u] = q < net.n_control ? co[q] : 0;
            }
            // usable controls = prefix of ctrl_in before its first -1 (:137; and :161, which tests
            // the IN list while walking the OUT list)
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const unsigned bal = __ballot_sync(kFull, u * 32 + lane < net.n_control && cin[u] == -1);
                if (bal && m == net.n_control) m = u * 32 + __ffs(bal) - 1;
            }
            m_out = m;
#pragma unroll
            for (int u = 0; u < 4; u++) { // the reference reads X[-1] here: flag + stop
                const unsigned bal = __ballot_sync(kFull, u * 32 + lane < m && cout[u] < 0);
                if (bal && m_out == m) {
                    m_out = u * 32 + __ffs(bal) - 1;
                    if (lane == 0) atomicOr(flags, 2u);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) { // :136-152 and :160-176; clamped gathers, masked sums
                const int q = u * 32 + lane;
                const bool vi = q < m, vo_ = q < m_out;
                double an, ao, bn, bo;
                eta_pair(vi ? cin[u] : j, true, an, ao);
                eta_pair(vo_ ? cout[u] : j, false, bn, bo);
                const double la = log1pexp(an), lb = log1pexp(ao), lc = log1pexp(bn), ld = log1pexp(bo);
                if (vi) { ci_n += la; ci_o += lb; }
                if (vo_) { co_n += lc; co_o += ld; }
            }
        } else {
            for (int base = 0; base < net.n_control; base += 32) {
                const int q = base + lane;
                const bool stop = (q < net.n_control) && (ci[q] == -1);
                const unsigned bal = __ballot_sync(kFull, stop);
                if (bal) { m = base + __ffs(bal) - 1; break; }
            }
            m_out = m;
            for (int base = 0; base < m; base += 32) { // the reference reads X[-1] here: flag + stop
                const int q = base + lane;
                const bool bad = (q < m) && (co[q] < 0);
                const unsigned bal = __ballot_sync(kFull, bad);
                if (bal) {
                    m_out = base + __ffs(bal) - 1;
                    if (lane == 0 && wteam == 0) atomicOr(flags, 2u);
                    break;
                }
            }
            for (int q = q0; q < m; q += qs) { // :136-152
                double vn, vo;
                eta_pair(ci[q], true, vn, vo);
                ci_n += log1pexp(vn);
                ci_o += log1pexp(vo);
            }
            for (int q = q0; q < m_out; q += qs) { // :160-176
                double vn, vo;
                eta_pair(co[q], false, vn, vo);
                co_n += log1pexp(vn);
                co_o += log1pexp(vo);
            }
        }
        e_n = warp_sum(e_n); e_o = warp_sum(e_o);
        ci_n = warp_sum(ci_n); ci_o = warp_sum(ci_o);
        co_n = warp_sum(co_n); co_o = warp_sum(co_o);
        const double adj_in = (double)(n - indeg - 1) / (double)m;       // :155
        const double adj_out = (double)(n - outdeg - 1) / (double)m_out; // :179
        ll_new = (e_n - adj_in * ci_n) - adj_out * co_n;
        ll_old = (e_o - adj_in * ci_o) - adj_out * co_o;
    }
}

// ---------------------------------------------------------------------------------------------
// Row-sum cache.  The MH ratio of node j needs l_j(x') - l_j(x) with l_j(.) = sum_i term(x_i, .).
// The reference evaluates both sums afresh (2 (n-1) pair terms per node-update).  Here the device
// loop keeps rows[j] = l_j(x_j) for every node of the slice: a node-update evaluates the
// PROPOSAL's terms only (node_loglik1, parking them in scr[]), takes l_j(x) from the cache, and
// only an ACCEPTED move pays the second pass (rows_apply): rows[i] += term(x_i, x') - term(x_i, x)
// for every other node and rows[j] = l_j(x').  (1 + p_accept)(n-1) pair terms instead of 2 (n-1).
// The cached sums differ from fresh ones by summation order only (<= ~1e-13 relative); whatever
// moves the whole network at once (an accepted intercept / radii proposal) replaces the rows by
// the ones k_rows computed for that proposal.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double ld_cg(const double *p) { return __ldcg(p); }
__device__ __forceinline__ void st_cg(double *p, double v) { __stcg(p, v); }

// proposal-only pass: returns l_j(xn) (warp-uniform); scr[i] = term(x_i, xn) for every i != j
// PAD: the slice is padded to a multiple of 128 rows with far-away dummy nodes whose adjacency bits
// are 0 (their terms are exactly 0), so only the trip that holds the self pair needs a mask
template <int LK, int DM, bool INTR, bool PAD = false>
__device__ __forceinline__ double node_loglik1(const NetView &net, const double *Xt, const double *rinv,
                                               int t, int j, const double (&xn)[DM], double b0, double b1,
                                               int lane, double *scr)
{
    const int n = net.n, d = latent_dim<DM>(net);
    if (LK == kUndirected) {
        // four 32-node chunks per trip = four independent sqrt/softplus chains per lane; the 128
        // adjacency bits of a trip arrive with one 128-bit load (rows are 16-byte aligned)
        const uint32_t *row = net.rowbits + ((size_t)t * n + j) * net.W;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        for (int base = 0; base < n; base += 128) {
            const uint4 w = __ldg(reinterpret_cast<const uint4 *>(row + (base >> 5)));
            const int i0 = base + lane, i1 = i0 + 32, i2 = i0 + 64, i3 = i0 + 96;
            const double y0 = ymask(w.x, lane), y1 = ymask(w.y, lane), y2 = ymask(w.z, lane), y3 = ymask(w.w, lane);
            double xa[DM], xb[DM], xc[DM], xd[DM];
            if ((PAD || (INTR && base + 128 <= n)) && (unsigned)(j - base) >= 128u) { // interior trip: no masks
                load_pos<DM>(Xt + (size_t)i0 * d, d, xa);
                load_pos<DM>(Xt + (size_t)i1 * d, d, xb);
                load_pos<DM>(Xt + (size_t)i2 * d, d, xc);
                load_pos<DM>(Xt + (size_t)i3 * d, d, xd);
                const double t0 = logit_term(y0, b0 - fast_dist<DM>(xa, xn, d));
                const double t1 = logit_term(y1, b0 - fast_dist<DM>(xb, xn, d));
                const double t2 = logit_term(y2, b0 - fast_dist<DM>(xc, xn, d));
                const double t3 = logit_term(y3, b0 - fast_dist<DM>(xd, xn, d));
                st_cg(scr + i0, t0); st_cg(scr + i1, t1); st_cg(scr + i2, t2); st_cg(scr + i3, t3);
                a0 += t0; a1 += t1; a2 += t2; a3 += t3;
                continue;
            }
            load_pos<DM>(Xt + (size_t)((PAD || i0 < n) ? i0 : n - 1) * d, d, xa);
            load_pos<DM>(Xt + (size_t)((PAD || i1 < n) ? i1 : n - 1) * d, d, xb);
            load_pos<DM>(Xt + (size_t)((PAD || i2 < n) ? i2 : n - 1) * d, d, xc);
            load_pos<DM>(Xt + (size_t)((PAD || i3 < n) ? i3 : n - 1) * d, d, xd);
            const double t0 = logit_term(y0, b0 - fast_dist<DM>(xa, xn, d));
            const double t1 = logit_term(y1, b0 - fast_dist<DM>(xb, xn, d));
            const double t2 = logit_term(y2, b0 - fast_dist<DM>(xc, xn, d));
            const double t3 = logit_term(y3, b0 - fast_dist<DM>(xd, xn, d));
            if (PAD || i0 < n) st_cg(scr + i0, t0);
            if (PAD || i1 < n) st_cg(scr + i1, t1);
            if (PAD || i2 < n) st_cg(scr + i2, t2);
            if (PAD || i3 < n) st_cg(scr + i3, t3);
            a0 = fma(vmask((PAD || i0 < n) && (i0 != j)), t0, a0);
            a1 = fma(vmask((PAD || i1 < n) && (i1 != j)), t1, a1);
            a2 = fma(vmask((PAD || i2 < n) && (i2 != j)), t2, a2);
            a3 = fma(vmask((PAD || i3 < n) && (i3 != j)), t3, a3);
        }
        return warp_sum((a0 + a1) + (a2 + a3));
    } else {
        // K2: two chunks per trip x two directions = four softplus chains per lane
        const uint32_t *row = net.rowbits + ((size_t)t * n + j) * net.W;
        const uint32_t *col = net.colbits + ((size_t)t * n + j) * net.W;
        const double rj = rinv[j];
        double a0 = 0.0, a1 = 0.0;
        for (int base = 0; base < n; base += 64) {
            const uint2 wr = __ldg(reinterpret_cast<const uint2 *>(row + (base >> 5)));
            const uint2 wc = __ldg(reinterpret_cast<const uint2 *>(col + (base >> 5)));
            const int i0 = base + lane, i1 = i0 + 32;
            const int c0 = i0 < n ? i0 : n - 1, c1 = i1 < n ? i1 : n - 1;
            double xa[DM], xb[DM];
            load_pos<DM>(Xt + (size_t)c0 * d, d, xa);
            load_pos<DM>(Xt + (size_t)c1 * d, d, xb);
            const double r0 = rinv[c0], r1 = rinv[c1];
            const double d0 = fast_dist<DM>(xa, xn, d), d1 = fast_dist<DM>(xb, xn, d);
            const double t0 = logit_term(ymask(wr.x, lane), eta_directed(b0, b1, d0, r0, rj)) +
                              logit_term(ymask(wc.x, lane), eta_directed(b0, b1, d0, rj, r0));
            const double t1 = logit_term(ymask(wr.y, lane), eta_directed(b0, b1, d1, r1, rj)) +
                              logit_term(ymask(wc.y, lane), eta_directed(b0, b1, d1, rj, r1));
            if (i0 < n) st_cg(scr + i0, t0);
            if (i1 < n) st_cg(scr + i1, t1);
            a0 = fma(vmask((i0 < n) && (i0 != j)), t0, a0);
            a1 = fma(vmask((i1 < n) && (i1 != j)), t1, a1);
        }
        return warp_sum(a0 + a1);
    }
}

// accepted move of node j (old position xo): every other row trades its old pair term for the
// parked new one, row j becomes the proposal's sum.  One warp; rows / scr point at slice t.
template <int LK, int DM>
__device__ __forceinline__ void rows_apply(const NetView &net, const double *Xt, const double *rinv, int t,
                                           int j, const double (&xo)[DM], double b0, double b1, int lane,
                                           const double *scr, double *rows, double ll_new)
{
    const int n = net.n, d = latent_dim<DM>(net);
    if (LK == kUndirected) {
        const uint32_t *row = net.rowbits + ((size_t)t * n + j) * net.W;
        for (int base = 0; base < n; base += 128) {
            const uint4 w = __ldg(reinterpret_cast<const uint4 *>(row + (base >> 5)));
            const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
            double told[4], tnew[4], cur[4];
            bool ok[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = base + lane + 32 * u;
                ok[u] = (i < n) && (i != j);
                const int ic = i < n ? i : n - 1;
                double xi[DM];
                load_pos<DM>(Xt + (size_t)ic * d, d, xi);
                tnew[u] = ld_cg(scr + ic);
                cur[u] = ld_cg(rows + ic);
                told[u] = logit_term(ymask(ww[u], lane), b0 - fast_dist<DM>(xi, xo, d));
            }
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (ok[u]) st_cg(rows + base + lane + 32 * u, cur[u] + (tnew[u] - told[u]));
        }
    } else {
        const uint32_t *row = net.rowbits + ((size_t)t * n + j) * net.W;
        const uint32_t *col = net.colbits + ((size_t)t * n + j) * net.W;
        const double rj = rinv[j];
        for (int base = 0; base < n; base += 64) {
            const uint2 wr = __ldg(reinterpret_cast<const uint2 *>(row + (base >> 5)));
            const uint2 wc = __ldg(reinterpret_cast<const uint2 *>(col + (base >> 5)));
            const uint32_t wrr[2] = {wr.x, wr.y}, wcc[2] = {wc.x, wc.y};
            double told[2], tnew[2], cur[2];
            bool ok[2];
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int i = base + lane + 32 * u;
                ok[u] = (i < n) && (i != j);
                const int ic = i < n ? i : n - 1;
                double xi[DM];
                load_pos<DM>(Xt + (size_t)ic * d, d, xi);
                const double ri = rinv[ic];
                tnew[u] = ld_cg(scr + ic);
                cur[u] = ld_cg(rows + ic);
                const double dd = fast_dist<DM>(xi, xo, d);
                told[u] = logit_term(ymask(wrr[u], lane), eta_directed(b0, b1, dd, ri, rj)) +
                          logit_term(ymask(wcc[u], lane), eta_directed(b0, b1, dd, rj, ri));
            }
#pragma unroll
            for (int u = 0; u < 2; u++)
                if (ok[u]) st_cg(rows + base + lane + 32 * u, cur[u] + (tnew[u] - told[u]));
        }
    }
    if (lane == 0) st_cg(rows + j, ll_new);
}

// ---------------------------------------------------------------------------------------------
// Prior terms of the closure `logp` (sample_latent_positions.py:131-140 LSM, :187-199 mixture).
// The reference subtracts them from loglik one after the other:
//     loglik -= term_prev(x | X[t-1,j])   (or the t == 0 term)
//     loglik -= term_next(X[t+1,j] | x)   (if t < T-1)
// term_next only needs the OLD X[t+1,j], so it is evaluated lane-parallel for 32 nodes at a time;
// term_prev needs this sweep's X[t-1,j] and is evaluated when the wavefront flag allows.
// ---------------------------------------------------------------------------------------------
template <int DM>
__device__ __forceinline__ double prior_next(const SweepParams &p, int c, int t, int j,
                                             const double (&x)[DM], const double (&xnext)[DM])
{
    const int d = latent_dim<DM>(p.net), T = p.net.T, n = p.net.n;
    double diff[DM];
    if (p.prior == 0) {
#pragma unroll
        for (int k = 0; k < DM; k++) diff[k] = (k < d) ? __dsub_rn(xnext[k], x[k]) : 0.0;
        return half_sumsq_over<DM>(diff, d, p.sigma_sq);
    }
    const double lm = p.lambda[c], oml = __dsub_rn(1.0, lm);
    const int zn = p.z[((size_t)c * T + (t + 1)) * n + j];
    const double *mu = p.mu + ((size_t)c * p.K + zn) * d;
#pragma unroll
    for (int k = 0; k < DM; k++)
        diff[k] = (k < d) ? __dsub_rn(__dsub_rn(xnext[k], __dmul_rn(oml, x[k])), __dmul_rn(lm, mu[k]))
                          : 0.0;
    return half_sumsq_over<DM>(diff, d, p.sigma[(size_t)c * p.K + zn]);
}

// 0.5 * np.sum(v*v) * (1/s): the serial part of a node-update multiplies by a reciprocal prepared
// once per 32-node block (1 ulp from the reference's division; decisions are unaffected)
template <int DM>
__device__ __forceinline__ double half_sumsq_times(const double (&v)[DM], int d, double inv)
{
    double sq[DM];
#pragma unroll
    for (int k = 0; k < DM; k++) sq[k] = (k < d) ? __dmul_rn(v[k], v[k]) : 0.0;
    return __dmul_rn(__dmul_rn(0.5, np_sum_small<DM>(sq, d)), inv);
}

template <int DM>
__device__ __forceinline__ double prior_prev(const SweepParams &p, int c, int t, int zc, double inv,
                                             const double (&x)[DM], const double (&xprev)[DM])
{
    const int d = latent_dim<DM>(p.net);
    double diff[DM];
    if (p.prior == 0) {
        if (t == 0) return half_sumsq_times<DM>(x, d, inv);
#pragma unroll
        for (int k = 0; k < DM; k++) diff[k] = (k < d) ? __dsub_rn(x[k], xprev[k]) : 0.0;
        return half_sumsq_times<DM>(diff, d, inv);
    }
    const double *mu = p.mu + ((size_t)c * p.K + zc) * d;
    if (t == 0) {
#pragma unroll
        for (int k = 0; k < DM; k++) diff[k] = (k < d) ? __dsub_rn(x[k], mu[k]) : 0.0;
    } else {
        const double lm = p.lambda[c], oml = __dsub_rn(1.0, lm);
#pragma unroll
        for (int k = 0; k < DM; k++)
            diff[k] = (k < d) ? __dsub_rn(__dsub_rn(x[k], __dmul_rn(oml, xprev[k])), __dmul_rn(lm, mu[k]))
                              : 0.0;
    }
    return half_sumsq_times<DM>(diff, d, inv);
}

// per-warp staging area of one 32-node block (shared memory)
__host__ __device__ inline size_t sweep_stage_doubles(int d) { return (size_t)32 * (d + 5); }
// rows per slice of the chain kernel's shared-memory copy (padding: see k_sweep)
__host__ __device__ inline int sweep_rows_padded(bool pad, int n) { return pad ? ((n + 127) & ~127) : n; }
constexpr double kFarAway = 1.0e6; // coordinate of the padding nodes: exp(-|beta - dist|) underflows to 0

// ---------------------------------------------------------------------------------------------
// k_sweep: one latent-position sweep of every chain.
// grid = C chains, block = 32 * min(T, 16) threads,
// dynamic smem = [T*n*d doubles if XS] + nwarps * 32*(d+5) doubles + T ints
//
// Each warp owns a time slice and walks its nodes in blocks of 32.  At the head of a block lane l
// prepares everything that does not depend on the in-flight wavefront for node jb+l -- the
// sampler state, the random draws (replay buffer or Philox + Box-Muller), the proposal
// x0 + step*eps, the "next" prior terms -- once per 32 nodes with all lanes busy and coalesced
// loads, and parks it in a per-warp shared-memory stage.  The serial part of a node-update is then:
// broadcast loads from the stage, the 32-lane pairwise reduction, the wavefront flag, the "prev"
// prior term and the accept/reject, which lane (j mod 32) commits.  The Metropolis bookkeeping of
// the 32 samplers runs lane-parallel at the end of the block.
// ---------------------------------------------------------------------------------------------
template <int LK, int D, bool XS, int MAXT, int MINB, bool RS = false>
__global__ void __launch_bounds__(MAXT, MINB) k_sweep(const SweepParams p)
{
    constexpr int DM = (D == 0) ? kMaxD : D;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int T = p.net.T, n = p.net.n, d = (D == 0) ? p.net.d : D;
    const int c = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const size_t chain_elems = (size_t)T * n * d;
    double *Xg = p.X + (size_t)c * chain_elems;
    double *Xc;
    double *stage_base;
    // rows per slice as laid out in Xc: undirected chains in shared memory are padded to a multiple of
    // 128 rows with far-away dummy nodes (no adjacency bits, terms exactly 0), see node_loglik1<PAD>
    const int ns = sweep_rows_padded(XS && LK == kUndirected, n);
    if (XS) {
        Xc = reinterpret_cast<double *>(smem_raw);
        stage_base = Xc + (size_t)T * ns * d;
        const int per = ns * d;
        for (int e = threadIdx.x; e < T * per; e += blockDim.x) {
            const int tt = e / per, r = e - tt * per;
            Xc[e] = (r < n * d) ? Xg[(size_t)tt * n * d + r] : kFarAway;
        }
    } else {
        Xc = Xg;
        stage_base = reinterpret_cast<double *>(smem_raw);
    }
    volatile int *progress =
        reinterpret_cast<volatile int *>(stage_base + (size_t)nwarps * sweep_stage_doubles(d));
    for (int t = threadIdx.x; t < T; t += blockDim.x) progress[t] = 0;
    __syncthreads();

    // stage layout: prop[32][d] | logu[32] | next_new[32] | next_old[32] | inv[32] | zc[32] (int)
    double *st_prop = stage_base + (size_t)warp * sweep_stage_doubles(d);
    double *st_logu = st_prop + 32 * d, *st_nn = st_logu + 32, *st_no = st_nn + 32, *st_inv = st_no + 32;
    int *st_zc = reinterpret_cast<int *>(st_inv + 32);

    const double b0 = p.intercept[c * 2 + 0], b1 = p.intercept[c * 2 + 1];
    const double *rinv = (LK == kUndirected) ? nullptr : p.rinv + (size_t)c * n;
    const uint32_t chain_id = (uint32_t)c + p.chain_offset;
    bool nonfinite = false;

    double full_acc = 0.0; // lower-triangle terms of this warp's slices at the kept states
    constexpr bool kPad = XS && LK == kUndirected;
    for (int t = warp; t < T; t += nwarps) {
        double *Xt = Xc + (size_t)t * ns * d;
        double *rows_t = RS ? p.rows + ((size_t)c * T + t) * n : nullptr; // row-sum cache of this slice
        double *scr_t = RS ? p.scr + ((size_t)c * T + t) * sweep_rows_padded(true, n) : nullptr;
        for (int jb = 0; jb < n; jb += 32) {
            // ---------------- lane-parallel preparation for node jl = jb + lane ----------------
            const int jl = jb + lane;
            const bool mine = jl < n;
            const size_t gs = ((size_t)c * T + t) * n + (mine ? jl : 0);
            double my_step = 0.0;
            int my_nacc = 0, my_nsteps = 0, my_until = 0, my_acc = 0;
            if (mine) {
                double eps[DM], x0[DM], x[DM], logu;
                load_pos<DM>(Xt + (size_t)jl * d, d, x0);
                my_step = p.step[gs]; my_nacc = p.nacc[gs]; my_nsteps = p.nsteps[gs]; my_until = p.until[gs];
                if (p.eps) {
#pragma unroll
                    for (int k = 0; k < DM; k++) eps[k] = (k < d) ? p.eps[gs * d + k] : 0.0;
                    logu = p.logu[gs];
                } else {
                    latent_draws<DM>(p.seed, (uint32_t)(t * n + jl), p.sweep, chain_id, d, eps, logu);
                }
                // metropolis.py:44  x = x0 + step_size * randn(d): separately rounded mul and add
#pragma unroll
                for (int k = 0; k < DM; k++) {
                    x[k] = (k < d) ? __dadd_rn(x0[k], __dmul_rn(my_step, eps[k])) : 0.0;
                    if (k < d) st_prop[lane * d + k] = x[k];
                }
                st_logu[lane] = logu;
                double inv = (t == 0) ? 1.0 / p.tau_sq : 1.0 / p.sigma_sq; // LSM prior scales
                int zc = 0;
                if (p.prior != 0) {
                    zc = p.z[((size_t)c * T + t) * n + jl];
                    inv = 1.0 / p.sigma[(size_t)c * p.K + zc];
                }
                st_inv[lane] = inv;
                st_zc[lane] = zc;
                double nn = 0.0, no = 0.0;
                if (t < T - 1) { // X[t+1, jl] is still last sweep's value: slice t+1 trails this one
                    double xnx[DM];
                    const volatile double *q = Xc + ((size_t)(t + 1) * ns + jl) * d;
#pragma unroll
                    for (int k = 0; k < DM; k++) xnx[k] = (k < d) ? q[k] : 0.0;
                    nn = prior_next<DM>(p, c, t, jl, x, xnx);
                    no = prior_next<DM>(p, c, t, jl, x0, xnx);
                }
                st_nn[lane] = nn;
                st_no[lane] = no;
            }
            __syncwarp();
            const int jend = (n - jb) < 32 ? (n - jb) : 32;
            // ---------------- serial node-updates of this block ----------------
            for (int jj = 0; jj < jend; jj++) {
                const int j = jb + jj;
                double x[DM], x0[DM];
                load_pos<DM>(st_prop + jj * d, d, x);
                load_pos<DM>(Xt + (size_t)j * d, d, x0);
                if (LK != kCaseControl && j + 1 < n && lane * 32 < p.net.W) { // next row -> L1
                    const size_t o = ((size_t)t * n + j + 1) * p.net.W + lane * 32;
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(p.net.rowbits + o));
                    if (LK == kDirected) asm volatile("prefetch.global.L1 [%0];" ::"l"(p.net.colbits + o));
                }
                double ll_new, ll_old, lo_n = 0.0, lo_o = 0.0;
                // the (320, 2) build is the one picked when shared memory admits two chains per SM,
                // i.e. for long rows: it gets the unmasked interior trips
                if (RS && LK != kCaseControl) {
                    // proposal only; the current position's sum comes from the row-sum cache
                    ll_old = ld_cg(rows_t + j);
                    ll_new = node_loglik1<LK == kCaseControl ? kDirected : LK, DM, (MAXT == 320 && MINB == 2), kPad>(
                        p.net, Xt, rinv, t, j, x, b0, b1, lane, scr_t);
                } else if (LK != kCaseControl)
                    node_loglik2<LK, DM, true, (MAXT == 320 && MINB == 2)>(p.net, Xt, rinv, c, t, j, x, x0, b0,
                                                                          b1, lane, ll_new, ll_old, p.flags, 0,
                                                                          1, -1, &lo_n, &lo_o);
                else
                    node_loglik2<LK, DM>(p.net, Xt, rinv, c, t, j, x, x0, b0, b1, lane, ll_new, ll_old,
                                         p.flags);

                double xp[DM];
#pragma unroll
                for (int k = 0; k < DM; k++) xp[k] = 0.0;
                if (t > 0) { // wavefront: X[t-1, j] must be this sweep's value (uniform poll)
                    while (progress[t - 1] <= j) { __nanosleep(DLSM_SPIN_NS); } /* poll shared memory, yielding issue slots */
                    __threadfence_block();
                    const volatile double *q = Xc + ((size_t)(t - 1) * ns + j) * d;
#pragma unroll
                    for (int k = 0; k < DM; k++) if (k < d) xp[k] = q[k];
                }
                const double inv = st_inv[jj];
                const int zc = st_zc[jj];
                double lp_new = __dsub_rn(ll_new, prior_prev<DM>(p, c, t, zc, inv, x, xp));
                double lp_old = __dsub_rn(ll_old, prior_prev<DM>(p, c, t, zc, inv, x0, xp));
                if (t < T - 1) {
                    lp_new = __dsub_rn(lp_new, st_nn[jj]);
                    lp_old = __dsub_rn(lp_old, st_no[jj]);
                }
                const double ratio = __dsub_rn(lp_new, lp_old);
                const int acc = (st_logu[jj] >= ratio) ? 0 : 1; // metropolis.py:50 (NaN accepts)
                const bool me = lane == jj;
                my_acc = me ? acc : my_acc;
                full_acc += acc ? lo_n : lo_o; // this lane's share of the dyads {i < j} at the kept state
                nonfinite |= me && (!(ratio == ratio) || ratio - ratio != 0.0);
                if (me) {
                    if (acc) {
#pragma unroll
                        for (int k = 0; k < DM; k++) if (k < d) Xt[(size_t)j * d + k] = x[k];
                    }
                    if (p.ratio) p.ratio[((size_t)c * T + t) * n + j] = ratio;
                    __threadfence_block();
                    progress[t] = j + 1;
                }
                if (RS && LK != kCaseControl && acc) // accepted: the other rows trade their pair terms
                    rows_apply<LK == kCaseControl ? kDirected : LK, DM>(p.net, Xt, rinv, t, j, x0, b0, b1, lane,
                                                                        scr_t, rows_t, ll_new);
                __syncwarp();
            }
            // ---------------- lane-parallel Metropolis bookkeeping + coalesced write-back --------
            if (mine) {
                metropolis_bookkeep(my_step, my_nacc, my_nsteps, my_until, p.tune, p.tune_interval,
                                    my_acc, false);
                p.step[gs] = my_step; p.nacc[gs] = my_nacc; p.nsteps[gs] = my_nsteps; p.until[gs] = my_until;
                if (p.accepted) p.accepted[gs] = my_acc;
            }
            __syncwarp();
        }
    }
    if (nonfinite) atomicOr(p.flags, 1u);
    if (p.ll_cur) { // full-network log-likelihood of the post-sweep state, summed in warp order
        if (RS && LK != kCaseControl) { // every dyad sits in the row sums of both its nodes
            full_acc = 0.0;
            for (int t = warp; t < T; t += nwarps) {
                const double *rows_t = p.rows + ((size_t)c * T + t) * n;
                for (int i = lane; i < n; i += 32) full_acc += 0.5 * ld_cg(rows_t + i);
            }
        }
        full_acc = warp_sum(full_acc);
        __syncthreads();
        double *wsum = stage_base; // the staging area is free now
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
        if (p.fuse_center) {
            // X -= np.mean(X, axis=(0,1)) (lsm.py:501) while the chain is still in shared memory:
            // serial per-column accumulation = numpy's order, bit-identical to k_center
            double *mean = stage_base; // the staging area is free now
            if ((int)threadIdx.x < d) {
                double sacc = 0.0;
                for (int tt = 0; tt < T; tt++) {
                    const double *Xr = Xc + (size_t)tt * ns * d + threadIdx.x;
                    for (int r = 0; r < n; r++) sacc = __dadd_rn(sacc, Xr[(size_t)r * d]);
                }
                mean[threadIdx.x] = __ddiv_rn(sacc, (double)((size_t)T * n));
            }
            __syncthreads();
            for (int e = threadIdx.x; e < T * n * d; e += blockDim.x) {
                const int tt = e / (n * d), r = e - tt * (n * d);
                Xg[e] = __dsub_rn(Xc[(size_t)tt * ns * d + r], mean[r % d]);
            }
        } else {
            for (int e = threadIdx.x; e < T * n * d; e += blockDim.x) {
                const int tt = e / (n * d), r = e - tt * (n * d);
                Xg[e] = Xc[(size_t)tt * ns * d + r];
            }
        }
    }
}

// wavefront flags across CTAs with acquire / release instead of full fences: the producer's
// st.release (after the CTA barrier that follows the commits) orders every member's stores before
// the flag, the consumer's ld.acquire orders the flag before its reads of the neighbour slice
__device__ __forceinline__ int ld_acquire_gpu(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int *p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---------------------------------------------------------------------------------------------
// k_sweep_slice: the same sweep with one CTA per (chain, time slice) -- for long rows and few
// chains (cfg 3: one chain, n = 2000, T = 20), where a warp per slice would leave the GPU idle.
// The team that shares a node's row is the whole CTA (NW warps): each warp reduces its chunks of
// the row, warp 0 adds the NW partials in a fixed order and finishes the node.  The wavefront
// crosses CTAs: progress flags and the neighbouring slices' positions travel through global
// memory (st + __threadfence + flag; volatile loads on the other side).  CTAs take their (chain,
// slice) from an atomic ticket, so a CTA only ever waits on a CTA that has already started:
// no co-residency assumption, no deadlock.
// grid = C*T, block = 32*NW, dynamic smem = [n*d (+ n) doubles if XS] + 32*(d+5) doubles + 2*NW doubles
// ---------------------------------------------------------------------------------------------
template <int LK, int D, bool XS>
__global__ void __launch_bounds__(512) k_sweep_slice(const SweepParams p, int *progress_g,
                                                     unsigned int *ticket)
{
    constexpr int DM = (D == 0) ? kMaxD : D;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_ticket;
    const int T = p.net.T, n = p.net.n, d = (D == 0) ? p.net.d : D;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    if (threadIdx.x == 0) s_ticket = (int)atomicAdd(ticket, 1u);
    __syncthreads();
    const int c = s_ticket / T, t = s_ticket % T;
    double *Xchain = p.X + (size_t)c * T * n * d;
    double *Xg = Xchain + (size_t)t * n * d;   // this slice in global memory
    double *Xt, *stage_base;
    if (XS) {
        Xt = reinterpret_cast<double *>(smem_raw);
        stage_base = Xt + (size_t)n * d;
        for (int e = threadIdx.x; e < n * d; e += blockDim.x) Xt[e] = Xg[e];
    } else {
        Xt = Xg;
        stage_base = reinterpret_cast<double *>(smem_raw);
    }
    double *st_prop = stage_base;
    double *st_logu = st_prop + 32 * d, *st_nn = st_logu + 32, *st_no = st_nn + 32, *st_inv = st_no + 32;
    int *st_zc = reinterpret_cast<int *>(st_inv + 32);
    double *part = stage_base + sweep_stage_doubles(d);     // [nwarps][2]
    int *prog = progress_g + (size_t)c * T;
    const double *rinv = (LK == kUndirected) ? nullptr : p.rinv + (size_t)c * n;
    if (XS && LK != kUndirected) { // reciprocal radii next to the positions (read once per pair)
        double *s_rinv = part + 2 * nwarps;
        for (int e = threadIdx.x; e < n; e += blockDim.x) s_rinv[e] = rinv[e];
        rinv = s_rinv;
    }
    __syncthreads();

    const double b0 = p.intercept[c * 2 + 0], b1 = p.intercept[c * 2 + 1];
    const uint32_t chain_id = (uint32_t)c + p.chain_offset;
    bool nonfinite = false;

    for (int jb = 0; jb < n; jb += 32) {
        // ---- lane-parallel preparation (warp 0) ----
        const int jl = jb + lane;
        const bool mine = (warp == 0) && (jl < n);
        const size_t gs = ((size_t)c * T + t) * n + (jl < n ? jl : 0);
        double my_step = 0.0;
        int my_nacc = 0, my_nsteps = 0, my_until = 0, my_acc = 0;
        if (mine) {
            double eps[DM], x0[DM], x[DM], logu;
            load_pos<DM>(Xt + (size_t)jl * d, d, x0);
            my_step = p.step[gs]; my_nacc = p.nacc[gs]; my_nsteps = p.nsteps[gs]; my_until = p.until[gs];
            if (p.eps) {
#pragma unroll
                for (int k = 0; k < DM; k++) eps[k] = (k < d) ? p.eps[gs * d + k] : 0.0;
                logu = p.logu[gs];
            } else {
                latent_draws<DM>(p.seed, (uint32_t)(t * n + jl), p.sweep, chain_id, d, eps, logu);
            }
#pragma unroll
            for (int k = 0; k < DM; k++) {
                x[k] = (k < d) ? __dadd_rn(x0[k], __dmul_rn(my_step, eps[k])) : 0.0;
                if (k < d) st_prop[lane * d + k] = x[k];
            }
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
            if (t < T - 1) { // slice t+1 (another CTA) cannot have touched nodes >= jb yet
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
        __syncthreads();
        const int jend = (n - jb) < 32 ? (n - jb) : 32;
        for (int jj = 0; jj < jend; jj++) {
            const int j = jb + jj;
            double x[DM], x0[DM];
            load_pos<DM>(st_prop + jj * d, d, x);
            load_pos<DM>(Xt + (size_t)j * d, d, x0);
            if (LK != kCaseControl && j + 1 < n && (int)threadIdx.x * 32 < p.net.W) { // next row -> L1
                const size_t o = ((size_t)t * n + j + 1) * p.net.W + threadIdx.x * 32;
                asm volatile("prefetch.global.L1 [%0];" ::"l"(p.net.rowbits + o));
                if (LK == kDirected) asm volatile("prefetch.global.L1 [%0];" ::"l"(p.net.colbits + o));
            }
            double ll_new, ll_old;
            node_loglik2<LK, DM>(p.net, Xt, rinv, c, t, j, x, x0, b0, b1, lane, ll_new, ll_old,
                                 p.flags, warp, nwarps);
            if (lane == 0) { part[warp * 2] = ll_new; part[warp * 2 + 1] = ll_old; }
            __syncthreads();
            if (warp == 0) {
                ll_new = 0.0; ll_old = 0.0;
                for (int w = 0; w < nwarps; w++) { ll_new += part[w * 2]; ll_old += part[w * 2 + 1]; }
                double xp[DM];
#pragma unroll
                for (int k = 0; k < DM; k++) xp[k] = 0.0;
                if (t > 0) {
                    while (ld_acquire_gpu(prog + t - 1) <= j) { /* spin on the L2-resident flag of slice t-1 */ }
                    const volatile double *q = Xchain + ((size_t)(t - 1) * n + j) * d;
#pragma unroll
                    for (int k = 0; k < DM; k++) if (k < d) xp[k] = q[k];
                }
                const double inv = st_inv[jj];
                const int zc = st_zc[jj];
                double lp_new = __dsub_rn(ll_new, prior_prev<DM>(p, c, t, zc, inv, x, xp));
                double lp_old = __dsub_rn(ll_old, prior_prev<DM>(p, c, t, zc, inv, x0, xp));
                if (t < T - 1) {
                    lp_new = __dsub_rn(lp_new, st_nn[jj]);
                    lp_old = __dsub_rn(lp_old, st_no[jj]);
                }
                const double ratio = __dsub_rn(lp_new, lp_old);
                const int acc = (st_logu[jj] >= ratio) ? 0 : 1;
                const bool me = lane == jj;
                my_acc = me ? acc : my_acc;
                nonfinite |= me && (!(ratio == ratio) || ratio - ratio != 0.0);
                if (me) {
                    if (acc) {
#pragma unroll
                        for (int k = 0; k < DM; k++)
                            if (k < d) {
                                Xt[(size_t)j * d + k] = x[k];
                                if (XS) Xg[(size_t)j * d + k] = x[k];
                            }
                    }
                    if (p.ratio) p.ratio[((size_t)c * T + t) * n + j] = ratio;
                    st_release_gpu(prog + t, j + 1); // this lane's stores above are ordered before the flag
                }
            }
            __syncthreads();
        }
        if (mine) {
            metropolis_bookkeep(my_step, my_nacc, my_nsteps, my_until, p.tune, p.tune_interval,
                                my_acc, false);
            p.step[gs] = my_step; p.nacc[gs] = my_nacc; p.nsteps[gs] = my_nsteps; p.until[gs] = my_until;
            if (p.accepted) p.accepted[gs] = my_acc;
        }
        __syncthreads();
    }
    if (nonfinite) atomicOr(p.flags, 1u);
}

// ---------------------------------------------------------------------------------------------
// Case-control sweeps of large sparse networks (cfg 5: n = 50 000, ~220 list entries per node).
// Node j's conditional only reads the positions of the nodes in ITS lists (in/out edges, in/out
// controls).  dep[j] = 1 + max{i in lists(j) : i < j} (0 if none) is precomputed by k_cc_deps; then
// the consecutive nodes b, b+1, ..., e-1 can be updated CONCURRENTLY with the result of the
// sequential sweep iff dep[j] <= b for all of them: none reads a batch-mate with a smaller index
// (whose new value it would need), and batch-mates with a larger index are read at their old value
// because commits wait for a CTA barrier.  With lists of ~220 out of 50 000 nodes batches are ~15-20
// nodes long, so one CTA per (chain, slice) runs a warp per node of the batch instead of a team per
// node: the sweep is the same Markov transition, ~10x faster.
// ---------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(256) k_cc_deps(NetView net, int sets, int32_t *dep)
{
    const int lane = threadIdx.x & 31;
    const size_t wid = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t per_set = (size_t)net.T * net.n;
    if (wid >= (size_t)sets * per_set) return;
    const size_t r = wid % per_set;
    const int j = (int)(r % net.n);
    int best = -1;
    auto scan = [&](const int32_t *lst, int len) {
        for (int q = lane; q < len; q += 32) {
            const int i = lst[q];
            if (i >= 0 && i < j && i > best) best = i;
        }
    };
    scan(net.in_edges + r * net.max_in, net.deg[r * 2 + 0]);
    scan(net.out_edges + r * net.max_out, net.deg[r * 2 + 1]);
    const size_t coff = ((wid / per_set) * per_set + r) * net.n_control;
    scan(net.ctrl_in + coff, net.n_control);
    scan(net.ctrl_out + coff, net.n_control);
    for (int o = 16; o > 0; o >>= 1) {
        const int other = __shfl_xor_sync(kFull, best, o);
        best = other > best ? other : best;
    }
    if (lane == 0) dep[wid] = best + 1;
}

// grid = C*T (atomic ticket), block = 32*NW (NW = largest batch), positions in global memory;
// dynamic smem = 32*(d+5) doubles (stage) + 32 ints (decisions)
template <int D>
__global__ void __launch_bounds__(512, 1) k_sweep_cc(const SweepParams p, int *progress_g,
                                                     unsigned int *ticket, const int32_t *dep_all)
{
    constexpr int DM = (D == 0) ? kMaxD : D;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_ticket;
    const int T = p.net.T, n = p.net.n, d = (D == 0) ? p.net.d : D;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    if (threadIdx.x == 0) s_ticket = (int)atomicAdd(ticket, 1u);
    __syncthreads();
    const int c = s_ticket / T, t = s_ticket % T;
    double *Xchain = p.X + (size_t)c * T * n * d;
    double *Xt = Xchain + (size_t)t * n * d;
    double *st_prop = reinterpret_cast<double *>(smem_raw);
    double *st_logu = st_prop + 32 * d, *st_nn = st_logu + 32, *st_no = st_nn + 32, *st_inv = st_no + 32;
    int *st_zc = reinterpret_cast<int *>(st_inv + 32);
    int *st_acc = reinterpret_cast<int *>(st_prop + sweep_stage_doubles(d));
    int *st_dep = st_acc + 32;
    int *prog = progress_g + (size_t)c * T;
    const double *rinv = p.rinv + (size_t)c * n;
    const int32_t *dep = dep_all + ((size_t)(p.net.ctrl_per_chain ? c : 0) * T + t) * n;
    const double b0 = p.intercept[c * 2 + 0], b1 = p.intercept[c * 2 + 1];
    const uint32_t chain_id = (uint32_t)c + p.chain_offset;
    bool nonfinite = false;

    for (int jb = 0; jb < n; jb += 32) {
        // ---- lane-parallel preparation of 32 nodes (warp 0), as in k_sweep_slice ----
        const int jl = jb + lane;
        const bool mine = (warp == 0) && (jl < n);
        const size_t gs = ((size_t)c * T + t) * n + (jl < n ? jl : 0);
        double my_step = 0.0;
        int my_nacc = 0, my_nsteps = 0, my_until = 0;
        if (mine) {
            double eps[DM], x0[DM], x[DM], logu;
            load_pos<DM>(Xt + (size_t)jl * d, d, x0);
            my_step = p.step[gs]; my_nacc = p.nacc[gs]; my_nsteps = p.nsteps[gs]; my_until = p.until[gs];
            if (p.eps) {
#pragma unroll
                for (int k = 0; k < DM; k++) eps[k] = (k < d) ? p.eps[gs * d + k] : 0.0;
                logu = p.logu[gs];
            } else {
                latent_draws<DM>(p.seed, (uint32_t)(t * n + jl), p.sweep, chain_id, d, eps, logu);
            }
#pragma unroll
            for (int k = 0; k < DM; k++) {
                x[k] = (k < d) ? __dadd_rn(x0[k], __dmul_rn(my_step, eps[k])) : 0.0;
                if (k < d) st_prop[lane * d + k] = x[k];
            }
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
            if (t < T - 1) { // slice t+1 (another CTA) cannot have touched nodes >= jb yet
                double xnx[DM];
                const volatile double *q = Xchain + ((size_t)(t + 1) * n + jl) * d;
#pragma unroll
                for (int k = 0; k < DM; k++) xnx[k] = (k < d) ? q[k] : 0.0;
                nn = prior_next<DM>(p, c, t, jl, x, xnx);
                no = prior_next<DM>(p, c, t, jl, x0, xnx);
            }
            st_nn[lane] = nn;
            st_no[lane] = no;
            st_dep[lane] = dep[jl];
        }
        __syncthreads();
        const int jend = (n - jb) < 32 ? (n - jb) : 32;
        int b = 0;
        while (b < jend) {
            // batch [b, b+len): the longest run whose members do not read an earlier member
            // (every warp derives the same length from the same dep entries)
            const int cand = b + lane;
            const bool fits = lane < nwarps && cand < jend && st_dep[cand] <= jb + b;
            const unsigned run = __ballot_sync(kFull, fits);
            const int len = (run == kFull) ? 32 : __ffs(~run) - 1; // >= 1: dep[j] <= j always
            int acc = 0;
            const int jj = b + warp, j = jb + jj;
            if (warp < len) {
                double x[DM], x0[DM];
                load_pos<DM>(st_prop + jj * d, d, x);
                load_pos<DM>(Xt + (size_t)j * d, d, x0);
                // X[t-1, j]: look at the wavefront flag and fetch the neighbour slice's position
                // BEFORE the list walk, so that their L2 round trips overlap it; in steady state
                // slice t-1 is ahead and the early copy is the one that is used
                double xp[DM];
#pragma unroll
                for (int k = 0; k < DM; k++) xp[k] = 0.0;
                bool early = false;
                if (t > 0) {
                    early = ld_acquire_gpu(prog + t - 1) > j;
                    const volatile double *q = Xchain + ((size_t)(t - 1) * n + j) * d;
#pragma unroll
                    for (int k = 0; k < DM; k++) if (k < d) xp[k] = q[k];
                }
                double ll_new, ll_old;
                node_loglik2<kCaseControl, DM>(p.net, Xt, rinv, c, t, j, x, x0, b0, b1, lane, ll_new,
                                               ll_old, p.flags);
                if (t > 0 && !early) {
                    while (ld_acquire_gpu(prog + t - 1) <= j) { /* spin on the L2-resident flag of slice t-1 */ }
                    const volatile double *q = Xchain + ((size_t)(t - 1) * n + j) * d;
#pragma unroll
                    for (int k = 0; k < DM; k++) if (k < d) xp[k] = q[k];
                }
                const double inv = st_inv[jj];
                const int zc = st_zc[jj];
                double lp_new = __dsub_rn(ll_new, prior_prev<DM>(p, c, t, zc, inv, x, xp));
                double lp_old = __dsub_rn(ll_old, prior_prev<DM>(p, c, t, zc, inv, x0, xp));
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
            if (warp < len && acc && lane < d) Xt[(size_t)j * d + lane] = st_prop[jj * d + lane];
            __syncthreads();
            if (threadIdx.x == 0) st_release_gpu(prog + t, jb + b + len);
            b += len;
        }
        if (mine) {
            metropolis_bookkeep(my_step, my_nacc, my_nsteps, my_until, p.tune, p.tune_interval,
                                st_acc[lane], false);
            p.step[gs] = my_step; p.nacc[gs] = my_nacc; p.nsteps[gs] = my_nsteps; p.until[gs] = my_until;
            if (p.accepted) p.accepted[gs] = st_acc[lane];
        }
        __syncthreads();
    }
    if (nonfinite) atomicOr(p.flags, 1u);
}

// the (j, i) dyad's contribution to node j's two log-likelihoods (proposal xn, current xo)
template <int LK, int DM>
__device__ __forceinline__ void pair_term2(const NetView &net, const double *Xt, const double *rinv,
                                           int t, int j, int i, const double (&xn)[DM],
                                           const double (&xo)[DM], double b0, double b1, double &tn,
                                           double &to)
{
    const int n = net.n, d = latent_dim<DM>(net);
    double xi[DM];
    load_pos<DM>(Xt + (size_t)i * d, d, xi);
    const size_t wo = ((size_t)t * n + j) * net.W + (i >> 5);
    const double dn = fast_dist<DM>(xi, xn, d), dd = fast_dist<DM>(xi, xo, d);
    if (LK == kUndirected) {
        const double y = ymask(__ldg(net.rowbits + wo), i & 31);
        tn = logit_term(y, b0 - dn);
        to = logit_term(y, b0 - dd);
    } else {
        const double y_ji = ymask(__ldg(net.rowbits + wo), i & 31);
        const double y_ij = ymask(__ldg(net.colbits + wo), i & 31);
        const double ri = rinv[i], rj = rinv[j];
        tn = logit_term(y_ji, eta_directed(b0, b1, dn, ri, rj)) + logit_term(y_ij, eta_directed(b0, b1, dn, rj, ri));
        to = logit_term(y_ji, eta_directed(b0, b1, dd, ri, rj)) + logit_term(y_ij, eta_directed(b0, b1, dd, rj, ri));
    }
}

// ---------------------------------------------------------------------------------------------
// k_sweep_slice_ws: the CTA-per-(chain, slice) sweep, warp-specialised and software-pipelined.
// Warp 0 is the CONTROL warp (wavefront flag, priors, accept/reject, commit); warps 1..NW-1 are
// COMPUTE warps (the pairwise reduction).  Node j+1's row depends on node j's decision through ONE
// dyad only, (j+1, j): the compute warps therefore start node j+1 as soon as node j-1 is decided,
// always leaving out i = j, and the control warp adds that single dyad once it has decided node j.
// The serial tail of a node (L2 round trips for the flag and the neighbour slice, fences, prior)
// then overlaps the next node's reduction instead of adding to it.  Exact (undirected / directed)
// likelihoods only; the case-control lists use k_sweep_slice.
// dynamic smem = [n*(d [+1]) doubles if XS] + 32*(d+5) doubles + 4*NW doubles + 2*NW ints
// ---------------------------------------------------------------------------------------------
template <int LK, int D, bool XS>
__global__ void __launch_bounds__(576) k_sweep_slice_ws(const SweepParams p, int *progress_g,
                                                        unsigned int *ticket)
{
    constexpr int DM = (D == 0) ? kMaxD : D;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_ticket;
    __shared__ volatile int s_decided; // nodes 0 .. s_decided-1 of this slice are final in Xt
    const int T = p.net.T, n = p.net.n, d = (D == 0) ? p.net.d : D;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int ncomp = nwarps - 1;
    if (threadIdx.x == 0) { s_ticket = (int)atomicAdd(ticket, 1u); s_decided = 0; }
    __syncthreads();
    const int c = s_ticket / T, t = s_ticket % T;
    double *Xchain = p.X + (size_t)c * T * n * d;
    double *Xg = Xchain + (size_t)t * n * d;
    double *Xt, *stage_base;
    const double *rinv = (LK == kUndirected) ? nullptr : p.rinv + (size_t)c * n;
    if (XS) {
        Xt = reinterpret_cast<double *>(smem_raw);
        stage_base = Xt + (size_t)n * d;
        for (int e = threadIdx.x; e < n * d; e += blockDim.x) Xt[e] = Xg[e];
        if (LK != kUndirected) {
            double *s_rinv = stage_base;
            stage_base += (n + 1) & ~1; // keep the stage 16-byte aligned (double2 loads)
            for (int e = threadIdx.x; e < n; e += blockDim.x) s_rinv[e] = rinv[e];
            rinv = s_rinv;
        }
    } else {
        Xt = Xg;
        stage_base = reinterpret_cast<double *>(smem_raw);
    }
    double *st_prop = stage_base;
    double *st_logu = st_prop + 32 * d, *st_nn = st_logu + 32, *st_no = st_nn + 32, *st_inv = st_no + 32;
    int *st_zc = reinterpret_cast<int *>(st_inv + 32);
    double *part = stage_base + sweep_stage_doubles(d);          // [2][nwarps][2]
    volatile int *done = reinterpret_cast<volatile int *>(part + 4 * nwarps); // [nwarps] nodes summed
    int *prog = progress_g + (size_t)c * T;
    for (int w = threadIdx.x; w < nwarps; w += blockDim.x) done[w] = 0;
    __syncthreads();

    const double b0 = p.intercept[c * 2 + 0], b1 = p.intercept[c * 2 + 1];
    const uint32_t chain_id = (uint32_t)c + p.chain_offset;
    bool nonfinite = false;

    for (int jb = 0; jb < n; jb += 32) {
        // ---- lane-parallel preparation of the block (control warp) ----
        const int jl = jb + lane;
        const bool mine = (warp == 0) && (jl < n);
        const size_t gs = ((size_t)c * T + t) * n + (jl < n ? jl : 0);
        double my_step = 0.0;
        int my_nacc = 0, my_nsteps = 0, my_until = 0, my_acc = 0;
        if (mine) {
            double eps[DM], x0[DM], x[DM], logu;
            load_pos<DM>(Xt + (size_t)jl * d, d, x0);
            my_step = p.step[gs]; my_nacc = p.nacc[gs]; my_nsteps = p.nsteps[gs]; my_until = p.until[gs];
            if (p.eps) {
#pragma unroll
                for (int k = 0; k < DM; k++) eps[k] = (k < d) ? p.eps[gs * d + k] : 0.0;
                logu = p.logu[gs];
            } else {
                latent_draws<DM>(p.seed, (uint32_t)(t * n + jl), p.sweep, chain_id, d, eps, logu);
            }
#pragma unroll
            for (int k = 0; k < DM; k++) {
                x[k] = (k < d) ? __dadd_rn(x0[k], __dmul_rn(my_step, eps[k])) : 0.0;
                if (k < d) st_prop[lane * d + k] = x[k];
            }
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
            if (t < T - 1) {
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
        __syncthreads();
        const int jend = (n - jb) < 32 ? (n - jb) : 32;
        if (warp > 0) {
            // ---------------- compute warps: pairwise sums, one node ahead of the decisions ------
            for (int jj = 0; jj < jend; jj++) {
                const int j = jb + jj;
                while (s_decided < j - 1) { /* nodes < j-1 must be final; node j-1 is left out */ }
                __threadfence_block();
                double x[DM], x0[DM];
                load_pos<DM>(st_prop + jj * d, d, x);
                load_pos<DM>(Xt + (size_t)j * d, d, x0);
                if (j + 1 < n && (warp - 1) == 0 && lane * 32 < p.net.W) {
                    const size_t o = ((size_t)t * n + j + 1) * p.net.W + lane * 32;
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(p.net.rowbits + o));
                    if (LK == kDirected) asm volatile("prefetch.global.L1 [%0];" ::"l"(p.net.colbits + o));
                }
                double ll_new, ll_old;
                node_loglik2<LK, DM>(p.net, Xt, rinv, c, t, j, x, x0, b0, b1, lane, ll_new, ll_old,
                                     p.flags, warp - 1, ncomp, j - 1);
                if (lane == 0) {
                    double *pp = part + ((j & 1) * nwarps + warp) * 2;
                    pp[0] = ll_new; pp[1] = ll_old;
                    __threadfence_block();
                    done[warp] = j + 1;
                }
                __syncwarp();
            }
        } else {
            // ---------------- control warp: finish node j while node j+1 is being summed --------
            for (int jj = 0; jj < jend; jj++) {
                const int j = jb + jj;
                double x[DM], x0[DM];
                load_pos<DM>(st_prop + jj * d, d, x);
                load_pos<DM>(Xt + (size_t)j * d, d, x0);
                double tn = 0.0, to = 0.0;
                if (j > 0) pair_term2<LK, DM>(p.net, Xt, rinv, t, j, j - 1, x, x0, b0, b1, tn, to);
                double xp[DM];
#pragma unroll
                for (int k = 0; k < DM; k++) xp[k] = 0.0;
                if (t > 0) {
                    while (ld_acquire_gpu(prog + t - 1) <= j) { /* spin on the L2-resident flag of slice t-1 */ }
                    const volatile double *q = Xchain + ((size_t)(t - 1) * n + j) * d;
#pragma unroll
                    for (int k = 0; k < DM; k++) if (k < d) xp[k] = q[k];
                }
                const double inv = st_inv[jj];
                const int zc = st_zc[jj];
                const double pr_new = prior_prev<DM>(p, c, t, zc, inv, x, xp);
                const double pr_old = prior_prev<DM>(p, c, t, zc, inv, x0, xp);
                // the row sums of the compute warps, in warp order, then the deferred dyad
                for (int w = 1; w < nwarps; w++)
                    while (done[w] <= j) { /* spin */ }
                __threadfence_block();
                double ll_new = 0.0, ll_old = 0.0;
                const double *pp = part + (j & 1) * nwarps * 2;
                for (int w = 1; w < nwarps; w++) { ll_new += pp[w * 2]; ll_old += pp[w * 2 + 1]; }
                ll_new += tn;
                ll_old += to;
                double lp_new = __dsub_rn(ll_new, pr_new), lp_old = __dsub_rn(ll_old, pr_old);
                if (t < T - 1) {
                    lp_new = __dsub_rn(lp_new, st_nn[jj]);
                    lp_old = __dsub_rn(lp_old, st_no[jj]);
                }
                const double ratio = __dsub_rn(lp_new, lp_old);
                const int acc = (st_logu[jj] >= ratio) ? 0 : 1;
                const bool me = lane == jj;
                my_acc = me ? acc : my_acc;
                nonfinite |= me && (!(ratio == ratio) || ratio - ratio != 0.0);
                if (me) {
                    if (acc) {
#pragma unroll
                        for (int k = 0; k < DM; k++)
                            if (k < d) {
                                Xt[(size_t)j * d + k] = x[k];
                                if (XS) Xg[(size_t)j * d + k] = x[k];
                            }
                    }
                    if (p.ratio) p.ratio[((size_t)c * T + t) * n + j] = ratio;
                    __threadfence_block();
                    s_decided = j + 1;       // releases the compute warps' node j+2
                    st_release_gpu(prog + t, j + 1); // releases slice t+1's node j
                }
                __syncwarp();
            }
        }
        __syncthreads();
        if (mine) {
            metropolis_bookkeep(my_step, my_nacc, my_nsteps, my_until, p.tune, p.tune_interval,
                                my_acc, false);
            p.step[gs] = my_step; p.nacc[gs] = my_nacc; p.nsteps[gs] = my_nsteps; p.until[gs] = my_until;
            if (p.accepted) p.accepted[gs] = my_acc;
        }
        __syncthreads();
    }
    if (nonfinite) atomicOr(p.flags, 1u);
}

// ---------------------------------------------------------------------------------------------
// k_sweep_slice_cl: a thread-block CLUSTER per (chain, slice) -- the single-chain mapping (cfg 3:
// one chain, n = 2 000, T = 20, where CTA-per-slice leaves 128 of the 148 SMs idle).
// The CS CTAs of a cluster each hold a copy of the slice's positions (+ reciprocal radii) in their
// own shared memory and reduce 1/CS of every node's row: compute warp w of CTA r is member
// r * NCOMP + w of a team of CS * NCOMP warps (node_loglik2's wteam / nteam).  Partial sums travel
// through DISTRIBUTED SHARED MEMORY: lane 0 of a compute warp stores its pair into the leader
// CTA's slot (st.shared::cluster) and arrives on the leader's mbarrier
// (mbarrier.arrive.release.cluster); the leader's control warp waits for the phase, adds the
// partials in team order, adds the deferred (j, j-1) dyad, decides, and broadcasts the decision:
// the accepted position into every CTA's copy of the slice, then st.release.cluster of that CTA's
// `decided` counter, which its compute warps poll locally (ld.acquire.cluster).  As in
// k_sweep_slice_ws the compute warps run one node ahead of the decisions, so the DSMEM round trip
// (~2 x 215 cycles) overlaps the next node's reduction.  The wavefront across slices still goes
// through the L2 flags (prog[], st.release.gpu / ld.acquire.gpu), and clusters take their (chain,
// slice) from an atomic ticket: a cluster only waits on clusters that have started.
// Every CTA's warp 0 stages the proposals of a 32-node block redundantly (same inputs, same
// arithmetic), so nothing but partial sums and decisions crosses the cluster.
// grid = C*T*CS, cluster = (CS,1,1), block = 32 * (1 + NCOMP);
// dynamic smem = n*(d [+1]) doubles + 32*(d+5) doubles + 2 * kMaxTeam * 2 doubles
// ---------------------------------------------------------------------------------------------
constexpr int kMaxTeam = 128; // CS * NCOMP <= 8 * 16

__device__ __forceinline__ uint32_t cl_rank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cl_size()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cl_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t cl_map(uint32_t addr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cl_st_f64(uint32_t addr, double v)
{
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ int cl_ld_s32(uint32_t addr)
{
    int v;
    asm volatile("ld.shared::cluster.s32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void cl_st_release_s32(uint32_t addr, int v)
{
    asm volatile("st.release.cluster.shared::cluster.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ int cl_ld_acquire_local_s32(uint32_t addr)
{
    int v;
    asm volatile("ld.acquire.cluster.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void mbar_init(uint32_t addr, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr)
{
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t addr, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t}" ::"r"(addr), "r"(parity) : "memory");
}

template <int LK, int D>
__global__ void __launch_bounds__(512, 1) k_sweep_slice_cl(const SweepParams p, int *progress_g,
                                                           unsigned int *ticket)
{
    constexpr int DM = (D == 0) ? kMaxD : D;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long s_bar[2]; // leader: partials of node j arrive on s_bar[j & 1]
    __shared__ int s_ticket;
    __shared__ int s_decided; // nodes 0 .. s_decided-1 of this slice are final in THIS CTA's copy
    const int T = p.net.T, n = p.net.n, d = (D == 0) ? p.net.d : D;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int ncomp = nwarps - 1;
    const int rank = (int)cl_rank(), CS = (int)cl_size();
    const int nteam = CS * ncomp;
    const bool leader = rank == 0;
    if (threadIdx.x == 0) {
        s_decided = 0;
        if (leader) {
            s_ticket = (int)atomicAdd(ticket, 1u);
            mbar_init(smem_addr(&s_bar[0]), (uint32_t)nteam);
            mbar_init(smem_addr(&s_bar[1]), (uint32_t)nteam);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    cl_sync();
    const int tk = cl_ld_s32(cl_map(smem_addr(&s_ticket), 0));
    const int c = tk / T, t = tk % T;
    double *Xchain = p.X + (size_t)c * T * n * d;
    double *Xg = Xchain + (size_t)t * n * d;
    double *Xt = reinterpret_cast<double *>(smem_raw);
    double *stage_base = Xt + (((size_t)n * d + 1) & ~(size_t)1);
    const double *rinv = nullptr;
    for (int e = threadIdx.x; e < n * d; e += blockDim.x) Xt[e] = Xg[e];
    if (LK != kUndirected) {
        double *s_rinv = stage_base;
        stage_base += (n + 1) & ~1; // keep the stage 16-byte aligned (double2 loads)
        const double *rg = p.rinv + (size_t)c * n;
        for (int e = threadIdx.x; e < n; e += blockDim.x) s_rinv[e] = rg[e];
        rinv = s_rinv;
    }
    double *st_prop = stage_base;
    double *st_logu = st_prop + 32 * d, *st_nn = st_logu + 32, *st_no = st_nn + 32, *st_inv = st_no + 32;
    int *st_zc = reinterpret_cast<int *>(st_inv + 32);
    double *part = stage_base + sweep_stage_doubles(d); // leader: [2][kMaxTeam][2]
    int *prog = progress_g + (size_t)c * T;
    __syncthreads();

    const double b0 = p.intercept[c * 2 + 0], b1 = p.intercept[c * 2 + 1];
    const uint32_t chain_id = (uint32_t)c + p.chain_offset;
    const uint32_t a_decided = smem_addr(&s_decided);
    const uint32_t a_part0 = cl_map(smem_addr(part), 0);          // the leader's slots and barriers
    const uint32_t a_bar0 = cl_map(smem_addr(&s_bar[0]), 0), a_bar1 = cl_map(smem_addr(&s_bar[1]), 0);
    bool nonfinite = false;

    for (int jb = 0; jb < n; jb += 32) {
        // ---- lane-parallel preparation of the block: every CTA stages the proposals, the leader
        //      also the sampler state, the uniforms and the "next" prior terms ----
        const int jl = jb + lane;
        const bool mine = (warp == 0) && (jl < n);
        const size_t gs = ((size_t)c * T + t) * n + (jl < n ? jl : 0);
        double my_step = 0.0;
        int my_nacc = 0, my_nsteps = 0, my_until = 0, my_acc = 0;
        if (mine) {
            double eps[DM], x0[DM], x[DM], logu;
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
            for (int k = 0; k < DM; k++) {
                x[k] = (k < d) ? __dadd_rn(x0[k], __dmul_rn(my_step, eps[k])) : 0.0;
                if (k < d) st_prop[lane * d + k] = x[k];
            }
            if (leader) {
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
                if (t < T - 1) {
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
        __syncthreads();
        const int jend = (n - jb) < 32 ? (n - jb) : 32;
        if (warp > 0) {
            // ---------------- compute warps of every CTA: 1 / (CS * NCOMP) of each row ----------
            const int wteam = rank * ncomp + (warp - 1);
            for (int jj = 0; jj < jend; jj++) {
                const int j = jb + jj;
                while (cl_ld_acquire_local_s32(a_decided) < j - 1) { /* node j-1 is left out */ }
                double x[DM], x0[DM];
                load_pos<DM>(st_prop + jj * d, d, x);
                load_pos<DM>(Xt + (size_t)j * d, d, x0);
                double ll_new, ll_old;
                node_loglik2<LK, DM>(p.net, Xt, rinv, c, t, j, x, x0, b0, b1, lane, ll_new, ll_old, p.flags,
                                     wteam, nteam, j - 1);
                if (lane == 0) {
                    const uint32_t slot = a_part0 + (uint32_t)(((j & 1) * kMaxTeam + wteam) * 2 * sizeof(double));
                    cl_st_f64(slot, ll_new);
                    cl_st_f64(slot + 8, ll_old);
                    mbar_arrive_remote((j & 1) ? a_bar1 : a_bar0); // release: the two stores come first
                }
                __syncwarp();
            }
        } else if (leader) {
            // ---------------- control warp of the leader CTA ------------------------------------
            for (int jj = 0; jj < jend; jj++) {
                const int j = jb + jj;
                double x[DM], x0[DM];
                load_pos<DM>(st_prop + jj * d, d, x);
                load_pos<DM>(Xt + (size_t)j * d, d, x0);
                double tn = 0.0, to = 0.0;
                if (j > 0) pair_term2<LK, DM>(p.net, Xt, rinv, t, j, j - 1, x, x0, b0, b1, tn, to);
                double xp[DM];
#pragma unroll
                for (int k = 0; k < DM; k++) xp[k] = 0.0;
                if (t > 0) {
                    while (ld_acquire_gpu(prog + t - 1) <= j) { /* spin on the L2-resident flag of slice t-1 */ }
                    const volatile double *q = Xchain + ((size_t)(t - 1) * n + j) * d;
#pragma unroll
                    for (int k = 0; k < DM; k++) if (k < d) xp[k] = q[k];
                }
                const double inv = st_inv[jj];
                const int zc = st_zc[jj];
                const double pr_new = prior_prev<DM>(p, c, t, zc, inv, x, xp);
                const double pr_old = prior_prev<DM>(p, c, t, zc, inv, x0, xp);
                mbar_wait(smem_addr(&s_bar[j & 1]), (uint32_t)((j >> 1) & 1)); // all partials of node j are in
                double ll_new = 0.0, ll_old = 0.0;
                const volatile double *pp = part + (size_t)(j & 1) * kMaxTeam * 2;
                for (int w = 0; w < nteam; w++) { ll_new += pp[w * 2]; ll_old += pp[w * 2 + 1]; }
                ll_new += tn;
                ll_old += to;
                double lp_new = __dsub_rn(ll_new, pr_new), lp_old = __dsub_rn(ll_old, pr_old);
                if (t < T - 1) {
                    lp_new = __dsub_rn(lp_new, st_nn[jj]);
                    lp_old = __dsub_rn(lp_old, st_no[jj]);
                }
                const double ratio = __dsub_rn(lp_new, lp_old);
                const int acc = (st_logu[jj] >= ratio) ? 0 : 1;
                my_acc = (lane == jj) ? acc : my_acc;
                nonfinite |= (lane == jj) && (!(ratio == ratio) || ratio - ratio != 0.0);
                if (lane < CS) { // lane r publishes the decision to CTA r (lane 0: this CTA and global)
                    if (acc) {
                        const uint32_t xr = cl_map(smem_addr(Xt + (size_t)j * d), (uint32_t)lane);
#pragma unroll
                        for (int k = 0; k < DM; k++)
                            if (k < d) cl_st_f64(xr + 8 * k, x[k]);
                    }
                    cl_st_release_s32(cl_map(a_decided, (uint32_t)lane), j + 1);
                } else if (lane == CS) {
                    if (acc) {
#pragma unroll
                        for (int k = 0; k < DM; k++)
                            if (k < d) Xg[(size_t)j * d + k] = x[k];
                    }
                    if (p.ratio) p.ratio[((size_t)c * T + t) * n + j] = ratio;
                    st_release_gpu(prog + t, j + 1); // releases slice t+1's node j
                }
                __syncwarp();
            }
        }
        __syncthreads();
        if (mine && leader) {
            metropolis_bookkeep(my_step, my_nacc, my_nsteps, my_until, p.tune, p.tune_interval,
                                my_acc, false);
            p.step[gs] = my_step; p.nacc[gs] = my_nacc; p.nsteps[gs] = my_nsteps; p.until[gs] = my_until;
            if (p.accepted) p.accepted[gs] = my_acc;
        }
        // (no cluster barrier per block: a block's step sizes are read by every CTA before any of its
        //  partial sums can reach the leader, and written back by the leader only after all of them)
    }
    if (nonfinite) atomicOr(p.flags, 1u);
    cl_sync(); // no CTA may exit while a peer can still write into its shared memory
}

// ---------------------------------------------------------------------------------------------
// parity probe: per-node log-likelihood at the current state, one warp per (c, t, j)
// ---------------------------------------------------------------------------------------------
template <int LK, int D>
__global__ void k_partial(const SweepParams p, double *out)
{
    constexpr int DM = (D == 0) ? kMaxD : D;
    const int T = p.net.T, n = p.net.n, d = (D == 0) ? p.net.d : D;
    const size_t gw = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= (size_t)p.C * T * n) return;
    const int c = (int)(gw / ((size_t)T * n));
    const int t = (int)((gw / n) % T), j = (int)(gw % n);
    const double *Xt = p.X + ((size_t)c * T + t) * n * d;
    double x0[DM];
    load_pos<DM>(Xt + (size_t)j * d, d, x0);
    double a, b;
    node_loglik2<LK, DM>(p.net, Xt, (LK == kUndirected) ? nullptr : p.rinv + (size_t)c * n, c, t,
                         j, x0, x0, p.intercept[c * 2], p.intercept[c * 2 + 1], lane, a, b,
                         p.flags);
    if (lane == 0) out[gw] = b;
}

// the raw Philox draws of the next native sweep (parity probe for the device RNG path)
template <int D>
__global__ void k_debug_draws(const SweepParams p, double *eps_out, double *logu_out)
{
    constexpr int DM = (D == 0) ? kMaxD : D;
    const int T = p.net.T, n = p.net.n, d = (D == 0) ? p.net.d : D;
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (size_t)p.C * T * n) return;
    const int c = (int)(g / ((size_t)T * n));
    const uint32_t site = (uint32_t)(g % ((size_t)T * n));
    double eps[DM], logu;
    latent_draws<DM>(p.seed, site, p.sweep, (uint32_t)c + p.chain_offset, d, eps, logu);
    for (int k = 0; k < d; k++) eps_out[g * d + k] = eps[k];
    logu_out[g] = logu;
}

// ---------------------------------------------------------------------------------------------
// X -= np.mean(X, axis=(0,1))  (lsm.py:501): serial accumulation per column = numpy's order
// ---------------------------------------------------------------------------------------------
static __global__ void k_center(double *X, int T, int n, int d)
{
    __shared__ double mean[kMaxD];
    double *Xc = X + (size_t)blockIdx.x * T * n * d;
    const size_t rows = (size_t)T * n;
    if ((int)threadIdx.x < d) {
        // eight loads in flight per step keep the serial DADD chain (8 cycles each) fed
        double s = 0.0;
        size_t r = 0;
        for (; r + 8 <= rows; r += 8) {
            double v[8];
#pragma unroll
            for (int q = 0; q < 8; q++) v[q] = Xc[(r + q) * d + threadIdx.x];
#pragma unroll
            for (int q = 0; q < 8; q++) s = __dadd_rn(s, v[q]);
        }
        for (; r < rows; r++) s = __dadd_rn(s, Xc[r * d + threadIdx.x]);
        mean[threadIdx.x] = __ddiv_rn(s, (double)rows);
    }
    __syncthreads();
    for (size_t e = threadIdx.x; e < rows * d; e += blockDim.x)
        Xc[e] = __dsub_rn(Xc[e], mean[e % d]);
}

// Long chains (cfg 5: T n = 500 000 rows): the column means in two shapes, the subtraction spread
// over (B, C) CTAs.
//   k_center_mean_exact  numpy's serial order, bit-identical to k_center: the CTA stages 2048-row
//                        chunks in shared memory (coalesced), thread k < d adds them in order --
//                        bounded by the serial DADD chain (~8 cycles per row), not by load latency
//   k_center_mean_tree   device loop only: per-(chain, block) partial sums + a fixed-order total;
//                        deterministic, but not numpy's rounding
static __global__ void __launch_bounds__(256) k_center_mean_exact(const double *X, int T, int n, int d, double *means)
{
    __shared__ double buf[2048 * kMaxD / 4]; // 2048 rows at d = 2, 512 rows at d = 8
    const double *Xc = X + (size_t)blockIdx.x * T * n * d;
    const size_t rows = (size_t)T * n;
    const int chunk = (2048 * kMaxD / 4) / d;
    double s = 0.0;
    for (size_t r0 = 0; r0 < rows; r0 += chunk) {
        const int cnt = (int)((rows - r0) < (size_t)chunk ? (rows - r0) : (size_t)chunk);
        __syncthreads();
        for (int e = threadIdx.x; e < cnt * d; e += blockDim.x) buf[e] = Xc[r0 * d + e];
        __syncthreads();
        if ((int)threadIdx.x < d) {
            int r = 0;
            for (; r + 8 <= cnt; r += 8) {
                double v[8];
#pragma unroll
                for (int q = 0; q < 8; q++) v[q] = buf[(r + q) * d + threadIdx.x];
#pragma unroll
                for (int q = 0; q < 8; q++) s = __dadd_rn(s, v[q]);
            }
            for (; r < cnt; r++) s = __dadd_rn(s, buf[r * d + threadIdx.x]);
        }
    }
    if ((int)threadIdx.x < d) means[blockIdx.x * kMaxD + threadIdx.x] = __ddiv_rn(s, (double)rows);
}

__device__ __forceinline__ double block_sum(double v, double *sh);

// grid (B, C): partial[c][b][k] = sum over the rows of block b
static __global__ void __launch_bounds__(256) k_center_partial(const double *X, int T, int n, int d, double *partial)
{
    __shared__ double sh[8];
    const int B = gridDim.x, b = blockIdx.x, c = blockIdx.y;
    const double *Xc = X + (size_t)c * T * n * d;
    const size_t rows = (size_t)T * n, per = (rows + B - 1) / B;
    const size_t lo = per * b, hi = (lo + per) < rows ? (lo + per) : rows;
    double acc[kMaxD];
#pragma unroll
    for (int k = 0; k < kMaxD; k++) acc[k] = 0.0;
    for (size_t r = lo + threadIdx.x; r < hi; r += blockDim.x)
#pragma unroll
        for (int k = 0; k < kMaxD; k++)
            if (k < d) acc[k] += Xc[r * d + k];
    for (int k = 0; k < d; k++) {
        const double tot = block_sum(acc[k], sh);
        if (threadIdx.x == 0) partial[((size_t)c * B + b) * kMaxD + k] = tot;
    }
}

static __global__ void k_center_total(const double *partial, int B, int d, double rows, double *means)
{
    const int c = blockIdx.x, k = threadIdx.x;
    if (k >= d) return;
    double s = 0.0;
    for (int b = 0; b < B; b++) s += partial[((size_t)c * B + b) * kMaxD + k];
    means[c * kMaxD + k] = s / rows;
}

// grid (B, C): X -= mean (exactly rounded subtraction, as numpy's in-place -=)
static __global__ void __launch_bounds__(256) k_center_apply(double *X, int T, int n, int d, const double *means)
{
    __shared__ double mean[kMaxD];
    const int c = blockIdx.y;
    if ((int)threadIdx.x < d) mean[threadIdx.x] = means[c * kMaxD + threadIdx.x];
    __syncthreads();
    double *Xc = X + (size_t)c * T * n * d;
    const size_t total = (size_t)T * n * d;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x)
        Xc[e] = __dsub_rn(Xc[e], mean[e % d]);
}

static __global__ void k_rinv(const double *radii, double *rinv, size_t total)
{
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < total) rinv[g] = 1.0 / radii[g];
}

// ---------------------------------------------------------------------------------------------
// adjacency bit-packing from the dense fp64 Y the reference's fit() takes (lsm.py:319-343)
// one warp per 32-column word, ballot over the lanes; bad[0] set if an entry is not 0/1
// ---------------------------------------------------------------------------------------------
static __global__ void k_pack_rows(const double *Y, int n, int W, uint32_t *bits, int *bad)
{
    // Y: one time slice [n][n]; bits [n][W]
    const size_t gw = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= (size_t)n * W) return;
    const int i = (int)(gw / W), w = (int)(gw % W);
    const int j = w * 32 + lane;
    double v = 0.0;
    if (j < n) v = Y[(size_t)i * n + j];
    if (v != 0.0 && v != 1.0) *bad = 1;
    const unsigned word = __ballot_sync(kFull, v == 1.0);
    if (lane == 0) bits[(size_t)i * W + w] = word;
}

static __global__ void k_pack_cols(const double *Y, int n, int W, uint32_t *bits)
{
    // bits[i][w] bit b = Y[w*32+b][i]; thread per (w, i) with i fastest (coalesced reads)
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (size_t)n * W) return;
    const int i = (int)(g % n), w = (int)(g / n);
    uint32_t word = 0;
    for (int b = 0; b < 32; b++) {
        const int j = w * 32 + b;
        if (j < n && Y[(size_t)j * n + i] == 1.0) word |= (1u << b);
    }
    bits[(size_t)i * W + w] = word;
}

// ---------------------------------------------------------------------------------------------
// Full-network log-likelihood, two parameter variants per pass.
//   variant v uses intercepts bvar[c][v][0..1] and reciprocal radii rinv_v[c][:]
// grid = (T * tiles, C), block = 256; CTA (t, tile) stages X[c, t] in shared memory and its warps
// take rows i = tile + k * tiles.  Output: partial[c][blockIdx.x][v].
// ---------------------------------------------------------------------------------------------
struct FullParams {
    NetView net;
    int C, tiles;
    const double *X;        // [C][T][n][d]
    const double *bvar;     // [C][2][2]
    const double *rinv0;    // [C][n] variant 0
    const double *rinv1;    // [C][n] variant 1
    double *partial;        // [C][T*tiles][2]
    unsigned int *flags;
    const double *gather;   // optional [C][T][n][4] = {x0, x1, 1/r (variant 0), 0}: d = 2 case-control rows
    int same_r;             // both variants use rinv0 (then the gather records serve NV = 2 as well)
};

// one 256-bit load (LDG.E.256 on sm_100a): a whole 32-byte gather record with a single L1 wavefront
__device__ __forceinline__ void ld256(const double *p, double &a, double &b, double &c, double &d)
{
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}

// G[c][t][j] = {X[c,t,j,0], X[c,t,j,1], rinv[c,j], 0}
static __global__ void k_pack_gather(const double *X, const double *rinv, double *G, int C, int T, int n)
{
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)C * T * n;
    if (g >= total) return;
    const size_t c = g / ((size_t)T * n);
    const int j = (int)(g % n);
    double2 x = reinterpret_cast<const double2 *>(X)[g];
    double4 o = make_double4(x.x, x.y, rinv[c * n + j], 0.0);
    reinterpret_cast<double4 *>(G)[g] = o;
}

// (the packed case-control rows inline the two wrappers of the table-driven softplus)
#ifdef DLSM_NAIVE_SOFTPLUS
constexpr bool kFusedSoftplusForms = false;
#else
constexpr bool kFusedSoftplusForms = true;
#endif

// NV = 2: proposal and current variants; NV = 1: proposal only (the device loop tracks the current
// state's log-likelihood itself, see SweepParams::ll_cur)
template <int LK, int D, int NV>
__global__ void __launch_bounds__(256) k_full(const FullParams p)
{
    constexpr int DM = (D == 0) ? kMaxD : D;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red[2][8];
    const int T = p.net.T, n = p.net.n, d = (D == 0) ? p.net.d : D, W = p.net.W;
    const int c = blockIdx.y;
    const int t = blockIdx.x / p.tiles, tile = blockIdx.x % p.tiles;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const double *Xg = p.X + ((size_t)c * T + t) * n * d;
    const double *Xt;
    if (LK != kCaseControl) {
        double *Xs = reinterpret_cast<double *>(smem_raw);
        for (int e = threadIdx.x; e < n * d; e += blockDim.x) Xs[e] = Xg[e];
        __syncthreads();
        Xt = Xs;
    } else {
        Xt = Xg;
    }
    const double *bv = p.bvar + (size_t)c * 4;
    const double b00 = bv[0], b01 = bv[1], b10 = bv[2], b11 = bv[3];
    const double *r0 = p.rinv0 ? p.rinv0 + (size_t)c * n : nullptr;
    const double *r1 = p.rinv1 ? p.rinv1 + (size_t)c * n : nullptr;
    double a0 = 0.0, a1 = 0.0;

    if (LK != kCaseControl) {
        // K5 network_likelihoods.py:26-33 / K4 directed_likelihoods_fast.pyx:185-205.
        // Unordered pairs a < b of one slice, visited as FOLDED rows: row i (n-1-i pairs) is glued
        // to row n-1-i (i pairs), so every folded row has n-1 pair slots and the lanes stay full
        // (a plain triangular sweep leaves ~30 % of the lane slots empty).  Two 32-slot chunks
        // per trip and two parameter variants per pair = four independent softplus chains.
        const int half = (n + 1) >> 1;
        for (int i = tile + warp * p.tiles; i < half; i += nwarps * p.tiles) {
            const int i2 = n - 1 - i;
            const int len1 = n - 1 - i;                       // pairs (i, i+1+m)
            const int slots = (i2 == i) ? len1 : n - 1;       // + pairs (i2, i2+1+k), k < i
            double xi[DM], xi2[DM];
            load_pos<DM>(Xt + (size_t)i * d, d, xi);
            load_pos<DM>(Xt + (size_t)i2 * d, d, xi2);
            double ra0 = 0, ra1 = 0, rb0 = 0, rb1 = 0;
            if (LK == kDirected) { ra0 = r0[i]; ra1 = r1[i]; rb0 = r0[i2]; rb1 = r1[i2]; }
            for (int base = 0; base < slots; base += 64) {
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    const int m = base + u * 32 + lane;
                    const bool first = m < len1;
                    const int arow = first ? i : i2;
                    int bcol = first ? (i + 1 + m) : (i2 + 1 + (m - len1));
                    const double v = vmask(m < slots);
                    bcol = bcol < n ? bcol : n - 1;
                    double xb[DM], xa[DM];
                    load_pos<DM>(Xt + (size_t)bcol * d, d, xb);
#pragma unroll
                    for (int k = 0; k < DM; k++) xa[k] = first ? xi[k] : xi2[k];
                    const double dist = fast_dist<DM>(xb, xa, d);
                    const size_t wo = ((size_t)t * n + arow) * W + (bcol >> 5);
                    if (LK == kUndirected) {
                        const double y = ymask(__ldg(p.net.rowbits + wo), bcol & 31);
                        a0 = fma(v, logit_term(y, b00 - dist), a0);
                        if (NV == 2) a1 = fma(v, logit_term(y, b10 - dist), a1);
                    } else {
                        const double y_ab = ymask(__ldg(p.net.rowbits + wo), bcol & 31);
                        const double y_ba = ymask(__ldg(p.net.colbits + wo), bcol & 31);
                        const double q0 = first ? ra0 : rb0, q1 = first ? ra1 : rb1;
                        const double s0 = r0[bcol], s1 = r1[bcol];
                        a0 = fma(v, logit_term(y_ab, eta_directed(b00, b01, dist, s0, q0)) +
                                        logit_term(y_ba, eta_directed(b00, b01, dist, q0, s0)), a0);
                        if (NV == 2)
                            a1 = fma(v, logit_term(y_ab, eta_directed(b10, b11, dist, s1, q1)) +
                                            logit_term(y_ba, eta_directed(b10, b11, dist, q1, s1)), a1);
                    }
                }
            }
        }
    } else {
        for (int i = tile + warp * p.tiles; i < n; i += nwarps * p.tiles) {
            double xi[DM];
            load_pos<DM>(Xt + (size_t)i * d, d, xi);
            // K6 directed_likelihoods_fast.pyx:208-270: out-edges and out-controls only
            const size_t r = (size_t)t * n + i;
            const int outdeg = p.net.deg[r * 2 + 1];
            const int32_t *oe = p.net.out_edges + r * p.net.max_out;
            const int32_t *co = p.net.ctrl_out +
                                ((size_t)(p.net.ctrl_per_chain ? c : 0) * T * n + r) * p.net.n_control;
            const double ri0 = r0[i], ri1 = (NV == 2) ? r1[i] : 0.0;
            double e0 = 0.0, e1 = 0.0, c0 = 0.0, c1 = 0.0;
            // gather records serve one variant, or two that share the radii (intercept MH: the proposal
            // and the current state differ in the intercepts only)
            const double *Gt = (D == 2 && p.gather && (NV == 1 || p.same_r))
                                   ? p.gather + ((size_t)c * T + t) * n * 4 : nullptr;
            auto terms = [&](int k, double &v0, double &v1) {
                double xk[DM], rk0, rk1;
                if (D == 2 && Gt) {
                    double pad;
                    ld256(Gt + (size_t)k * 4, xk[0], xk[DM > 1 ? 1 : 0], rk0, pad);
                    rk1 = rk0;
                } else {
                    load_pos<DM>(Xt + (size_t)k * d, d, xk);
                    rk0 = r0[k];
                    rk1 = (NV == 2) ? r1[k] : 0.0;
                }
                const double dist = fast_dist<DM>(xk, xi, d);
                v0 = eta_directed(b00, b01, dist, rk0, ri0);
                v1 = (NV == 2) ? eta_directed(b10, b11, dist, rk1, ri1) : 0.0;
            };
            int m = p.net.n_control;
            if (kFusedSoftplusForms && p.net.n_control + outdeg <= 128) {
                // One slot space per row: controls [0, nc), then the out-edges [nc, nc + outdeg) -- 110 entries
                // of cfg 5 fill 4 trips of 32 lanes instead of 2 (edges) + 4 (controls).  Every list index
                // of the row is in registers after ONE round trip, then clamped, masked gathers: the L2
                // latencies of a row overlap instead of chaining.  A lane evaluates the expensive part
                // L = log1p(e^-|eta|) once and wraps it as an edge term or as a control term.
                const int nc = p.net.n_control, slots = nc + outdeg;
                int kq[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int q = u * 32 + lane;
                    kq[u] = q < nc ? co[q] : (q < slots ? oe[q - nc] : i);
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const unsigned bal = __ballot_sync(kFull, u * 32 + lane < nc && kq[u] == -1);
                    if (bal && m == nc) m = u * 32 + __ffs(bal) - 1;
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int q = u * 32 + lane;
                    const bool is_ctrl = q < m, is_edge = q >= nc && q < slots;
                    double v0, v1;
                    terms((is_ctrl || is_edge) ? kq[u] : i, v0, v1);
                    {
                        const double a = fabs(v0), L = DLSM_L1PEN(a);
                        const double te = fma(0.5, v0, fma(-0.5, a, -L)); // logit_term(0.5, v0)
                        const double tc = fma(0.5, a, 0.5 * v0) + L;      // log1pexp(v0)
                        if (is_edge) e0 += te;
                        if (is_ctrl) c0 += tc;
                    }
                    if (NV == 2) {
                        const double a = fabs(v1), L = DLSM_L1PEN(a);
                        const double te = fma(0.5, v1, fma(-0.5, a, -L));
                        const double tc = fma(0.5, a, 0.5 * v1) + L;
                        if (is_edge) e1 += te;
                        if (is_ctrl) c1 += tc;
                    }
                }
            } else {
                for (int q = lane; q < outdeg; q += 32) {
                    double v0, v1;
                    terms(oe[q], v0, v1);
                    e0 += logit_term(0.5, v0);
                    if (NV == 2) e1 += logit_term(0.5, v1);
                }
                for (int base = 0; base < p.net.n_control; base += 32) {
                    const int q = base + lane;
                    const bool stop = (q < p.net.n_control) && (co[q] == -1);
                    const unsigned bal = __ballot_sync(kFull, stop);
                    if (bal) { m = base + __ffs(bal) - 1; break; }
                }
                for (int q = lane; q < m; q += 32) {
                    double v0, v1;
                    terms(co[q], v0, v1);
                    c0 += log1pexp(v0);
                    if (NV == 2) c1 += log1pexp(v1);
                }
            }
            warp_sum2(e0, c0, lane);
            if (NV == 2) warp_sum2(e1, c1, lane);
            if (lane == 0) {
                const double adj = (double)(n - outdeg - 1) / (double)m;
                a0 += e0 - adj * c0;
                a1 += e1 - adj * c1;
            }
        }
    }
    a0 = warp_sum(a0);
    a1 = warp_sum(a1);
    if (lane == 0) { red[0][warp] = a0; red[1][warp] = a1; }
    __syncthreads();
    if (threadIdx.x < 2) {
        double s = 0.0;
        for (int w = 0; w < nwarps; w++) s += red[threadIdx.x][w];
        p.partial[((size_t)c * gridDim.x + blockIdx.x) * 2 + threadIdx.x] = s;
    }
}

// ---------------------------------------------------------------------------------------------
// k_rows: the full-network log-likelihood of ONE parameter variant (an intercept / radii proposal,
// or the current state) together with its ROW SUMS rows'[j] = sum_i term(i, j), every unordered
// pair evaluated once (exact likelihoods; feeds the row-sum cache of the sweep kernels).
// Pairs are visited as 32 x 32 tiles (row block I, column block J >= I).  Within a tile lane l
// keeps node a = 32 I + l and its row accumulator; at step s it meets node b = 32 J + (l + s) mod 32,
// whose column accumulator ROTATES through the lanes (one 64-bit shuffle per step), so both nodes
// of a pair are credited from one evaluation and nothing is reduced through atomics: the column
// vector of every tile is written once (part[J][I][32]), the row vectors once per run
// (own[f][r][h][32]), and k_rows_commit adds them in a fixed order -> reproducible bit for bit.
// Work is dealt in equal pieces: row block I is glued to row block nb-1-I (together nb + 1 tiles)
// and the glued row is cut into R runs of <= ~17 tiles; a warp takes one run, a CTA eight
// consecutive runs of one chain (they may span several short slices, all staged in shared memory).
// grid = (ceil(T * ipc / 8), C), block = 256; dynamic smem = ns * n * (d [+ 1]) doubles
// ---------------------------------------------------------------------------------------------
struct RowsParams {
    NetView net;
    int C, nb, half, R, L, ipc, ns;  // row blocks, glued rows, runs per glued row, tiles per run,
                                     // items per slice, slices staged per CTA (max)
    const double *X;        // [C][T][n][d]
    const double *bvar;     // [C][2][2]: variant 0 is evaluated
    const double *rinv0;    // [C][n]
    double *partial;        // [C][gridDim.x][2] (slot 0)
    double *own;            // [C][T][half][R][2][32]
    double *part;           // [C][T][nb (nb-1) / 2][32]
};

template <int LK, int D>
#ifndef DLSM_ROWS_MINB
#define DLSM_ROWS_MINB 2
#endif
__global__ void __launch_bounds__(256, DLSM_ROWS_MINB) k_rows(const RowsParams p)
{
    constexpr int DM = (D == 0) ? kMaxD : D;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red[8];
    const int T = p.net.T, n = p.net.n, d = (D == 0) ? p.net.d : D, W = p.net.W, nb = p.nb;
    const int c = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int items = T * p.ipc;
    const int first = blockIdx.x * 8, last = (first + 7 < items ? first + 7 : items - 1);
    const int t0 = first / p.ipc, t1 = last / p.ipc;
    const int per = (n * (d + (LK == kDirected ? 1 : 0)) + 1) & ~1; // doubles staged per slice (16-byte aligned)
    double *Xs = reinterpret_cast<double *>(smem_raw);
    for (int sl = 0; sl <= t1 - t0; sl++) {
        const double *Xg = p.X + ((size_t)c * T + t0 + sl) * n * d;
        double *dst = Xs + (size_t)sl * per;
        for (int e = threadIdx.x; e < n * d; e += blockDim.x) dst[e] = Xg[e];
        if (LK == kDirected)
            for (int e = threadIdx.x; e < n; e += blockDim.x) dst[n * d + e] = p.rinv0[(size_t)c * n + e];
    }
    __syncthreads();
    const double b0 = p.bvar[(size_t)c * 4], b1 = p.bvar[(size_t)c * 4 + 1];
    double once = 0.0; // every pair of this warp's tiles counted once
    const int item = first + warp;
    if (item < items) {
        const int t = item / p.ipc, rem = item % p.ipc, f = rem / p.R, r = rem % p.R;
        const int I1 = f, I2 = nb - 1 - f;
        const int len1 = nb - I1;                         // tiles of row block I1: J = I1 .. nb-1
        const int tp = (I2 == I1) ? len1 : nb + 1;        // + those of row block I2: J = I2 .. nb-1
        const int u0 = r * p.L, u1 = (u0 + p.L < tp) ? u0 + p.L : tp;
        const double *Xt = Xs + (size_t)(t - t0) * per;
        const double *rv = Xt + (size_t)n * d;
        const size_t slice = (size_t)c * T + t;
        double *part_t = p.part + slice * ((size_t)nb * (nb - 1) / 2) * 32;
        double acc[2] = {0.0, 0.0};
        int cur = -1, a = 0, ac = 0;
        bool va = false;
        double xa[DM], ra = 0.0;
        const uint32_t *rowa = nullptr, *cola = nullptr;
#pragma unroll
        for (int k = 0; k < DM; k++) xa[k] = 0.0;
        auto pair_term = [&](int J, int bl, uint32_t wr, uint32_t wc) {
            const int b = 32 * J + bl;
            const int bc = b < n ? b : n - 1;
            double xb[DM];
            load_pos<DM>(Xt + (size_t)bc * d, d, xb);
            const double dist = fast_dist<DM>(xb, xa, d);
            if (LK == kUndirected) return logit_term(ymask(wr, bl), b0 - dist);
            const double rb = rv[bc];
            return logit_term(ymask(wr, bl), eta_directed(b0, b1, dist, rb, ra)) +  // a sends to b
                   logit_term(ymask(wc, bl), eta_directed(b0, b1, dist, ra, rb));   // b sends to a
        };
        for (int u = u0; u < u1; u++) {
            const int h = u >= len1 ? 1 : 0;
            const int I = h ? I2 : I1, J = h ? I2 + (u - len1) : I1 + u;
            if (I != cur) { // (re)load this lane's row node
                cur = I;
                a = 32 * I + lane;
                va = a < n;
                ac = va ? a : n - 1;
                load_pos<DM>(Xt + (size_t)ac * d, d, xa);
                if (LK == kDirected) ra = rv[ac];
                rowa = p.net.rowbits + ((size_t)t * n + ac) * W;
                if (LK == kDirected) cola = p.net.colbits + ((size_t)t * n + ac) * W;
            }
            const uint32_t wr = __ldg(rowa + J);
            const uint32_t wc = (LK == kDirected) ? __ldg(cola + J) : 0u;
            double R = 0.0, tile = 0.0;
            if (J == I) {
                // diagonal tile: pairs at cyclic distance 1..16 (distance 16 is met from both ends)
#pragma unroll 4
                for (int s = 1; s <= 16; s++) {
                    const int bl = (lane + s) & 31;
                    const double v = vmask(va && (32 * J + bl < n) && (s < 16 || lane < 16));
                    const double term = pair_term(J, bl, wr, wc);
                    R = __shfl_sync(kFull, R, (lane + 1) & 31);
                    R = fma(v, term, R);
                    tile = fma(v, term, tile);
                }
                R = __shfl_sync(kFull, R, (lane + 16) & 31); // column accumulators back to their nodes
                acc[h] += tile + R;
            } else {
#pragma unroll 4
                for (int s = 0; s < 32; s++) {
                    const int bl = (lane + s) & 31;
                    const double v = vmask(va && (32 * J + bl < n));
                    const double term = pair_term(J, bl, wr, wc);
                    if (s) R = __shfl_sync(kFull, R, (lane + 1) & 31);
                    R = fma(v, term, R);
                    tile = fma(v, term, tile);
                }
                R = __shfl_sync(kFull, R, (lane + 1) & 31);
                part_t[((size_t)J * (J - 1) / 2 + I) * 32 + lane] = R;
                acc[h] += tile;
            }
            once += tile;
        }
        double *own = p.own + (((slice * p.half + f) * p.R + r) * 2) * 32;
        own[lane] = acc[0];
        own[32 + lane] = acc[1];
    }
    once = warp_sum(once);
    if (lane == 0) red[warp] = once;
    __syncthreads();
    if (threadIdx.x == 0) {
        double sacc = 0.0;
        for (int w = 0; w < 8; w++) sacc += red[w];
        p.partial[((size_t)c * gridDim.x + blockIdx.x) * 2] = sacc;
    }
}

// rows[c][t][j] <- the row sums k_rows left in (own, part), for every chain whose flag is set
// (flag == nullptr: all chains); one warp per (slice, row block), fixed summation order
static __global__ void __launch_bounds__(256) k_rows_commit(const int32_t *flag, int T, int n, int nb, int half, int R,
                                                     const double *own, const double *part, double *rows)
{
    const int c = blockIdx.y, lane = threadIdx.x & 31;
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= T * nb || (flag && !flag[c])) return;
    const int t = w / nb, Q = w % nb;
    const size_t slice = (size_t)c * T + t;
    const int f = Q < nb - 1 - Q ? Q : nb - 1 - Q, h = (Q == f) ? 0 : 1;
    double s = 0.0;
    for (int r = 0; r < R; r++) s += own[(((slice * half + f) * R + r) * 2 + h) * 32 + lane];
    const double *pt = part + slice * ((size_t)nb * (nb - 1) / 2) * 32 + ((size_t)Q * (Q - 1) / 2) * 32;
    for (int I = 0; I < Q; I++) s += pt[(size_t)I * 32 + lane];
    const int i = 32 * Q + lane;
    if (i < n) rows[slice * n + i] = s;
}

// ---------------------------------------------------------------------------------------------
// scalar MH on the full-network likelihood: propose / finalize (one thread per chain)
// ---------------------------------------------------------------------------------------------
struct ScalarMH {
    int C, which, nblk, tune, tune_interval;
    double *intercept;       // [C][2]
    double *bvar;            // [C][2][2]
    double *prop;            // [C]
    const double *partial;   // [C][nblk][2]
    double prior_mean, prior_var;
    double *step;            // [C][2]
    int32_t *nacc, *nsteps, *until;
    const double *eps, *logu;   // replay [C][m] with stride m (nullptr -> Philox)
    int m;
    uint64_t seed;
    uint32_t sweep, chain_offset, site;
    int32_t *accepted;       // optional [C][m]
    double *ratio;           // optional [C][m]
    double *ll_out;          // optional [C][2] (variant sums) for probes
    double *ll_cur;          // optional [C]: log-likelihood of the current state, kept up to date
    int use_cur;             // 1: partial[..][1] was not computed, take ll_cur instead
    unsigned int *flags;
    int32_t *accflag;        // optional [C]: this step's decision (k_rows_commit reads it)
};

static __global__ void k_intercept_propose(const ScalarMH p)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= p.C) return;
    const double b0 = p.intercept[c * 2], b1 = p.intercept[c * 2 + 1];
    double eps;
    if (p.eps) {
        eps = p.eps[(size_t)c * p.m + p.which];
    } else {
        double z0, z1;
        box_muller(philox_u2(p.seed, p.site + p.which, p.sweep, (uint32_t)c + p.chain_offset,
                             kRngIntercept, 1), z0, z1);
        eps = z0;
    }
    const double x0 = p.which == 0 ? b0 : b1;
    const double x = __dadd_rn(x0, __dmul_rn(p.step[c * 2 + p.which], eps));
    p.prop[c] = x;
    double *bv = p.bvar + (size_t)c * 4;
    bv[0] = p.which == 0 ? x : b0; bv[1] = p.which == 1 ? x : b1; // variant 0: proposal
    bv[2] = b0; bv[3] = b1;                                       // variant 1: current
}

// the two variant sums of a chain's per-CTA partials, by one warp: lane-strided sums, then a fixed
// xor tree (deterministic; a chain can have thousands of partials: T * n / 64 with the case-control lists)
__device__ __forceinline__ void warp_sum_partials(const double *pp, int nblk, int lane, double &s0, double &s1)
{
    double a0 = 0.0, a1 = 0.0;
    for (int b = lane; b < nblk; b += 32) {
        const double2 v = *reinterpret_cast<const double2 *>(pp + (size_t)b * 2);
        a0 += v.x; a1 += v.y;
    }
    warp_sum2(a0, a1, lane);
    s0 = a0; s1 = a1;
}

// one warp per chain
static __global__ void k_intercept_finalize(const ScalarMH p)
{
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (c >= p.C) return;
    double s0, s1;
    warp_sum_partials(p.partial + (size_t)c * p.nblk * 2, p.nblk, lane, s0, s1);
    if (lane != 0) return;
    if (p.use_cur) s1 = p.ll_cur[c];
    const double x = p.prop[c], x0 = p.intercept[c * 2 + p.which];
    // sample_coefficients.py:39-41 / :84-85:  loglik -= (x - prior) ** 2 / (2 * variance)
    double df = __dsub_rn(x, p.prior_mean);
    const double lp_new = __dsub_rn(s0, __ddiv_rn(__dmul_rn(df, df), __dmul_rn(2.0, p.prior_var)));
    df = __dsub_rn(x0, p.prior_mean);
    const double lp_old = __dsub_rn(s1, __ddiv_rn(__dmul_rn(df, df), __dmul_rn(2.0, p.prior_var)));
    const double ratio = __dsub_rn(lp_new, lp_old);
    double logu;
    if (p.logu) logu = p.logu[(size_t)c * p.m + p.which];
    else logu = log(philox_u2(p.seed, p.site + p.which, p.sweep, (uint32_t)c + p.chain_offset,
                              kRngIntercept, 0).a);
    const int acc = (logu >= ratio) ? 0 : 1;
    if (acc) p.intercept[c * 2 + p.which] = x;
    if (p.accflag) p.accflag[c] = acc;
    if (p.ll_cur) p.ll_cur[c] = acc ? s0 : s1;
    const int o = c * 2 + p.which;
    double st = p.step[o];
    int na = p.nacc[o], ns = p.nsteps[o], un = p.until[o];
    metropolis_bookkeep(st, na, ns, un, p.tune, p.tune_interval, acc, false);
    p.step[o] = st; p.nacc[o] = na; p.nsteps[o] = ns; p.until[o] = un;
    if (p.accepted) p.accepted[(size_t)c * p.m + p.which] = acc;
    if (p.ratio) p.ratio[(size_t)c * p.m + p.which] = ratio;
    if (p.ll_out) { p.ll_out[c * 2] = s0; p.ll_out[c * 2 + 1] = s1; }
    if (!(ratio == ratio) || ratio - ratio != 0.0) atomicOr(p.flags, 1u);
}

// current-state probe: bvar = current intercepts for both variants
static __global__ void k_bvar_current(int C, const double *intercept, double *bvar)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    bvar[c * 4 + 0] = bvar[c * 4 + 2] = intercept[c * 2];
    bvar[c * 4 + 1] = bvar[c * 4 + 3] = intercept[c * 2 + 1];
}

// one warp per chain
static __global__ void k_sum_partials(int C, int nblk, const double *partial, double *out2, double *out_first = nullptr)
{
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (c >= C) return;
    double s0, s1;
    warp_sum_partials(partial + (size_t)c * nblk * 2, nblk, lane, s0, s1);
    if (lane != 0) return;
    out2[c * 2] = s0; out2[c * 2 + 1] = s1;
    if (out_first) out_first[c] = s0;
}

// ---------------------------------------------------------------------------------------------
// radii MH (sample_coefficients.py:91-121 + metropolis.py:57-82): one CTA per chain finalises:
// Hastings correction log Dir(r; s r') - log Dir(r'; s r), accept, copy
// ---------------------------------------------------------------------------------------------
struct RadiiMH {
    int C, n, nblk, tune, tune_interval;
    double *radii, *rinv;            // [C][n] current
    const double *prop, *prop_rinv;  // [C][n] proposal
    const double *partial;           // [C][nblk][2]  variant 0 = proposal, 1 = current
    double *step;                    // [C]
    int32_t *nacc, *nsteps, *until;
    const double *logu;              // replay [C] or nullptr
    uint64_t seed;
    uint32_t sweep, chain_offset, site;
    int32_t *accepted;
    double *ratio;
    unsigned int *flags;
    double *ll_cur;                  // optional [C], see ScalarMH
    int use_cur;
    int32_t *accflag;                // optional [C], see ScalarMH
};

__device__ __forceinline__ double block_sum(double v, double *sh)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < nw; w++) s += sh[w];
    return s;
}

// Dirichlet terms of the Hastings correction, kRadiiChunk nodes per CTA (grid = chunks x chains):
// terms[c][chunk][6] = {sum alpha_f, sum lgamma(alpha_f), sum (alpha_f - 1) log r, same three backward}.
// With one CTA per chain (round 1) the 2 n lgamma + 2 n log evaluations of a 50 000-node chain took
// 0.9 ms on 8 SMs.
constexpr int kRadiiChunk = 2048;
static __global__ void __launch_bounds__(256) k_radii_terms(const RadiiMH p, double *terms)
{
    __shared__ double sh[8];
    const int c = blockIdx.y, n = p.n;
    const double *r = p.radii + (size_t)c * n;
    const double *q = p.prop + (size_t)c * n;
    const double s = p.step[c];
    // scipy.stats.dirichlet.logpdf(x, alpha) = -[sum lgamma(alpha) - lgamma(sum alpha)]
    //                                          + sum (alpha - 1) log x
    double sa_f = 0, sl_f = 0, sx_f = 0; // forward:  x = r (current), alpha = s * q
    double sa_b = 0, sl_b = 0, sx_b = 0; // backward: x = q (proposal), alpha = s * r
    const int lo = blockIdx.x * kRadiiChunk, hi = lo + kRadiiChunk < n ? lo + kRadiiChunk : n;
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const double af = s * q[i], ab = s * r[i];
        sa_f += af; sl_f += lgamma(af); sx_f += (af - 1.0) * log(r[i]);
        sa_b += ab; sl_b += lgamma(ab); sx_b += (ab - 1.0) * log(q[i]);
    }
    sa_f = block_sum(sa_f, sh); sl_f = block_sum(sl_f, sh); sx_f = block_sum(sx_f, sh);
    sa_b = block_sum(sa_b, sh); sl_b = block_sum(sl_b, sh); sx_b = block_sum(sx_b, sh);
    if (threadIdx.x == 0) {
        double *o = terms + ((size_t)c * gridDim.x + blockIdx.x) * 6;
        o[0] = sa_f; o[1] = sl_f; o[2] = sx_f; o[3] = sa_b; o[4] = sl_b; o[5] = sx_b;
    }
}

// one CTA per chain: adds the chunks' terms and the likelihood partials, decides, copies
static __global__ void __launch_bounds__(256) k_radii_finalize(const RadiiMH p, const double *terms, int chunks)
{
    __shared__ double sh[8];
    __shared__ int s_acc;
    const int c = blockIdx.x, n = p.n;
    double *r = p.radii + (size_t)c * n;
    const double *q = p.prop + (size_t)c * n;
    const double s = p.step[c];
    double tm[6] = {0, 0, 0, 0, 0, 0};
    for (int b = threadIdx.x; b < chunks; b += blockDim.x)
#pragma unroll
        for (int k = 0; k < 6; k++) tm[k] += terms[((size_t)c * chunks + b) * 6 + k];
#pragma unroll
    for (int k = 0; k < 6; k++) tm[k] = block_sum(tm[k], sh);
    double s0 = 0.0, s1 = 0.0;
    const double *pp = p.partial + (size_t)c * p.nblk * 2;
    for (int b = threadIdx.x; b < p.nblk; b += blockDim.x) { s0 += pp[b * 2]; s1 += pp[b * 2 + 1]; }
    s0 = block_sum(s0, sh); s1 = block_sum(s1, sh);
    if (threadIdx.x == 0) {
        if (p.use_cur) s1 = p.ll_cur[c];
        const double lf = -(tm[1] - lgamma(tm[0])) + tm[2];
        const double lb = -(tm[4] - lgamma(tm[3])) + tm[5];
        const double ratio = (s0 - s1) + (lf - lb);
        double logu;
        if (p.logu) logu = p.logu[c];
        else logu = log(philox_u2(p.seed, p.site, p.sweep, (uint32_t)c + p.chain_offset, kRngRadii, 0).a);
        const int acc = (logu >= ratio) ? 0 : 1;
        s_acc = acc;
        if (p.accflag) p.accflag[c] = acc;
        if (p.ll_cur) p.ll_cur[c] = acc ? s0 : s1;
        double st = s;
        int na = p.nacc[c], ns = p.nsteps[c], un = p.until[c];
        metropolis_bookkeep(st, na, ns, un, p.tune, p.tune_interval, acc, true);
        p.step[c] = st; p.nacc[c] = na; p.nsteps[c] = ns; p.until[c] = un;
        if (p.accepted) p.accepted[c] = acc;
        if (p.ratio) p.ratio[c] = ratio;
        if (!(ratio == ratio) || ratio - ratio != 0.0) atomicOr(p.flags, 1u);
    }
    __syncthreads();
    if (s_acc) {
        double *ri = p.rinv + (size_t)c * n;
        const double *qi = p.prop_rinv + (size_t)c * n;
        for (int i = threadIdx.x; i < n; i += blockDim.x) { r[i] = q[i]; ri[i] = qi[i]; }
    }
}

// native Dirichlet proposal r' ~ Dir(s * r) via Marsaglia-Tsang gamma variates (one thread per
// node, Philox stream per (chain, node)); zero guard of metropolis.py:65-67
__device__ inline double gamma_mt(double shape, uint64_t seed, uint32_t site, uint32_t sweep,
                                  uint32_t chain)
{
    uint32_t blk = 1;
    double boost = 1.0;
    if (shape < 1.0) {
        const U2 u = philox_u2(seed, site, sweep, chain, kRngRadii, blk++);
        boost = pow(u.a, 1.0 / shape);
        shape += 1.0;
    }
    const double dd = shape - 1.0 / 3.0, cc = 1.0 / sqrt(9.0 * dd);
    for (int it = 0; it < 64; it++) {
        double z0, z1;
        box_muller(philox_u2(seed, site, sweep, chain, kRngRadii, blk++), z0, z1);
        const U2 u = philox_u2(seed, site, sweep, chain, kRngRadii, blk++);
        double v = 1.0 + cc * z0;
        if (v <= 0.0) continue;
        v = v * v * v;
        if (log(u.a) < 0.5 * z0 * z0 + dd - dd * v + dd * log(v)) return boost * dd * v;
    }
    return boost * dd;
}

// stage 1 (grid = chunks x chains): the gamma variates of kRadiiChunk nodes and their total
static __global__ void __launch_bounds__(256) k_radii_gammas(int n, const double *radii, const double *step,
                                                             double *prop, double *chunk_tot, uint64_t seed,
                                                             uint32_t sweep, uint32_t chain_offset, uint32_t site0)
{
    __shared__ double sh[8];
    const int c = blockIdx.y;
    const double *r = radii + (size_t)c * n;
    double *q = prop + (size_t)c * n;
    const double sc = step[c];
    const int lo = blockIdx.x * kRadiiChunk, hi = lo + kRadiiChunk < n ? lo + kRadiiChunk : n;
    double tot = 0.0;
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const double g = gamma_mt(sc * r[i], seed, site0 + 1 + i, sweep, (uint32_t)c + chain_offset);
        q[i] = g;
        tot += g;
    }
    tot = block_sum(tot, sh);
    if (threadIdx.x == 0) chunk_tot[(size_t)c * gridDim.x + blockIdx.x] = tot;
}

// stage 2 (one CTA per chain): normalise, zero guard, reciprocals
static __global__ void __launch_bounds__(1024) k_radii_propose(int n, int chunks, const double *chunk_tot,
                                                               double *prop, double *prop_rinv)
{
    __shared__ double sh[32];
    __shared__ int any_zero;
    const int c = blockIdx.x;
    double *q = prop + (size_t)c * n;
    if (threadIdx.x == 0) any_zero = 0;
    double tot = 0.0;
    for (int b = threadIdx.x; b < chunks; b += blockDim.x) tot += chunk_tot[(size_t)c * chunks + b];
    tot = block_sum(tot, sh);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        q[i] = q[i] / tot;
        // metropolis.py:65 tests `== 0`; a subnormal component is treated the same way: its
        // reciprocal is inf and the likelihood NaN (in the reference as well)
        if (q[i] < 2.2250738585072014e-308) any_zero = 1;
    }
    __syncthreads();
    if (any_zero) {
        double t2 = 0.0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) { q[i] += 1e-5; t2 += q[i]; }
        t2 = block_sum(t2, sh);
        for (int i = threadIdx.x; i < n; i += blockDim.x) q[i] /= t2;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) prop_rinv[(size_t)c * n + i] = 1.0 / q[i];
}

// ---------------------------------------------------------------------------------------------
// k_ffbs: HDP-HMM label block sampler, one warp per (chain, node)
// (sample_labels.py:134-190 + gaussian_likelihood_fast.pyx:17-54)
// dynamic smem per warp: (T*K + 3*K) doubles
// ---------------------------------------------------------------------------------------------
struct LabelParams {
    int C, T, n, d, K;
    const double *X;        // [C][T][n][d]
    const double *mu;       // [C][K][d]
    const double *sigma;    // [C][K]
    const double *lambda;   // [C]
    const double *w;        // [C][T][K][K]
    const double *U;        // replay [C][n][T] or nullptr
    uint64_t seed;
    uint32_t sweep, chain_offset;
    int32_t *z;             // [C][T][n]  out
    double *ncount;         // [C][T][K][K] out (zeroed by the caller)
    int32_t *nk;            // [C][T][K]  out (zeroed by the caller)
    double *lik_out;        // optional probe [C][n][T][K]
    double *gstage;         // optional global stage for k_ffbs_t: (T*K + 2K) * TPB doubles per CTA
    int sample;             // 0: emission densities only
};

// numpy pairwise_sum restricted to n <= 128 (K <= 128)
__device__ inline double np_sum_k(const double *a, int n)
{
    if (n < 8) {
        double r = -0.0;
        for (int i = 0; i < n; i++) r = __dadd_rn(r, a[i]);
        return r;
    }
    double r[8];
    for (int k = 0; k < 8; k++) r[k] = a[k];
    int i;
    for (i = 8; i < n - (n % 8); i += 8)
        for (int k = 0; k < 8; k++) r[k] = __dadd_rn(r[k], a[i + k]);
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; i++) res = __dadd_rn(res, a[i]);
    return res;
}

static __global__ void __launch_bounds__(128) k_ffbs(const LabelParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int T = p.T, n = p.n, d = p.d, K = p.K;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t gw = (size_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (gw >= (size_t)p.C * n) return;
    const int c = (int)(gw / n), i = (int)(gw % n);
    double *pm = reinterpret_cast<double *>(smem_raw) + (size_t)warp * (T * K + 3 * K);
    double *bw0 = pm + T * K, *bw1 = bw0 + K, *cdf = bw1 + K;
    const double *X = p.X + (size_t)c * T * n * d;
    const double *mu = p.mu + (size_t)c * K * d;
    const double *sg = p.sigma + (size_t)c * K;
    const double *w = p.w + (size_t)c * T * K * K;
    const double lm = p.lambda[c], oml = __dsub_rn(1.0, lm);

    // K7: emission densities exp(loglik), un-normalised (normalize=False, sample_labels.py:159)
    for (int idx = lane; idx < T * K; idx += 32) {
        const int t = idx / K, k = idx % K;
        const double *x = X + ((size_t)t * n + i) * d;
        const double *xp = X + ((size_t)(t > 0 ? t - 1 : 0) * n + i) * d;
        double sum_sq = 0.0;
        for (int q = 0; q < d; q++) {
            const double mean = (t == 0) ? mu[k * d + q]
                                         : __dadd_rn(__dmul_rn(lm, mu[k * d + q]),
                                                     __dmul_rn(oml, xp[q]));
            const double df = __dsub_rn(x[q], mean);
            sum_sq = __dadd_rn(sum_sq, __dmul_rn(df, df));
        }
        const double var = sg[k];
        sum_sq = __dmul_rn(sum_sq, __dmul_rn(0.5, __ddiv_rn(1.0, var)));
        const double ll = __dsub_rn(
            __dmul_rn(__dmul_rn(-0.5, (double)d), log(__dmul_rn(6.283185307179586, var))), sum_sq);
        const double L = exp(ll);
        pm[idx] = L;
        if (p.lik_out) p.lik_out[(((size_t)c * n + i) * T + t) * K + k] = L;
    }
    __syncwarp();
    if (!p.sample) return;

    // backward messages (:164-169); bwds_msg[T-1] is all ones (allocated once, never written)
    double *bcur = bw0, *bprev = bw1;
    for (int k = lane; k < K; k += 32) bcur[k] = 1.0;
    __syncwarp();
    for (int t = T - 1; t > 0; t--) {
        for (int k = lane; k < K; k += 32) pm[t * K + k] = __dmul_rn(pm[t * K + k], bcur[k]);
        __syncwarp();
        for (int j = lane; j < K; j += 32) {
            const double *wr = w + ((size_t)t * K + j) * K;
            double s = 0.0;
            for (int k = 0; k < K; k++) s = __dadd_rn(s, __dmul_rn(wr[k], pm[t * K + k]));
            bprev[j] = s;
        }
        __syncwarp();
        const double tot = np_sum_k(bprev, K);
        __syncwarp();
        for (int j = lane; j < K; j += 32) bprev[j] = __ddiv_rn(bprev[j], tot);
        __syncwarp();
        double *tmp = bcur; bcur = bprev; bprev = tmp;
    }
    for (int k = lane; k < K; k += 32) pm[k] = __dmul_rn(pm[k], bcur[k]);
    __syncwarp();

    // forward sampling (:173-188); cumsum and the comparison are serial, lane 0
    if (lane == 0) {
        int zp = 0;
        for (int t = 0; t < T; t++) {
            const double *wr = (t == 0) ? w : w + ((size_t)t * K + zp) * K;
            double cs = 0.0;
            for (int k = 0; k < K; k++) {
                const double pr = __dmul_rn(wr[k], pm[t * K + k]);
                cs = (k == 0) ? pr : __dadd_rn(cs, pr);
                cdf[k] = cs;
            }
            double U;
            if (p.U) U = p.U[((size_t)c * n + i) * T + t];
            else U = philox_u2(p.seed, (uint32_t)(i * T + t), p.sweep, (uint32_t)c + p.chain_offset,
                               kRngLabels, 0).a;
            const double u = __dmul_rn(cdf[K - 1], U);
            int zz = 0;
            for (int k = 0; k < K; k++) zz += (u > cdf[k]) ? 1 : 0;
            if (zz >= K) zz = K - 1;
            p.z[((size_t)c * T + t) * n + i] = zz;
            if (t == 0) atomicAdd(&p.ncount[(size_t)c * T * K * K + zz], 1.0);
            else atomicAdd(&p.ncount[(((size_t)c * T + t) * K + zp) * K + zz], 1.0);
            atomicAdd(&p.nk[((size_t)c * T + t) * K + zz], 1);
            zp = zz;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// k_ffbs_r: the label block sampler for K <= 16, one THREAD per (chain, node), with everything
// that is indexed by the component held in REGISTERS (arrays of KC = 4/8/12/16 >= K entries, all
// loops over components fully unrolled): the emission row, the backward message and the matrix-
// vector product of a step never touch memory, the transition weights of all T steps sit in
// shared memory (broadcast reads), and the only per-thread stage is pm[t][k] = L[t][k] bwd[t][k]
// (T*K doubles, written once and read once per pass).  That stage lives in a global scratch
// indexed by CTA *slot*: the grid is persistent (a few CTAs per SM, each looping over (chain,
// node-tile) items), so the scratch stays a few tens of MB and L2-resident.
// Arithmetic and its order are those of k_ffbs_t / the oracle (separate multiplies and adds,
// serial-in-k rows, numpy's pairwise total): labels are bit-identical.
// grid = min(items, 148 * resident), block = 64;
// dynamic smem = (T*K*K + K*d + 2K) doubles (+ T*K*64 doubles when the stage is in shared memory)
// ---------------------------------------------------------------------------------------------
template <int KC>
__device__ __forceinline__ double np_sum_regs(const double (&a)[KC], int K)
{
    if (K < 8) {
        double r = -0.0;
#pragma unroll
        for (int k = 0; k < KC; k++)
            if (k < K && k < 8) r = __dadd_rn(r, a[k]);
        return r;
    }
    double res = 0.0;
    if (KC >= 8) {
        double r8[8];
#pragma unroll
        for (int q = 0; q < 8; q++) r8[q] = a[q < KC ? q : 0];
        if (KC >= 16 && K >= 16) {
#pragma unroll
            for (int q = 0; q < 8; q++) r8[q] = __dadd_rn(r8[q], a[(8 + q) < KC ? 8 + q : 0]);
        }
        res = __dadd_rn(__dadd_rn(__dadd_rn(r8[0], r8[1]), __dadd_rn(r8[2], r8[3])),
                        __dadd_rn(__dadd_rn(r8[4], r8[5]), __dadd_rn(r8[6], r8[7])));
        const int done = K - (K % 8);
#pragma unroll
        for (int k = 8; k < KC; k++)
            if (k >= done && k < K) res = __dadd_rn(res, a[k]);
    }
    return res;
}

template <int KC, int D>
__global__ void __launch_bounds__(64) k_ffbs_r(const LabelParams p, int tiles, int items)
{
    constexpr int TPB = 64;
    constexpr int DM = (D == 0) ? kMaxD : D;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int T = p.T, n = p.n, K = p.K, d = (D == 0) ? p.d : D, KK = K * K;
    const int tid = threadIdx.x, lane = tid & 31;
    double *s_w = reinterpret_cast<double *>(smem_raw);   // [T][K][K]
    double *s_mu = s_w + (size_t)T * KK;                  // [K][d]
    double *s_ln = s_mu + (size_t)K * d;                  // [K]  -(d/2) log(2 pi var)
    double *s_hv = s_ln + K;                              // [K]  0.5 * (1 / var)
    double *pm = p.gstage ? p.gstage + (size_t)blockIdx.x * T * K * TPB : s_hv + K; // [T*K][TPB]
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int c = item / tiles, i = (item % tiles) * TPB + tid;
        const bool valid = i < n;
        const int ic = valid ? i : n - 1;
        const double *X = p.X + (size_t)c * T * n * d;
        const double *w = p.w + (size_t)c * T * KK;
        const double lm = p.lambda[c], oml = __dsub_rn(1.0, lm);
        __syncthreads(); // the previous item is done with the shared tables
        for (int e = tid; e < T * KK; e += TPB) s_w[e] = w[e];
        for (int k = tid; k < K; k += TPB) {
            const double var = p.sigma[(size_t)c * K + k];
            s_ln[k] = __dmul_rn(__dmul_rn(-0.5, (double)d), log(__dmul_rn(6.283185307179586, var)));
            s_hv[k] = __dmul_rn(0.5, __ddiv_rn(1.0, var));
        }
        for (int e = tid; e < K * d; e += TPB) s_mu[e] = p.mu[(size_t)c * K * d + e];
        __syncthreads();

        // K7 emission densities (gaussian_likelihood_fast.pyx:17-54), normalize=False
        double xprev[DM], xt[DM];
#pragma unroll
        for (int q = 0; q < DM; q++) xprev[q] = 0.0;
        for (int t = 0; t < T; t++) {
            const double *xg = X + ((size_t)t * n + ic) * d;
#pragma unroll
            for (int q = 0; q < DM; q++) xt[q] = (q < d) ? xg[q] : 0.0;
#pragma unroll
            for (int k = 0; k < KC; k++) {
                if (k < K) {
                    double sum_sq = 0.0;
#pragma unroll
                    for (int q = 0; q < DM; q++)
                        if (q < d) {
                            const double m = s_mu[k * d + q];
                            const double mean = (t == 0) ? m : __dadd_rn(__dmul_rn(lm, m), __dmul_rn(oml, xprev[q]));
                            const double df = __dsub_rn(xt[q], mean);
                            sum_sq = __dadd_rn(sum_sq, __dmul_rn(df, df));
                        }
                    const double L = fast_exp(__dsub_rn(s_ln[k], __dmul_rn(sum_sq, s_hv[k])));
                    pm[(size_t)(t * K + k) * TPB + tid] = L;
                    if (p.lik_out && valid) p.lik_out[(((size_t)c * n + i) * T + t) * K + k] = L;
                }
            }
#pragma unroll
            for (int q = 0; q < DM; q++) xprev[q] = xt[q];
        }
        if (!p.sample) continue;

        // backward messages (sample_labels.py:164-169); bwds_msg[T-1] stays all ones
        double b[KC], v[KC];
#pragma unroll
        for (int k = 0; k < KC; k++) b[k] = 1.0;
        for (int t = T - 1; t > 0; t--) {
            const double *wt = s_w + (size_t)t * KK;
#pragma unroll
            for (int k = 0; k < KC; k++) {
                v[k] = 0.0;
                if (k < K) {
                    const size_t o = (size_t)(t * K + k) * TPB + tid;
                    v[k] = __dmul_rn(pm[o], b[k]);
                    pm[o] = v[k];
                }
            }
            // bwd[t-1][j] = sum_k w[t][j][k] pm[t][k], serial in k per row j (the oracle's order);
            // the rows are independent chains the scheduler interleaves
#pragma unroll
            for (int j = 0; j < KC; j++) {
                double sacc = 0.0;
                if (j < K) {
#pragma unroll
                    for (int k = 0; k < KC; k++)
                        if (k < K) sacc = __dadd_rn(sacc, __dmul_rn(wt[j * K + k], v[k]));
                }
                b[j] = sacc;
            }
            const double itot = __ddiv_rn(1.0, np_sum_regs<KC>(b, K)); // one division per step
#pragma unroll
            for (int j = 0; j < KC; j++) b[j] = __dmul_rn(b[j], itot);
        }
#pragma unroll
        for (int k = 0; k < KC; k++) v[k] = (k < K) ? __dmul_rn(pm[(size_t)k * TPB + tid], b[k]) : 0.0;

        // forward sampling (:173-188): cumsum, u = cdf[-1] * U, z = #{k : u > cdf[k]}
        int zp = 0;
        for (int t = 0; t < T; t++) {
            if (t > 0) {
#pragma unroll
                for (int k = 0; k < KC; k++) v[k] = (k < K) ? pm[(size_t)(t * K + k) * TPB + tid] : 0.0;
            }
            const double *wr = (t == 0) ? s_w : s_w + (size_t)t * KK + zp * K;
            double cdf[KC];
            double cs = 0.0;
#pragma unroll
            for (int k = 0; k < KC; k++) {
                if (k < K) {
                    const double pr = __dmul_rn(wr[k], v[k]);
                    cs = (k == 0) ? pr : __dadd_rn(cs, pr);
                }
                cdf[k] = cs;
            }
            double U;
            if (p.U) U = p.U[((size_t)c * n + ic) * T + t];
            else U = philox_u2(p.seed, (uint32_t)(ic * T + t), p.sweep, (uint32_t)c + p.chain_offset,
                               kRngLabels, 0).a;
            const double u = __dmul_rn(cs, U);
            int zz = 0;
#pragma unroll
            for (int k = 0; k < KC; k++) zz += (k < K && u > cdf[k]) ? 1 : 0;
            if (zz >= K) zz = K - 1;
            if (valid) p.z[((size_t)c * T + t) * n + i] = zz;
            // counts: one atomic per distinct (from, to) pair / label in the warp
            const int from = (t == 0) ? 0 : zp;
            const int key = valid ? from * K + zz : -1;
            const unsigned same = __match_any_sync(kFull, key);
            if (valid && lane == __ffs(same) - 1)
                atomicAdd(&p.ncount[((size_t)c * T + t) * KK + key], (double)__popc(same));
            const unsigned same_z = __match_any_sync(kFull, valid ? zz : -1);
            if (valid && lane == __ffs(same_z) - 1)
                atomicAdd(&p.nk[((size_t)c * T + t) * K + zz], __popc(same_z));
            zp = zz;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// k_ffbs_t: HDP-HMM label block sampler, one THREAD per (chain, node) -- the production mapping.
// The per-node recursion is strictly sequential in t and only K wide, so a warp per node leaves
// two thirds of the lanes idle and pays a shuffle/sync per step (k_ffbs above, kept as the
// fallback for very large T*K).  Here every lane runs its own node; the T*K partial marginals and
// the two K-vectors of backward messages live in shared memory, laid out [entry][thread] so that a
// warp's accesses are conflict-free.  grid = (ceil(n/TPB), C), block = TPB (32 or 64).
// dynamic smem = (T*K + 2K) * TPB doubles + K*(d+2) doubles + K*K doubles (w[t] stage).
// ---------------------------------------------------------------------------------------------
template <int TPB>
__global__ void __launch_bounds__(TPB) k_ffbs_t(const LabelParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int T = p.T, n = p.n, d = p.d, K = p.K;
    const int c = blockIdx.y, tid = threadIdx.x;
    const int i = blockIdx.x * TPB + tid;
    const bool valid = i < n;
    const int ic = valid ? i : n - 1;
    // The per-thread stage (T*K partial marginals + two K-vectors) lives either in shared memory
    // (8 warps per SM at cfg 2) or, when p.gstage is set, in an L2-resident global scratch laid out
    // [entry][thread] per CTA (coalesced), which lifts the occupancy cap.
    double *pm, *s_mu;
    if (p.gstage) {
        pm = p.gstage + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * ((size_t)T * K + 2 * K) * TPB;
        s_mu = reinterpret_cast<double *>(smem_raw);
    } else {
        pm = reinterpret_cast<double *>(smem_raw);      // [T*K][TPB]
        s_mu = pm + ((size_t)T * K + 2 * K) * TPB;      // [K][d]
    }
    double *bwA = pm + (size_t)T * K * TPB;             // [K][TPB]
    double *bwB = bwA + (size_t)K * TPB;                // [K][TPB]
    double *s_ln = s_mu + (size_t)K * d;                // [K]  -(d/2) log(2 pi var)
    double *s_hv = s_ln + K;                            // [K]  0.5 * (1 / var)
    double *s_w = s_hv + K;                             // [K][K] transition weights of one step
    const double *X = p.X + (size_t)c * T * n * d;
    const double *w = p.w + (size_t)c * T * K * K;
    const double lm = p.lambda[c], oml = __dsub_rn(1.0, lm);
    for (int k = tid; k < K; k += TPB) {
        const double var = p.sigma[(size_t)c * K + k];
        s_ln[k] = __dmul_rn(__dmul_rn(-0.5, (double)d), log(__dmul_rn(6.283185307179586, var)));
        s_hv[k] = __dmul_rn(0.5, __ddiv_rn(1.0, var));
    }
    for (int e = tid; e < K * d; e += TPB) s_mu[e] = p.mu[(size_t)c * K * d + e];
    __syncthreads();

    // K7 emission densities (gaussian_likelihood_fast.pyx:17-54), normalize=False
    double xprev[kMaxD], xt[kMaxD];
#pragma unroll
    for (int q = 0; q < kMaxD; q++) xprev[q] = 0.0;
    for (int t = 0; t < T; t++) {
        const double *xg = X + ((size_t)t * n + ic) * d;
#pragma unroll
        for (int q = 0; q < kMaxD; q++) xt[q] = (q < d) ? xg[q] : 0.0;
        for (int k = 0; k < K; k++) {
            double sum_sq = 0.0;
#pragma unroll
            for (int q = 0; q < kMaxD; q++)
                if (q < d) {
                    const double m = s_mu[k * d + q];
                    const double mean = (t == 0) ? m : __dadd_rn(__dmul_rn(lm, m), __dmul_rn(oml, xprev[q]));
                    const double df = __dsub_rn(xt[q], mean);
                    sum_sq = __dadd_rn(sum_sq, __dmul_rn(df, df));
                }
            const double L = fast_exp(__dsub_rn(s_ln[k], __dmul_rn(sum_sq, s_hv[k])));
            pm[(size_t)(t * K + k) * TPB + tid] = L;
            if (p.lik_out && valid) p.lik_out[(((size_t)c * n + i) * T + t) * K + k] = L;
        }
#pragma unroll
        for (int q = 0; q < kMaxD; q++) xprev[q] = xt[q];
    }
    if (!p.sample) return;

    // backward messages (sample_labels.py:164-169); bwds_msg[T-1] stays all ones
    double *bcur = bwA, *bprev = bwB;
    for (int k = 0; k < K; k++) bcur[k * TPB + tid] = 1.0;
    for (int t = T - 1; t > 0; t--) {
        __syncthreads();
        for (int e = tid; e < K * K; e += TPB) s_w[e] = w[(size_t)t * K * K + e]; // w[t] -> smem
        for (int k = 0; k < K; k++) {
            const size_t o = (size_t)(t * K + k) * TPB + tid;
            pm[o] = __dmul_rn(pm[o], bcur[k * TPB + tid]);
        }
        __syncthreads();
        // bwd[t-1][j] = sum_k w[t][j][k] pm[t][k]: serial in k per j (the oracle's order), four
        // rows j at a time so four dependent add chains are in flight per thread
        int j = 0;
        for (; j + 4 <= K; j += 4) {
            const double *wr = s_w + j * K;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            for (int k = 0; k < K; k++) {
                const double v = pm[(size_t)(t * K + k) * TPB + tid];
                s0 = __dadd_rn(s0, __dmul_rn(wr[k], v));
                s1 = __dadd_rn(s1, __dmul_rn(wr[K + k], v));
                s2 = __dadd_rn(s2, __dmul_rn(wr[2 * K + k], v));
                s3 = __dadd_rn(s3, __dmul_rn(wr[3 * K + k], v));
            }
            bprev[j * TPB + tid] = s0; bprev[(j + 1) * TPB + tid] = s1;
            bprev[(j + 2) * TPB + tid] = s2; bprev[(j + 3) * TPB + tid] = s3;
        }
        for (; j < K; j++) {
            const double *wr = s_w + j * K;
            double sacc = 0.0;
            for (int k = 0; k < K; k++)
                sacc = __dadd_rn(sacc, __dmul_rn(wr[k], pm[(size_t)(t * K + k) * TPB + tid]));
            bprev[j * TPB + tid] = sacc;
        }
        // np.sum over K entries: numpy's pairwise order (serial below 8, 8 accumulators above)
        double tot;
        if (K < 8) {
            tot = -0.0;
            for (int k = 0; k < K; k++) tot = __dadd_rn(tot, bprev[k * TPB + tid]);
        } else {
            double r8[8];
#pragma unroll
            for (int q = 0; q < 8; q++) r8[q] = bprev[q * TPB + tid];
            int k = 8;
            for (; k < K - (K % 8); k += 8)
#pragma unroll
                for (int q = 0; q < 8; q++) r8[q] = __dadd_rn(r8[q], bprev[(k + q) * TPB + tid]);
            tot = __dadd_rn(__dadd_rn(__dadd_rn(r8[0], r8[1]), __dadd_rn(r8[2], r8[3])),
                            __dadd_rn(__dadd_rn(r8[4], r8[5]), __dadd_rn(r8[6], r8[7])));
            for (; k < K; k++) tot = __dadd_rn(tot, bprev[k * TPB + tid]);
        }
        const double itot = __ddiv_rn(1.0, tot); // one division per step (1 ulp from x / tot)
        for (int j = 0; j < K; j++) bprev[j * TPB + tid] = __dmul_rn(bprev[j * TPB + tid], itot);
        double *tmp = bcur; bcur = bprev; bprev = tmp;
    }
    for (int k = 0; k < K; k++) pm[(size_t)k * TPB + tid] = __dmul_rn(pm[(size_t)k * TPB + tid], bcur[k * TPB + tid]);

    // forward sampling (:173-188): cumsum, u = cdf[-1] * U, z = #{k : u > cdf[k]}
    double *cdf = bprev; // reuse
    int zp = 0;
    for (int t = 0; t < T; t++) {
        __syncthreads();
        for (int e = tid; e < K * K; e += TPB) s_w[e] = w[(size_t)t * K * K + e];
        __syncthreads();
        const double *wr = (t == 0) ? s_w : s_w + zp * K;
        double cs = 0.0;
        for (int k = 0; k < K; k++) {
            const double pr = __dmul_rn(wr[k], pm[(size_t)(t * K + k) * TPB + tid]);
            cs = (k == 0) ? pr : __dadd_rn(cs, pr);
            cdf[k * TPB + tid] = cs;
        }
        double U;
        if (p.U) U = p.U[((size_t)c * n + ic) * T + t];
        else U = philox_u2(p.seed, (uint32_t)(ic * T + t), p.sweep, (uint32_t)c + p.chain_offset,
                           kRngLabels, 0).a;
        const double u = __dmul_rn(cs, U);
        int zz = 0;
        for (int k = 0; k < K; k++) zz += (u > cdf[k * TPB + tid]) ? 1 : 0;
        if (zz >= K) zz = K - 1;
        if (valid) {
            p.z[((size_t)c * T + t) * n + i] = zz;
            if (t == 0) atomicAdd(&p.ncount[(size_t)c * T * K * K + zz], 1.0);
            else atomicAdd(&p.ncount[(((size_t)c * T + t) * K + zp) * K + zz], 1.0);
            atomicAdd(&p.nk[((size_t)c * T + t) * K + zz], 1);
        }
        zp = zz;
    }
}

} // namespace dlsm
