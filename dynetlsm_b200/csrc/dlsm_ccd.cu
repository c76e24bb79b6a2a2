// dlsm_ccd.cu -- k_sweep_ccd: the case-control sweep of large sparse networks as a DATAFLOW over nodes
// (cfg 5: n = 50 000, T = 10, ~220 list entries per node).
//
// The reference updates the nodes of a slice one after the other (sample_latent_positions.py:92-146
// around directed_likelihoods_fast.pyx:83-182).  Node j's conditional reads only the positions of
// the nodes in ITS lists: for i < j it must see i's post-update position, for i > j the pre-sweep
// one.  The batch kernels (k_sweep_cc, _cc2, _cc3) cut the sweep into runs of mutually independent
// consecutive nodes -- ~4 350 runs of ~11.5 nodes per slice, each run a barrier-delimited ~10 us,
// because a run ends at the FIRST node that reads any earlier member and because a committed
// position overwrites the old one (so later readers have to wait for earlier readers too).
//
// Here both constraints go:
//   * positions are double-buffered for the duration of the sweep: G[c][t][i] = {x, y, 1/r, 0} keeps
//     the PRE-sweep state read-only, an accepted move goes to Nw[c][t][i] and the node's state word
//     becomes (epoch << 1 | accepted).  A reader takes G[i] for i > j without any synchronisation
//     (no anti-dependencies), and for i < j waits for i's state word, then reads Nw[i] or G[i];
//   * nodes are handed out in index order by an atomic ticket per (chain, slice) to ~37 resident
//     warps per pair; a warp only waits for the specific earlier nodes its node reads.  The
//     dependency graph of a slice (edge probability ~220 / 50 000 per pair) has a critical path of
//     ~e * 220 = 600 nodes instead of 4 350 runs, so the sweep is bound by throughput (issue slots,
//     L2 gather sectors), not by the latency of one node.  Deadlock-free: tickets are taken in index
//     order by resident warps, so the smallest unfinished node of the smallest unfinished pair always
//     has all its inputs.
// The same mechanism carries the wavefront over the slices: node (t, j) waits for the state word of
// (t-1, j) (its new position enters the prior) and reads (t+1, j) from the pre-sweep record.
//
// Everything that does not depend on in-sweep state is hoisted into k_ccd_prep (all nodes in
// parallel: Metropolis state, Philox / replay draws, proposal, the "next slice" prior terms, the
// packed records); the Metropolis bookkeeping and the commit of accepted moves into X follow in
// k_ccd_post.  Decisions are those of the sequential sweep, bit for bit (the per-node arithmetic and
// its order of accumulation are k_sweep_cc's); tests/test_gpu_edge_cases.py, test_gpu_operating_points.py.
#include "dlsm_kernels.cuh"
#include "dlsm_ccd.h"

#include <cstdlib>
#include <cstring>

namespace dlsm {

constexpr int kCcdThreads = 640; // 20 warps per SM at <= 102 registers
constexpr int kPrepStride = 8;   // doubles per prep record

struct CcdWork {
    double *Nw = nullptr;    // [C][T][n][4] accepted moves {x', y', 1/r, 0}
    double *prep = nullptr;  // [C][T][n][8] {x', y', log u, next-prior(new), next-prior(old), 1/var, z, -}
    int *state = nullptr;    // [C][T][n]    (epoch << 1) | accepted once the node is decided
    int *next = nullptr;     // [C][T]       ticket counters
    size_t cells = 0, pairs = 0;
    int epoch = 0;
    int grid = 0;
    int group = 0;           // chains per launch of the sweep kernel (0 = all)
    int hints = 1;
};

struct CcdView {
    const double *G;
    double *Nw;
    double *prep;
    int *state;
    int *next;
    int epoch;
    int hints;               // 1: L2 eviction priorities (records / state words last, streamed lists first)
};

// L2 residency: the records and state words of the slices in flight are gathered at random ~220 times
// per node-update and must stay in L2 (evict_last), the list indices and prep records stream through
// once (evict_first, no L1 allocation).
__device__ __forceinline__ uint64_t policy_evict_last()
{
    uint64_t pol;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t policy_evict_normal()
{
    uint64_t pol;
    asm("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t policy_evict_first()
{
    uint64_t pol;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

__device__ __forceinline__ int ld_relaxed_gpu(const int *p, uint64_t pol)
{
    int v;
    asm volatile("ld.relaxed.gpu.global.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol) : "memory");
    return v;
}

// 256-bit load served by L2 (the point of coherence): Nw records are written by other CTAs during
// the kernel.  volatile + memory clobber: never speculated above the state word it depends on.
__device__ __forceinline__ void ld256cg(const double *p, double &a, double &b, double &c, double &d, uint64_t pol)
{
    asm volatile("ld.global.cg.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;"
                 : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p), "l"(pol) : "memory");
}

__device__ __forceinline__ int ld_stream_s32(const int32_t *p, uint64_t pol)
{
    int v;
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}

// read-only data of this kernel (prep records, the pre-sweep record of the node itself)
__device__ __forceinline__ double4 ld256nc(const double *p, uint64_t pol)
{
    double4 v;
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;"
        : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p), "l"(pol));
    return v;
}

__device__ __forceinline__ void st256(double *p, double a, double b, double c, double d, uint64_t pol)
{
    asm volatile("st.global.L2::cache_hint.v4.f64 [%0], {%1,%2,%3,%4}, %5;"
                 ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d), "l"(pol) : "memory");
}

__device__ __forceinline__ void st_release_gpu_hint(int *p, int v, uint64_t pol)
{
    asm volatile("st.release.gpu.global.L2::cache_hint.s32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(pol) : "memory");
}

// ---- all nodes in parallel: draws, proposals, next-slice prior terms, packed pre-sweep records ----
static __global__ void __launch_bounds__(256) k_ccd_prep(const SweepParams p, double *G, CcdView B)
{
    constexpr int DM = 2;
    const int T = p.net.T, n = p.net.n, d = 2;
    const size_t cell = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)p.C * T * n;
    if (cell >= total) return;
    const int c = (int)(cell / ((size_t)T * n));
    const int t = (int)((cell / n) % T), j = (int)(cell % n);
    double x0[DM], x[DM], eps[DM], logu;
    load_pos<DM>(p.X + cell * d, d, x0);
    const double step = p.step[cell];
    if (p.eps) {
        eps[0] = p.eps[cell * d]; eps[1] = p.eps[cell * d + 1];
        logu = p.logu[cell];
    } else {
        latent_draws<DM>(p.seed, (uint32_t)(t * n + j), p.sweep, (uint32_t)c + p.chain_offset, d, eps, logu);
    }
    x[0] = __dadd_rn(x0[0], __dmul_rn(step, eps[0]));
    x[1] = __dadd_rn(x0[1], __dmul_rn(step, eps[1]));
    double inv = (t == 0) ? 1.0 / p.tau_sq : 1.0 / p.sigma_sq;
    int zc = 0;
    if (p.prior != 0) {
        zc = p.z[cell];
        inv = 1.0 / p.sigma[(size_t)c * p.K + zc];
    }
    double nn = 0.0, no = 0.0;
    if (t < T - 1) { // X[t+1, j] enters at its pre-sweep value (sample_latent_positions.py:138-140)
        double xnx[DM];
        load_pos<DM>(p.X + (cell + n) * d, d, xnx);
        nn = prior_next<DM>(p, c, t, j, x, xnx);
        no = prior_next<DM>(p, c, t, j, x0, xnx);
    }
    double *pr = B.prep + cell * kPrepStride;
    reinterpret_cast<double4 *>(pr)[0] = make_double4(x[0], x[1], logu, nn);
    reinterpret_cast<double4 *>(pr)[1] = make_double4(no, inv, __longlong_as_double((long long)zc), 0.0);
    reinterpret_cast<double4 *>(G)[cell] = make_double4(x0[0], x0[1], p.rinv[(size_t)c * n + j], 0.0);
}

// ---- the sweep ----
// pairs [pair0, pair0 + CT) (whole chains) are served by this launch
__global__ void __launch_bounds__(kCcdThreads, 1) k_sweep_ccd(const SweepParams p, const CcdView B, int pair0, int CT)
{
    constexpr int DM = 2;
    const int T = p.net.T, n = p.net.n, d = 2, nc = p.net.n_control;
    const int lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
    const int gw = blockIdx.x * wpc + (threadIdx.x >> 5);
    const int Wtot = gridDim.x * wpc;
    const int epoch = B.epoch;
    const uint64_t keep = B.hints ? policy_evict_last() : policy_evict_normal();
    const uint64_t once = B.hints ? policy_evict_first() : policy_evict_normal();
    bool nonfinite = false;

    for (int pair = pair0 + gw % CT; pair < pair0 + CT; pair += Wtot) {
        const int c = pair / T, t = pair % T;
        const size_t slice = (size_t)pair * n;
        const double *Gt = B.G + slice * 4;
        double *Nt = B.Nw + slice * 4;
        int *St = B.state + slice;
        const double b0 = p.intercept[c * 2 + 0], b1 = p.intercept[c * 2 + 1];
        const size_t ctrl_slice = ((size_t)(p.net.ctrl_per_chain ? c : 0) * T + t) * n;

        for (;;) {
            int j = 0;
            if (lane == 0) j = atomicAdd(B.next + pair, 1);
            j = __shfl_sync(kFull, j, 0);
            if (j >= n) break;

            // ---- everything that is independent of the sweep's progress ----
            const size_t r = (size_t)t * n + j;
            const int indeg = ld_stream_s32(p.net.deg + r * 2 + 0, once), outdeg = ld_stream_s32(p.net.deg + r * 2 + 1, once);
            const int32_t *ie = p.net.in_edges + r * p.net.max_in;
            const int32_t *oe = p.net.out_edges + r * p.net.max_out;
            const int32_t *ci = p.net.ctrl_in + (ctrl_slice + j) * nc;
            const int32_t *co = p.net.ctrl_out + (ctrl_slice + j) * nc;
            const bool short_lists = indeg <= 16 && outdeg <= 16;
            const bool out_side = lane >= 16;
            int ke = j;       // the lane's edge-list entry when both lists fit one trip
            bool live = false;
            if (short_lists) {
                const int q = lane & 15;
                live = q < (out_side ? outdeg : indeg);
                if (live) ke = ld_stream_s32(out_side ? oe + q : ie + q, once);
            }
            int cin[4], cout[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int q = u * 32 + lane;
                cin[u] = q < nc ? ld_stream_s32(ci + q, once) : 0;
                cout[u] = q < nc ? ld_stream_s32(co + q, once) : 0;
            }
            const double4 pa = ld256nc(B.prep + (slice + j) * kPrepStride, once);
            const double4 pb = ld256nc(B.prep + (slice + j) * kPrepStride + 4, once);
            const double x[DM] = {pa.x, pa.y};
            const double4 own = ld256nc(Gt + (size_t)j * 4, keep);
            const double x0[DM] = {own.x, own.y};
            const double rj = own.z;
            // usable controls = prefix of ctrl_in before its first -1 (:137; and :161, which tests
            // the IN list while walking the OUT list)
            int m = nc, m_out;
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
            for (int u = 0; u < 4; u++) { // masked lanes gather the node's own record (no wait)
                if (!(u * 32 + lane < m)) cin[u] = j;
                if (!(u * 32 + lane < m_out)) cout[u] = j;
            }

            // ---- wait for the earlier nodes this node reads (and for node j of slice t-1) ----
            // bit u of `newer`: entry u comes from an accepted move of this sweep (read Nw, not G)
            unsigned newer = 0;
            {
                int s[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
                bool pend = false;
                auto probe = [&](int u, const int *addr, bool need) {
                    if (need && (s[u] >> 1) != epoch) {
                        s[u] = ld_relaxed_gpu(addr, keep);
                        pend |= (s[u] >> 1) != epoch;
                    }
                };
                for (;;) {
                    pend = false;
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        probe(u, St + cin[u], cin[u] < j);
                        probe(4 + u, St + cout[u], cout[u] < j);
                    }
                    probe(8, St + ke, ke < j);
                    probe(9, St - n + j, t > 0);
                    if (!__any_sync(kFull, pend)) break;
                    __nanosleep(DLSM_SPIN_NS);
                }
#pragma unroll
                for (int u = 0; u < 10; u++) newer |= (unsigned)(((s[u] >> 1) == epoch) & (s[u] & 1)) << u;
            }

            auto fetch = [&](int k, unsigned is_new, double (&xk)[DM], double &rk) {
                double pad;
                const double *src = (is_new ? (const double *)Nt : Gt) + (size_t)k * 4;
                ld256cg(src, xk[0], xk[1], rk, pad, keep);
            };
            auto eta_pair = [&](int k, unsigned is_new, bool k_sends, double &vn, double &vo) {
                double xk[DM], rk;
                fetch(k, is_new, xk, rk);
                const double dn = fast_dist<DM>(xk, x, d);
                const double dd = fast_dist<DM>(xk, x0, d);
                const double r_recv = k_sends ? rj : rk, r_send = k_sends ? rk : rj;
                vn = eta_directed(b0, b1, dn, r_recv, r_send);
                vo = eta_directed(b0, b1, dd, r_recv, r_send);
            };
            // an entry of a long edge list: its own wait (rare path)
            auto eta_pair_wait = [&](int k, bool k_sends, double &vn, double &vo) {
                unsigned is_new = 0;
                if (k < j) {
                    int sv = ld_relaxed_gpu(St + k, keep);
                    while ((sv >> 1) != epoch) { __nanosleep(DLSM_SPIN_NS); sv = ld_relaxed_gpu(St + k, keep); }
                    is_new = sv & 1;
                }
                eta_pair(k, is_new, k_sends, vn, vo);
            };

            double e_n = 0.0, e_o = 0.0;   // edge terms
            double ci_n = 0.0, ci_o = 0.0; // control sums over the in lists
            double co_n = 0.0, co_o = 0.0; // control sums over the out lists
            if (short_lists) { // in-list on lanes 0-15, out-list on lanes 16-31 (:108-133)
                double vn, vo;
                eta_pair(ke, (newer >> 8) & 1u, !out_side, vn, vo);
                const double tn = logit_term(0.5, vn), to = logit_term(0.5, vo);
                if (live) { e_n += tn; e_o += to; }
            } else {
                for (int q = lane; q < indeg; q += 32) {
                    double vn, vo;
                    eta_pair_wait(ie[q], true, vn, vo);
                    e_n += logit_term(0.5, vn);
                    e_o += logit_term(0.5, vo);
                }
                for (int q = lane; q < outdeg; q += 32) {
                    double vn, vo;
                    eta_pair_wait(oe[q], false, vn, vo);
                    e_n += logit_term(0.5, vn);
                    e_o += logit_term(0.5, vo);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) { // :136-152 and :160-176; masked sums
                const int q = u * 32 + lane;
                const bool vi = q < m, vo_ = q < m_out;
                double an, ao, bn, bo;
                eta_pair(cin[u], (newer >> u) & 1u, true, an, ao);
                eta_pair(cout[u], (newer >> (4 + u)) & 1u, false, bn, bo);
                const double la = log1pexp(an), lb = log1pexp(ao), lc = log1pexp(bn), ld = log1pexp(bo);
                if (vi) { ci_n += la; ci_o += lb; }
                if (vo_) { co_n += lc; co_o += ld; }
            }
            e_n = warp_sum(e_n); e_o = warp_sum(e_o);
            ci_n = warp_sum(ci_n); ci_o = warp_sum(ci_o);
            co_n = warp_sum(co_n); co_o = warp_sum(co_o);
            const double adj_in = (double)(n - indeg - 1) / (double)m;       // :155
            const double adj_out = (double)(n - outdeg - 1) / (double)m_out; // :179
            const double ll_new = (e_n - adj_in * ci_n) - adj_out * co_n;
            const double ll_old = (e_o - adj_in * ci_o) - adj_out * co_o;

            // ---- priors, decision (sample_latent_positions.py:131-146) ----
            double xp[DM] = {0.0, 0.0};
            if (t > 0) {
                double rp;
                const ptrdiff_t off = ((ptrdiff_t)j - n) * 4; // same node, slice t-1
                double pad;
                ld256cg(((newer >> 9) & 1u ? (const double *)Nt : Gt) + off, xp[0], xp[1], rp, pad, keep);
            }
            const double inv = pb.y;
            const int zc = (int)__double_as_longlong(pb.z);
            double lp_new = __dsub_rn(ll_new, prior_prev<DM>(p, c, t, zc, inv, x, xp));
            double lp_old = __dsub_rn(ll_old, prior_prev<DM>(p, c, t, zc, inv, x0, xp));
            if (t < T - 1) {
                lp_new = __dsub_rn(lp_new, pa.w);
                lp_old = __dsub_rn(lp_old, pb.x);
            }
            const double ratio = __dsub_rn(lp_new, lp_old);
            const int acc = (pa.z >= ratio) ? 0 : 1;
            if (lane == 0) {
                nonfinite |= !(ratio == ratio) || ratio - ratio != 0.0;
                if (p.ratio) p.ratio[slice + j] = ratio;
                if (acc) st256(Nt + (size_t)j * 4, x[0], x[1], rj, 0.0, keep);
                st_release_gpu_hint(St + j, (epoch << 1) | acc, keep); // orders the record before the state word
            }
        }
    }
    if (nonfinite) atomicOr(p.flags, 1u);
}

// ---- all nodes in parallel: Metropolis bookkeeping, accepted moves into X ----
static __global__ void __launch_bounds__(256) k_ccd_post(const SweepParams p, const CcdView B)
{
    const size_t cell = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)p.C * p.net.T * p.net.n;
    if (cell >= total) return;
    const int acc = B.state[cell] & 1;
    double step = p.step[cell];
    int nacc = p.nacc[cell], nsteps = p.nsteps[cell], until = p.until[cell];
    metropolis_bookkeep(step, nacc, nsteps, until, p.tune, p.tune_interval, acc, false);
    p.step[cell] = step; p.nacc[cell] = nacc; p.nsteps[cell] = nsteps; p.until[cell] = until;
    if (p.accepted) p.accepted[cell] = acc;
    if (acc) {
        const double4 v = reinterpret_cast<const double4 *>(B.Nw)[cell];
        reinterpret_cast<double2 *>(p.X)[cell] = make_double2(v.x, v.y);
    }
}

void ccd_free(CcdWork *w)
{
    if (!w) return;
    cudaFree(w->Nw); cudaFree(w->prep); cudaFree(w->state); cudaFree(w->next);
    delete w;
}

cudaError_t ccd_launch(const SweepParams &p, double *G, CcdWork **work, int sm_count, cudaStream_t stream,
                       int *launches)
{
    const size_t pairs = (size_t)p.C * p.net.T, cells = pairs * p.net.n;
    cudaError_t e;
    CcdWork *w = *work;
    if (w && (w->cells != cells || w->pairs != pairs)) { ccd_free(w); w = nullptr; *work = nullptr; }
    if (!w) {
        w = new CcdWork;
        w->cells = cells; w->pairs = pairs;
        *work = w;
        if ((e = cudaMalloc((void **)&w->Nw, cells * 4 * sizeof(double))) != cudaSuccess) return e;
        if ((e = cudaMalloc((void **)&w->prep, cells * kPrepStride * sizeof(double))) != cudaSuccess) return e;
        if ((e = cudaMalloc((void **)&w->state, cells * sizeof(int))) != cudaSuccess) return e;
        if ((e = cudaMalloc((void **)&w->next, pairs * sizeof(int))) != cudaSuccess) return e;
        if ((e = cudaMemsetAsync(w->state, 0, cells * sizeof(int), stream)) != cudaSuccess) return e;
        int per_sm = 0;
        if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sweep_ccd, kCcdThreads, 0)) != cudaSuccess)
            return e;
        if (per_sm < 1) return cudaErrorLaunchOutOfResources;
        w->grid = sm_count * per_sm; // all CTAs resident: a waiting warp's inputs are always being produced
        if (const char *g = getenv("DLSM_CCD_GROUP")) w->group = atoi(g);
        if (const char *g = getenv("DLSM_CCD_HINTS")) w->hints = atoi(g);
    }
    if (w->epoch >= (1 << 30)) { // state words must never match a stale epoch
        if ((e = cudaMemsetAsync(w->state, 0, cells * sizeof(int), stream)) != cudaSuccess) return e;
        w->epoch = 0;
    }
    w->epoch += 1;
    CcdView B;
    B.G = G; B.Nw = w->Nw; B.prep = w->prep; B.state = w->state; B.next = w->next; B.epoch = w->epoch; B.hints = w->hints;
    if ((e = cudaMemsetAsync(w->next, 0, pairs * sizeof(int), stream)) != cudaSuccess) return e;
    const unsigned nb = (unsigned)((cells + 255) / 256);
    k_ccd_prep<<<nb, 256, 0, stream>>>(p, G, B);
    // chains per launch: the records + state words of the pairs in flight (3.4 MB per pair at n = 50 000)
    // are gathered at random and should stay in L2; more warps per pair, on the other hand, wait more
    // often for a node that is still in flight
    int group = w->group > 0 ? w->group : p.C;
    int nl = 2;
    for (int c0 = 0; c0 < p.C; c0 += group, nl++) {
        const int gc = c0 + group <= p.C ? group : p.C - c0;
        k_sweep_ccd<<<(unsigned)w->grid, kCcdThreads, 0, stream>>>(p, B, c0 * p.net.T, gc * p.net.T);
    }
    k_ccd_post<<<nb, 256, 0, stream>>>(p, B);
    if (launches) *launches = nl;
    return cudaGetLastError();
}

} // namespace dlsm
