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
//   * positions are double-buffered for the duration of the sweep.  G[c][t][i] = {x, y, 1/r, 0} keeps
//     the PRE-sweep state read-only; Nw[c][t][i] starts as {SENT, SENT, 1/r, 0} (SENT: a NaN pattern no
//     position takes) and the warp that decides node i writes i's FINAL position (the proposal if
//     accepted, the old one if not) over the two sentinels.  A reader takes G[i] for i > j without
//     any synchronisation (no anti-dependencies); for i < j it loads Nw[i] and retries while either
//     coordinate is still the sentinel.  Every 8-byte word is written exactly once per sweep, by a
//     relaxed (strong) store, and read by relaxed loads: a non-sentinel word IS the final value, so
//     no ordering between the words, no fences and no separate flag are needed -- one 32-byte gather
//     per list entry;
//   * nodes are handed out in index order by an atomic ticket per (chain, slice) to the resident
//     warps of the pair; a warp only waits for the specific earlier nodes its node reads.  The
//     dependency graph of a slice (edge probability ~220 / 50 000 per pair) has a critical path of
//     ~e * 220 = 600 nodes instead of 4 350 runs, so the sweep is bound by throughput (issue slots,
//     gather sectors), not by the latency of one node.  Deadlock-free: tickets are taken in index
//     order by resident warps, so the smallest unfinished node of the smallest unfinished pair always
//     has all its inputs.
// The same mechanism carries the wavefront over the slices: node (t, j) waits for the record of
// (t-1, j) (its new position enters the prior) and reads (t+1, j) from the pre-sweep state.
//
// Everything that does not depend on in-sweep state is hoisted into k_ccd_prep (all nodes in
// parallel: Metropolis state, Philox / replay draws, proposal, the "next slice" prior terms, the
// packed records); the Metropolis bookkeeping and the commit of accepted moves into X follow in
// k_ccd_post.  Decisions are those of the sequential sweep, bit for bit (the per-node arithmetic and
// its order of accumulation are k_sweep_cc's); tests/test_gpu_edge_cases.py, test_gpu_operating_points.py.
// Measured (cfg 5, 8 chains, B200): 43.1 ms (k_sweep_cc) -> 19.7 ms (dataflow with per-node state words
// + release/acquire) -> 13.4 ms (self-validating records); profiles/r2d_*.
#include "dlsm_kernels.cuh"
#include "dlsm_ccd.h"

#include <cstring>

namespace dlsm {

constexpr int kCcdThreads = 640; // 20 warps per SM at <= 102 registers (768 threads at 80 registers measured: +-0)
constexpr int kPrepStride = 8;   // doubles per prep record
constexpr int kCcdSlots = 256;   // list entries of a node whose indices are held in registers (8 per lane)
constexpr unsigned long long kSentinel = 0xFFF8DEADBEEF0001ull; // "undecided" coordinate of a k_sweep_ccd record

struct CcdWork {
    double *Nw = nullptr;    // [C][T][n][4] this sweep's final positions {x, y, 1/r, 0} (sentinels until decided)
    double *prep = nullptr;  // [C][T][n][8] {x', y', log u, next-prior(new), next-prior(old), 1/var, z, -}
    int *state = nullptr;    // [C][T][n]    1 if the node's proposal was accepted (read by k_ccd_post)
    int *next = nullptr;     // [C][T]       ticket counters
    size_t cells = 0, pairs = 0;
    int grid = 0;
};

struct CcdView {
    const double *G;
    double *Nw;
    double *prep;
    int *state;
    int *next;
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
__device__ __forceinline__ uint64_t policy_evict_first()
{
    uint64_t pol;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

// 256-bit relaxed (strong, gpu scope) load, served by L2 -- the point of coherence: Nw records are
// written by other CTAs during the kernel.  Each 8-byte element is single-copy atomic.
__device__ __forceinline__ void ld256_relaxed(const double *p, double &a, double &b, double &c, double &d, uint64_t pol)
{
    asm volatile("ld.relaxed.gpu.global.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;"
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

__device__ __forceinline__ void st128_relaxed(double *p, double a, double b, uint64_t pol)
{
    asm volatile("st.relaxed.gpu.global.L2::cache_hint.v2.f64 [%0], {%1,%2}, %3;" ::"l"(p), "d"(a), "d"(b), "l"(pol) : "memory");
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
    const double rinv = p.rinv[(size_t)c * n + j];
    reinterpret_cast<double4 *>(G)[cell] = make_double4(x0[0], x0[1], rinv, 0.0);
    const double sent = __longlong_as_double((long long)kSentinel);
    reinterpret_cast<double4 *>(B.Nw)[cell] = make_double4(sent, sent, rinv, 0.0);
}

// ---- the sweep: pairs [pair0, pair0 + CT) (whole chains) are served by one launch ----
// (the test looks at the high word only: any double with these 32 bits is a NaN)
__device__ __forceinline__ bool is_sentinel(double v) { return (unsigned)__double2hiint(v) == (unsigned)(kSentinel >> 32); }
__device__ __forceinline__ double not_sentinel(double v)
{
    return is_sentinel(v) ? __longlong_as_double(0x7FF8000000000000ll) : v; // (a NaN position: the chain is lost anyway)
}

struct CcdRec { double x, y, r; };

template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k_sweep_ccd(const SweepParams p, const CcdView B, int pair0, int CT)
{
    constexpr int DM = 2;
    const int T = p.net.T, n = p.net.n, d = 2, nc = p.net.n_control;
    const int lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
    const int gw = blockIdx.x * wpc + (threadIdx.x >> 5);
    const int Wtot = gridDim.x * wpc;
    const uint64_t keep = policy_evict_last(), once = policy_evict_first();
    bool nonfinite = false;

    for (int pair = pair0 + gw % CT; pair < pair0 + CT; pair += Wtot) {
        const int c = pair / T, t = pair % T;
        const size_t slice = (size_t)pair * n;
        const double *Gt = B.G + slice * 4;
        double *Nt = B.Nw + slice * 4;
        const double b0 = p.intercept[c * 2 + 0], b1 = p.intercept[c * 2 + 1];
        const size_t ctrl_slice = ((size_t)(p.net.ctrl_per_chain ? c : 0) * T + t) * n;

        for (;;) {
            int j = 0;
            if (lane == 0) j = atomicAdd(B.next + pair, 1);
            j = __shfl_sync(kFull, j, 0);
            if (j >= n) break;

            // the record of list entry k as node j must see it: final (this sweep's) for k < j, pre-sweep for k > j
            auto load = [&](int k, CcdRec &o) {
                double pad;
                ld256_relaxed((k < j ? (const double *)Nt : Gt) + (size_t)k * 4, o.x, o.y, o.r, pad, keep);
            };
            auto settle = [&](int k, CcdRec &o) { // warp-collective: retry while node k < j is undecided
                bool pend = k < j && (is_sentinel(o.x) || is_sentinel(o.y));
                while (__any_sync(kFull, pend)) {
                    __nanosleep(DLSM_SPIN_NS);
                    if (pend) {
                        load(k, o);
                        pend = is_sentinel(o.x) || is_sentinel(o.y);
                    }
                }
            };

            // ---- everything that is independent of the sweep's progress ----
            const size_t r = (size_t)t * n + j;
            const int indeg = ld_stream_s32(p.net.deg + r * 2 + 0, once), outdeg = ld_stream_s32(p.net.deg + r * 2 + 1, once);
            const int32_t *ie = p.net.in_edges + r * p.net.max_in;
            const int32_t *oe = p.net.out_edges + r * p.net.max_out;
            const int32_t *ci = p.net.ctrl_in + (ctrl_slice + j) * nc;
            const int32_t *co = p.net.ctrl_out + (ctrl_slice + j) * nc;
            CcdRec prev = {0.0, 0.0, 0.0};
            if (t > 0) { // X[t-1, j] of THIS sweep: the same node's record one slice back
                double pad;
                ld256_relaxed(Nt + ((ptrdiff_t)j - n) * 4, prev.x, prev.y, prev.r, pad, keep);
            }
            const double4 pa = ld256nc(B.prep + (slice + j) * kPrepStride, once);
            const double4 pb = ld256nc(B.prep + (slice + j) * kPrepStride + 4, once);
            const double x[DM] = {pa.x, pa.y};
            const double4 own = ld256nc(Gt + (size_t)j * 4, keep);
            const double x0[DM] = {own.x, own.y};
            const double rj = own.z;

            // both MH evaluations of one list entry: eta at the proposal and at the current position
            auto eta_pair = [&](const CcdRec &o, bool k_sends, double &vn, double &vo) {
                const double xk[DM] = {o.x, o.y};
                const double dn = fast_dist<DM>(xk, x, d);
                const double dd = fast_dist<DM>(xk, x0, d);
                const double r_recv = k_sends ? rj : o.r, r_send = k_sends ? o.r : rj;
                vn = eta_directed(b0, b1, dn, r_recv, r_send);
                vo = eta_directed(b0, b1, dd, r_recv, r_send);
            };

            double e_n = 0.0, e_o = 0.0;   // edge terms (:108-133)
            double ci_n = 0.0, ci_o = 0.0; // control sums over the in lists (:136-152)
            double co_n = 0.0, co_o = 0.0; // control sums over the out lists (:160-176)
            int m = nc, m_out;             // usable controls
            // One slot space per node: in-controls [0, nc), out-controls [nc, 2 nc), in-edges, out-edges.
            // The ~220 entries of a cfg-5 node fill 7 trips of 32 lanes (the per-list layout of
            // k_sweep_cc needs 9).  All list indices are in registers after one round trip.
            const int B1 = nc, B2 = 2 * nc, B3 = B2 + indeg, total = B3 + outdeg;
            if (kFusedSoftplusForms && total <= kCcdSlots) {
                int kq[kCcdSlots / 32];
#pragma unroll
                for (int u = 0; u < kCcdSlots / 32; u++) {
                    const int q = u * 32 + lane;
                    const int32_t *src = q < B1 ? ci + q : (q < B2 ? co + (q - B1) : (q < B3 ? ie + (q - B2) : oe + (q - B3)));
                    kq[u] = q < total ? ld_stream_s32(src, once) : j;
                }
                // usable controls = prefix of ctrl_in before its first -1 (:137; and :161, which tests
                // the IN list while walking the OUT list)
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const unsigned bal = __ballot_sync(kFull, u * 32 + lane < B1 && kq[u] == -1);
                    if (bal && m == nc) m = u * 32 + __ffs(bal) - 1;
                }
                m_out = m;
#pragma unroll
                for (int u = 0; u < kCcdSlots / 32; u++) { // the reference reads X[-1] here: flag + stop
                    const int qo = u * 32 + lane - B1;
                    const unsigned bal = __ballot_sync(kFull, qo >= 0 && qo < m && kq[u] < 0);
                    if (bal && m_out == m) {
                        m_out = u * 32 + __ffs(bal) - 1 - B1;
                        if (lane == 0) atomicOr(p.flags, 2u);
                    }
                }
                unsigned valid = 0; // bit u: this lane's slot of trip u enters a sum
#pragma unroll
                for (int u = 0; u < kCcdSlots / 32; u++) {
                    const int q = u * 32 + lane;
                    const bool ok = q < B1 ? q < m : (q < B2 ? q - B1 < m_out : q < total);
                    if (ok) valid |= 1u << u;
                    else kq[u] = j; // masked lanes gather the node's own pre-sweep record (no wait)
                }
                // Fully unrolled over the trips, the next trip's gathers in flight while one is evaluated.
                // (Measured alternatives: three trips ahead 7.7 ms per launch instead of 6.7; a compact loop
                // with the indices in shared memory 8.5 ms -- no instruction-fetch stalls any more, but 17 %
                // more instructions and twice the long-scoreboard stalls.)
                CcdRec rec[2];
                load(kq[0], rec[0]);
#pragma unroll
                for (int u = 0; u < kCcdSlots / 32; u++) {
                    if (u * 32 >= total) break; // (warp-uniform)
                    if (u + 1 < kCcdSlots / 32 && (u + 1) * 32 < total) load(kq[u + 1], rec[(u + 1) & 1]);
                    settle(kq[u], rec[u & 1]);
                    const int q = u * 32 + lane;
                    const bool is_edge = q >= B2, in_side = q < B1 || (is_edge && q < B3); // in-lists: k sends to j
                    double vn, vo;
                    eta_pair(rec[u & 1], in_side, vn, vo);
                    // the expensive part L = log1p(e^-|eta|) once, wrapped as an edge term logit_term(0.5, eta)
                    // or as a control term log1pexp(eta)
                    const double an = fabs(vn), ao = fabs(vo);
                    const double Ln = DLSM_L1PEN(an), Lo = DLSM_L1PEN(ao);
                    const double tn = is_edge ? fma(0.5, vn, fma(-0.5, an, -Ln)) : fma(0.5, an, 0.5 * vn) + Ln;
                    const double to = is_edge ? fma(0.5, vo, fma(-0.5, ao, -Lo)) : fma(0.5, ao, 0.5 * vo) + Lo;
                    if ((valid >> u) & 1u) {
                        if (is_edge) { e_n += tn; e_o += to; }
                        else if (q < B1) { ci_n += tn; ci_o += to; }
                        else { co_n += tn; co_o += to; }
                    }
                }
            } else {
                // long lists: one list after the other, a round trip per 32 entries
                for (int base = 0; base < nc; base += 32) {
                    const int q = base + lane;
                    const unsigned bal = __ballot_sync(kFull, q < nc && ci[q] == -1);
                    if (bal) { m = base + __ffs(bal) - 1; break; }
                }
                m_out = m;
                for (int base = 0; base < m; base += 32) { // the reference reads X[-1] here: flag + stop
                    const int q = base + lane;
                    const unsigned bal = __ballot_sync(kFull, q < m && co[q] < 0);
                    if (bal) {
                        m_out = base + __ffs(bal) - 1;
                        if (lane == 0) atomicOr(p.flags, 2u);
                        break;
                    }
                }
                auto walk = [&](const int32_t *lst, int len, bool k_sends, bool is_edge, double &sn, double &so) {
                    for (int base = 0; base < len; base += 32) {
                        const int q = base + lane;
                        const int k = q < len ? lst[q] : j;
                        CcdRec o;
                        load(k, o);
                        settle(k, o);
                        double vn, vo;
                        eta_pair(o, k_sends, vn, vo);
                        const double tn = is_edge ? logit_term(0.5, vn) : log1pexp(vn);
                        const double to = is_edge ? logit_term(0.5, vo) : log1pexp(vo);
                        if (q < len) { sn += tn; so += to; }
                    }
                };
                walk(ie, indeg, true, true, e_n, e_o);
                walk(oe, outdeg, false, true, e_n, e_o);
                walk(ci, m, true, false, ci_n, ci_o);
                walk(co, m_out, false, false, co_n, co_o);
            }
            warp_sum2(e_n, e_o, lane);
            warp_sum2(ci_n, ci_o, lane);
            warp_sum2(co_n, co_o, lane);
            const double adj_in = (double)(n - indeg - 1) / (double)m;       // :155
            const double adj_out = (double)(n - outdeg - 1) / (double)m_out; // :179
            const double ll_new = (e_n - adj_in * ci_n) - adj_out * co_n;
            const double ll_old = (e_o - adj_in * ci_o) - adj_out * co_o;

            // ---- priors, decision (sample_latent_positions.py:131-146) ----
            if (t > 0) { // (uniform over the warp)
                bool pend = is_sentinel(prev.x) || is_sentinel(prev.y);
                while (pend) {
                    __nanosleep(DLSM_SPIN_NS);
                    double pad;
                    ld256_relaxed(Nt + ((ptrdiff_t)j - n) * 4, prev.x, prev.y, prev.r, pad, keep);
                    pend = is_sentinel(prev.x) || is_sentinel(prev.y);
                }
            }
            const double xp[DM] = {prev.x, prev.y};
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
                // the node's final position replaces the two sentinels (each word written once, atomically)
                const double fx = not_sentinel(acc ? x[0] : x0[0]), fy = not_sentinel(acc ? x[1] : x0[1]);
                st128_relaxed(Nt + (size_t)j * 4, fx, fy, keep);
                B.state[slice + j] = acc;
                nonfinite |= !(ratio == ratio) || ratio - ratio != 0.0;
                if (p.ratio) p.ratio[slice + j] = ratio;
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
    const int acc = B.state[cell];
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

cudaError_t ccd_launch(const SweepParams &p, double *G, CcdWork **work, int sm_count, int chains_per_launch,
                       cudaStream_t stream, int *launches)
{
    const size_t pairs = (size_t)p.C * p.net.T, cells = pairs * p.net.n;
    cudaError_t e;
    CcdWork *w = *work;
    if (w && (w->cells != cells || w->pairs != pairs)) { ccd_free(w); w = nullptr; *work = nullptr; }
    if (!w) {
        w = new CcdWork;
        w->cells = cells; w->pairs = pairs;
        int per_sm = 0;
        e = cudaMalloc((void **)&w->Nw, cells * 4 * sizeof(double));
        if (e == cudaSuccess) e = cudaMalloc((void **)&w->prep, cells * kPrepStride * sizeof(double));
        if (e == cudaSuccess) e = cudaMalloc((void **)&w->state, cells * sizeof(int));
        if (e == cudaSuccess) e = cudaMalloc((void **)&w->next, pairs * sizeof(int));
        if (e == cudaSuccess) {
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sweep_ccd<kCcdThreads>, kCcdThreads, 0);
        }
        if (e == cudaSuccess && per_sm < 1) e = cudaErrorLaunchOutOfResources;
        if (e != cudaSuccess) { ccd_free(w); return e; } // nothing half-built is kept
        w->grid = sm_count * per_sm; // all CTAs resident: a waiting warp's inputs are always being produced
        *work = w;
    }
    CcdView B;
    B.G = G; B.Nw = w->Nw; B.prep = w->prep; B.state = w->state; B.next = w->next;
    if ((e = cudaMemsetAsync(w->next, 0, pairs * sizeof(int), stream)) != cudaSuccess) return e;
    const unsigned nb = (unsigned)((cells + 255) / 256);
    k_ccd_prep<<<nb, 256, 0, stream>>>(p, G, B);
    // Chains per launch: the records of the pairs in flight (3.2 MB per pair at n = 50 000) are gathered
    // at random and should stay in L2, and a pair wants >= ~64 warps to cover its dependency graph's
    // parallelism; more warps per pair wait more often for a node that is still in flight.  Measured at
    // cfg 5: 8 chains per launch 13.6 ms, 4 chains 13.4 ms (profiles/r2d_ab_cfg5.json).
    const int warps = w->grid * (kCcdThreads / 32);
    int group = chains_per_launch > 0 ? chains_per_launch : (warps / 64) / p.net.T;
    group = group < 1 ? 1 : (group > p.C ? p.C : group);
    int nl = 2;
    for (int c0 = 0; c0 < p.C; c0 += group, nl++) {
        const int gc = c0 + group <= p.C ? group : p.C - c0;
        k_sweep_ccd<kCcdThreads><<<(unsigned)w->grid, kCcdThreads, 0, stream>>>(p, B, c0 * p.net.T, gc * p.net.T);
    }
    k_ccd_post<<<nb, 256, 0, stream>>>(p, B);
    if (launches) *launches = nl;
    return cudaGetLastError();
}

} // namespace dlsm
