// dlsm_fullr.cu -- k_full_lr: the exact full-network log-likelihood (K5 network_likelihoods.py:26-33 /
// K4 directed_likelihoods_fast.pyx:185-205) of ONE parameter variant with lanes = rows.
//
// k_full (dlsm_kernels.cuh) walks folded rows with lanes = columns: every pair pays its own index
// arithmetic (row / column of the folded slot, clamps, a 64-bit address and two loads for the adjacency
// words, selects between the two glued rows) -- ~2 non-fp64 instructions per fp64 one, on a path whose
// time is 2 F + O issue cycles.  Here lane l of a warp keeps row node
// a = 32 k + l of row block k in registers and the warp walks 32 columns b > a per work item: x_b and
// 1/r_b are broadcast shared-memory loads, the adjacency bits of the 32 columns are ONE word of the
// lane's own bit-row (and one of its column-major copy), the sum stays in the lane.  Only the first
// item of a row block (its own 32 columns) needs a mask (b > a).  Work items (row block, 32-column
// word) are dealt round-robin to the warps of the slice's CTAs.
// Same partial-sum layout as k_full (p.partial[c][T * tiles][2], slot 0), d = 2, one variant.
#include "dlsm_kernels.cuh"
#include "dlsm_fullr.h"

namespace dlsm {

template <int LK>
__global__ void __launch_bounds__(256) k_full_lr(const FullParams p)
{
    constexpr int DM = 2;
    constexpr bool kDir = LK != kUndirected;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red[8];
    const int T = p.net.T, n = p.net.n, d = 2, W = p.net.W;
    const int c = blockIdx.y;
    const int t = blockIdx.x / p.tiles, tile = blockIdx.x % p.tiles;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    double *Xs = reinterpret_cast<double *>(smem_raw);     // [n][2]
    double *Rs = Xs + (size_t)n * d;                       // [n]   (directed)
    const double *Xg = p.X + ((size_t)c * T + t) * n * d;
    for (int e = threadIdx.x; e < n * d; e += blockDim.x) Xs[e] = Xg[e];
    if (kDir) {
        const double *rg = p.rinv0 + (size_t)c * n;
        for (int e = threadIdx.x; e < n; e += blockDim.x) Rs[e] = rg[e];
    }
    __syncthreads();
    const double b0 = p.bvar[(size_t)c * 4], b1 = p.bvar[(size_t)c * 4 + 1];
    const int nb = (n + 31) >> 5;
    const int me = tile * nwarps + warp, stride = p.tiles * nwarps;
    double acc0 = 0.0, acc1 = 0.0;
    int id = 0; // running index of the work items (row block k, column word k + ch)
    for (int k = 0; k < nb; k++) {
        const int nch = nb - k; // column words k .. nb-1
        int ch = (me - id) % stride;
        if (ch < 0) ch += stride;
        if (ch < nch) {
            const int a = 32 * k + lane;
            const bool va = a < n;
            const int ac = va ? a : n - 1;
            double xa[DM];
            load_pos<DM>(Xs + (size_t)ac * d, d, xa);
            const double ra = kDir ? Rs[ac] : 0.0;
            const uint32_t *rowb = p.net.rowbits + ((size_t)t * n + ac) * W;
            const uint32_t *colb = kDir ? p.net.colbits + ((size_t)t * n + ac) * W : nullptr;
            for (; ch < nch; ch += stride) {
                const int w = k + ch, base = 32 * w;
                const uint32_t wr = __ldg(rowb + w), wc = kDir ? __ldg(colb + w) : 0u;
                const int cnt = (n - base) < 32 ? (n - base) : 32;
                auto pair = [&](int bb, double &acc) {
                    const int b = base + bb;
                    double xb[DM];
                    load_pos<DM>(Xs + (size_t)b * d, d, xb);
                    const double dist = fast_dist<DM>(xb, xa, d);
                    double term;
                    if (kDir) {
                        const double rb = Rs[b];
                        term = logit_term(ymask(wr, bb), eta_directed(b0, b1, dist, rb, ra)) +
                               logit_term(ymask(wc, bb), eta_directed(b0, b1, dist, ra, rb));
                    } else {
                        term = logit_term(ymask(wr, bb), b0 - dist);
                    }
                    const bool ok = va && (ch > 0 || bb > lane); // the block's own columns: b > a only
                    acc = fma(vmask(ok), term, acc);
                };
                int bb = 0;
                for (; bb + 1 < cnt; bb += 2) { // two columns per trip: two (directed: four) softplus chains
                    pair(bb, acc0);
                    pair(bb + 1, acc1);
                }
                if (bb < cnt) pair(bb, acc0);
            }
        }
        id += nch;
    }
    double s = warp_sum(acc0 + acc1);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < nwarps; w++) tot += red[w];
        p.partial[((size_t)c * gridDim.x + blockIdx.x) * 2] = tot;
        p.partial[((size_t)c * gridDim.x + blockIdx.x) * 2 + 1] = 0.0;
    }
}

size_t fullr_smem_bytes(int n, bool directed) { return (size_t)n * (directed ? 3 : 2) * sizeof(double); }

cudaError_t fullr_launch(const FullParams &p, bool directed, dim3 grid, cudaStream_t stream)
{
    const size_t smem = fullr_smem_bytes(p.net.n, directed);
    cudaError_t e;
    if (directed) {
        if ((e = cudaFuncSetAttribute(k_full_lr<kDirected>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        k_full_lr<kDirected><<<grid, 256, smem, stream>>>(p);
    } else {
        if ((e = cudaFuncSetAttribute(k_full_lr<kUndirected>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        k_full_lr<kUndirected><<<grid, 256, smem, stream>>>(p);
    }
    return cudaGetLastError();
}

} // namespace dlsm
