// dlsm_graph.cu -- sparse network input: the case-control bookkeeping of
// DirectedCaseControlSampler.init (case_control_likelihood.py:37-73) built ON THE DEVICE from an
// edge list, for networks whose dense (T, n, n) tensor cannot exist (cfg 5: 200 GB).
//   degrees[t, i] = (in-degree, out-degree); out_edges[t, i, :] = receivers of i's ties in ascending
//   order, in_edges[t, i, :] = senders of the ties i receives, ascending; both zero-padded to the
//   largest degree -- exactly what the reference derives from np.where(Y[t, i, :] == 1) /
//   np.where(Y[t, :, i] == 1).
// Three passes over the E ties (count, fill through per-list cursors, sort each short list in place):
// O(E) HBM traffic instead of the reference's O(T n^2) Python loop over a dense Y.
#include "dlsm_graph.h"

#include <cstdint>

namespace dlsm {

// edges: (E, 3) int32 rows (t, sender i, receiver j)
static __global__ void k_edges_count(const int32_t *edges, size_t E, int T, int n, int32_t *deg, int *bad)
{
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int t = edges[e * 3], i = edges[e * 3 + 1], j = edges[e * 3 + 2];
    if ((unsigned)t >= (unsigned)T || (unsigned)i >= (unsigned)n || (unsigned)j >= (unsigned)n || i == j) {
        *bad = 1; // out of range or a self tie (the diagonal of Y is ignored by every likelihood)
        return;
    }
    atomicAdd(&deg[((size_t)t * n + i) * 2 + 1], 1);
    atomicAdd(&deg[((size_t)t * n + j) * 2 + 0], 1);
}

static __global__ void k_edges_max(const int32_t *deg, size_t cells, int *max_in, int *max_out)
{
    int mi = 0, mo = 0;
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < cells; c += (size_t)gridDim.x * blockDim.x) {
        mi = max(mi, deg[c * 2]);
        mo = max(mo, deg[c * 2 + 1]);
    }
    for (int o = 16; o > 0; o >>= 1) {
        mi = max(mi, __shfl_xor_sync(0xffffffffu, mi, o));
        mo = max(mo, __shfl_xor_sync(0xffffffffu, mo, o));
    }
    if ((threadIdx.x & 31) == 0) { atomicMax(max_in, mi); atomicMax(max_out, mo); }
}

static __global__ void k_edges_fill(const int32_t *edges, size_t E, int n, int32_t *cursor, int32_t *in_edges,
                                    int max_in, int32_t *out_edges, int max_out)
{
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int t = edges[e * 3], i = edges[e * 3 + 1], j = edges[e * 3 + 2];
    const size_t ci = (size_t)t * n + i, cj = (size_t)t * n + j;
    out_edges[ci * max_out + atomicAdd(&cursor[ci * 2 + 1], 1)] = j;
    in_edges[cj * max_in + atomicAdd(&cursor[cj * 2 + 0], 1)] = i;
}

// one thread per (cell, direction): insertion sort of a list of <= max degree entries; a repeated
// entry means the same tie was listed twice
static __global__ void k_edges_sort(const int32_t *deg, size_t cells, int32_t *in_edges, int max_in,
                                    int32_t *out_edges, int max_out, int *bad)
{
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= cells * 2) return;
    const size_t c = g >> 1;
    const int dir = (int)(g & 1);
    int32_t *lst = dir ? out_edges + c * max_out : in_edges + c * max_in;
    const int len = deg[c * 2 + dir];
    for (int a = 1; a < len; a++) {
        const int32_t v = lst[a];
        int b = a - 1;
        while (b >= 0 && lst[b] > v) { lst[b + 1] = lst[b]; b--; }
        lst[b + 1] = v;
    }
    for (int a = 1; a < len; a++)
        if (lst[a] == lst[a - 1]) *bad = 2;
}

cudaError_t graph_build_edge_lists(const int32_t *d_edges, size_t E, int T, int n, int32_t **deg_out,
                                   int32_t **in_out, int *max_in, int32_t **out_out, int *max_out,
                                   int *status, cudaStream_t stream)
{
    const size_t cells = (size_t)T * n;
    int32_t *deg = nullptr, *cursor = nullptr, *in_e = nullptr, *out_e = nullptr;
    int *d_small = nullptr; // [bad, max_in, max_out]
    cudaError_t e;
#define TRY(x) if ((e = (x)) != cudaSuccess) goto fail
    TRY(cudaMalloc((void **)&deg, cells * 2 * sizeof(int32_t)));
    TRY(cudaMalloc((void **)&cursor, cells * 2 * sizeof(int32_t)));
    TRY(cudaMalloc((void **)&d_small, 3 * sizeof(int)));
    TRY(cudaMemsetAsync(deg, 0, cells * 2 * sizeof(int32_t), stream));
    TRY(cudaMemsetAsync(cursor, 0, cells * 2 * sizeof(int32_t), stream));
    TRY(cudaMemsetAsync(d_small, 0, 3 * sizeof(int), stream));
    {
        const unsigned eb = (unsigned)((E + 255) / 256);
        if (E) k_edges_count<<<eb, 256, 0, stream>>>(d_edges, E, T, n, deg, d_small);
        k_edges_max<<<148 * 4, 256, 0, stream>>>(deg, cells, d_small + 1, d_small + 2);
        int small[3];
        TRY(cudaMemcpyAsync(small, d_small, sizeof(small), cudaMemcpyDeviceToHost, stream));
        TRY(cudaStreamSynchronize(stream));
        *status = small[0];
        *max_in = small[1];
        *max_out = small[2];
        if (small[0] == 0) {
            TRY(cudaMalloc((void **)&in_e, (cells * (size_t)small[1] + 1) * sizeof(int32_t)));
            TRY(cudaMalloc((void **)&out_e, (cells * (size_t)small[2] + 1) * sizeof(int32_t)));
            TRY(cudaMemsetAsync(in_e, 0, (cells * (size_t)small[1] + 1) * sizeof(int32_t), stream));
            TRY(cudaMemsetAsync(out_e, 0, (cells * (size_t)small[2] + 1) * sizeof(int32_t), stream));
            if (E) k_edges_fill<<<eb, 256, 0, stream>>>(d_edges, E, n, cursor, in_e, small[1], out_e, small[2]);
            k_edges_sort<<<(unsigned)((cells * 2 + 255) / 256), 256, 0, stream>>>(deg, cells, in_e, small[1], out_e,
                                                                                 small[2], d_small);
            TRY(cudaMemcpyAsync(small, d_small, sizeof(int), cudaMemcpyDeviceToHost, stream));
            TRY(cudaStreamSynchronize(stream));
            *status = small[0];
        }
    }
    TRY(cudaGetLastError());
    cudaFree(cursor);
    cudaFree(d_small);
    if (*status != 0) { cudaFree(deg); cudaFree(in_e); cudaFree(out_e); deg = in_e = out_e = nullptr; }
    *deg_out = deg; *in_out = in_e; *out_out = out_e;
    return cudaSuccess;
fail:
    cudaFree(deg); cudaFree(cursor); cudaFree(in_e); cudaFree(out_e); cudaFree(d_small);
    return e;
#undef TRY
}

} // namespace dlsm
