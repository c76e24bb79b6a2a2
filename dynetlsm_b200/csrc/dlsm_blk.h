// dlsm_blk.h -- launcher of the block-speculative cluster sweep kernel (dlsm_blk.cu)
#pragma once
#include <cuda_runtime.h>

namespace dlsm {
struct SweepParams;
// Launches k_sweep_blk on clusters of CS CTAs x nwarps warps, one cluster per (chain, slice); with
// max_active != nullptr nothing is launched and *max_active receives the number of such clusters
// that can be resident at once.
cudaError_t blk_launch(const SweepParams &p, bool directed, int CS, int nwarps, int *progress,
                       unsigned int *ticket, cudaStream_t stream, int *max_active);
size_t blk_smem_bytes(int n, int d, bool directed, int W);
// k_sweep_blkw: the same sweep with a two-block window (the leader resolves block b while the cluster's
// work warps run the parallel phase of block b+1); clusters of 2 <= CS <= 8 CTAs x 16 warps
// ll_slices: optional [C][T]: the slice's dyads {i < j} at the post-sweep state (their sum over t is the
// full-network log-likelihood the sweep leaves behind)
cudaError_t blkw_launch(const SweepParams &p, bool directed, int CS, int *progress, unsigned int *ticket,
                        double *ll_slices, cudaStream_t stream, int *max_active);
size_t blkw_smem_bytes(int n, int d, bool directed, int W);
// k_sweep_cb: block-speculative sweep, one CTA per chain / one warp per slice (many chains)
// allow_pair: chains that have an SM to themselves (C <= 148, d = 2, T <= 15) run on k_sweep_cbp
cudaError_t cb_launch(const SweepParams &p, bool directed, cudaStream_t stream, bool allow_pair = false);
// k_sweep_cbp (dlsm_cbp.cu): the same sweep with TWO warps per slice, one chain per SM
bool cbp_applicable(int T, int n, int d);
cudaError_t cbp_launch(const SweepParams &p, bool directed, cudaStream_t stream);
size_t cb_smem_bytes(int T, int n, int d, bool xs);
} // namespace dlsm
