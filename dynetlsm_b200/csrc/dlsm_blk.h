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
// k_sweep_cb: block-speculative sweep, one CTA per chain / one warp per slice (many chains)
cudaError_t cb_launch(const SweepParams &p, bool directed, cudaStream_t stream);
size_t cb_smem_bytes(int T, int n, int d, bool xs);
} // namespace dlsm
