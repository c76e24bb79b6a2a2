// dlsm_fullr.h -- launcher of the lanes-are-rows full-network log-likelihood kernel (dlsm_fullr.cu)
#pragma once
#include <cuda_runtime.h>

namespace dlsm {
struct FullParams;
// exact (undirected / directed) likelihood, d = 2, one parameter variant (p.bvar[c][0..1], p.rinv0);
// grid = (T * p.tiles, C) as for k_full; partial sums land in p.partial[c][T * tiles][2] (slot 0)
cudaError_t fullr_launch(const FullParams &p, bool directed, dim3 grid, cudaStream_t stream);
size_t fullr_smem_bytes(int n, bool directed);
} // namespace dlsm
