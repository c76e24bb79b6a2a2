// dlsm_cc.h -- launcher of the second-generation case-control sweep kernel (dlsm_cc.cu)
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace dlsm {
struct SweepParams;
// G: packed gather records [C][T][n][4] = {x, y, 1/r, 0} in sync with X and the radii (k_pack_gather);
// the kernel keeps them in sync with the moves it accepts.  d = 2, n_control <= 128.
cudaError_t cc2_launch(const SweepParams &p, double *G, int *progress, unsigned int *ticket, const int32_t *dep,
                       cudaStream_t stream);
size_t cc2_smem_bytes(int max_in, int max_out, int n_control);
// k_sweep_cc3: a 2-CTA cluster per (chain, slice), in-lists on one CTA, out-lists on the other (d = 2,
// n_control <= 128); max_active != nullptr: only report how many clusters can be resident
cudaError_t cc3_launch(const SweepParams &p, int *progress, unsigned int *ticket, const int32_t *dep,
                       cudaStream_t stream, int *max_active);
} // namespace dlsm
