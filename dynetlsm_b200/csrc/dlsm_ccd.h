// dlsm_ccd.h -- launcher of the dataflow case-control sweep (dlsm_ccd.cu)
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace dlsm {
struct SweepParams;
struct CcdWork; // device buffers of the dataflow sweep, owned by dlsm_ccd.cu

// One latent-position sweep of every chain with the case-control likelihood (d = 2, n_control <= 128).
// G: [C][T][n][4] scratch for the packed {x, y, 1/r, 0} records of the pre-sweep state (rewritten here).
// *work is allocated on first use (ccd_free releases it).  chains_per_launch: 0 = heuristic (DLSM_OPT_CCD_GROUP).
// launches: number of kernels launched.
cudaError_t ccd_launch(const SweepParams &p, double *G, CcdWork **work, int sm_count, int chains_per_launch,
                       cudaStream_t stream, int *launches);
void ccd_free(CcdWork *work);
} // namespace dlsm
