// dlsm_dyad.cuh -- the dyad term shared by the block-speculative kernels (dlsm_blk.cu, dlsm_cbp.cu)
#pragma once
#include "dlsm_kernels.cuh"

namespace dlsm {

// the dyad {row node j, column node i}: both directions for the directed model
//   yr: bit Y[j, i] (j sends), yc: bit Y[i, j] (i sends), as y - 1/2
template <int LK, int DM>
__device__ __forceinline__ double dyad(const double (&xi)[DM], double ri, const double (&xj)[DM], double rj,
                                       double yr, double yc, double b0, double b1, int d)
{
    const double dist = fast_dist<DM>(xi, xj, d);
    if (LK == kUndirected) return logit_term(yr, b0 - dist);
    return logit_term(yr, eta_directed(b0, b1, dist, ri, rj)) + logit_term(yc, eta_directed(b0, b1, dist, rj, ri));
}

} // namespace dlsm
