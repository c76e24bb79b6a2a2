// dlsm_graph.h -- device-side construction of the case-control edge lists (dlsm_graph.cu)
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace dlsm {
// d_edges: device (E, 3) int32 rows (t, sender, receiver).  On success the three device arrays are
// allocated here and handed to the caller; *status: 0 ok, 1 = an index out of range or a self tie,
// 2 = a tie listed twice.
cudaError_t graph_build_edge_lists(const int32_t *d_edges, size_t E, int T, int n, int32_t **deg_out,
                                   int32_t **in_out, int *max_in, int32_t **out_out, int *max_out,
                                   int *status, cudaStream_t stream);
} // namespace dlsm
