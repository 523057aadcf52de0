// internal entry of the dense Sinkhorn loop (sinkhorn_dense.cu), shared with the point-cloud front end
#pragma once
#include "otk_common.cuh"
namespace otk {
int sinkhorn_dense_f32(const float* a, const float* b, const float* C, int64_t L, int64_t N, int64_t M, double reg,
                       int max_iter, double threshold, int poll_every, float* u, float* v, float* plan,
                       int* iters_done_host, void* workspace, size_t workspace_bytes, bool warm_start, cudaStream_t st);
// single half-steps on a materialised slab (used by the row-sharded path)
int dense_col_partial_f32(const float* C, const float* u, int64_t N, int64_t M, double reg, float* col_max, float* col_sum,
                          void* workspace, size_t workspace_bytes, cudaStream_t st);
int dense_row_step_f32(const float* C, const float* v, int64_t N, int64_t M, double reg, const float* a, float* u,
                       float* diff, void* workspace, size_t workspace_bytes, cudaStream_t st);
}  // namespace otk
