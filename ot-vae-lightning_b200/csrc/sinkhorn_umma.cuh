// fused tcgen05 Sinkhorn engine hooks (sinkhorn_umma.cu)
#pragma once
#include "otk_common.cuh"
namespace otk {
bool sk_umma_eligible(int64_t N, int64_t M, int64_t dim, int cost_kind);
size_t sk_umma_workspace_bytes(int64_t N, int64_t M, int64_t dim);
int sk_umma_cost_max(const float* x, const float* y, int64_t N, int64_t M, int64_t dim, float* out, void* workspace,
                     size_t workspace_bytes, cudaStream_t st);
int sk_umma_solve(const float* x, const float* y, int64_t N, int64_t M, int64_t dim, const float* a, const float* b,
                  double scale, int scale_inv_max, double reg, int max_iter, double threshold, int poll_every, int precision,
                  float* u, float* v, double* summary, float* row_marginal, float* col_marginal, int* iters_done_host,
                  void* workspace, size_t workspace_bytes, cudaStream_t st);
int sk_umma_colstep(const float* x_local, const float* y, int64_t n_local, int64_t M, int64_t dim, const float* u_local,
                    double scale, double reg, int reuse_prepared, float* col_max, float* col_sum, void* workspace,
                    size_t workspace_bytes, cudaStream_t st);
int sk_umma_rowstep(const float* x_local, const float* y, int64_t n_local, int64_t M, int64_t dim, const float* a_local,
                    const float* v, double scale, double reg, int reuse_prepared, float* u_local, float* diff, void* workspace,
                    size_t workspace_bytes, cudaStream_t st);
int sk_umma_summary(const float* x_local, const float* y, int64_t n_local, int64_t M, int64_t dim, const float* a_local,
                    const float* b, const float* u_local, const float* v, double scale, double reg, int reuse_prepared,
                    double* part, float* row_marginal, float* col_partial, void* workspace, size_t workspace_bytes,
                    cudaStream_t st);
size_t sk_umma_exchange_bytes(int world, int64_t M);
int sk_umma_colstep_push(const float* x_local, const float* y, int64_t n_local, int64_t M, int64_t dim, const float* u_local,
                         double scale, double reg, int reuse_prepared, void* const* peers_dev, int world, int rank, int* ctrl,
                         void* workspace, size_t workspace_bytes, cudaStream_t st);
int sk_combine_wait(void* xchg, int world, int64_t M, const float* b, float* v, float* diff, int* ctrl, cudaStream_t st);
int sk_umma_sharded_step(const float* x_local, const float* y, int64_t n_local, int64_t M, int64_t dim, const float* a_local,
                         const float* b, float* u_local, float* v, double scale, double reg, int stage, void* const* peers_dev,
                         int world, int rank, void* xchg_local, int* ctrl, float* diffs, void* workspace, size_t workspace_bytes,
                         cudaStream_t st);
}  // namespace otk
