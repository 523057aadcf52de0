// K8/K9 (point clouds): log-domain Sinkhorn with the cost recomputed from the points.
// Reference pair: cost producers (ot/w2_utils.py:121-125 squared distance; CodebookModel.energy,
// ot/distribution_models/codebook_model.py:155-160) + sinkhorn_log (ot/w2_utils.py:276-319).
//
// Two engines behind the same entry points:
//   * fused tcgen05 engine (sinkhorn_umma.cu): cost tiles on the tensor cores, online LSE out of TMEM, no N x M
//     matrix in HBM - used when sk_umma_eligible();
//   * streaming engine: the cost slab is materialised in the caller's workspace and the dense HBM-bound kernels run
//     on it (any dim / cost kind).
#include "sinkhorn_dense.cuh"
#include "sinkhorn_umma.cuh"
#include "gemm.cuh"
#include <cfloat>

namespace otk {

__global__ void row_sqnorm_kernel(const float* __restrict__ x, int64_t n, int64_t d, float* __restrict__ out) {
  const int lane = threadIdx.x % 32;
  const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / 32;
  if (row >= n) return;
  float acc = 0;
  for (int64_t k = lane; k < d; k += 32) { float v = x[row * d + k]; acc = fmaf(v, v, acc); }
  acc = warp_sum(acc);
  if (lane == 0) out[row] = acc;
}

__device__ __forceinline__ float cost_from_dot(float dot, float nx, float ny, int kind) {
  float sq = fmaxf(nx + ny - 2.f * dot, 0.f);
  return kind == OTK_COST_SQEUCLIDEAN ? sq : 1.f / (sqrtf(sq) + 1e-8f);
}

// 64x64 tile of cost(x_i, y_j); MODE 0: write scale*cost to C ; MODE 1: max-reduce into *out_max
constexpr int CT = 64, CT_BK = 16;
template <int MODE>
__global__ void __launch_bounds__(256)
cost_tile_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ nx,
                 const float* __restrict__ ny, int64_t N, int64_t M, int64_t d, int kind, const float* scale_dev,
                 float scale_host, float* __restrict__ C, float* out_max) {
  __shared__ float As[CT_BK][CT + 4], Bs[CT_BK][CT + 4];
  const int64_t i0 = (int64_t)blockIdx.y * CT, j0 = (int64_t)blockIdx.x * CT;
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  float acc[4][4] = {};
  for (int64_t k0 = 0; k0 < d; k0 += CT_BK) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      int e = tid + r * 256, kk = e % CT_BK, mm = e / CT_BK;
      int64_t k = k0 + kk;
      As[kk][mm] = (i0 + mm < N && k < d) ? x[(i0 + mm) * d + k] : 0.f;
      Bs[kk][mm] = (j0 + mm < M && k < d) ? y[(j0 + mm) * d + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < CT_BK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  const float sc = scale_dev ? *scale_dev : scale_host;
  float mx = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t gi = i0 + ty * 4 + i;
    if (gi >= N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int64_t gj = j0 + tx * 4 + j;
      if (gj >= M) continue;
      float c = cost_from_dot(acc[i][j], nx[gi], ny[gj], kind);
      if (MODE == 0) C[gi * M + gj] = c * sc;
      else mx = fmaxf(mx, c);
    }
  }
  if (MODE == 1) {
    mx = warp_max(mx);
    if (tid % 32 == 0) atomicMax(reinterpret_cast<unsigned*>(out_max), __float_as_uint(mx));  // costs are >= 0
  }
}

__global__ void inv_kernel(const float* in, float* out) { *out = 1.f / *in; }
__global__ void set_kernel(float* out, float v) { *out = v; }

// summary of the plan pi = exp(u_i + v_j - C_ij/reg) without storing it:
//   rows: row sums (-> max |row - (a+1e-8)| ... compared by the caller against a), <C,pi>, total mass
__global__ void __launch_bounds__(256)
plan_row_summary_kernel(const float* __restrict__ C, const float* __restrict__ u, const float* __restrict__ v,
                        const float* __restrict__ a, int64_t N, int64_t M, float nir, double* summary,
                        float* __restrict__ colsum, float* __restrict__ rowsum_out) {
  // one block per row; also accumulates column sums with atomics (M floats)
  __shared__ double red_c[8], red_m[8];
  const int64_t i = blockIdx.x;
  const float ui = u[i];
  double cost = 0, mass = 0;
  for (int64_t j = threadIdx.x; j < M; j += 256) {
    float c = C[i * M + j];
    float p = __expf(fmaf(c, nir, ui + v[j]));
    cost += (double)c * p;
    mass += p;
    atomicAdd(&colsum[j], p);
  }
  cost = warp_sum(cost); mass = warp_sum(mass);
  if (threadIdx.x % 32 == 0) { red_c[threadIdx.x / 32] = cost; red_m[threadIdx.x / 32] = mass; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tc = 0, tm = 0;
    for (int w = 0; w < 8; ++w) { tc += red_c[w]; tm += red_m[w]; }
    atomicAdd(&summary[0], tc);
    atomicAdd(&summary[1], tm);
    if (rowsum_out) rowsum_out[i] = (float)tm;
    double err = fabs(tm - (double)a[i]);
    // max via atomicMax on the bit pattern of a non-negative double
    atomicMax(reinterpret_cast<unsigned long long*>(&summary[2]), (unsigned long long)__double_as_longlong(err));
  }
}
__global__ void plan_col_err_kernel(const float* colsum, const float* b, int64_t M, double* summary) {
  double mx = 0;
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < M; j += (int64_t)gridDim.x * blockDim.x)
    mx = fmax(mx, fabs((double)colsum[j] - (double)b[j]));
  mx = warp_max(mx);
  if (threadIdx.x % 32 == 0)
    atomicMax(reinterpret_cast<unsigned long long*>(&summary[3]), (unsigned long long)__double_as_longlong(mx));
}

// C_ij (holding -2 x_i.y_j + |y_j|^2 from the tensor-core product) -> scale * cost
__global__ void cost_from_gram_kernel(float* __restrict__ C, const float* __restrict__ nx, int64_t N, int64_t M, int kind,
                                      const float* scale_dev, float scale_host) {
  const float sc = scale_dev ? *scale_dev : scale_host;
  const int64_t total = N * M;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const float sq = fmaxf(C[e] + nx[e / M], 0.f);
    C[e] = (kind == OTK_COST_SQEUCLIDEAN ? sq : 1.f / (sqrtf(sq) + 1e-8f)) * sc;
  }
}

// scratch floats the tensor-core cost producer needs (TF32 hi/lo planes of both clouds)
static size_t cost_umma_scratch_floats(int64_t N, int64_t M, int64_t d) { return (size_t)2 * (N + M) * d; }
static bool cost_umma_eligible(const float* x, const float* y, const float* C, int64_t N, int64_t M, int64_t d) {
  return N >= 128 && M >= 128 && d >= 8 && d % 4 == 0 && M % 4 == 0 && N * M >= (1 << 20) &&
         reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(y) % 16 == 0 &&
         reinterpret_cast<uintptr_t>(C) % 16 == 0;
}

// The contraction of the cost (K9: x.y^T of `cdist` / CodebookModel.energy) on tcgen05 (3xTF32, fp32-accurate): the product
// kernel's epilogue emits -2 x_i.y_j + |y_j|^2, one elementwise pass finishes the cost.  `scratch` (cost_umma_scratch_floats)
// may be null: small or unaligned problems, or a caller without the scratch, take the FFMA tile kernel.
static int build_cost(const float* x, const float* y, int64_t N, int64_t M, int64_t d, int kind, const float* scale_dev,
                      float scale_host, float* nx, float* ny, float* C, cudaStream_t st, float* scratch = nullptr) {
  row_sqnorm_kernel<<<(unsigned)ceil_div(N * 32, 256), 256, 0, st>>>(x, N, d, nx);
  row_sqnorm_kernel<<<(unsigned)ceil_div(M * 32, 256), 256, 0, st>>>(y, M, d, ny);
  if (scratch && cost_umma_eligible(x, y, C, N, M, d)) {
    GemmArgs<float> g = nt_args(x, y, C, N, M, d, d, d, M, 0, 0, 0, -2.f, 0.f);
    g.bias = ny;
    g.scratch = scratch;
    const int r = gemm_umma_try(g, 1, 3, st);
    if (r < 0) return r;
    if (r == 1) {
      int64_t blocks = ceil_div(N * M, 256 * 4);
      if (blocks > (int64_t)sm_count() * 16) blocks = (int64_t)sm_count() * 16;
      cost_from_gram_kernel<<<(unsigned)blocks, 256, 0, st>>>(C, nx, N, M, kind, scale_dev, scale_host);
      count_launch(2);
      OTK_LAUNCH_CHECK();
      return OTK_OK;
    }
  }
  dim3 grid((unsigned)ceil_div(M, CT), (unsigned)ceil_div(N, CT));
  cost_tile_kernel<0><<<grid, 256, 0, st>>>(x, y, nx, ny, N, M, d, kind, scale_dev, scale_host, C, nullptr);
  count_launch(2);
  OTK_LAUNCH_CHECK();
  return OTK_OK;
}

static size_t stream_ws_bytes(int64_t N, int64_t M) {
  return align_up((size_t)N * M * 4, 256) + 2 * align_up((size_t)N * 4, 256) + 3 * align_up((size_t)M * 4, 256) +
         otk_sinkhorn_dense_workspace_bytes(1, N, M) + 4096;
}

}  // namespace otk
using namespace otk;

// precision 0 = automatic: the fused tcgen05 engine (FP16 operand planes) when the shape is eligible, else the streaming
// engine; precision 1 = exact fp32 cost tiles (streaming engine) whatever the shape - the small codebook / mixture problems
// of DiscreteTransport and batch_ot_gmm, whose plans are compared entry by entry.
static bool use_fused(int64_t N, int64_t M, int64_t dim, int cost_kind, int precision) {
  return precision == 0 && sk_umma_eligible(N, M, dim, cost_kind);
}

extern "C" size_t otk_sinkhorn_points_workspace_bytes(int64_t N, int64_t M, int64_t dim, int cost_kind) {
  // sized for either engine, so that the same workspace serves every `precision`
  const size_t a = sk_umma_eligible(N, M, dim, cost_kind) ? sk_umma_workspace_bytes(N, M, dim) : 0;
  const size_t b = (double)N * (double)M <= 3.0e9 ? stream_ws_bytes(N, M) : 0;
  return a > b ? a : (b ? b : stream_ws_bytes(N, M));
}

extern "C" int otk_cost_matrix(const float* x, const float* y, int64_t N, int64_t M, int64_t dim, int cost_kind,
                               double scale, float* C, void* workspace, size_t workspace_bytes, otk_stream_t stream) {
  OTK_TRY(require_device());
  OTK_REQUIRE(x && y && C && N > 0 && M > 0 && dim > 0, "cost_matrix: bad arguments");
  if (!workspace || workspace_bytes < (size_t)(N + M) * 4 + 512) return OTK_ERR_WORKSPACE;
  Arena ar(workspace, workspace_bytes);
  float* nx = ar.take<float>((size_t)N);
  float* ny = ar.take<float>((size_t)M);
  // with the larger workspace of otk_cost_workspace_bytes the contraction runs on the tensor cores
  float* scratch = ar.take<float>(cost_umma_scratch_floats(N, M, dim));
  if (!ar.ok()) scratch = nullptr;
  return build_cost(x, y, N, M, dim, cost_kind, nullptr, (float)scale, nx, ny, C, as_stream(stream), scratch);
}

extern "C" size_t otk_cost_workspace_bytes(int64_t N, int64_t M, int64_t dim) {
  return 2 * align_up((size_t)(N > M ? N : M) * 4, 256) + align_up(cost_umma_scratch_floats(N, M, dim) * 4, 256) + 1024;
}

extern "C" int otk_cost_max(const float* x, const float* y, int64_t N, int64_t M, int64_t dim, int cost_kind, float* out,
                            void* workspace, size_t workspace_bytes, otk_stream_t stream) {
  OTK_TRY(require_device());
  OTK_REQUIRE(x && y && out && N > 0 && M > 0 && dim > 0, "cost_max: bad arguments");
  if (!workspace || workspace_bytes < (size_t)(N + M) * 4 + 512) return OTK_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  if (sk_umma_eligible(N, M, dim, cost_kind) && workspace_bytes >= sk_umma_workspace_bytes(N, M, dim))
    return sk_umma_cost_max(x, y, N, M, dim, out, workspace, workspace_bytes, st);
  Arena ar(workspace, workspace_bytes);
  float* nx = ar.take<float>((size_t)N);
  float* ny = ar.take<float>((size_t)M);
  row_sqnorm_kernel<<<(unsigned)ceil_div(N * 32, 256), 256, 0, st>>>(x, N, dim, nx);
  row_sqnorm_kernel<<<(unsigned)ceil_div(M * 32, 256), 256, 0, st>>>(y, M, dim, ny);
  set_kernel<<<1, 1, 0, st>>>(out, 0.f);
  dim3 grid((unsigned)ceil_div(M, CT), (unsigned)ceil_div(N, CT));
  cost_tile_kernel<1><<<grid, 256, 0, st>>>(x, y, nx, ny, N, M, dim, cost_kind, nullptr, 1.f, nullptr, out);
  count_launch(3);
  OTK_LAUNCH_CHECK();
  return OTK_OK;
}

extern "C" int otk_sinkhorn_points(const float* x, const float* y, int64_t N, int64_t M, int64_t dim, const float* a,
                                   const float* b, int cost_kind, double scale, int scale_inv_max, double reg,
                                   int max_iter, double threshold, int poll_every, int precision, float* u, float* v,
                                   double* summary, float* row_marginal, float* col_marginal, int* iters_done_host,
                                   void* workspace, size_t workspace_bytes, otk_stream_t stream) {
  OTK_TRY(require_device());
  OTK_REQUIRE(x && y && a && b && u && v && N > 0 && M > 0 && dim > 0, "sinkhorn_points: bad arguments");
  OTK_REQUIRE(reg > 0 && max_iter >= 0, "sinkhorn_points: reg must be > 0");
  if (!workspace || workspace_bytes < otk_sinkhorn_points_workspace_bytes(N, M, dim, cost_kind)) return OTK_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  if (use_fused(N, M, dim, cost_kind, precision))
    return sk_umma_solve(x, y, N, M, dim, a, b, scale, scale_inv_max, reg, max_iter, threshold, poll_every, precision, u,
                         v, summary, row_marginal, col_marginal, iters_done_host, workspace, workspace_bytes, st);
  // streaming engine
  Arena ar(workspace, workspace_bytes);
  float* C = ar.take<float>((size_t)N * M);
  float* nx = ar.take<float>((size_t)N);
  float* ny = ar.take<float>((size_t)M);
  float* sc = ar.take<float>(64);
  float* colsum = ar.take<float>((size_t)M);
  size_t used = align_up(ar.off, 256);
  const float* scale_dev = nullptr;
  if (scale_inv_max) {
    row_sqnorm_kernel<<<(unsigned)ceil_div(N * 32, 256), 256, 0, st>>>(x, N, dim, nx);
    row_sqnorm_kernel<<<(unsigned)ceil_div(M * 32, 256), 256, 0, st>>>(y, M, dim, ny);
    set_kernel<<<1, 1, 0, st>>>(sc + 1, 0.f);
    dim3 grid((unsigned)ceil_div(M, CT), (unsigned)ceil_div(N, CT));
    cost_tile_kernel<1><<<grid, 256, 0, st>>>(x, y, nx, ny, N, M, dim, cost_kind, nullptr, 1.f, nullptr, sc + 1);
    inv_kernel<<<1, 1, 0, st>>>(sc + 1, sc);
    count_launch(5);
    scale_dev = sc;
  }
  OTK_TRY(build_cost(x, y, N, M, dim, cost_kind, scale_dev, (float)scale, nx, ny, C, st));
  OTK_TRY(sinkhorn_dense_f32(a, b, C, 1, N, M, reg, max_iter, threshold, poll_every, u, v, nullptr, iters_done_host,
                             (char*)workspace + used, workspace_bytes - used, false, st));
  if (summary) {
    OTK_CUDA(cudaMemsetAsync(summary, 0, 4 * sizeof(double), st));
    OTK_CUDA(cudaMemsetAsync(colsum, 0, (size_t)M * 4, st));
    plan_row_summary_kernel<<<(unsigned)N, 256, 0, st>>>(C, u, v, a, N, M, (float)(-1.0 / reg), summary, colsum, row_marginal);
    plan_col_err_kernel<<<(unsigned)ceil_div(M, 256), 256, 0, st>>>(colsum, b, M, summary);
    if (col_marginal) OTK_CUDA(cudaMemcpyAsync(col_marginal, colsum, (size_t)M * 4, cudaMemcpyDeviceToDevice, st));
    count_launch(1);
    OTK_LAUNCH_CHECK();
  }
  return OTK_OK;
}

extern "C" int otk_sinkhorn_points_colstep(const float* x_local, const float* y, int64_t n_local, int64_t M, int64_t dim,
                                           const float* u_local, int cost_kind, double scale, double reg, int precision,
                                           int reuse_prepared, float* col_max, float* col_sum, void* workspace,
                                           size_t workspace_bytes, otk_stream_t stream) {
  OTK_TRY(require_device());
  OTK_REQUIRE(x_local && y && u_local && col_max && col_sum && n_local > 0 && M > 0 && dim > 0, "colstep: bad arguments");
  if (!workspace || workspace_bytes < otk_sinkhorn_points_workspace_bytes(n_local, M, dim, cost_kind)) return OTK_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  if (use_fused(n_local, M, dim, cost_kind, precision))
    return sk_umma_colstep(x_local, y, n_local, M, dim, u_local, scale, reg, reuse_prepared, col_max, col_sum, workspace,
                           workspace_bytes, st);
  Arena ar(workspace, workspace_bytes);
  float* C = ar.take<float>((size_t)n_local * M);
  float* nx = ar.take<float>((size_t)n_local);
  float* ny = ar.take<float>((size_t)M);
  size_t used = align_up(ar.off, 256);
  OTK_TRY(build_cost(x_local, y, n_local, M, dim, cost_kind, nullptr, (float)scale, nx, ny, C, st));
  return dense_col_partial_f32(C, u_local, n_local, M, reg, col_max, col_sum, (char*)workspace + used,
                               workspace_bytes - used, st);
}

namespace otk {
__global__ void lse_combine_kernel(const float* pm, const float* ps, int64_t parts, int64_t stride, int64_t M, const float* b,
                                   float* v, float* diff) {
  __shared__ float red[32];
  float acc = 0;
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < M; j += (int64_t)gridDim.x * blockDim.x) {
    float mm = pm[j], ss = ps[j];
    for (int64_t p = 1; p < parts; ++p) {
      float m2 = pm[p * stride + j], s2 = ps[p * stride + j];
      if (m2 > mm) { ss = ss * __expf(mm - m2) + s2; mm = m2; } else ss += s2 * __expf(m2 - mm);
    }
    float vn = logf(b[j] + 1e-8f) - (mm + logf(ss));
    acc += fabsf(vn - v[j]);
    v[j] = vn;
  }
  acc = warp_sum(acc);
  if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = acc;
  __syncthreads();
  if (threadIdx.x == 0 && diff) { float t = 0; for (int w = 0; w < blockDim.x / 32; ++w) t += red[w]; atomicAdd(diff, t); }
}
}  // namespace otk

extern "C" int otk_lse_combine(const float* part_max, const float* part_sum, int64_t parts, int64_t part_stride, int64_t M,
                               const float* b, float* v, float* diff, otk_stream_t stream) {
  OTK_TRY(require_device());
  OTK_REQUIRE(part_max && part_sum && b && v && parts > 0 && M > 0 && part_stride >= M, "lse_combine: bad arguments");
  int64_t blocks = ceil_div(M, 256);
  if (blocks > (int64_t)sm_count() * 4) blocks = (int64_t)sm_count() * 4;
  lse_combine_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(part_max, part_sum, parts, part_stride, M, b, v, diff);
  OTK_LAUNCH_CHECK();
  return OTK_OK;
}

extern "C" int otk_sinkhorn_points_rowstep(const float* x_local, const float* y, int64_t n_local, int64_t M, int64_t dim,
                                           const float* a_local, const float* v, int cost_kind, double scale, double reg,
                                           int precision, int reuse_prepared, float* u_local, float* diff, void* workspace,
                                           size_t workspace_bytes, otk_stream_t stream) {
  OTK_TRY(require_device());
  OTK_REQUIRE(x_local && y && a_local && v && u_local && n_local > 0 && M > 0 && dim > 0, "rowstep: bad arguments");
  if (!workspace || workspace_bytes < otk_sinkhorn_points_workspace_bytes(n_local, M, dim, cost_kind)) return OTK_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  if (use_fused(n_local, M, dim, cost_kind, precision))
    return sk_umma_rowstep(x_local, y, n_local, M, dim, a_local, v, scale, reg, reuse_prepared, u_local, diff, workspace,
                           workspace_bytes, st);
  Arena ar(workspace, workspace_bytes);
  float* C = ar.take<float>((size_t)n_local * M);
  float* nx = ar.take<float>((size_t)n_local);
  float* ny = ar.take<float>((size_t)M);
  size_t used = align_up(ar.off, 256);
  OTK_TRY(build_cost(x_local, y, n_local, M, dim, cost_kind, nullptr, (float)scale, nx, ny, C, st));
  return dense_row_step_f32(C, v, n_local, M, reg, a_local, u_local, diff, (char*)workspace + used, workspace_bytes - used,
                            st);
}

extern "C" int otk_sinkhorn_points_summary(const float* x_local, const float* y, int64_t n_local, int64_t M, int64_t dim,
                                           const float* a_local, const float* b, const float* u_local, const float* v,
                                           int cost_kind, double scale, double reg, int precision, int reuse_prepared,
                                           double* part, float* row_marginal, float* col_partial, void* workspace,
                                           size_t workspace_bytes, otk_stream_t stream) {
  OTK_TRY(require_device());
  OTK_REQUIRE(x_local && y && a_local && b && u_local && v && part && col_partial && n_local > 0 && M > 0 && dim > 0,
              "sinkhorn_points_summary: bad arguments");
  if (!workspace || workspace_bytes < otk_sinkhorn_points_workspace_bytes(n_local, M, dim, cost_kind)) return OTK_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  if (use_fused(n_local, M, dim, cost_kind, precision))
    return sk_umma_summary(x_local, y, n_local, M, dim, a_local, b, u_local, v, scale, reg, reuse_prepared, part,
                           row_marginal, col_partial, workspace, workspace_bytes, st);
  Arena ar(workspace, workspace_bytes);
  float* C = ar.take<float>((size_t)n_local * M);
  float* nx = ar.take<float>((size_t)n_local);
  float* ny = ar.take<float>((size_t)M);
  OTK_TRY(build_cost(x_local, y, n_local, M, dim, cost_kind, nullptr, (float)scale, nx, ny, C, st));
  OTK_CUDA(cudaMemsetAsync(part, 0, 4 * sizeof(double), st));
  OTK_CUDA(cudaMemsetAsync(col_partial, 0, (size_t)M * 4, st));
  plan_row_summary_kernel<<<(unsigned)n_local, 256, 0, st>>>(C, u_local, v, a_local, n_local, M, (float)(-1.0 / reg), part,
                                                           col_partial, row_marginal);
  plan_col_err_kernel<<<(unsigned)ceil_div(M, 256), 256, 0, st>>>(col_partial, b, M, part);
  count_launch(1);
  OTK_LAUNCH_CHECK();
  return OTK_OK;
}

namespace otk {
// in place: C_ij (already scale * cost) -> exp(u_i + v_j - C_ij / reg)
__global__ void plan_from_cost_kernel(float* __restrict__ C, const float* __restrict__ u, const float* __restrict__ v, int64_t N,
                                      int64_t M, float nir) {
  const int64_t total = N * M;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x)
    C[e] = __expf(fmaf(C[e], nir, u[e / M] + v[e % M]));
}
}  // namespace otk

extern "C" int otk_sinkhorn_points_plan(const float* x, const float* y, int64_t N, int64_t M, int64_t dim, const float* u,
                                        const float* v, int cost_kind, double scale, double reg, float* plan, void* workspace,
                                        size_t workspace_bytes, otk_stream_t stream) {
  OTK_TRY(require_device());
  OTK_REQUIRE(x && y && u && v && plan && N > 0 && M > 0 && dim > 0 && reg > 0, "sinkhorn_points_plan: bad arguments");
  if (!workspace || workspace_bytes < (size_t)(N + M) * 4 + 512) return OTK_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  Arena ar(workspace, workspace_bytes);
  float* nx = ar.take<float>((size_t)N);
  float* ny = ar.take<float>((size_t)M);
  OTK_TRY(build_cost(x, y, N, M, dim, cost_kind, nullptr, (float)scale, nx, ny, plan, st));
  int64_t blocks = ceil_div(N * M, 256);
  if (blocks > (int64_t)sm_count() * 16) blocks = (int64_t)sm_count() * 16;
  plan_from_cost_kernel<<<(unsigned)blocks, 256, 0, st>>>(plan, u, v, N, M, (float)(-1.0 / reg));
  OTK_LAUNCH_CHECK();
  return OTK_OK;
}

// ---- peer-memory exchange entry points (fused tcgen05 engine only; see sinkhorn_umma.cu) -----------------------------
extern "C" size_t otk_sinkhorn_exchange_bytes(int world, int64_t M) { return sk_umma_exchange_bytes(world, M); }

extern "C" int otk_sinkhorn_points_colstep_push(const float* x_local, const float* y, int64_t n_local, int64_t M, int64_t dim,
                                                const float* u_local, int cost_kind, double scale, double reg, int precision,
                                                int reuse_prepared, void* const* peer_buffers_dev, int world, int rank,
                                                int* ctrl, void* workspace, size_t workspace_bytes, otk_stream_t stream) {
  OTK_TRY(require_device());
  OTK_REQUIRE(x_local && y && u_local && peer_buffers_dev && ctrl && n_local > 0 && M > 0 && dim > 0 && world >= 1 &&
                  world <= 32 && rank >= 0 && rank < world, "colstep_push: bad arguments");
  OTK_REQUIRE(use_fused(n_local, M, dim, cost_kind, precision), "colstep_push: shape / cost not eligible for the fused engine");
  if (!workspace || workspace_bytes < otk_sinkhorn_points_workspace_bytes(n_local, M, dim, cost_kind)) return OTK_ERR_WORKSPACE;
  return sk_umma_colstep_push(x_local, y, n_local, M, dim, u_local, scale, reg, reuse_prepared, peer_buffers_dev, world, rank,
                              ctrl, workspace, workspace_bytes, as_stream(stream));
}

extern "C" int otk_lse_combine_wait(void* exchange_local, int world, int64_t M, const float* b, float* v, float* diff, int* ctrl,
                                    otk_stream_t stream) {
  OTK_TRY(require_device());
  OTK_REQUIRE(exchange_local && b && v && ctrl && world >= 1 && world <= 32 && M > 0, "lse_combine_wait: bad arguments");
  return sk_combine_wait(exchange_local, world, M, b, v, diff, ctrl, as_stream(stream));
}

extern "C" int otk_sinkhorn_points_sharded_step(const float* x_local, const float* y, int64_t n_local, int64_t M, int64_t dim,
                                                const float* a_local, const float* b, float* u_local, float* v, int cost_kind,
                                                double scale, double reg, int precision, int stage,
                                                void* const* peer_buffers_dev, int world, int rank, void* exchange_local,
                                                int* ctrl, float* diffs, void* workspace, size_t workspace_bytes,
                                                otk_stream_t stream) {
  OTK_TRY(require_device());
  OTK_REQUIRE(x_local && y && a_local && b && u_local && v && peer_buffers_dev && exchange_local && ctrl && n_local > 0 &&
                  M > 0 && dim > 0 && world >= 1 && world <= 32 && rank >= 0 && rank < world && stage >= 0,
              "sharded_step: bad arguments");
  OTK_REQUIRE(use_fused(n_local, M, dim, cost_kind, precision), "sharded_step: shape / cost not eligible for the fused engine");
  if (!workspace || workspace_bytes < otk_sinkhorn_points_workspace_bytes(n_local, M, dim, cost_kind)) return OTK_ERR_WORKSPACE;
  return sk_umma_sharded_step(x_local, y, n_local, M, dim, a_local, b, u_local, v, scale, reg, stage, peer_buffers_dev, world,
                              rank, exchange_local, ctrl, diffs, workspace, workspace_bytes, as_stream(stream));
}
