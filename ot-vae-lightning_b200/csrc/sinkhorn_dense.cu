// K8 (dense): log-domain Sinkhorn on a materialised cost matrix, HBM-bound streaming kernels.
// Reference: sinkhorn_log, ot/w2_utils.py:276-319.  Exact recurrences kept (v first, then u; +1e-8 inside the logs;
// stop when the MIN over the batch of sum|du|+sum|dv| < threshold).  One read of C per half-iteration.
#include "sinkhorn_dense.cuh"
#include <cfloat>
#include <algorithm>

namespace otk {

struct SkState { int done; int iters; };

template <typename T> struct SkMath;
template <> struct SkMath<float> {
  static __device__ __forceinline__ float ex(float x) { return __expf(x); }
  static __device__ __forceinline__ float lg(float x) { return logf(x); }
  static __device__ __forceinline__ float ninf() { return -FLT_MAX; }
};
template <> struct SkMath<double> {
  static __device__ __forceinline__ double ex(double x) { return exp(x); }
  static __device__ __forceinline__ double lg(double x) { return log(x); }
  static __device__ __forceinline__ double ninf() { return -DBL_MAX; }
};

// online (max, sum exp) update with one exponential per element
template <typename T>
__device__ __forceinline__ void lse_push(T& m, T& s, T t) {
  if (t > m) { s = s * SkMath<T>::ex(m - t) + T(1); m = t; }
  else s += SkMath<T>::ex(t - m);
}
template <typename T>
__device__ __forceinline__ void lse_merge(T& m, T& s, T m2, T s2) {
  if (m2 > m) { s = s * SkMath<T>::ex(m - m2) + s2; m = m2; }
  else s += s2 * SkMath<T>::ex(m2 - m);
}

template <typename T>
__global__ void sk_prep_kernel(const T* a, const T* b, int64_t LN, int64_t LM, T* log_a, T* log_b, T* u, T* v,
                               SkState* state, bool warm) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e == 0) { state->done = 0; state->iters = 0; }
  for (; e < LN + LM; e += (int64_t)gridDim.x * blockDim.x) {
    if (e < LN) { log_a[e] = SkMath<T>::lg(a[e] + T(1e-8)); if (!warm) u[e] = T(0); }
    else { log_b[e - LN] = SkMath<T>::lg(b[e - LN] + T(1e-8)); if (!warm) v[e - LN] = T(0); }
  }
}

// partial column LSE over a slab of rows:  (pm, ps)[l, slab, j] = online-LSE_i( u_i - C_ij/reg )
constexpr int SKC_COLS = 128, SKC_WARPS = 8;
template <typename T>
__global__ void __launch_bounds__(SKC_WARPS * 32)
sk_col_partial_kernel(const T* __restrict__ C, const T* __restrict__ u, int64_t N, int64_t M, int64_t rows_per_slab,
                      T neg_inv_reg, T* __restrict__ pm, T* __restrict__ ps, const SkState* state) {
  if (state->done) return;
  __shared__ T sm_m[SKC_WARPS][SKC_COLS], sm_s[SKC_WARPS][SKC_COLS];
  const int64_t l = blockIdx.z, slab = blockIdx.y, slabs = gridDim.y;
  const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
  const int64_t j0 = (int64_t)blockIdx.x * SKC_COLS;
  const int64_t r0 = slab * rows_per_slab, r1 = min(N, r0 + rows_per_slab);
  const T* Cl = C + l * N * M;
  const T* ul = u + l * N;
  T m[4], s[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) { m[c] = SkMath<T>::ninf(); s[c] = T(0); }
  for (int64_t i = r0 + warp; i < r1; i += SKC_WARPS) {
    const T ui = ul[i];
    const T* row = Cl + i * M + j0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      int64_t j = j0 + lane + 32 * c;
      if (j < M) lse_push(m[c], s[c], fma(row[lane + 32 * c], neg_inv_reg, ui));
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) { sm_m[warp][lane + 32 * c] = m[c]; sm_s[warp][lane + 32 * c] = s[c]; }
  __syncthreads();
  if (threadIdx.x < SKC_COLS) {
    int cidx = threadIdx.x;
    T mm = sm_m[0][cidx], ss = sm_s[0][cidx];
#pragma unroll
    for (int w = 1; w < SKC_WARPS; ++w) lse_merge(mm, ss, sm_m[w][cidx], sm_s[w][cidx]);
    int64_t j = j0 + cidx;
    if (j < M) { pm[(l * slabs + slab) * M + j] = mm; ps[(l * slabs + slab) * M + j] = ss; }
  }
}

// v_j = log b_j - LSE over slabs ; dv_j = |v_j - v_old|
template <typename T>
__global__ void sk_col_final_kernel(const T* __restrict__ pm, const T* __restrict__ ps, int slabs, int64_t L, int64_t M,
                                    const T* __restrict__ log_b, T* __restrict__ v, T* __restrict__ dv,
                                    const SkState* state) {
  if (state->done) return;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < L * M; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t l = e / M, j = e % M;
    T mm = pm[(l * slabs) * M + j], ss = ps[(l * slabs) * M + j];
    for (int sidx = 1; sidx < slabs; ++sidx) lse_merge(mm, ss, pm[(l * slabs + sidx) * M + j], ps[(l * slabs + sidx) * M + j]);
    T vn = log_b[e] - (mm + SkMath<T>::lg(ss));
    dv[e] = fabs(vn - v[e]);
    v[e] = vn;
  }
}

// u_i = log a_i - LSE_j( v_j - C_ij/reg ) ; TPR threads cooperate on one row
template <typename T, int TPR>
__global__ void __launch_bounds__(256)
sk_row_kernel(const T* __restrict__ C, const T* __restrict__ v, int64_t N, int64_t M, T neg_inv_reg,
              const T* __restrict__ log_a, T* __restrict__ u, T* __restrict__ du, const SkState* state) {
  if (state->done) return;
  constexpr int ROWS = 256 / TPR;
  __shared__ T sm_m[8], sm_s[8];
  const int64_t l = blockIdx.y;
  const int sub = threadIdx.x / TPR, t = threadIdx.x % TPR;
  const int64_t i = (int64_t)blockIdx.x * ROWS + sub;
  T m = SkMath<T>::ninf(), s = T(0);
  if (i < N) {
    const T* row = C + (l * N + i) * M;
    const T* vl = v + l * M;
    for (int64_t j = t; j < M; j += TPR) lse_push(m, s, fma(row[j], neg_inv_reg, vl[j]));
  }
  // warp combine
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    T m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    lse_merge(m, s, m2, s2);
  }
  if (TPR == 256) {
    if (threadIdx.x % 32 == 0) { sm_m[threadIdx.x / 32] = m; sm_s[threadIdx.x / 32] = s; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < 8; ++w) lse_merge(m, s, sm_m[w], sm_s[w]);
    }
  }
  if (t == 0 && i < N) {
    T un = log_a[l * N + i] - (m + SkMath<T>::lg(s));
    du[l * N + i] = fabs(un - u[l * N + i]);
    u[l * N + i] = un;
  }
}

// diff[l] = sum du + sum dv ; then (last block) done = min_l diff < threshold
template <typename T>
__global__ void sk_check_kernel(const T* du, const T* dv, int64_t L, int64_t N, int64_t M, double threshold, double* diff,
                                unsigned* ticket, SkState* state) {
  if (state->done) return;
  __shared__ double red[32];
  __shared__ bool last;
  const int64_t l = blockIdx.x;
  double acc = 0;
  for (int64_t e = threadIdx.x; e < N; e += blockDim.x) acc += (double)du[l * N + e];
  for (int64_t e = threadIdx.x; e < M; e += blockDim.x) acc += (double)dv[l * M + e];
  acc = warp_sum(acc);
  if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0;
    for (int w = 0; w < blockDim.x / 32; ++w) tot += red[w];
    diff[l] = tot;
    __threadfence();
    last = (atomicAdd(ticket, 1u) == (unsigned)(L - 1));
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    double mn = DBL_MAX;
    for (int64_t k = 0; k < L; ++k) mn = fmin(mn, ((volatile double*)diff)[k]);
    *ticket = 0;
    state->iters += 1;
    if (mn < threshold) state->done = 1;
  }
}

template <typename T>
__global__ void sk_plan_kernel(const T* __restrict__ C, const T* __restrict__ u, const T* __restrict__ v, int64_t L,
                               int64_t N, int64_t M, T neg_inv_reg, T* __restrict__ plan) {
  const int64_t total = L * N * M;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t l = e / (N * M), r = e % (N * M), i = r / M, j = r % M;
    plan[e] = SkMath<T>::ex(fma(C[e], neg_inv_reg, u[l * N + i] + v[l * M + j]));
  }
}

static int sk_slabs(int64_t L, int64_t N, int64_t M) {
  int64_t col_blocks = ceil_div(M, SKC_COLS) * L;
  int64_t want = ceil_div((int64_t)sm_count() * 8, col_blocks);
  int64_t max_by_rows = ceil_div(N, 64);  // at least 64 rows per slab
  if (want > max_by_rows) want = max_by_rows;
  if (want < 1) want = 1;
  if (want > 64) want = 64;
  return (int)want;
}

template <typename T>
static int sinkhorn_dense_impl(const T* a, const T* b, const T* C, int64_t L, int64_t N, int64_t M, double reg,
                               int max_iter, double threshold, int poll_every, T* u, T* v, T* plan, int* iters_done_host,
                               void* workspace, size_t workspace_bytes, cudaStream_t st, bool warm = false) {
  const int slabs = sk_slabs(L, N, M);
  Arena ar(workspace, workspace_bytes);
  T* log_a = ar.take<T>((size_t)L * N);
  T* log_b = ar.take<T>((size_t)L * M);
  T* du = ar.take<T>((size_t)L * N);
  T* dv = ar.take<T>((size_t)L * M);
  T* pm = ar.take<T>((size_t)L * slabs * M);
  T* ps = ar.take<T>((size_t)L * slabs * M);
  double* diff = ar.take<double>((size_t)L);
  SkState* state = ar.take<SkState>(1);
  unsigned* ticket = ar.take<unsigned>(1);
  if (!ar.ok()) return OTK_ERR_WORKSPACE;
  OTK_CUDA(cudaMemsetAsync(ticket, 0, sizeof(unsigned), st));
  const T nir = (T)(-1.0 / reg);
  unsigned pg = (unsigned)std::min<int64_t>(ceil_div(L * (N + M), 256), (int64_t)sm_count() * 8);
  sk_prep_kernel<T><<<pg ? pg : 1, 256, 0, st>>>(a, b, L * N, L * M, log_a, log_b, u, v, state, warm);
  OTK_LAUNCH_CHECK();
  const int64_t rows_per_slab = ceil_div(N, slabs);
  dim3 gcol((unsigned)ceil_div(M, SKC_COLS), (unsigned)slabs, (unsigned)L);
  unsigned gfin = (unsigned)std::min<int64_t>(ceil_div(L * M, 256), (int64_t)sm_count() * 8);
  if (poll_every <= 0) poll_every = 16;
  SkState host_state{0, 0};
  for (int it = 0; it < max_iter; ++it) {
    sk_col_partial_kernel<T><<<gcol, SKC_WARPS * 32, 0, st>>>(C, u, N, M, rows_per_slab, nir, pm, ps, state);
    sk_col_final_kernel<T><<<gfin, 256, 0, st>>>(pm, ps, slabs, L, M, log_b, v, dv, state);
    if (M <= 2048) {
      dim3 g((unsigned)ceil_div(N, 8), (unsigned)L);
      sk_row_kernel<T, 32><<<g, 256, 0, st>>>(C, v, N, M, nir, log_a, u, du, state);
    } else {
      dim3 g((unsigned)N, (unsigned)L);
      sk_row_kernel<T, 256><<<g, 256, 0, st>>>(C, v, N, M, nir, log_a, u, du, state);
    }
    sk_check_kernel<T><<<(unsigned)L, 256, 0, st>>>(du, dv, L, N, M, threshold, diff, ticket, state);
    count_launch(3);
    OTK_LAUNCH_CHECK();
    if (threshold > 0 && (it + 1) % poll_every == 0 && it + 1 < max_iter) {
      OTK_CUDA(cudaMemcpyAsync(&host_state, state, sizeof(SkState), cudaMemcpyDeviceToHost, st));
      OTK_CUDA(cudaStreamSynchronize(st));
      if (host_state.done) break;
    }
  }
  if (plan) {
    unsigned g = (unsigned)std::min<int64_t>(ceil_div(L * N * M, 256), (int64_t)sm_count() * 32);
    sk_plan_kernel<T><<<g ? g : 1, 256, 0, st>>>(C, u, v, L, N, M, nir, plan);
    OTK_LAUNCH_CHECK();
  }
  if (iters_done_host) {
    OTK_CUDA(cudaMemcpyAsync(&host_state, state, sizeof(SkState), cudaMemcpyDeviceToHost, st));
    OTK_CUDA(cudaStreamSynchronize(st));
    *iters_done_host = host_state.iters;
  }
  return OTK_OK;
}


int sinkhorn_dense_f32(const float* a, const float* b, const float* C, int64_t L, int64_t N, int64_t M, double reg,
                       int max_iter, double threshold, int poll_every, float* u, float* v, float* plan,
                       int* iters_done_host, void* workspace, size_t workspace_bytes, bool warm_start, cudaStream_t st) {
  return sinkhorn_dense_impl<float>(a, b, C, L, N, M, reg, max_iter, threshold, poll_every, u, v, plan, iters_done_host,
                                    workspace, workspace_bytes, st, warm_start);
}

__global__ void sk_slab_merge_kernel(const float* pm, const float* ps, int slabs, int64_t M, float* col_max, float* col_sum) {
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < M; j += (int64_t)gridDim.x * blockDim.x) {
    float mm = pm[j], ss = ps[j];
    for (int s = 1; s < slabs; ++s) lse_merge(mm, ss, pm[(int64_t)s * M + j], ps[(int64_t)s * M + j]);
    col_max[j] = mm; col_sum[j] = ss;
  }
}
__global__ void sk_zero_state_kernel(SkState* st) { st->done = 0; st->iters = 0; }
__global__ void sk_log_eps_kernel(const float* a, int64_t n, float* out) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
    out[e] = logf(a[e] + 1e-8f);
}
__global__ void sk_sum_abs_kernel(const float* d, int64_t n, float* acc) {
  __shared__ float red[32];
  float a = 0;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) a += d[e];
  a = warp_sum(a);
  if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = a;
  __syncthreads();
  if (threadIdx.x == 0) { float t = 0; for (int w = 0; w < blockDim.x / 32; ++w) t += red[w]; atomicAdd(acc, t); }
}

int dense_col_partial_f32(const float* C, const float* u, int64_t N, int64_t M, double reg, float* col_max, float* col_sum,
                          void* workspace, size_t workspace_bytes, cudaStream_t st) {
  const int slabs = sk_slabs(1, N, M);
  Arena ar(workspace, workspace_bytes);
  float* pm = ar.take<float>((size_t)slabs * M);
  float* ps = ar.take<float>((size_t)slabs * M);
  SkState* state = ar.take<SkState>(1);
  if (!ar.ok()) return OTK_ERR_WORKSPACE;
  sk_zero_state_kernel<<<1, 1, 0, st>>>(state);
  dim3 gcol((unsigned)ceil_div(M, SKC_COLS), (unsigned)slabs, 1);
  sk_col_partial_kernel<float><<<gcol, SKC_WARPS * 32, 0, st>>>(C, u, N, M, ceil_div(N, slabs), (float)(-1.0 / reg), pm, ps, state);
  sk_slab_merge_kernel<<<(unsigned)ceil_div(M, 256), 256, 0, st>>>(pm, ps, slabs, M, col_max, col_sum);
  count_launch(2);
  OTK_LAUNCH_CHECK();
  return OTK_OK;
}

int dense_row_step_f32(const float* C, const float* v, int64_t N, int64_t M, double reg, const float* a, float* u,
                       float* diff, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  Arena ar(workspace, workspace_bytes);
  float* log_a = ar.take<float>((size_t)N);
  float* du = ar.take<float>((size_t)N);
  SkState* state = ar.take<SkState>(1);
  if (!ar.ok()) return OTK_ERR_WORKSPACE;
  sk_zero_state_kernel<<<1, 1, 0, st>>>(state);
  unsigned g1 = (unsigned)std::min<int64_t>(ceil_div(N, 256), (int64_t)sm_count() * 8);
  sk_log_eps_kernel<<<g1, 256, 0, st>>>(a, N, log_a);
  const float nir = (float)(-1.0 / reg);
  if (M <= 2048) {
    dim3 g((unsigned)ceil_div(N, 8), 1);
    sk_row_kernel<float, 32><<<g, 256, 0, st>>>(C, v, N, M, nir, log_a, u, du, state);
  } else {
    dim3 g((unsigned)N, 1);
    sk_row_kernel<float, 256><<<g, 256, 0, st>>>(C, v, N, M, nir, log_a, u, du, state);
  }
  if (diff) sk_sum_abs_kernel<<<g1, 256, 0, st>>>(du, N, diff);
  count_launch(diff ? 3 : 2);
  OTK_LAUNCH_CHECK();
  return OTK_OK;
}
}  // namespace otk
using namespace otk;

extern "C" size_t otk_sinkhorn_dense_workspace_bytes(int64_t L, int64_t N, int64_t M) {
  size_t per = 8;  // sized for fp64
  return 2 * (align_up((size_t)L * N * per, 256) + align_up((size_t)L * M * per, 256)) +
         2 * align_up((size_t)L * 64 * M * per, 256) + align_up((size_t)L * 8, 256) + 2048;
}

extern "C" int otk_sinkhorn_dense(const void* a, const void* b, const void* C, int64_t L, int64_t N, int64_t M, int dtype,
                                  double reg, int max_iter, double threshold, int poll_every, void* u, void* v,
                                  void* plan, int* iters_done_host, void* workspace, size_t workspace_bytes,
                                  otk_stream_t stream) {
  OTK_TRY(require_device());
  OTK_REQUIRE(a && b && C && u && v && L > 0 && N > 0 && M > 0, "sinkhorn_dense: bad arguments");
  OTK_REQUIRE(reg > 0 && max_iter >= 0, "sinkhorn_dense: reg must be > 0");
  OTK_REQUIRE(L <= 65535, "sinkhorn_dense: batch too large");
  if (!workspace || workspace_bytes < otk_sinkhorn_dense_workspace_bytes(L, N, M)) return OTK_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  if (dtype == OTK_F64)
    return sinkhorn_dense_impl<double>((const double*)a, (const double*)b, (const double*)C, L, N, M, reg, max_iter,
                                       threshold, poll_every, (double*)u, (double*)v, (double*)plan, iters_done_host,
                                       workspace, workspace_bytes, st);
  return sinkhorn_dense_impl<float>((const float*)a, (const float*)b, (const float*)C, L, N, M, reg, max_iter, threshold,
                                    poll_every, (float*)u, (float*)v, (float*)plan, iters_done_host, workspace,
                                    workspace_bytes, st);
}
