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

// ---- vectorised streaming kernels -------------------------------------------------------------------------------
// Both half-steps read C exactly once with 16-byte loads (VEC = 4 floats / 2 doubles per load; VEC = 1 when a row of C is
// not 16-byte aligned), UNROLL independent loads in flight per thread, and a branch-free online log-sum-exp: a group of
// values is folded into the running (max, sum) pair with one rescaling exponential per GROUP instead of a data-dependent
// branch per element.  Grids are persistent-sized (a few CTAs per SM, grid-stride over rows / slabs).
template <typename T, int VEC> struct VecLoad;
template <typename T> struct VecLoad<T, 1> {
  static __device__ __forceinline__ void ld(const T* p, T (&o)[1]) { o[0] = __ldg(p); }
  static __device__ __forceinline__ void ld_stream(const T* p, T (&o)[1]) { o[0] = __ldcs(p); }
};
template <> struct VecLoad<float, 4> {
  static __device__ __forceinline__ void ld(const float* p, float (&o)[4]) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p)); o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
  }
  static __device__ __forceinline__ void ld_stream(const float* p, float (&o)[4]) {   // C is read once: evict first
    const float4 t = __ldcs(reinterpret_cast<const float4*>(p)); o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
  }
};
template <> struct VecLoad<double, 2> {
  static __device__ __forceinline__ void ld(const double* p, double (&o)[2]) {
    const double2 t = __ldg(reinterpret_cast<const double2*>(p)); o[0] = t.x; o[1] = t.y;
  }
  static __device__ __forceinline__ void ld_stream(const double* p, double (&o)[2]) {
    const double2 t = __ldcs(reinterpret_cast<const double2*>(p)); o[0] = t.x; o[1] = t.y;
  }
};
template <typename T> __device__ __forceinline__ T tmax(T a, T b) { return a > b ? a : b; }

// partial column LSE over a slab of rows:  (pm, ps)[l, slab, j] = online-LSE_i( u_i - C_ij/reg )
// Block = TX x TY threads: thread (tx, ty) owns the VEC columns j0 + tx*VEC .. and the rows r0 + ty, r0 + ty + TY, ...
// of its slab (so a warp reads >= 512 contiguous bytes of one row); the TY row phases are merged through shared memory.
constexpr int SKC_THREADS = 256, SKC_UNROLL = 4, SK_MAX_SLABS = 256;
template <typename T, int VEC>
__global__ void __launch_bounds__(SKC_THREADS)
sk_col_partial_kernel(const T* __restrict__ C, const T* __restrict__ u, int64_t N, int64_t M, int tx_count, int slabs,
                      int64_t rows_per_slab, T neg_inv_reg, T* __restrict__ pm, T* __restrict__ ps, const SkState* state) {
  if (state->done) return;
  extern __shared__ unsigned char sk_smem_raw[];
  T* sm_m = reinterpret_cast<T*>(sk_smem_raw);               // [TY][TX * VEC]
  T* sm_s = sm_m + SKC_THREADS * VEC;
  const int ty_count = SKC_THREADS / tx_count;
  const int tx = threadIdx.x % tx_count, ty = threadIdx.x / tx_count;
  const int64_t col_blocks = (M + (int64_t)tx_count * VEC - 1) / ((int64_t)tx_count * VEC);
  const int64_t l = blockIdx.y;
  const T* Cl = C + l * N * M;
  const T* ul = u + l * N;
  for (int64_t item = blockIdx.x; item < col_blocks * slabs; item += gridDim.x) {
    const int64_t cb = item % col_blocks, slab = item / col_blocks;
    const int64_t j = cb * tx_count * VEC + (int64_t)tx * VEC;
    const int64_t r0 = slab * rows_per_slab, r1 = min(N, r0 + rows_per_slab);
    T m[VEC], s[VEC];
#pragma unroll
    for (int c = 0; c < VEC; ++c) { m[c] = SkMath<T>::ninf(); s[c] = T(0); }
    if (j < M) {
      int64_t i = r0 + ty;
      for (; i + (int64_t)(SKC_UNROLL - 1) * ty_count < r1; i += (int64_t)SKC_UNROLL * ty_count) {
        T cv[SKC_UNROLL][VEC], ui[SKC_UNROLL];
#pragma unroll
        for (int k = 0; k < SKC_UNROLL; ++k) {
          VecLoad<T, VEC>::ld_stream(Cl + (i + (int64_t)k * ty_count) * M + j, cv[k]);
          ui[k] = __ldg(ul + i + (int64_t)k * ty_count);
        }
#pragma unroll
        for (int c = 0; c < VEC; ++c) {
          T t[SKC_UNROLL];
#pragma unroll
          for (int k = 0; k < SKC_UNROLL; ++k) t[k] = fma(cv[k][c], neg_inv_reg, ui[k]);
          T nm = m[c];
#pragma unroll
          for (int k = 0; k < SKC_UNROLL; ++k) nm = tmax(nm, t[k]);
          T acc = s[c] * SkMath<T>::ex(m[c] - nm);
#pragma unroll
          for (int k = 0; k < SKC_UNROLL; ++k) acc += SkMath<T>::ex(t[k] - nm);
          m[c] = nm; s[c] = acc;
        }
      }
      for (; i < r1; i += ty_count) {
        T cv[VEC];
        VecLoad<T, VEC>::ld_stream(Cl + i * M + j, cv);
        const T ui = __ldg(ul + i);
#pragma unroll
        for (int c = 0; c < VEC; ++c) {
          const T t = fma(cv[c], neg_inv_reg, ui), nm = tmax(m[c], t);
          s[c] = s[c] * SkMath<T>::ex(m[c] - nm) + SkMath<T>::ex(t - nm);
          m[c] = nm;
        }
      }
    }
    if (ty_count > 1) {
      __syncthreads();     // the previous item's merge has finished reading the arrays
#pragma unroll
      for (int c = 0; c < VEC; ++c) { sm_m[(ty * tx_count + tx) * VEC + c] = m[c]; sm_s[(ty * tx_count + tx) * VEC + c] = s[c]; }
      __syncthreads();
      if (ty == 0) {
        for (int w = 1; w < ty_count; ++w) {
#pragma unroll
          for (int c = 0; c < VEC; ++c) lse_merge(m[c], s[c], sm_m[(w * tx_count + tx) * VEC + c], sm_s[(w * tx_count + tx) * VEC + c]);
        }
      }
    }
    if (ty == 0 && j < M) {
#pragma unroll
      for (int c = 0; c < VEC; ++c) {
        pm[(l * slabs + slab) * M + j + c] = m[c];
        ps[(l * slabs + slab) * M + j + c] = s[c];
      }
    }
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

// u_i = log a_i - LSE_j( v_j - C_ij/reg ) ; TPR threads cooperate on one row, rows are dealt grid-stride
template <typename T, int TPR, int VEC>
__global__ void __launch_bounds__(256)
sk_row_kernel(const T* __restrict__ C, const T* __restrict__ v, int64_t N, int64_t M, T neg_inv_reg,
              const T* __restrict__ log_a, T* __restrict__ u, T* __restrict__ du, const SkState* state) {
  if (state->done) return;
  constexpr int ROWS = 256 / TPR, UNR = 4;
  __shared__ T sm_m[8], sm_s[8];
  const int64_t l = blockIdx.y;
  const int sub = threadIdx.x / TPR, t = threadIdx.x % TPR;
  const T* vl = v + l * M;
  for (int64_t ib = (int64_t)blockIdx.x * ROWS; ib < N; ib += (int64_t)gridDim.x * ROWS) {
    const int64_t i = ib + sub;
    T m = SkMath<T>::ninf(), s = T(0);
    if (i < N) {
      const T* row = C + (l * N + i) * M;
      int64_t j = (int64_t)t * VEC;
      for (; j + (int64_t)(UNR - 1) * TPR * VEC < M; j += (int64_t)UNR * TPR * VEC) {
        T cv[UNR][VEC], vv[UNR][VEC];
#pragma unroll
        for (int k = 0; k < UNR; ++k) {
          VecLoad<T, VEC>::ld_stream(row + j + (int64_t)k * TPR * VEC, cv[k]);
          VecLoad<T, VEC>::ld(vl + j + (int64_t)k * TPR * VEC, vv[k]);
        }
        T nm = m;
#pragma unroll
        for (int k = 0; k < UNR; ++k)
#pragma unroll
          for (int c = 0; c < VEC; ++c) { cv[k][c] = fma(cv[k][c], neg_inv_reg, vv[k][c]); nm = tmax(nm, cv[k][c]); }
        T acc = s * SkMath<T>::ex(m - nm);
#pragma unroll
        for (int k = 0; k < UNR; ++k)
#pragma unroll
          for (int c = 0; c < VEC; ++c) acc += SkMath<T>::ex(cv[k][c] - nm);
        m = nm; s = acc;
      }
      for (; j < M; j += (int64_t)TPR * VEC) {
        T cv[VEC], vv[VEC];
        VecLoad<T, VEC>::ld_stream(row + j, cv);
        VecLoad<T, VEC>::ld(vl + j, vv);
        T nm = m;
#pragma unroll
        for (int c = 0; c < VEC; ++c) { cv[c] = fma(cv[c], neg_inv_reg, vv[c]); nm = tmax(nm, cv[c]); }
        T acc = s * SkMath<T>::ex(m - nm);
#pragma unroll
        for (int c = 0; c < VEC; ++c) acc += SkMath<T>::ex(cv[c] - nm);
        m = nm; s = acc;
      }
    }
    // warp combine
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      T m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
      lse_merge(m, s, m2, s2);
    }
    if (TPR == 256) {
      __syncthreads();
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

// launch geometry shared by the full loop and the single half-steps
struct SkGeom {
  int vec;            // elements per 16-byte load (1 = scalar path: a row of C is not 16-byte aligned)
  int tx, slabs;      // column kernel: threads along a row, row slabs
  int64_t rows_per_slab;
  unsigned col_grid, row_grid;
  bool wide_rows;     // row kernel: 256 threads per row instead of a warp
};
template <typename T>
static SkGeom sk_geometry(const T* C, const T* v, int64_t L, int64_t N, int64_t M) {
  SkGeom g;
  const int full = 16 / (int)sizeof(T);
  const bool aligned = M % full == 0 && reinterpret_cast<uintptr_t>(C) % 16 == 0 && reinterpret_cast<uintptr_t>(v) % 16 == 0;
  g.vec = aligned ? full : 1;
  int tx = 1;
  while (tx < SKC_THREADS && (int64_t)tx * g.vec < M) tx <<= 1;
  g.tx = tx;
  const int ty = SKC_THREADS / tx;
  const int64_t col_blocks = ceil_div(M, (int64_t)tx * g.vec);
  int64_t want = ceil_div((int64_t)sm_count() * 8, col_blocks * L);
  const int64_t max_by_rows = ceil_div(N, (int64_t)ty * SKC_UNROLL * 4);   // >= 4 unrolled groups per thread
  if (want > max_by_rows) want = max_by_rows;
  if (want > SK_MAX_SLABS) want = SK_MAX_SLABS;
  if (want < 1) want = 1;
  g.slabs = (int)want;
  g.rows_per_slab = ceil_div(N, want);
  g.col_grid = (unsigned)std::min<int64_t>(col_blocks * want, (int64_t)sm_count() * 8);
  g.wide_rows = M > 4096;
  const int64_t row_blocks = g.wide_rows ? N : ceil_div(N, 8);
  g.row_grid = (unsigned)std::min<int64_t>(row_blocks, (int64_t)sm_count() * 8);
  return g;
}
template <typename T>
static void sk_launch_col(const SkGeom& g, const T* C, const T* u, int64_t L, int64_t N, int64_t M, T nir, T* pm, T* ps,
                          const SkState* state, cudaStream_t st) {
  constexpr int FULL = 16 / (int)sizeof(T);
  const dim3 grid(g.col_grid, (unsigned)L);
  const size_t smem = (size_t)2 * SKC_THREADS * g.vec * sizeof(T);
  if (g.vec == FULL)
    sk_col_partial_kernel<T, FULL><<<grid, SKC_THREADS, smem, st>>>(C, u, N, M, g.tx, g.slabs, g.rows_per_slab, nir, pm, ps, state);
  else
    sk_col_partial_kernel<T, 1><<<grid, SKC_THREADS, smem, st>>>(C, u, N, M, g.tx, g.slabs, g.rows_per_slab, nir, pm, ps, state);
}
template <typename T>
static void sk_launch_row(const SkGeom& g, const T* C, const T* v, int64_t L, int64_t N, int64_t M, T nir, const T* log_a,
                          T* u, T* du, const SkState* state, cudaStream_t st) {
  constexpr int FULL = 16 / (int)sizeof(T);
  const dim3 grid(g.row_grid, (unsigned)L);
  if (g.wide_rows) {
    if (g.vec == FULL) sk_row_kernel<T, 256, FULL><<<grid, 256, 0, st>>>(C, v, N, M, nir, log_a, u, du, state);
    else sk_row_kernel<T, 256, 1><<<grid, 256, 0, st>>>(C, v, N, M, nir, log_a, u, du, state);
  } else {
    if (g.vec == FULL) sk_row_kernel<T, 32, FULL><<<grid, 256, 0, st>>>(C, v, N, M, nir, log_a, u, du, state);
    else sk_row_kernel<T, 32, 1><<<grid, 256, 0, st>>>(C, v, N, M, nir, log_a, u, du, state);
  }
}

template <typename T>
static int sinkhorn_dense_impl(const T* a, const T* b, const T* C, int64_t L, int64_t N, int64_t M, double reg,
                               int max_iter, double threshold, int poll_every, T* u, T* v, T* plan, int* iters_done_host,
                               void* workspace, size_t workspace_bytes, cudaStream_t st, bool warm = false) {
  const SkGeom g = sk_geometry<T>(C, v, L, N, M);
  const int slabs = g.slabs;
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
  unsigned gfin = (unsigned)std::min<int64_t>(ceil_div(L * M, 256), (int64_t)sm_count() * 8);
  if (poll_every <= 0) poll_every = 16;
  SkState host_state{0, 0};
  for (int it = 0; it < max_iter; ++it) {
    sk_launch_col<T>(g, C, u, L, N, M, nir, pm, ps, state, st);
    sk_col_final_kernel<T><<<gfin, 256, 0, st>>>(pm, ps, slabs, L, M, log_b, v, dv, state);
    sk_launch_row<T>(g, C, v, L, N, M, nir, log_a, u, du, state, st);
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
    unsigned gp = (unsigned)std::min<int64_t>(ceil_div(L * N * M, 256), (int64_t)sm_count() * 32);
    sk_plan_kernel<T><<<gp ? gp : 1, 256, 0, st>>>(C, u, v, L, N, M, nir, plan);
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
  const SkGeom g = sk_geometry<float>(C, C, 1, N, M);
  const int slabs = g.slabs;
  Arena ar(workspace, workspace_bytes);
  float* pm = ar.take<float>((size_t)slabs * M);
  float* ps = ar.take<float>((size_t)slabs * M);
  SkState* state = ar.take<SkState>(1);
  if (!ar.ok()) return OTK_ERR_WORKSPACE;
  sk_zero_state_kernel<<<1, 1, 0, st>>>(state);
  sk_launch_col<float>(g, C, u, 1, N, M, (float)(-1.0 / reg), pm, ps, state, st);
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
  sk_launch_row<float>(sk_geometry<float>(C, v, 1, N, M), C, v, 1, N, M, nir, log_a, u, du, state, st);
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
         2 * align_up((size_t)L * SK_MAX_SLABS * M * per, 256) + align_up((size_t)L * 8, 256) + 2048;
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
