// Generic-shape FFMA/DFMA GEMM (any M, N, K, strides).  This is the engine for shapes the tcgen05 path
// cannot take (dims that are not multiples of 4, fp64 polish steps, tiny matrices such as the reference
// tests' d = 3) and the on-device fp32 yardstick the tcgen05 kernels are tested against.
//
//   C[m,n] = alpha * sum_k (A(m,k) - a_off[k]) * B(n,k) + beta * C[m,n] + bias[n] + diag_add * delta_mn
//   A(m,k) = A[m*sam + k*sak],  B(n,k) = B[n*sbn + k*sbk],  C row-major with leading dim ldc.
#pragma once
#include "otk_common.cuh"

namespace otk {

template <typename T>
struct GemmArgs {
  const T* A; const T* B; T* C;
  int64_t M, N, K;
  int64_t sam, sak, sbn, sbk, ldc;
  int64_t strideA, strideB, strideC;  // batch strides (elements)
  T alpha, beta;
  const T* a_off; int64_t stride_aoff;  // optional [K] per batch
  const T* bias;  int64_t stride_bias;  // optional [N] per batch
  T diag_add;                           // added to C[m,m] (after alpha/beta)
  double* resid;                        // optional [batch]: += sum_mn (acc[m,n] - delta_mn)^2 of the raw product
  // --- tcgen05 3xTF32 engine only (ignored by the FFMA/DFMA engine) ---
  const T* A_lo = nullptr;              // if set, A is the TF32 "hi" plane and A_lo the "lo" plane (same layout)
  const T* B_lo = nullptr;
  T* C_hi = nullptr;                    // optional split output planes (layout of C); C itself may then be null
  T* C_lo = nullptr;
  T* scratch = nullptr;                 // >= 2*batch*(M*K + N*K) elements: used to split A / B when no lo plane is given
  // --- both engines: optional addend  C += add_scale * (add + add_lo), layout / batch stride of C (add_lo: the "lo" plane
  //     when the addend is stored as TF32 planes) - the b*M term of the accelerated Newton-Schulz polynomial
  const T* add = nullptr;
  const T* add_lo = nullptr;
  T add_scale = T(0);
};

constexpr int SG_BM = 64, SG_BN = 64, SG_BK = 16, SG_THREADS = 256;

template <typename T>
__global__ void __launch_bounds__(SG_THREADS) gemm_simt_kernel(GemmArgs<T> g) {
  __shared__ T As[SG_BK][SG_BM + 4];
  __shared__ T Bs[SG_BK][SG_BN + 4];
  const int64_t batch = blockIdx.z;
  const T* A = g.A + batch * g.strideA;
  const T* B = g.B + batch * g.strideB;
  T* C = g.C ? g.C + batch * g.strideC : nullptr;
  const T* a_off = g.a_off ? g.a_off + batch * g.stride_aoff : nullptr;
  const T* bias = g.bias ? g.bias + batch * g.stride_bias : nullptr;
  const T* add = g.add ? g.add + batch * g.strideC : nullptr;
  const T* add_lo = g.add_lo ? g.add_lo + batch * g.strideC : nullptr;
  const int64_t m0 = (int64_t)blockIdx.y * SG_BM, n0 = (int64_t)blockIdx.x * SG_BN;
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;  // 16 x 16 threads, 4 x 4 outputs each
  T acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = T(0);

  // loader mapping: choose the unit-stride direction for coalescing
  const bool a_kfast = (g.sak == 1), b_kfast = (g.sbk == 1);
  for (int64_t k0 = 0; k0 < g.K; k0 += SG_BK) {
#pragma unroll
    for (int r = 0; r < (SG_BM * SG_BK) / SG_THREADS; ++r) {
      int e = tid + r * SG_THREADS;
      int kk = a_kfast ? e % SG_BK : e / SG_BM;
      int mm = a_kfast ? e / SG_BK : e % SG_BM;
      int64_t m = m0 + mm, k = k0 + kk;
      T v = T(0);
      if (m < g.M && k < g.K) {
        v = A[m * g.sam + k * g.sak];
        if (a_off) v -= a_off[k];
      }
      As[kk][mm] = v;
    }
#pragma unroll
    for (int r = 0; r < (SG_BN * SG_BK) / SG_THREADS; ++r) {
      int e = tid + r * SG_THREADS;
      int kk = b_kfast ? e % SG_BK : e / SG_BN;
      int nn = b_kfast ? e / SG_BK : e % SG_BN;
      int64_t n = n0 + nn, k = k0 + kk;
      Bs[kk][nn] = (n < g.N && k < g.K) ? B[n * g.sbn + k * g.sbk] : T(0);
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SG_BK; ++kk) {
      T a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  double res = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int64_t n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      T r = g.alpha * acc[i][j];
      if (g.beta != T(0) && C) r += g.beta * C[m * g.ldc + n];
      if (bias) r += bias[n];
      if (add) r += g.add_scale * (add[m * g.ldc + n] + (add_lo ? add_lo[m * g.ldc + n] : T(0)));
      if (m == n) r += g.diag_add;
      if (C) C[m * g.ldc + n] = r;
      if (g.resid) { double e = (double)acc[i][j] - (m == n ? 1.0 : 0.0); res += e * e; }
    }
  }
  if (g.resid) {
    res = warp_sum(res);
    if (tid % 32 == 0) atomicAdd(&g.resid[batch], res);
  }
}

// fp64 engine for products of 128 rows / columns and more: 128 x 128 tiles, 8 x 8 outputs per thread (as 2 x 2 blocks of
// 4 x 4, so that a warp reads two broadcast addresses of A and 16 consecutive pairs of B per step), K in steps of 16.
// 64 DFMA per 16 shared-memory doubles read: bound by the DFMA pipe (measured ceiling, otk_microbench_peak: 36 TFLOP/s),
// where the 64 x 64 / 4 x 4 kernel above is bound by shared-memory bandwidth and barriers (~5 TFLOP/s).  Same epilogue.
constexpr int DG_BM = 128, DG_BN = 128, DG_BK = 16, DG_THREADS = 256;

static __global__ void __launch_bounds__(DG_THREADS, 1) gemm_dfma_kernel(GemmArgs<double> g) {
  __shared__ double As[DG_BK][DG_BM + 2];
  __shared__ double Bs[DG_BK][DG_BN + 2];
  const int64_t batch = blockIdx.z;
  const double* A = g.A + batch * g.strideA;
  const double* B = g.B + batch * g.strideB;
  double* C = g.C ? g.C + batch * g.strideC : nullptr;
  const double* a_off = g.a_off ? g.a_off + batch * g.stride_aoff : nullptr;
  const double* bias = g.bias ? g.bias + batch * g.stride_bias : nullptr;
  const double* add = g.add ? g.add + batch * g.strideC : nullptr;
  const double* add_lo = g.add_lo ? g.add_lo + batch * g.strideC : nullptr;
  const int64_t m0 = (int64_t)blockIdx.y * DG_BM, n0 = (int64_t)blockIdx.x * DG_BN;
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  double acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;
  const bool a_kfast = (g.sak == 1), b_kfast = (g.sbk == 1);
  for (int64_t k0 = 0; k0 < g.K; k0 += DG_BK) {
#pragma unroll
    for (int r = 0; r < (DG_BM * DG_BK) / DG_THREADS; ++r) {
      const int e = tid + r * DG_THREADS;
      const int kk = a_kfast ? e % DG_BK : e / DG_BM, mm = a_kfast ? e / DG_BK : e % DG_BM;
      const int64_t m = m0 + mm, k = k0 + kk;
      double v = 0.0;
      if (m < g.M && k < g.K) {
        v = A[m * g.sam + k * g.sak];
        if (a_off) v -= a_off[k];
      }
      As[kk][mm] = v;
    }
#pragma unroll
    for (int r = 0; r < (DG_BN * DG_BK) / DG_THREADS; ++r) {
      const int e = tid + r * DG_THREADS;
      const int kk = b_kfast ? e % DG_BK : e / DG_BN, nn = b_kfast ? e / DG_BK : e % DG_BN;
      const int64_t n = n0 + nn, k = k0 + kk;
      Bs[kk][nn] = (n < g.N && k < g.K) ? B[n * g.sbn + k * g.sbk] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < DG_BK; ++kk) {
      double a[8], b[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; a[4 + i] = As[kk][64 + ty * 4 + i]; }
#pragma unroll
      for (int j = 0; j < 4; ++j) { b[j] = Bs[kk][tx * 4 + j]; b[4 + j] = Bs[kk][64 + tx * 4 + j]; }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  double res = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int64_t n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= g.N) continue;
      double r = g.alpha * acc[i][j];
      if (g.beta != 0.0 && C) r += g.beta * C[m * g.ldc + n];
      if (bias) r += bias[n];
      if (add) r += g.add_scale * (add[m * g.ldc + n] + (add_lo ? add_lo[m * g.ldc + n] : 0.0));
      if (m == n) r += g.diag_add;
      if (C) C[m * g.ldc + n] = r;
      if (g.resid) { const double e = acc[i][j] - (m == n ? 1.0 : 0.0); res += e * e; }
    }
  }
  if (g.resid) {
    res = warp_sum(res);
    if (tid % 32 == 0) atomicAdd(&g.resid[batch], res);
  }
}

template <typename T>
inline int gemm_simt(const GemmArgs<T>& g, int64_t batch, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0 || batch <= 0) return OTK_OK;
  if constexpr (sizeof(T) == 8) {
    if (g.M >= DG_BM && g.N >= DG_BN && batch <= 65535) {
      dim3 grid((unsigned)ceil_div(g.N, DG_BN), (unsigned)ceil_div(g.M, DG_BM), (unsigned)batch);
      gemm_dfma_kernel<<<grid, DG_THREADS, 0, st>>>(g);
      OTK_LAUNCH_CHECK();
      return OTK_OK;
    }
  }
  dim3 grid((unsigned)ceil_div(g.N, SG_BN), (unsigned)ceil_div(g.M, SG_BM), (unsigned)batch);
  gemm_simt_kernel<T><<<grid, SG_THREADS, 0, st>>>(g);
  OTK_LAUNCH_CHECK();
  return OTK_OK;
}

// convenience: C = alpha * A * B^T (+ beta C) for row-major square-ish operands
template <typename T>
inline int gemm_nt_simt(const T* A, const T* B, T* C, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb,
                        int64_t ldc, int64_t batch, int64_t sA, int64_t sB, int64_t sC, T alpha, T beta,
                        cudaStream_t st) {
  GemmArgs<T> g{A, B, C, M, N, K, lda, 1, ldb, 1, ldc, sA, sB, sC, alpha, beta, nullptr, 0, nullptr, 0, T(0), nullptr};
  return gemm_simt(g, batch, st);
}

}  // namespace otk
