// fp32-accurate GEMM front end: tcgen05 3xTF32 when the shape is eligible, FFMA otherwise.
#pragma once
#include "gemm_simt.cuh"
namespace otk {
enum { ENGINE_AUTO = 0, ENGINE_SIMT = 1, ENGINE_UMMA_3X = 2, ENGINE_UMMA_1X = 3 };
// returns 1 if launched on tcgen05, 0 if not eligible, <0 on error   (gemm_umma.cu)
int gemm_umma_try(const GemmArgs<float>& g, int64_t batch, int passes, cudaStream_t st);
int gemm_f32(const GemmArgs<float>& g, int64_t batch, int engine, cudaStream_t st);
// one or two independent 3xTF32 products (operands given as hi/lo planes) in one launch; with ctrl != nullptr the launch
// is a no-op when ctrl[0] <= ctrl_index (device-side iteration limit of the Newton-Schulz loop)   (gemm_umma.cu)
// Optional stopping rule of the Newton-Schulz loop, evaluated by ONE CTA at the end of the launch (instead of a
// separate one-block kernel between two products): reads the residuals resid_k[0..L) of iteration k and lowers the
// iteration limit ctrl[0] / sets the verdict ctrl[1], ctrl[2] exactly as ns_ctrl_kernel (matfun.cu) does.
struct NsCtrlEval { const double* resid_k; int64_t L; int k, max_iters; double tol_done, tol_near; int* ctrl; };
int gemm_umma_dual(const GemmArgs<float>& g0, const GemmArgs<float>* g1, int64_t batch, const int* ctrl, int ctrl_index,
                   cudaStream_t st, const NsCtrlEval* eval = nullptr);
// precision-generic NT product: float -> tcgen05 / FFMA dispatch, double -> DFMA
inline int gemm_any(const GemmArgs<float>& g, int64_t batch, cudaStream_t st) { return gemm_f32(g, batch, ENGINE_AUTO, st); }
inline int gemm_any(const GemmArgs<double>& g, int64_t batch, cudaStream_t st) { return gemm_simt<double>(g, batch, st); }
// square NN product C = alpha * A * B (row-major d x d, batch stride dd).  NB: the Newton-Schulz iterates are
// symmetric only up to round-off and the iteration is UNSTABLE if B^T is substituted for B (the antisymmetric error
// component is amplified every step; scratch/ns_sim.py), so B is read as a true [K,N] operand (N-major).
template <typename T>
inline GemmArgs<T> nn_args_t(const T* A, const T* B, T* C, int64_t d, int64_t dd, T alpha) {
  return GemmArgs<T>{A, B, C, d, d, d, d, 1, 1, d, d, dd, dd, dd, alpha, T(0), nullptr, 0, nullptr, 0, T(0), nullptr};
}
inline GemmArgs<float> nt_args(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K, int64_t lda,
                               int64_t ldb, int64_t ldc, int64_t sA, int64_t sB, int64_t sC, float alpha, float beta) {
  return GemmArgs<float>{A, B, C, M, N, K, lda, 1, ldb, 1, ldc, sA, sB, sC, alpha, beta, nullptr, 0, nullptr, 0, 0.f, nullptr};
}
}  // namespace otk
