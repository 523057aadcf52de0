#include "gemm.cuh"
namespace otk {
int gemm_f32(const GemmArgs<float>& g, int64_t batch, int engine, cudaStream_t st) {
  if (engine != ENGINE_SIMT) {
    int r = gemm_umma_try(g, batch, engine == ENGINE_UMMA_1X ? 1 : 3, st);
    if (r != 0) return r < 0 ? r : OTK_OK;
    if (engine == ENGINE_UMMA_3X || engine == ENGINE_UMMA_1X) {
      set_last_error_msg("gemm: shape not eligible for the tcgen05 engine");
      return OTK_ERR_INVALID_ARGUMENT;
    }
  }
  return gemm_simt<float>(g, batch, st);
}
}  // namespace otk
namespace otk { extern int g_apply_force_cg; extern int g_stats_force_cg; extern int g_stats_dbg; extern int g_apply_dbg; extern int g_fs_fast; }
using namespace otk;
// tuning aids (not part of include/otk.h)
// tuning aid: 0 = automatic, 1 / 2 = force the single-CTA / CTA-pair apply kernel
namespace otk { extern int g_fast_counters[4]; }
extern "C" void otkdbg_fast_counters(int* out) { for (int i = 0; i < 4; ++i) out[i] = otk::g_fast_counters[i]; }
extern "C" void otkdbg_set_apply_cg(int cg) { g_apply_force_cg = cg; }
extern "C" void otkdbg_set_stats_cg(int cg) { g_stats_force_cg = cg; }
extern "C" void otkdbg_set_stats_dbg(int m) { g_stats_dbg = m; }
extern "C" void otkdbg_set_apply_dbg(int m) { g_apply_dbg = m; }
extern "C" void otkdbg_set_sinkhorn_fast(int m) { g_fs_fast = m; }
static int gemm_export(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb,
                       int64_t ldc, int64_t batch, int64_t strideA, int64_t strideB, int64_t strideC, float alpha, float beta,
                       int engine, bool nn, void* workspace, size_t workspace_bytes, otk_stream_t stream) {
  OTK_TRY(require_device());
  OTK_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0 && batch > 0, "gemm: bad arguments");
  OTK_REQUIRE(lda >= K && ldb >= (nn ? N : K) && ldc >= N, "gemm: leading dimension too small");
  GemmArgs<float> g = nt_args(A, B, C, M, N, K, lda, ldb, ldc, strideA, strideB, strideC, alpha, beta);
  if (nn) { g.sbn = 1; g.sbk = ldb; }
  const size_t need = (size_t)2 * batch * (M * K + N * K) * sizeof(float);
  if (workspace && workspace_bytes >= need) g.scratch = static_cast<float*>(workspace);
  else if (engine == ENGINE_UMMA_3X) return OTK_ERR_WORKSPACE;
  return gemm_f32(g, batch, engine, as_stream(stream));
}
extern "C" size_t otk_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K, int64_t batch) {
  return (size_t)2 * batch * (M * K + N * K) * sizeof(float) + 256;
}
extern "C" int otk_gemm_nt(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K, int64_t lda,
                           int64_t ldb, int64_t ldc, int64_t batch, int64_t strideA, int64_t strideB, int64_t strideC,
                           float alpha, float beta, int engine, void* workspace, size_t workspace_bytes,
                           otk_stream_t stream) {
  return gemm_export(A, B, C, M, N, K, lda, ldb, ldc, batch, strideA, strideB, strideC, alpha, beta, engine, false,
                     workspace, workspace_bytes, stream);
}
extern "C" int otk_gemm_nn(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K, int64_t lda,
                           int64_t ldb, int64_t ldc, int64_t batch, int64_t strideA, int64_t strideB, int64_t strideC,
                           float alpha, float beta, int engine, void* workspace, size_t workspace_bytes,
                           otk_stream_t stream) {
  return gemm_export(A, B, C, M, N, K, lda, ldb, ldc, batch, strideA, strideB, strideC, alpha, beta, engine, true,
                     workspace, workspace_bytes, stream);
}
