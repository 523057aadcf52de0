// K7: y = T (x - mean_s) + mean_t for batches of latents.
// Reference: apply_transport, ot/w2_utils.py:464-527 (deterministic full-matrix branch: :517-520), as called by
// W2Mixin.apply_transport (:581-597) and GaussianTransport.transport (transport/gaussian_transport.py:80-95).
// The reference runs B broadcast fp64 mat-vecs; here it is one GEMM  Y = (X - 1 mean_s^T) T^T + 1 mean_t^T.
#include "gemm.cuh"
#include "apply_umma.cuh"

namespace otk {
__global__ void cast3_kernel(const void* ms, const void* mt, const void* T, int dt, int64_t nvec, int64_t nmat, float* ms32,
                             float* mt32, float* T32) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < nmat; e += (int64_t)gridDim.x * blockDim.x) {
    T32[e] = (float)load_real(T, e, dt);
    if (e < nvec) {
      ms32[e] = (float)load_real(ms, e, dt);
      mt32[e] = (float)load_real(mt, e, dt);
    }
  }
}
}  // namespace otk
using namespace otk;

extern "C" size_t otk_apply_transport_workspace_bytes(int64_t L, int64_t rows, int64_t dim) {
  (void)rows;
  return 3 * align_up((size_t)L * dim * dim * 4, 256) + 3 * align_up((size_t)L * dim * 4, 256) + 1024 +
         apply_h_workspace_bytes(L, dim) + 2048;
}

extern "C" int otk_apply_transport(const float* x, int64_t L, int64_t rows, int64_t dim, const void* mean_s,
                                   const void* mean_t, const void* T, int dtype, float* y, void* workspace,
                                   size_t workspace_bytes, otk_stream_t stream) {
  OTK_TRY(require_device());
  OTK_REQUIRE(L > 0 && dim > 0 && rows >= 0 && mean_s && mean_t && T, "apply_transport: bad arguments");
  if (rows == 0) return OTK_OK;
  OTK_REQUIRE(x && y, "apply_transport: null latents");
  if (!workspace || workspace_bytes < otk_apply_transport_workspace_bytes(L, rows, dim)) return OTK_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  Arena ar(workspace, workspace_bytes);
  float* T32 = ar.take<float>((size_t)L * dim * dim);
  float* Thi = ar.take<float>((size_t)L * dim * dim);
  float* Tlo = ar.take<float>((size_t)L * dim * dim);
  float* ms32 = ar.take<float>((size_t)L * dim);
  float* mt32 = ar.take<float>((size_t)L * dim);
  int64_t blocks = ceil_div(L * dim * dim, 256);
  if (blocks > (int64_t)sm_count() * 8) blocks = (int64_t)sm_count() * 8;
  cast3_kernel<<<(unsigned)blocks, 256, 0, st>>>(mean_s, mean_t, T, dtype, L * dim, L * dim * dim, ms32, mt32, T32);
  OTK_LAUNCH_CHECK();
  int used = apply_umma_try(x, L, rows, dim, ms32, mt32, T32, Thi, Tlo, y, ar, st);
  if (used < 0) return used;
  if (used) return OTK_OK;
  GemmArgs<float> g{x, T32, y, rows, dim, dim, dim, 1, dim, 1, dim, rows * dim, dim * dim, rows * dim,
                    1.f, 0.f, ms32, dim, mt32, dim, 0.f, nullptr};
  return gemm_simt<float>(g, L, st);
}
