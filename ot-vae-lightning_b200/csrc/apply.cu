// K7: y = T (x - mean_s) + mean_t for batches of latents.
// Reference: apply_transport, ot/w2_utils.py:464-527 (deterministic full-matrix branch: :517-520), as called by
// W2Mixin.apply_transport (:581-597) and GaussianTransport.transport (transport/gaussian_transport.py:80-95).
// The reference runs B broadcast fp64 mat-vecs; here it is one GEMM  Y = (X - 1 mean_s^T) T^T + 1 mean_t^T.
#include <cstdlib>
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
__global__ void cast_var_kernel(const void* var, int dt, int64_t n, float* out) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e < n) out[e] = (float)load_real(var, e, dt);
}
// layout of the prepared operator state (both entry points carve it in this order)
struct PreparedOp { float *T32, *Thi, *Tlo, *ms32, *mt32, *var32; };
static bool carve_prepared(Arena& ar, int64_t L, int64_t dim, PreparedOp* p) {
  p->T32 = ar.take<float>((size_t)L * dim * dim);
  p->Thi = ar.take<float>((size_t)L * dim * dim);
  p->Tlo = ar.take<float>((size_t)L * dim * dim);
  p->ms32 = ar.take<float>((size_t)L * dim);
  p->mt32 = ar.take<float>((size_t)L * dim);
  p->var32 = ar.take<float>((size_t)L * dim);
  return ar.ok();
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

extern "C" size_t otk_transport_prepared_bytes(int64_t L, int64_t dim) {
  return 3 * align_up((size_t)L * dim * dim * 4, 256) + 3 * align_up((size_t)L * dim * 4, 256) + apply_h_workspace_bytes(L, dim) +
         4096;
}

extern "C" int otk_transport_prepare(const void* mean_s, const void* mean_t, const void* T, const void* var_s, int dtype,
                                     int64_t L, int64_t dim, void* state, size_t state_bytes, otk_stream_t stream) {
  OTK_TRY(require_device());
  OTK_REQUIRE(L > 0 && dim > 0 && mean_s && mean_t && T && var_s, "transport_prepare: bad arguments");
  if (!state || state_bytes < otk_transport_prepared_bytes(L, dim)) return OTK_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  Arena ar(state, state_bytes);
  PreparedOp p;
  if (!carve_prepared(ar, L, dim, &p)) return OTK_ERR_WORKSPACE;
  int64_t blocks = ceil_div(L * dim * dim, 256);
  if (blocks > (int64_t)sm_count() * 8) blocks = (int64_t)sm_count() * 8;
  cast3_kernel<<<(unsigned)blocks, 256, 0, st>>>(mean_s, mean_t, T, dtype, L * dim, L * dim * dim, p.ms32, p.mt32, p.T32);
  OTK_LAUNCH_CHECK();
  cast_var_kernel<<<(unsigned)ceil_div(L * dim, 256), 256, 0, st>>>(var_s, dtype, L * dim, p.var32);
  OTK_LAUNCH_CHECK();
  if (dim >= 64 && dim % 4 == 0) {
    apply_umma_split(p.T32, L * dim * dim, p.Thi, p.Tlo, st);
    OTK_CUDA(cudaGetLastError());
  }
  int r = apply_h_prepare(L, dim, p.ms32, p.T32, p.var32, ar, st);
  return r < 0 ? r : OTK_OK;
}

static int apply_prepared_impl(const float* x, int64_t L, int64_t rows, int64_t dim, int64_t xrs, int64_t xbs, const void* state,
                               size_t state_bytes, float* y, otk_stream_t stream) {
  OTK_TRY(require_device());
  OTK_REQUIRE(L > 0 && dim > 0 && rows >= 0, "apply_transport_prepared: bad arguments");
  if (rows == 0) return OTK_OK;
  OTK_REQUIRE(x && y, "apply_transport_prepared: null latents");
  OTK_REQUIRE(xrs >= dim && (L == 1 || xbs > 0), "apply_transport_prepared: bad latent strides");
  if (!state || state_bytes < otk_transport_prepared_bytes(L, dim)) return OTK_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  Arena ar(const_cast<void*>(state), state_bytes);
  PreparedOp p;
  if (!carve_prepared(ar, L, dim, &p)) return OTK_ERR_WORKSPACE;
  // TMA reads the view in place when its strides are 16-byte multiples; anything else goes through the FFMA engine,
  // which takes arbitrary element strides
  // latency regime (the reference's batches of 250, tests/test_latent_transport.py:66-98): ONE launch of the FFMA engine
  // (fp32, centring and bias fused) instead of flag reset + FP16-split tcgen05 kernel + gated fallback - the call is
  // bound by launches, not by the 2 rows dim^2 flops (cfg1: 24.5 -> ~19 us per transport call)
  static const bool small_simt = [] { const char* e = getenv("OTK_APPLY_SMALL_SIMT"); return !(e && e[0] == '0'); }();   // tuning aid
  const bool latency_regime = small_simt && rows <= 256 && L * rows * dim <= (int64_t)1 << 16;
  if (!latency_regime && apply_umma_eligible(x, y, L, rows, dim) && xrs % 4 == 0 && xbs % 4 == 0) {
    const bool pair = apply_umma_pair(rows, dim);
    int* flag = nullptr;
    int used = apply_h_run_prepared(x, L, rows, dim, p.mt32, y, ar, pair, st, &flag, xrs, xbs);
    if (used < 0) return used;
    if (!used) flag = nullptr;
    used = apply_umma_run_planes(x, L, rows, dim, p.ms32, p.mt32, p.Thi, p.Tlo, y, pair, flag, st, xrs, xbs);
    if (used < 0) return used;
    if (used) return OTK_OK;
  }
  GemmArgs<float> g{x, p.T32, y, rows, dim, dim, xrs, 1, dim, 1, dim, xbs, dim * dim, rows * dim,
                    1.f, 0.f, p.ms32, dim, p.mt32, dim, 0.f, nullptr};
  return gemm_simt<float>(g, L, st);
}

extern "C" int otk_apply_transport_prepared(const float* x, int64_t L, int64_t rows, int64_t dim, const void* state,
                                            size_t state_bytes, float* y, otk_stream_t stream) {
  return apply_prepared_impl(x, L, rows, dim, dim, rows * dim, state, state_bytes, y, stream);
}

extern "C" int otk_apply_transport_prepared_strided(const float* x, int64_t L, int64_t rows, int64_t dim, int64_t row_stride,
                                                    int64_t batch_stride, const void* state, size_t state_bytes, float* y,
                                                    otk_stream_t stream) {
  return apply_prepared_impl(x, L, rows, dim, row_stride, batch_stride, state, state_bytes, y, stream);
}
