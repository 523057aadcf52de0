// K3-K6: SPD matrix square roots by coupled Newton-Schulz, Gaussian W2^2 and the transport operator; K4 min_eig.
// Reference: ot/matrix_utils.py:37-76,91-98 ; ot/w2_utils.py:40-80,756-768.
//
// All d x d products run through gemm_f32 (tcgen05 3xTF32 when eligible).  Every iterate is a polynomial in the
// (symmetric) input, so row-major operands are used as their own transposes: A*B is issued as the NT product A*B^T.
#include "gemm.cuh"

namespace otk {

constexpr int NS_MAX_ITERS = 60;

__device__ __forceinline__ double block_sum(double v, double* red) {
  v = warp_sum(v);
  if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = v;
  __syncthreads();
  double r = 0;
  if (threadIdx.x < 32) {
    r = threadIdx.x < blockDim.x / 32 ? red[threadIdx.x] : 0.0;
    r = warp_sum(r);
    if (threadIdx.x == 0) red[0] = r;
  }
  __syncthreads();
  r = red[0];
  __syncthreads();
  return r;
}

// c[l] = || A_l + ridge I ||_F
__global__ void frob_kernel(const void* a, int dt, int64_t dim, double ridge, float* c) {
  __shared__ double red[32];
  const int64_t l = blockIdx.x;
  double acc = 0;
  for (int64_t e = threadIdx.x; e < dim * dim; e += blockDim.x) {
    double v = load_real(a, l * dim * dim + e, dt);
    if (e / dim == e % dim) v += ridge;
    acc += v * v;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) c[l] = (float)sqrt(acc);
}

// Y = (A + ridge I)/c, Z = I
__global__ void ns_init_kernel(const void* a, int dt, int64_t L, int64_t dim, double ridge, const float* c, float* Y,
                               float* Z) {
  const int64_t total = L * dim * dim;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t l = e / (dim * dim), r = e % (dim * dim);
    bool diag = (r / dim == r % dim);
    double v = load_real(a, e, dt) + (diag ? ridge : 0.0);
    Y[e] = (float)(v / (double)c[l]);
    Z[e] = diag ? 1.f : 0.f;
  }
}

// out = scale(l) * (in + in^T)/2 [+ diag], scale = s0 * c[l]^pw ; optionally cast to dt
__global__ void sym_scale_kernel(const float* in, int64_t L, int64_t dim, const float* c, double pw, double s0,
                                 double diag_add, void* out, int out_dt) {
  const int64_t total = L * dim * dim;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t l = e / (dim * dim), r = e % (dim * dim), i = r / dim, j = r % dim;
    double sc = s0 * (c ? pow((double)c[l], pw) : 1.0);
    double v = 0.5 * ((double)in[e] + (double)in[l * dim * dim + j * dim + i]) * sc;
    if (i == j) v += diag_add;
    store_real(out, e, out_dt, v);
  }
}

__global__ void cast_f32_kernel(const void* a, int dt, int64_t n, float* out) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
    out[e] = (float)load_real(a, e, dt);
}

// w2[l] = |ms - mt|^2 + tr(Cs) + tr(Ct) - 2 * sqrt(c[l]) * tr(Y_l)      (one block per l)
__global__ void w2_trace_kernel(const void* ms, const void* mt, const void* cs, const void* ct, int dt, int64_t dim,
                                const float* Y, const float* c, double* w2) {
  __shared__ double red[32];
  const int64_t l = blockIdx.x;
  double acc = 0;
  for (int64_t i = threadIdx.x; i < dim; i += blockDim.x) {
    double dm = load_real(ms, l * dim + i, dt) - load_real(mt, l * dim + i, dt);
    int64_t dgl = l * dim * dim + i * dim + i;
    acc += dm * dm + load_real(cs, dgl, dt) + load_real(ct, dgl, dt) - 2.0 * sqrt((double)c[l]) * (double)Y[dgl];
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) w2[l] = acc;
}

static inline unsigned ew_grid(int64_t total) {
  int64_t b = ceil_div(total, 256), cap = (int64_t)sm_count() * 16;
  return (unsigned)(b < cap ? (b > 0 ? b : 1) : cap);
}

struct NsWork {
  float *Y[2], *Z[2], *T, *c;
  double* resid;  // [NS_MAX_ITERS][L]
  static size_t bytes(int64_t L, int64_t d) {
    return 5 * align_up((size_t)L * d * d * 4, 256) + align_up((size_t)L * 4, 256) +
           align_up((size_t)NS_MAX_ITERS * L * 8, 256);
  }
  void carve(Arena& ar, int64_t L, int64_t d) {
    for (int i = 0; i < 2; ++i) { Y[i] = ar.take<float>((size_t)L * d * d); Z[i] = ar.take<float>((size_t)L * d * d); }
    T = ar.take<float>((size_t)L * d * d);
    c = ar.take<float>((size_t)L);
    resid = ar.take<double>((size_t)NS_MAX_ITERS * L);
  }
};

// Coupled Newton-Schulz on (A + ridge I)/c:  T = (3I - Z Y)/2, Y <- Y T, Z <- T Z.
// On return w.Y[*cur] ~ sqrt(A/c), w.Z[*cur] ~ (A/c)^-1/2 (unsymmetrised), c in w.c.
static int ns_solve(const void* a, int dt, int64_t L, int64_t d, double ridge, int iters, NsWork& w, int* cur_out,
                    cudaStream_t st) {
  const int64_t dd = d * d;
  frob_kernel<<<(unsigned)L, 256, 0, st>>>(a, dt, d, ridge, w.c);
  OTK_LAUNCH_CHECK();
  ns_init_kernel<<<ew_grid(L * dd), 256, 0, st>>>(a, dt, L, d, ridge, w.c, w.Y[0], w.Z[0]);
  OTK_LAUNCH_CHECK();
  OTK_CUDA(cudaMemsetAsync(w.resid, 0, (size_t)NS_MAX_ITERS * L * 8, st));
  const bool adaptive = iters <= 0;
  const int max_iters = adaptive ? (L <= 64 ? NS_MAX_ITERS : 32) : (iters < NS_MAX_ITERS ? iters : NS_MAX_ITERS);
  int cur = 0, stop_at = max_iters;
  double host_res[64];
  for (int k = 0; k < max_iters && k < stop_at; ++k) {
    GemmArgs<float> g = nt_args(w.Z[cur], w.Y[cur], w.T, d, d, d, d, d, d, dd, dd, dd, -0.5f, 0.f);
    g.diag_add = 1.5f;
    g.resid = w.resid + (size_t)k * L;
    OTK_TRY(gemm_f32(g, L, ENGINE_AUTO, st));
    OTK_TRY(gemm_f32(nt_args(w.Y[cur], w.T, w.Y[cur ^ 1], d, d, d, d, d, d, dd, dd, dd, 1.f, 0.f), L, ENGINE_AUTO, st));
    OTK_TRY(gemm_f32(nt_args(w.T, w.Z[cur], w.Z[cur ^ 1], d, d, d, d, d, d, dd, dd, dd, 1.f, 0.f), L, ENGINE_AUTO, st));
    cur ^= 1;
    if (adaptive && (k % 2 == 1) && L <= 64) {
      OTK_CUDA(cudaMemcpyAsync(host_res, w.resid + (size_t)k * L, (size_t)L * 8, cudaMemcpyDeviceToHost, st));
      OTK_CUDA(cudaStreamSynchronize(st));
      double worst = 0;
      for (int64_t l = 0; l < L; ++l) {
        if (!(host_res[l] == host_res[l])) { set_last_error_msg("sqrtm: Newton-Schulz produced NaN"); return OTK_ERR_NOT_CONVERGED; }
        if (host_res[l] > worst) worst = host_res[l];
      }
      // resid holds ||I - Z_k Y_k||_F^2 of the state *before* update k; convergence is quadratic.
      if (worst < 1e-8) stop_at = k + 1;            // already at the fp32 floor
      else if (worst < 9e-4) stop_at = k + 2;       // ||.||_F < 0.03 -> one more update reaches the floor
    }
  }
  *cur_out = cur;
  return OTK_OK;
}

}  // namespace otk
using namespace otk;

extern "C" size_t otk_sqrtm_workspace_bytes(int64_t L, int64_t dim) { return NsWork::bytes(L, dim) + 4096; }

extern "C" int otk_sqrtm(const void* a, int64_t L, int64_t dim, int dtype, double ridge, int iters, int polish, void* root,
                         void* iroot, void* workspace, size_t workspace_bytes, otk_stream_t stream) {
  (void)polish;
  OTK_TRY(require_device());
  OTK_REQUIRE(a && L > 0 && dim > 0 && (root || iroot), "sqrtm: bad arguments");
  if (!workspace || workspace_bytes < otk_sqrtm_workspace_bytes(L, dim)) return OTK_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  Arena ar(workspace, workspace_bytes);
  NsWork w; w.carve(ar, L, dim);
  int cur = 0;
  OTK_TRY(ns_solve(a, dtype, L, dim, ridge, iters, w, &cur, st));
  if (root) {
    sym_scale_kernel<<<ew_grid(L * dim * dim), 256, 0, st>>>(w.Y[cur], L, dim, w.c, 0.5, 1.0, 0.0, root, dtype);
    OTK_LAUNCH_CHECK();
  }
  if (iroot) {
    sym_scale_kernel<<<ew_grid(L * dim * dim), 256, 0, st>>>(w.Z[cur], L, dim, w.c, -0.5, 1.0, 0.0, iroot, dtype);
    OTK_LAUNCH_CHECK();
  }
  return OTK_OK;
}

// shared by w2_gaussian and transport_operator: given covariances P (rooted) and Q, computes
//   S = P^1/2, Zp = P^-1/2 (fp32, symmetrised, in `S`/`Zp`) and leaves sqrt(S Q S)/sqrt(c2) in w.Y[cur2] with c2 in w.c.
static int rooted_mix(const void* P, const void* Q, int dt, int64_t L, int64_t d, double ridge, int iters, NsWork& w,
                      float* S, float* Zp, float* Q32, float* G, float* mix, int* cur2, cudaStream_t st) {
  const int64_t dd = d * d;
  int cur = 0;
  OTK_TRY(ns_solve(P, dt, L, d, ridge, iters, w, &cur, st));
  sym_scale_kernel<<<ew_grid(L * dd), 256, 0, st>>>(w.Y[cur], L, d, w.c, 0.5, 1.0, 0.0, S, OTK_F32);
  OTK_LAUNCH_CHECK();
  if (Zp) {
    sym_scale_kernel<<<ew_grid(L * dd), 256, 0, st>>>(w.Z[cur], L, d, w.c, -0.5, 1.0, 0.0, Zp, OTK_F32);
    OTK_LAUNCH_CHECK();
  }
  cast_f32_kernel<<<ew_grid(L * dd), 256, 0, st>>>(Q, dt, L * dd, Q32);
  OTK_LAUNCH_CHECK();
  OTK_TRY(gemm_f32(nt_args(S, Q32, G, d, d, d, d, d, d, dd, dd, dd, 1.f, 0.f), L, ENGINE_AUTO, st));     // S Q
  OTK_TRY(gemm_f32(nt_args(G, S, Q32, d, d, d, d, d, d, dd, dd, dd, 1.f, 0.f), L, ENGINE_AUTO, st));     // (S Q) S
  sym_scale_kernel<<<ew_grid(L * dd), 256, 0, st>>>(Q32, L, d, nullptr, 0.0, 1.0, 0.0, mix, OTK_F32);
  OTK_LAUNCH_CHECK();
  OTK_TRY(ns_solve(mix, OTK_F32, L, d, 0.0, iters, w, cur2, st));
  return OTK_OK;
}

extern "C" size_t otk_w2_gaussian_workspace_bytes(int64_t L, int64_t dim) {
  return NsWork::bytes(L, dim) + 5 * align_up((size_t)L * dim * dim * 4, 256) + 4096;
}

extern "C" int otk_w2_gaussian(const void* mean_s, const void* mean_t, const void* cov_s, const void* cov_t, int64_t L,
                               int64_t dim, int dtype, int iters, int polish, double* w2, void* workspace,
                               size_t workspace_bytes, otk_stream_t stream) {
  (void)polish;
  OTK_TRY(require_device());
  OTK_REQUIRE(mean_s && mean_t && cov_s && cov_t && w2 && L > 0 && dim > 0, "w2_gaussian: bad arguments");
  if (!workspace || workspace_bytes < otk_w2_gaussian_workspace_bytes(L, dim)) return OTK_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  Arena ar(workspace, workspace_bytes);
  NsWork w; w.carve(ar, L, dim);
  const size_t n = (size_t)L * dim * dim;
  float *S = ar.take<float>(n), *Q32 = ar.take<float>(n), *G = ar.take<float>(n), *mix = ar.take<float>(n);
  int cur2 = 0;
  // the reference roots the TARGET covariance for the distance (w2_utils.py:70-71)
  OTK_TRY(rooted_mix(cov_t, cov_s, dtype, L, dim, 0.0, iters, w, S, nullptr, Q32, G, mix, &cur2, st));
  w2_trace_kernel<<<(unsigned)L, 256, 0, st>>>(mean_s, mean_t, cov_s, cov_t, dtype, dim, w.Y[cur2], w.c, w2);
  OTK_LAUNCH_CHECK();
  return OTK_OK;
}

extern "C" size_t otk_transport_operator_workspace_bytes(int64_t L, int64_t dim) {
  return NsWork::bytes(L, dim) + 6 * align_up((size_t)L * dim * dim * 4, 256) + 4096;
}

extern "C" int otk_transport_operator(const void* cov_s, const void* cov_t, int64_t L, int64_t dim, int dtype,
                                      double pg_star, int iters, int polish, void* T, const void* mean_s,
                                      const void* mean_t, double* w2, void* workspace, size_t workspace_bytes,
                                      otk_stream_t stream) {
  (void)polish;
  OTK_TRY(require_device());
  OTK_REQUIRE(cov_s && cov_t && T && L > 0 && dim > 0, "transport_operator: bad arguments");
  OTK_REQUIRE(!w2 || (mean_s && mean_t), "transport_operator: w2 requested without means");
  if (!workspace || workspace_bytes < otk_transport_operator_workspace_bytes(L, dim)) return OTK_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  Arena ar(workspace, workspace_bytes);
  NsWork w; w.carve(ar, L, dim);
  const size_t n = (size_t)L * dim * dim;
  const int64_t d = dim, dd = dim * dim;
  float *S = ar.take<float>(n), *Zp = ar.take<float>(n), *Q32 = ar.take<float>(n), *G = ar.take<float>(n),
        *mix = ar.take<float>(n);
  int cur2 = 0;
  // the map roots the SOURCE covariance (w2_utils.py:766-767); the 1e-8 ridge of the inverse root is kept
  OTK_TRY(rooted_mix(cov_s, cov_t, dtype, L, dim, 1e-8, iters, w, S, Zp, Q32, G, mix, &cur2, st));
  if (w2) {
    w2_trace_kernel<<<(unsigned)L, 256, 0, st>>>(mean_s, mean_t, cov_s, cov_t, dtype, dim, w.Y[cur2], w.c, w2);
    OTK_LAUNCH_CHECK();
  }
  // R = sqrt(mix) = sqrt(c) * sym(Y);  T = (1-p) Zp R Zp + p I
  sym_scale_kernel<<<ew_grid(L * dd), 256, 0, st>>>(w.Y[cur2], L, d, w.c, 0.5, 1.0, 0.0, mix, OTK_F32);
  OTK_LAUNCH_CHECK();
  OTK_TRY(gemm_f32(nt_args(Zp, mix, G, d, d, d, d, d, d, dd, dd, dd, 1.f, 0.f), L, ENGINE_AUTO, st));
  OTK_TRY(gemm_f32(nt_args(G, Zp, Q32, d, d, d, d, d, d, dd, dd, dd, 1.f, 0.f), L, ENGINE_AUTO, st));
  sym_scale_kernel<<<ew_grid(L * dd), 256, 0, st>>>(Q32, L, d, nullptr, 0.0, 1.0 - pg_star, pg_star, T, dtype);
  OTK_LAUNCH_CHECK();
  return OTK_OK;
}
