// K4: smallest eigenvalue of symmetric matrices by Lanczos (full re-orthogonalisation) + Sturm multisection.
// Reference: min_eig, ot/matrix_utils.py:91-98 (`eigh` reads the LOWER triangle).  Used for is_pd / make_psd.
//
// The recurrence runs until the smallest Ritz pair has converged (residual bound |beta_m s_m| <= 1e-13 * ||T|| and a
// stagnant Ritz value) or, failing that, for all `dim` steps: with full re-orthogonalisation that is a complete
// tridiagonalisation, so the result is then the exact lambda_min (to fp64 round-off), not an upper bound.
#include "otk_common.cuh"

namespace otk {

constexpr int LZ_THREADS = 1024, LZ_CHECK_EVERY = 8, LZ_FIRST_CHECK = 16;

__global__ void lower_to_full_kernel(const void* a, int dt, int64_t L, int64_t dim, double* out) {
  const int64_t total = L * dim * dim;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t l = e / (dim * dim), r = e % (dim * dim), i = r / dim, j = r % dim;
    out[e] = load_real(a, l * dim * dim + (i >= j ? i * dim + j : j * dim + i), dt);
  }
}

__device__ __forceinline__ double lz_block_sum(double v, double* red) {
  v = warp_sum(v);
  if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = v;
  __syncthreads();
  double r = 0;
  if (threadIdx.x < 32) {
    r = threadIdx.x < LZ_THREADS / 32 ? red[threadIdx.x] : 0.0;
    r = warp_sum(r);
    if (threadIdx.x == 0) red[0] = r;
  }
  __syncthreads();
  r = red[0];
  __syncthreads();
  return r;
}

// true iff the m x m tridiagonal (alpha, beta) has an eigenvalue < x (first negative Sturm pivot)
__device__ __forceinline__ bool lz_has_eig_below(const double* alpha, const double* beta, int m, double x) {
  double q = alpha[0] - x;
  if (q < 0) return true;
  for (int i = 1; i < m; ++i) {
    const double qq = fabs(q) < 1e-300 ? 1e-300 : q;
    q = alpha[i] - x - beta[i - 1] * beta[i - 1] / qq;
    if (q < 0) return true;
  }
  return false;
}

// smallest eigenvalue of the leading m x m tridiagonal: every thread tests one shift per round (multisection), so a round
// narrows the bracket LZ_THREADS-fold.  All threads return the same value.
__device__ double lz_min_ritz(const double* alpha, const double* beta, int m, double* red, int* ired) {
  const int tid = threadIdx.x;
  double lo = 1e300, hi = -1e300;
  for (int i = tid; i < m; i += LZ_THREADS) {
    const double r = (i > 0 ? fabs(beta[i - 1]) : 0.0) + (i + 1 < m ? fabs(beta[i]) : 0.0);
    lo = fmin(lo, alpha[i] - r);
    hi = fmax(hi, alpha[i] + r);
  }
  lo = -warp_max(-lo);
  hi = warp_max(hi);
  if (tid % 32 == 0) { red[tid / 32] = lo; red[32 + tid / 32] = hi; }
  __syncthreads();
  lo = red[0]; hi = red[32];
  for (int w = 1; w < LZ_THREADS / 32; ++w) { lo = fmin(lo, red[w]); hi = fmax(hi, red[32 + w]); }
  __syncthreads();
  const double pad = 1e-14 * fmax(fabs(lo), fabs(hi)) + 1e-300;
  lo -= pad; hi += pad;            // lambda_min in (lo, hi]: nothing below lo, something below hi
  for (int round = 0; round < 12; ++round) {
    const double step = (hi - lo) / (LZ_THREADS + 1);
    if (!(step > 0) || lo + step == lo) break;
    const double x = lo + step * (tid + 1);
    const bool below = lz_has_eig_below(alpha, beta, m, x);
    if (tid == 0) *ired = LZ_THREADS;                       // index of the first shift with an eigenvalue below it
    __syncthreads();
    if (below) atomicMin(ired, tid);
    __syncthreads();
    const int first = *ired;
    __syncthreads();
    const double nlo = lo + step * first, nhi = first < LZ_THREADS ? lo + step * (first + 1) : hi;
    lo = nlo; hi = nhi;
  }
  return 0.5 * (lo + hi);
}

// |last component| of the unit eigenvector of the m x m tridiagonal for the eigenvalue theta (three-term recurrence from the
// top, rescaled against overflow); thread 0 only.
__device__ double lz_last_component(const double* alpha, const double* beta, int m, double theta) {
  double sp = 0.0, s = 1.0, nrm2 = 1.0;
  for (int i = 0; i + 1 < m; ++i) {
    const double b = fabs(beta[i]) < 1e-300 ? 1e-300 : beta[i];
    double sn = -((alpha[i] - theta) * s + (i > 0 ? beta[i - 1] * sp : 0.0)) / b;
    sp = s; s = sn;
    nrm2 += s * s;
    if (nrm2 > 1e200) { sp *= 1e-100; s *= 1e-100; nrm2 *= 1e-200; }
  }
  return fabs(s) / sqrt(nrm2);
}

__global__ void __launch_bounds__(LZ_THREADS)
lanczos_min_eig_kernel(const double* __restrict__ A, int64_t d, int steps, int adaptive, double* __restrict__ V,
                       double* __restrict__ ab, double* out) {
  __shared__ double red[64];
  __shared__ int ired, stop;
  extern __shared__ double proj[];            // [steps]
  const int64_t l = blockIdx.x;
  const double* Al = A + l * d * d;
  double* Vl = V + l * (int64_t)(steps + 1) * d;
  double* alpha = ab + l * 2 * (int64_t)steps;
  double* beta = alpha + steps;
  const int tid = threadIdx.x;
  // deterministic start vector
  double nrm = 0;
  for (int64_t i = tid; i < d; i += LZ_THREADS) {
    uint32_t h = (uint32_t)(i * 2654435761u) ^ 0x9e3779b9u;
    h ^= h >> 15; h *= 0x85ebca6bu; h ^= h >> 13;
    double v = 0.5 + (double)(h & 0xffff) / 65536.0;
    Vl[i] = v;
    nrm += v * v;
  }
  nrm = sqrt(lz_block_sum(nrm, red));
  for (int64_t i = tid; i < d; i += LZ_THREADS) Vl[i] /= nrm;
  if (tid == 0) stop = 0;
  __syncthreads();
  int m = steps;
  double theta_prev = 1e300, theta = 0.0;
  bool have_theta = false;
  for (int j = 0; j < steps; ++j) {
    const double* vj = Vl + (int64_t)j * d;
    double* w = Vl + (int64_t)(j + 1) * d;
    double dot = 0;
    for (int64_t i = tid; i < d; i += LZ_THREADS) {
      double a0 = 0, a1 = 0, a2 = 0, a3 = 0;  // symmetric: column access is coalesced across threads
      int64_t k = 0;
      for (; k + 3 < d; k += 4) {
        a0 += Al[k * d + i] * vj[k];
        a1 += Al[(k + 1) * d + i] * vj[k + 1];
        a2 += Al[(k + 2) * d + i] * vj[k + 2];
        a3 += Al[(k + 3) * d + i] * vj[k + 3];
      }
      for (; k < d; ++k) a0 += Al[k * d + i] * vj[k];
      const double acc = (a0 + a1) + (a2 + a3);
      w[i] = acc;
      dot += acc * vj[i];
    }
    dot = lz_block_sum(dot, red);
    if (tid == 0) alpha[j] = dot;
    // full re-orthogonalisation against v_0..v_j: two sweeps of block classical Gram-Schmidt (all j+1 projections
    // per sweep are computed together - one warp per basis vector - so a sweep costs two block barriers)
    for (int sweep = 0; sweep < 2; ++sweep) {
      for (int t = tid / 32; t <= j; t += LZ_THREADS / 32) {
        const double* vt = Vl + (int64_t)t * d;
        double p = 0;
        for (int64_t i = tid % 32; i < d; i += 32) p += w[i] * vt[i];
        p = warp_sum(p);
        if (tid % 32 == 0) proj[t] = p;
      }
      __syncthreads();
      for (int64_t i = tid; i < d; i += LZ_THREADS) {
        double acc = w[i];
        for (int t = 0; t <= j; ++t) acc -= proj[t] * Vl[(int64_t)t * d + i];
        w[i] = acc;
      }
      __syncthreads();
    }
    double nn = 0;
    for (int64_t i = tid; i < d; i += LZ_THREADS) nn += w[i] * w[i];
    nn = sqrt(lz_block_sum(nn, red));
    if (tid == 0) beta[j] = nn;
    __syncthreads();
    // invariant subspace (breakdown) or out of steps: the tridiagonal is final
    if (nn < 1e-13 * (fabs(dot) + 1e-300) || j + 1 == steps) { m = j + 1; have_theta = false; break; }
    // convergence test of the smallest Ritz pair
    if (adaptive && j + 1 >= LZ_FIRST_CHECK && (j + 1) % LZ_CHECK_EVERY == 0) {
      theta = lz_min_ritz(alpha, beta, j + 1, red, &ired);
      if (tid == 0) {
        double scale = 0;
        for (int i = 0; i <= j; ++i) scale = fmax(scale, fmax(fabs(alpha[i]), fabs(beta[i])));
        const double tol = 1e-13 * scale + 1e-300;
        const double resid = nn * lz_last_component(alpha, beta, j + 1, theta);
        stop = (resid <= tol && fabs(theta - theta_prev) <= tol) ? 1 : 0;
      }
      __syncthreads();
      theta_prev = theta;
      if (stop) { m = j + 1; have_theta = true; break; }
    }
    for (int64_t i = tid; i < d; i += LZ_THREADS) w[i] /= nn;
    __syncthreads();
  }
  if (!have_theta) theta = lz_min_ritz(alpha, beta, m, red, &ired);
  if (tid == 0) out[l] = theta;
}

}  // namespace otk
using namespace otk;

// steps <= 0: adaptive, up to `dim` steps (exact at the cap); steps > 0: exactly min(steps, dim) steps, no early stop
static int lz_steps(int64_t dim, int steps) {
  int64_t s = steps > 0 ? steps : dim;
  if (s > dim) s = dim;
  return (int)s;
}

extern "C" size_t otk_min_eig_workspace_bytes(int64_t L, int64_t dim, int steps) {
  int s = lz_steps(dim, steps);
  return align_up((size_t)L * dim * dim * 8, 256) + align_up((size_t)L * (s + 1) * dim * 8, 256) +
         align_up((size_t)L * 2 * s * 8, 256) + 512;
}

extern "C" int otk_min_eig(const void* a, int64_t L, int64_t dim, int dtype, int steps, double* lam_min, void* workspace,
                           size_t workspace_bytes, otk_stream_t stream) {
  OTK_TRY(require_device());
  OTK_REQUIRE(a && lam_min && L > 0 && dim > 0, "min_eig: bad arguments");
  OTK_REQUIRE(dim <= 16384, "min_eig: dim > 16384 is not supported");
  if (!workspace || workspace_bytes < otk_min_eig_workspace_bytes(L, dim, steps)) return OTK_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  const int s = lz_steps(dim, steps);
  Arena ar(workspace, workspace_bytes);
  double* full = ar.take<double>((size_t)L * dim * dim);
  double* V = ar.take<double>((size_t)L * (s + 1) * dim);
  double* ab = ar.take<double>((size_t)L * 2 * s);
  int64_t blocks = ceil_div(L * dim * dim, 256);
  if (blocks > (int64_t)sm_count() * 16) blocks = (int64_t)sm_count() * 16;
  lower_to_full_kernel<<<(unsigned)blocks, 256, 0, st>>>(a, dtype, L, dim, full);
  OTK_LAUNCH_CHECK();
  const size_t dyn = (size_t)s * sizeof(double);
  if (dyn > 48 * 1024)
    OTK_CUDA(cudaFuncSetAttribute(lanczos_min_eig_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  lanczos_min_eig_kernel<<<(unsigned)L, LZ_THREADS, dyn, st>>>(full, dim, s, steps > 0 ? 0 : 1, V, ab, lam_min);
  OTK_LAUNCH_CHECK();
  return OTK_OK;
}
