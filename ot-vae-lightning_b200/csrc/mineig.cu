// K4: smallest eigenvalue of symmetric matrices by Lanczos (full re-orthogonalisation) + Sturm bisection.
// Reference: min_eig, ot/matrix_utils.py:91-98 (`eigh` reads the LOWER triangle).  Used for is_pd / make_psd.
#include "otk_common.cuh"

namespace otk {

constexpr int LZ_MAX_STEPS = 128, LZ_THREADS = 512;

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

__global__ void __launch_bounds__(LZ_THREADS)
lanczos_min_eig_kernel(const double* __restrict__ A, int64_t d, int steps, double* __restrict__ V, double* out) {
  __shared__ double alpha[LZ_MAX_STEPS], beta[LZ_MAX_STEPS], proj[LZ_MAX_STEPS], red[32];
  __shared__ int m_eff;
  const int64_t l = blockIdx.x;
  const double* Al = A + l * d * d;
  double* Vl = V + l * (int64_t)(steps + 1) * d;
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
  if (tid == 0) m_eff = steps;
  __syncthreads();
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
    if (nn < 1e-13 * (fabs(dot) + 1e-300) || j + 1 == steps) {
      if (tid == 0) m_eff = j + 1;
      __syncthreads();
      break;
    }
    for (int64_t i = tid; i < d; i += LZ_THREADS) w[i] /= nn;
    __syncthreads();
  }
  if (tid == 0) {
    const int m = m_eff;
    double lo = 1e300, hi = -1e300;
    for (int i = 0; i < m; ++i) {
      double r = (i > 0 ? fabs(beta[i - 1]) : 0.0) + (i + 1 < m ? fabs(beta[i]) : 0.0);
      lo = fmin(lo, alpha[i] - r);
      hi = fmax(hi, alpha[i] + r);
    }
    // smallest eigenvalue of the tridiagonal: bisection on the Sturm count
    for (int it = 0; it < 200; ++it) {
      double x = 0.5 * (lo + hi);
      if (x == lo || x == hi) break;
      int neg = 0;
      double q = alpha[0] - x;
      if (q < 0) ++neg;
      for (int i = 1; i < m && neg == 0; ++i) {
        double qq = fabs(q) < 1e-300 ? 1e-300 : q;
        q = alpha[i] - x - beta[i - 1] * beta[i - 1] / qq;
        if (q < 0) ++neg;
      }
      if (neg > 0) hi = x; else lo = x;
    }
    out[l] = 0.5 * (lo + hi);
  }
}

}  // namespace otk
using namespace otk;

static int lz_steps(int64_t dim, int steps) {
  int s = steps > 0 ? steps : 96;
  if (s > dim) s = (int)dim;
  if (s > LZ_MAX_STEPS) s = LZ_MAX_STEPS;
  return s;
}

extern "C" size_t otk_min_eig_workspace_bytes(int64_t L, int64_t dim, int steps) {
  int s = lz_steps(dim, steps);
  return align_up((size_t)L * dim * dim * 8, 256) + align_up((size_t)L * (s + 1) * dim * 8, 256) + 512;
}

extern "C" int otk_min_eig(const void* a, int64_t L, int64_t dim, int dtype, int steps, double* lam_min, void* workspace,
                           size_t workspace_bytes, otk_stream_t stream) {
  OTK_TRY(require_device());
  OTK_REQUIRE(a && lam_min && L > 0 && dim > 0, "min_eig: bad arguments");
  if (!workspace || workspace_bytes < otk_min_eig_workspace_bytes(L, dim, steps)) return OTK_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  const int s = lz_steps(dim, steps);
  Arena ar(workspace, workspace_bytes);
  double* full = ar.take<double>((size_t)L * dim * dim);
  double* V = ar.take<double>((size_t)L * (s + 1) * dim);
  int64_t blocks = ceil_div(L * dim * dim, 256);
  if (blocks > (int64_t)sm_count() * 16) blocks = (int64_t)sm_count() * 16;
  lower_to_full_kernel<<<(unsigned)blocks, 256, 0, st>>>(a, dtype, L, dim, full);
  OTK_LAUNCH_CHECK();
  lanczos_min_eig_kernel<<<(unsigned)L, LZ_THREADS, 0, st>>>(full, dim, s, V, lam_min);
  OTK_LAUNCH_CHECK();
  return OTK_OK;
}
