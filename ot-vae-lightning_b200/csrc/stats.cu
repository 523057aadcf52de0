// K1/K2: streaming sufficient statistics (sum x, sum x x^T, n) and the mean/cov finaliser.
// Reference: ot/distribution_models/gaussian_model.py:99-108,144-157 ; utils/__init__.py:204-206 ;
//            metrics/fid.py:99-122 ; ot/matrix_utils.py:145-158.
#include "otk_common.cuh"
#include "stats_umma.cuh"

namespace otk {

// ------------------------------------------------------------------------------------------------
// DFMA engine (any dim): one CTA = one upper-triangular 64x64 tile pair x one chunk of rows.
// fp32 latents are widened to fp64 and accumulated with DFMA (products of fp32 values are exact in fp64), i.e. the
// same arithmetic as the reference's fp64 einsum on `samples.type_as(buffer)` (gaussian_model.py:103,148); fp64
// atomics across chunks.  This is the engine for small / unaligned dims; aligned dims go to the tcgen05 kernel.
// ------------------------------------------------------------------------------------------------
constexpr int ST_T = 64, ST_BK = 16, ST_THREADS = 256;
constexpr int ST_FUSED_ROWS = 256;   // batches up to this many rows take the single-launch fused path

// FUSED (small batches, one chunk = all rows): the CTA owns its output tile, so it applies the accumulate / EMA rule to the
// running buffers itself - one launch per update, no staging area, no atomics (latency mode: batches of a few hundred).
// (StatsRunning: stats_umma.cuh)

// T = tile width (64: 4 x 4 outputs per thread; 32: 2 x 2 - four times as many CTAs for the small batches of the latency
// mode, where a 128-wide model has only three 64-wide tile pairs)
template <bool FUSED, typename X, int T = ST_T>
__global__ void __launch_bounds__(ST_THREADS)
stats_simt_kernel(const X* __restrict__ x, int64_t rows, int64_t dim, int64_t row_stride, int64_t batch_stride,
                  int64_t chunk_rows, int n_tiles, double* __restrict__ ws_cov, double* __restrict__ ws_sum,
                  StatsRunning run, const X* __restrict__ x_b, StatsRunning run_b) {
  // FUSED with a second problem (otk_stats_update_pair: the source and the target batch of one GaussianTransport.update):
  // blockIdx.y selects the problem, both go through one launch
  if constexpr (FUSED) {
    if (blockIdx.y == 1) { x = x_b; run = run_b; }
  }
  // rows per step: the narrow-tile variant runs on a handful of CTAs and is bound by the load -> barrier -> compute round
  // trip of a step, so it takes 64 rows per step (16 independent loads in flight per thread) instead of 16
  constexpr int BK = T == 32 ? 64 : ST_BK;
  // the tiles are widened to fp64 ONCE, on the way into shared memory: every element is read by 16 threads, and the
  // F2F.F64.F32 conversions (a quarter-rate pipe) were what bound the kernel when they sat in the inner loop
  __shared__ double As[BK][T + 4];
  __shared__ double Bs[BK][T + 4];
  // decode the upper-triangular tile pair (ti <= tj) from blockIdx.x
  int p = blockIdx.x, ti = 0;
  while (p >= n_tiles - ti) { p -= n_tiles - ti; ++ti; }
  const int tj = ti + p;
  const int64_t l = blockIdx.z;
  const int64_t r0 = FUSED ? 0 : (int64_t)blockIdx.y * chunk_rows;
  const int64_t r1 = min(rows, r0 + chunk_rows);
  const X* xb = x + l * batch_stride;
  constexpr int TH = T / 16;      // outputs per thread and direction
  const int64_t i0 = (int64_t)ti * T, j0 = (int64_t)tj * T;
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  double acc[TH][TH];
#pragma unroll
  for (int i = 0; i < TH; ++i)
#pragma unroll
    for (int j = 0; j < TH; ++j) acc[i][j] = 0.0;
  double colsum = 0.0;  // threads 0..63 of a diagonal CTA own one column each

  for (int64_t k0 = r0; k0 < r1; k0 += BK) {
#pragma unroll
    for (int r = 0; r < (T * BK) / ST_THREADS; ++r) {
      int e = tid + r * ST_THREADS;
      int kk = e / T, cc = e % T;
      int64_t row = k0 + kk;
      bool rok = row < r1;
      As[kk][cc] = (rok && i0 + cc < dim) ? (double)xb[row * row_stride + i0 + cc] : 0.0;
      Bs[kk][cc] = (rok && j0 + cc < dim) ? (double)xb[row * row_stride + j0 + cc] : 0.0;
    }
    __syncthreads();
    if (ti == tj && tid < T) {
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) colsum += As[kk][tid];
    }
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      double a[TH], b[TH];
#pragma unroll
      for (int i = 0; i < TH; ++i) a[i] = As[kk][ty * TH + i];
#pragma unroll
      for (int j = 0; j < TH; ++j) b[j] = Bs[kk][tx * TH + j];
#pragma unroll
      for (int i = 0; i < TH; ++i)
#pragma unroll
        for (int j = 0; j < TH; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  if constexpr (FUSED) {
    const double keep = run.decay < 0 ? 1.0 : run.decay, gain = run.decay < 0 ? 1.0 : 1.0 - run.decay;
#pragma unroll
    for (int i = 0; i < TH; ++i) {
      int64_t gi = i0 + ty * TH + i;
      if (gi >= dim) continue;
#pragma unroll
      for (int j = 0; j < TH; ++j) {
        int64_t gj = j0 + tx * TH + j;
        if (gj >= dim) continue;
        const int64_t e = l * dim * dim + gi * dim + gj;
        store_real(run.sum_cov, e, run.buf_dtype, load_real(run.sum_cov, e, run.buf_dtype) * keep + acc[i][j] * gain);
        if (ti != tj) {      // mirror the off-diagonal tile (a diagonal tile holds both triangles, computed identically)
          const int64_t e2 = l * dim * dim + gj * dim + gi;
          store_real(run.sum_cov, e2, run.buf_dtype, load_real(run.sum_cov, e2, run.buf_dtype) * keep + acc[i][j] * gain);
        }
      }
    }
    if (ti == tj && tid < T && i0 + tid < dim) {
      const int64_t e = l * dim + i0 + tid;
      store_real(run.sum, e, run.buf_dtype, load_real(run.sum, e, run.buf_dtype) * keep + colsum * gain);
    }
    if (blockIdx.x == 0 && tid == 0)
      store_real(run.n_obs, l, run.n_dtype, load_real(run.n_obs, l, run.n_dtype) * keep + (double)rows * gain);
    return;
  }
  double* cov = ws_cov + l * dim * dim;
#pragma unroll
  for (int i = 0; i < TH; ++i) {
    int64_t gi = i0 + ty * TH + i;
    if (gi >= dim) continue;
#pragma unroll
    for (int j = 0; j < TH; ++j) {
      int64_t gj = j0 + tx * TH + j;
      if (gj < dim) atomicAdd(&cov[gi * dim + gj], acc[i][j]);
    }
  }
  if (ti == tj && tid < T && i0 + tid < dim) atomicAdd(&ws_sum[l * dim + i0 + tid], colsum);
}

// merge the fp64 staging area into the running buffers (mirror the lower triangle, apply the EMA rule).
// pivot == nullptr: the staging area holds the raw sums, element (i <= j) at [i][j].
// pivot != nullptr: it holds the pivot-shifted sums of the tcgen05 kernel, P' transposed (element (i <= j) at [j][i]) and
//   S'; the raw sums are rebuilt exactly in fp64:  sum x = S' + n c ,  sum x x^T = P' + c S'^T + S' c^T + n c c^T .
__global__ void stats_merge_kernel(const double* __restrict__ ws_cov, const double* __restrict__ ws_sum,
                                   const float* __restrict__ pivot, int64_t L, int64_t dim, double rows, double decay,
                                   void* n_obs, int n_dtype, void* sum, void* sum_cov, int buf_dtype) {
  const int64_t total = L * dim * dim;
  const double keep = decay < 0 ? 1.0 : decay, gain = decay < 0 ? 1.0 : 1.0 - decay;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t l = e / (dim * dim), r = e % (dim * dim), i = r / dim, j = r % dim;
    // only tiles with tile(i) <= tile(j) were accumulated; always reading the (min, max) element also makes the
    // result exactly symmetric
    const int64_t lo_i = i < j ? i : j, hi_j = i < j ? j : i;
    double v;
    if (pivot) {
      const double ci = pivot[l * dim + lo_i], cj = pivot[l * dim + hi_j];
      const double si = ws_sum[l * dim + lo_i], sj = ws_sum[l * dim + hi_j];
      v = ws_cov[l * dim * dim + hi_j * dim + lo_i] + ci * sj + si * cj + rows * ci * cj;
    } else {
      v = ws_cov[l * dim * dim + lo_i * dim + hi_j];
    }
    store_real(sum_cov, e, buf_dtype, load_real(sum_cov, e, buf_dtype) * keep + v * gain);
    if (r < dim) {
      int64_t s = l * dim + r;
      const double sv = ws_sum[s] + (pivot ? rows * (double)pivot[s] : 0.0);
      store_real(sum, s, buf_dtype, load_real(sum, s, buf_dtype) * keep + sv * gain);
    }
    if (r == 0) store_real(n_obs, l, n_dtype, load_real(n_obs, l, n_dtype) * keep + rows * gain);
  }
}

__global__ void mean_cov_kernel(const void* sum, const void* sum_cov, const void* n_obs, int n_dtype, int64_t L,
                                int64_t dim, void* mean, void* cov, int dt) {
  const int64_t total = L * dim * dim;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t l = e / (dim * dim), r = e % (dim * dim), i = r / dim, j = r % dim;
    double n = load_real(n_obs, l, n_dtype);
    double mi = load_real(sum, l * dim + i, dt) / n, mj = load_real(sum, l * dim + j, dt) / n;
    store_real(cov, e, dt, load_real(sum_cov, e, dt) / n - mi * mj);
    if (j == 0) store_real(mean, l * dim + i, dt, mi);
  }
}

// fit of one model in one launch (GaussianModel.fit -> _compute_mean_cov -> _update_mean / _update_cov,
// gaussian_model.py:110-183): for every leading index with n > 1e-8, mean = sum / n and the raw covariance
// sum_cov / n - mean mean^T (biased) are written in place; optionally also the operand GaussianTransport.compute() needs,
// triu-mirror(raw) + shift I (the `Symmetric` + strict `MakePositiveDefinite` read of a PD matrix, :204-229).
__global__ void gaussian_fit_kernel(const void* sum, const void* sum_cov, int buf_dt, const void* n_obs, int n_dt, int64_t L,
                                    int64_t dim, void* mean, void* cov_raw, void* cov_sym, double shift, int out_dt) {
  const int64_t total = L * dim * dim;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t l = e / (dim * dim), r = e % (dim * dim), i = r / dim, j = r % dim;
    const double n = load_real(n_obs, l, n_dt);
    const bool seen = n > 1e-8;
    double raw, sym;
    if (seen) {
      const double mi = load_real(sum, l * dim + i, buf_dt) / n, mj = load_real(sum, l * dim + j, buf_dt) / n;
      raw = load_real(sum_cov, e, buf_dt) / n - mi * mj;
      const int64_t up = l * dim * dim + (i <= j ? i * dim + j : j * dim + i);
      sym = load_real(sum_cov, up, buf_dt) / n - mi * mj;      // the upper-triangle element, mirrored
      store_real(cov_raw, e, out_dt, raw);
      if (j == 0) store_real(mean, l * dim + i, out_dt, mi);
    } else {
      sym = load_real(cov_raw, l * dim * dim + (i <= j ? i * dim + j : j * dim + i), out_dt);
    }
    if (cov_sym) store_real(cov_sym, e, out_dt, sym + (i == j ? shift : 0.0));
  }
}

__global__ void symmetrize_shift_kernel(const void* a, const void* shift, int64_t L, int64_t dim, void* out, int dt) {
  const int64_t total = L * dim * dim;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t l = e / (dim * dim), r = e % (dim * dim), i = r / dim, j = r % dim;
    double v = load_real(a, l * dim * dim + (i <= j ? i * dim + j : j * dim + i), dt);
    if (i == j && shift) v += load_real(shift, l, dt);
    store_real(out, e, dt, v);
  }
}

__global__ void asymmetry_kernel(const void* a, int64_t L, int64_t dim, int dt, double* asym) {
  // one block per matrix
  const int64_t l = blockIdx.x;
  double acc = 0;
  for (int64_t e = threadIdx.x; e < dim * dim; e += blockDim.x) {
    int64_t i = e / dim, j = e % dim;
    double d = load_real(a, l * dim * dim + e, dt) - load_real(a, l * dim * dim + j * dim + i, dt);
    acc += d * d;
  }
  __shared__ double red[32];
  acc = warp_sum(acc);
  if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < blockDim.x / 32 ? red[threadIdx.x] : 0.0;
    v = warp_sum(v);
    if (threadIdx.x == 0) asym[l] = v;
  }
}

static inline unsigned ew_grid(int64_t total) {
  int64_t b = ceil_div(total, 256);
  int64_t cap = (int64_t)sm_count() * 16;
  return (unsigned)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace otk

namespace otk { extern int g_stats_force_cg; }
using namespace otk;

extern "C" size_t otk_stats_update_workspace_bytes(int64_t L, int64_t rows, int64_t dim) {
  (void)rows;
  return align_up((size_t)L * dim * dim * 8, 256) + align_up((size_t)L * dim * 8, 256) + stats_umma_extra_workspace(L, dim);
}

// X = float: the product path (tcgen05 kernels, DFMA engine for small / unaligned shapes and small batches).
// X = double: fp64 latents (the reference casts `samples.type_as(buffer)`, gaussian_model.py:103, and FID takes
//   `features.double()`, fid.py:101) always run on the DFMA engine - fp64 products, fp64 accumulation - so that sums of
//   outer products stay positive semi-definite to fp64 round-off (a rank-deficient covariance is then singular, not
//   indefinite at the 1e-7 level of an fp32-accurate product).
template <typename X>
static int stats_update_impl(const X* x, int64_t L, int64_t rows, int64_t dim, int64_t row_stride, int64_t batch_stride,
                             double decay, void* n_obs, int n_dtype, void* sum, void* sum_cov, int buf_dtype,
                             void* workspace, size_t workspace_bytes, otk_stream_t stream) {
  OTK_TRY(require_device());
  OTK_REQUIRE(L > 0 && dim > 0 && rows >= 0, "stats_update: bad shape");
  OTK_REQUIRE(rows == 0 || x, "stats_update: null latents");
  OTK_REQUIRE(n_obs && sum && sum_cov, "stats_update: null running buffer");
  OTK_REQUIRE(row_stride >= dim, "stats_update: row_stride < dim");
  if (workspace_bytes < otk_stats_update_workspace_bytes(L, rows, dim) || !workspace) return OTK_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  constexpr bool kF32 = sizeof(X) == 4;
  if (rows > 0 && rows <= ST_FUSED_ROWS && !(kF32 && g_stats_force_cg)) {
    // latency mode: one launch, every CTA merges its own output tile (exact fp64 products, as the reference's einsum)
    const int n_tiles = (int)ceil_div(dim, ST_T);
    const int64_t pairs = (int64_t)n_tiles * (n_tiles + 1) / 2;
    if (pairs <= 65535 && L <= 65535) {
      StatsRunning run{n_obs, sum, sum_cov, n_dtype, buf_dtype, decay};
      if (pairs * L * 2 <= sm_count()) {        // too few 64-wide tile pairs to occupy the machine: 32-wide tiles
        const int n32 = (int)ceil_div(dim, 32);
        const int64_t pairs32 = (int64_t)n32 * (n32 + 1) / 2;
        stats_simt_kernel<true, X, 32><<<dim3((unsigned)pairs32, 1, (unsigned)L), ST_THREADS, 0, st>>>(
            x, rows, dim, row_stride, batch_stride, rows, n32, nullptr, nullptr, run, nullptr, StatsRunning{});
      } else {
        stats_simt_kernel<true, X><<<dim3((unsigned)pairs, 1, (unsigned)L), ST_THREADS, 0, st>>>(x, rows, dim, row_stride, batch_stride,
                                                                                               rows, n_tiles, nullptr, nullptr, run, nullptr, StatsRunning{});
      }
      OTK_LAUNCH_CHECK();
      return OTK_OK;
    }
  }
  Arena ar(workspace, workspace_bytes);
  double* ws_cov = ar.take<double>((size_t)L * dim * dim);
  double* ws_sum = ar.take<double>((size_t)L * dim);
  const size_t staging_bytes = align_up((size_t)L * dim * dim * 8, 256) + (size_t)L * dim * 8;
  int tile = ST_T;
  const float* pivot = nullptr;
  if (rows == 0) OTK_CUDA(cudaMemsetAsync(workspace, 0, staging_bytes, st));
  if (rows > 0) {
    int used = 0;
    if constexpr (kF32) {
      used = stats_umma_try(x, L, rows, dim, row_stride, batch_stride, ws_cov, ws_sum, ar, st, &tile, &pivot,
                            StatsRunning{n_obs, sum, sum_cov, n_dtype, buf_dtype, decay});
      if (used < 0) return used;
      if (used == 2) return OTK_OK;          // FP16-split kernels: already merged into the running buffers
    }
    if (!used) {
      OTK_CUDA(cudaMemsetAsync(workspace, 0, staging_bytes, st));
      tile = ST_T;
      pivot = nullptr;
      int n_tiles = (int)ceil_div(dim, ST_T);
      int64_t pairs = (int64_t)n_tiles * (n_tiles + 1) / 2;
      // enough chunks for ~4 waves, chunk length a multiple of ST_BK
      int64_t want = ceil_div((int64_t)sm_count() * 4, pairs * L);
      int64_t chunk = ceil_div(ceil_div(rows, want), ST_BK) * ST_BK;
      if (chunk < 64) chunk = 64;
      if (chunk > 65536) chunk = 65536;
      dim3 grid((unsigned)pairs, (unsigned)ceil_div(rows, chunk), (unsigned)L);
      OTK_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "stats_update: too many row chunks / batches");
      stats_simt_kernel<false, X><<<grid, ST_THREADS, 0, st>>>(x, rows, dim, row_stride, batch_stride, chunk, n_tiles, ws_cov,
                                                               ws_sum, StatsRunning{}, nullptr, StatsRunning{});
      OTK_LAUNCH_CHECK();
    }
  }
  stats_merge_kernel<<<ew_grid(L * dim * dim), 256, 0, st>>>(ws_cov, ws_sum, pivot, L, dim, (double)rows, decay, n_obs,
                                                            n_dtype, sum, sum_cov, buf_dtype);
  OTK_LAUNCH_CHECK();
  return OTK_OK;
}

extern "C" int otk_stats_update(const float* x, int64_t L, int64_t rows, int64_t dim, int64_t row_stride,
                                int64_t batch_stride, double decay, void* n_obs, int n_dtype, void* sum, void* sum_cov,
                                int buf_dtype, void* workspace, size_t workspace_bytes, otk_stream_t stream) {
  return stats_update_impl<float>(x, L, rows, dim, row_stride, batch_stride, decay, n_obs, n_dtype, sum, sum_cov, buf_dtype,
                                  workspace, workspace_bytes, stream);
}

// source + target batch of one GaussianTransport.update (reference ot/transport/base.py: two GaussianModel.update calls,
// gaussian_model.py:99-108) in ONE launch when the batch is in the latency regime; otherwise the two updates run back to back
extern "C" int otk_stats_update_pair(const float* x_a, const float* x_b, int64_t rows, int64_t dim, int64_t row_stride,
                                     double decay, void* n_obs_a, void* sum_a, void* sum_cov_a, void* n_obs_b, void* sum_b,
                                     void* sum_cov_b, int n_dtype, int buf_dtype, void* workspace, size_t workspace_bytes,
                                     otk_stream_t stream) {
  OTK_TRY(require_device());
  OTK_REQUIRE(dim > 0 && rows > 0 && x_a && x_b, "stats_update_pair: bad arguments");
  OTK_REQUIRE(n_obs_a && sum_a && sum_cov_a && n_obs_b && sum_b && sum_cov_b, "stats_update_pair: null running buffer");
  OTK_REQUIRE(row_stride >= dim, "stats_update_pair: row_stride < dim");
  const int n32 = (int)ceil_div(dim, 32);
  const int64_t pairs32 = (int64_t)n32 * (n32 + 1) / 2;
  if (rows <= ST_FUSED_ROWS && pairs32 <= 65535 && !g_stats_force_cg) {
    const StatsRunning ra{n_obs_a, sum_a, sum_cov_a, n_dtype, buf_dtype, decay}, rb{n_obs_b, sum_b, sum_cov_b, n_dtype, buf_dtype, decay};
    const int n_tiles = (int)ceil_div(dim, ST_T);
    const int64_t pairs = (int64_t)n_tiles * (n_tiles + 1) / 2;
    cudaStream_t st = as_stream(stream);
    if (pairs * 4 <= sm_count())        // too few 64-wide tile pairs (two problems) to occupy the machine: 32-wide tiles
      stats_simt_kernel<true, float, 32><<<dim3((unsigned)pairs32, 2, 1), ST_THREADS, 0, st>>>(
          x_a, rows, dim, row_stride, rows * row_stride, rows, n32, nullptr, nullptr, ra, x_b, rb);
    else
      stats_simt_kernel<true, float><<<dim3((unsigned)pairs, 2, 1), ST_THREADS, 0, st>>>(
          x_a, rows, dim, row_stride, rows * row_stride, rows, n_tiles, nullptr, nullptr, ra, x_b, rb);
    OTK_LAUNCH_CHECK();
    return OTK_OK;
  }
  OTK_TRY(stats_update_impl<float>(x_a, 1, rows, dim, row_stride, rows * row_stride, decay, n_obs_a, n_dtype, sum_a, sum_cov_a,
                                   buf_dtype, workspace, workspace_bytes, stream));
  return stats_update_impl<float>(x_b, 1, rows, dim, row_stride, rows * row_stride, decay, n_obs_b, n_dtype, sum_b, sum_cov_b,
                                  buf_dtype, workspace, workspace_bytes, stream);
}

extern "C" int otk_stats_update_f64(const double* x, int64_t L, int64_t rows, int64_t dim, int64_t row_stride,
                                    int64_t batch_stride, double decay, void* n_obs, int n_dtype, void* sum, void* sum_cov,
                                    int buf_dtype, void* workspace, size_t workspace_bytes, otk_stream_t stream) {
  return stats_update_impl<double>(x, L, rows, dim, row_stride, batch_stride, decay, n_obs, n_dtype, sum, sum_cov, buf_dtype,
                                   workspace, workspace_bytes, stream);
}

extern "C" int otk_mean_cov(const void* sum, const void* sum_cov, const void* n_obs, int n_dtype, int64_t L,
                            int64_t dim, void* mean, void* cov, int dtype, otk_stream_t stream) {
  OTK_TRY(require_device());
  OTK_REQUIRE(L > 0 && dim > 0 && sum && sum_cov && n_obs && mean && cov, "mean_cov: bad arguments");
  mean_cov_kernel<<<ew_grid(L * dim * dim), 256, 0, as_stream(stream)>>>(sum, sum_cov, n_obs, n_dtype, L, dim, mean,
                                                                        cov, dtype);
  OTK_LAUNCH_CHECK();
  return OTK_OK;
}

extern "C" int otk_gaussian_fit(const void* sum, const void* sum_cov, int buf_dtype, const void* n_obs, int n_dtype, int64_t L,
                                int64_t dim, void* mean, void* cov_raw, void* cov_sym, double shift, int dtype,
                                otk_stream_t stream) {
  OTK_TRY(require_device());
  OTK_REQUIRE(L > 0 && dim > 0 && sum && sum_cov && n_obs && mean && cov_raw, "gaussian_fit: bad arguments");
  OTK_REQUIRE(cov_sym != cov_raw, "gaussian_fit: cov_sym must not alias cov_raw");
  gaussian_fit_kernel<<<ew_grid(L * dim * dim), 256, 0, as_stream(stream)>>>(sum, sum_cov, buf_dtype, n_obs, n_dtype, L, dim, mean,
                                                                           cov_raw, cov_sym, shift, dtype);
  OTK_LAUNCH_CHECK();
  return OTK_OK;
}

extern "C" int otk_symmetrize_shift(const void* a, const void* shift, int64_t L, int64_t dim, void* out, int dtype,
                                    otk_stream_t stream) {
  OTK_TRY(require_device());
  OTK_REQUIRE(L > 0 && dim > 0 && a && out && a != out, "symmetrize_shift: bad arguments (in-place not allowed)");
  symmetrize_shift_kernel<<<ew_grid(L * dim * dim), 256, 0, as_stream(stream)>>>(a, shift, L, dim, out, dtype);
  OTK_LAUNCH_CHECK();
  return OTK_OK;
}

extern "C" int otk_asymmetry(const void* a, int64_t L, int64_t dim, int dtype, double* asym, otk_stream_t stream) {
  OTK_TRY(require_device());
  OTK_REQUIRE(L > 0 && dim > 0 && a && asym, "asymmetry: bad arguments");
  asymmetry_kernel<<<(unsigned)L, 256, 0, as_stream(stream)>>>(a, L, dim, dtype, asym);
  OTK_LAUNCH_CHECK();
  return OTK_OK;
}
