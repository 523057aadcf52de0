// Streaming k-means step of the codebook model: nearest-codeword assignment and the per-codeword
// sums, without the [B, K] energy / one-hot matrices.
// Reference: MixtureMixin.assign + kmean_iteration, ot/distribution_models/base.py:206-253 with the Euclidean energy
// 1 / (|x - c|_2 + 1e-8) of CodebookModel.energy (codebook_model.py:155-160) in 'argmax' mode: softmax is monotone and the
// energy is a decreasing function of the distance, so the one-hot weights select argmin_k |x_b - c_k| (first index on
// ties, as torch.argmax), `weights.sum(-2)` is the per-codeword count and `weights^T @ samples` the per-codeword sum.
#include "gemm.cuh"
#include "otk_common.cuh"

namespace otk {

__global__ void km_sqnorm_kernel(const float* __restrict__ x, int64_t n, int64_t d, float* __restrict__ out) {
  const int lane = threadIdx.x % 32;
  const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / 32;
  if (row >= n) return;
  float acc = 0;
  for (int64_t k = lane; k < d; k += 32) { const float v = x[row * d + k]; acc = fmaf(v, v, acc); }
  acc = warp_sum(acc);
  if (lane == 0) out[row] = acc;
}

// best[l, b] = min over this CTA's codeword range of (bits(|x_b - c_k|^2) << 32 | k): 64 x 64 FFMA tiles of x . c^T,
// one packed 64-bit atomicMin per row and CTA (squared distances are >= 0, so their bit patterns order like the values).
constexpr int KM_T = 64, KM_BK = 16;
__global__ void __launch_bounds__(256)
km_nearest_kernel(const float* __restrict__ x, const float* __restrict__ cb, const float* __restrict__ nx,
                  const float* __restrict__ nc, int64_t B, int64_t K, int64_t d, int64_t k_per_cta,
                  unsigned long long* __restrict__ best) {
  __shared__ float As[KM_BK][KM_T + 4], Bs[KM_BK][KM_T + 4];
  const int64_t l = blockIdx.z;
  const float* xl = x + l * B * d;
  const float* cl = cb + l * K * d;
  const int64_t i0 = (int64_t)blockIdx.y * KM_T;
  const int64_t k_lo = (int64_t)blockIdx.x * k_per_cta, k_hi = min(K, k_lo + k_per_cta);
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  unsigned long long mine[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) mine[i] = ~0ull;
  for (int64_t j0 = k_lo; j0 < k_hi; j0 += KM_T) {
    float acc[4][4] = {};
    for (int64_t k0 = 0; k0 < d; k0 += KM_BK) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int e = tid + r * 256, kk = e % KM_BK, mm = e / KM_BK;
        const int64_t k = k0 + kk;
        As[kk][mm] = (i0 + mm < B && k < d) ? xl[(i0 + mm) * d + k] : 0.f;
        Bs[kk][mm] = (j0 + mm < k_hi && k < d) ? cl[(j0 + mm) * d + k] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < KM_BK; ++kk) {
        float a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t gi = i0 + ty * 4 + i;
      if (gi >= B) continue;
      const float nxi = nx[l * B + gi];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t gj = j0 + tx * 4 + j;
        if (gj >= k_hi) continue;
        const float sq = fmaxf(nxi + nc[l * K + gj] - 2.f * acc[i][j], 0.f);
        const unsigned long long key = ((unsigned long long)__float_as_uint(sq) << 32) | (unsigned long long)gj;
        mine[i] = key < mine[i] ? key : mine[i];
      }
    }
  }
  // the 16 threads that share a row are 16 consecutive lanes
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    unsigned long long v = mine[i];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      const unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
      v = w < v ? w : v;
    }
    const int64_t gi = i0 + ty * 4 + i;
    if (tx == 0 && gi < B) atomicMin(&best[l * B + gi], v);
  }
}

// tensor-core path: `scores` [rows, K] holds -2 x_b . c_k + |c_k|^2 (epilogue of the tcgen05 product); one warp per row
// finds min_k of (bits(max(|x_b|^2 + score, 0)) << 32 | k) - the key of km_nearest_kernel, so ties resolve identically
__global__ void km_rowmin_kernel(const float* __restrict__ scores, const float* __restrict__ nx, int64_t rows, int64_t K,
                                 unsigned long long* __restrict__ best) {
  const int lane = threadIdx.x % 32;
  const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / 32;
  if (row >= rows) return;
  const float4* p = reinterpret_cast<const float4*>(scores + row * K);
  const float nxi = nx[row];
  unsigned long long m = ~0ull;
  for (int64_t c = lane; c < K / 4; c += 32) {
    const float4 v = p[c];
    const float s4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float sq = fmaxf(nxi + s4[j], 0.f);
      const unsigned long long key = ((unsigned long long)__float_as_uint(sq) << 32) | (unsigned long long)(4 * c + j);
      m = key < m ? key : m;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long w = __shfl_xor_sync(0xffffffffu, m, o);
    m = w < m ? w : m;
  }
  if (lane == 0) best[row] = m;
}

// one warp per sample: index out, count and feature sums of its codeword (atomics in the buffer dtype)
__global__ void km_scatter_kernel(const float* __restrict__ x, const unsigned long long* __restrict__ best, int64_t L, int64_t B,
                                  int64_t K, int64_t d, int64_t* __restrict__ index, void* wsum, void* ssum, int dt) {
  const int lane = threadIdx.x % 32;
  const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / 32;
  if (row >= L * B) return;
  const int64_t l = row / B;
  const int64_t k = (int64_t)(best[row] & 0xffffffffull);
  if (lane == 0 && index) index[row] = k;
  if (!ssum) return;
  const float* xr = x + row * d;
  if (dt == OTK_F64) {
    double* dst = static_cast<double*>(ssum) + (l * K + k) * d;
    for (int64_t j = lane; j < d; j += 32) atomicAdd(dst + j, (double)xr[j]);
    if (lane == 0) atomicAdd(static_cast<double*>(wsum) + l * K + k, 1.0);
  } else {
    float* dst = static_cast<float*>(ssum) + (l * K + k) * d;
    for (int64_t j = lane; j < d; j += 32) atomicAdd(dst + j, xr[j]);
    if (lane == 0) atomicAdd(static_cast<float*>(wsum) + l * K + k, 1.f);
  }
}

}  // namespace otk
using namespace otk;

extern "C" size_t otk_kmeans_assign_workspace_bytes(int64_t L, int64_t B, int64_t K) {
  return align_up((size_t)L * B * 8, 256) + align_up((size_t)L * B * 4, 256) + align_up((size_t)L * K * 4, 256) + 1024;
}

// tensor-core path: rows per chunk such that the fp32 score chunk [rows, K] (48 MB) stays resident in the 126 MB L2
// between the product that writes it and the row pass that reads it
constexpr size_t KM_SCORE_BYTES = (size_t)48 << 20;
static int64_t km_chunk_rows(int64_t B, int64_t K) {
  int64_t rc = (int64_t)(KM_SCORE_BYTES / 4) / K / 128 * 128;
  if (rc < 128) rc = 128;
  return rc < B ? rc : B;
}
static bool km_umma_eligible(const float* x, const float* cb, int64_t B, int64_t K, int64_t d) {
  static const bool on = [] { const char* e = getenv("OTK_KMEANS_TC"); return !(e && e[0] == '0'); }();   // tuning aid
  // measured (B200): 4096 x 8192, d = 256: 0.90 -> 0.48 ms; at d = 128 the product's one-row-per-thread epilogue costs what the
  // short contraction saves (65536 x 8192: 5.3 ms on the FFMA tiles, 5.7 ms here), so narrow codewords stay on the FFMA tiles
  return on && B >= 256 && K >= 256 && d >= 192 && d % 4 == 0 && K % 4 == 0 && B * K >= ((int64_t)1 << 22) &&
         reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(cb) % 16 == 0;
}
extern "C" size_t otk_kmeans_workspace_bytes(int64_t L, int64_t B, int64_t K, int64_t dim) {
  const int64_t rc = km_chunk_rows(B, K);
  return otk_kmeans_assign_workspace_bytes(L, B, K) + align_up((size_t)rc * K * 4, 256) +
         align_up((size_t)2 * (rc + K) * dim * 4, 256) + 512;
}

extern "C" int otk_kmeans_assign(const float* x, int64_t L, int64_t B, int64_t K, int64_t dim, const float* codebook,
                                 int64_t* index, void* weights_sum, void* samples_sum, int buf_dtype, void* workspace,
                                 size_t workspace_bytes, otk_stream_t stream) {
  OTK_TRY(require_device());
  OTK_REQUIRE(x && codebook && L > 0 && B > 0 && K > 0 && dim > 0 && L <= 65535, "kmeans_assign: bad arguments");
  OTK_REQUIRE((weights_sum == nullptr) == (samples_sum == nullptr), "kmeans_assign: weights_sum and samples_sum go together");
  OTK_REQUIRE(index || samples_sum, "kmeans_assign: nothing to compute");
  OTK_REQUIRE(K < (1ll << 32), "kmeans_assign: too many codewords");
  if (!workspace || workspace_bytes < otk_kmeans_assign_workspace_bytes(L, B, K)) return OTK_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  Arena ar(workspace, workspace_bytes);
  unsigned long long* best = ar.take<unsigned long long>((size_t)L * B);
  float* nx = ar.take<float>((size_t)L * B);
  float* nc = ar.take<float>((size_t)L * K);
  OTK_CUDA(cudaMemsetAsync(best, 0xff, (size_t)L * B * 8, st));
  km_sqnorm_kernel<<<(unsigned)ceil_div(L * B * 32, 256), 256, 0, st>>>(x, L * B, dim, nx);
  km_sqnorm_kernel<<<(unsigned)ceil_div(L * K * 32, 256), 256, 0, st>>>(codebook, L * K, dim, nc);
  count_launch(1);
  // split the codewords over enough CTAs to fill the machine (row tiles x L alone are few at batch sizes of a few hundred)
  const int64_t row_tiles = ceil_div(B, KM_T);
  int64_t splits = ceil_div((int64_t)sm_count() * 2, row_tiles * L);
  const int64_t k_tiles = ceil_div(K, KM_T);
  if (splits > k_tiles) splits = k_tiles;
  if (splits < 1) splits = 1;
  const int64_t k_per_cta = ceil_div(k_tiles, splits) * KM_T;
  dim3 grid((unsigned)ceil_div(K, k_per_cta), (unsigned)row_tiles, (unsigned)L);
  OTK_REQUIRE(grid.y <= 65535, "kmeans_assign: batch too large (split it)");
  // tensor cores when the caller brought the larger workspace: per leading index and row chunk, scores = -2 x . c^T + |c|^2
  // (3xTF32 product, bias epilogue), then one row pass; a tail of < 64 rows (and anything the product declines) takes the
  // FFMA tiles
  bool done = false;
  if (km_umma_eligible(x, codebook, B, K, dim) && workspace_bytes >= otk_kmeans_workspace_bytes(L, B, K, dim)) {
    const int64_t rc = km_chunk_rows(B, K);
    float* scores = ar.take<float>((size_t)rc * K);
    float* planes = ar.take<float>((size_t)2 * (rc + K) * dim);
    done = ar.ok();
    for (int64_t l = 0; l < L && done; ++l) {
      for (int64_t r0 = 0; r0 < B && done; r0 += rc) {
        const int64_t rows = B - r0 < rc ? B - r0 : rc;
        const float* xc = x + (l * B + r0) * dim;
        if (rows >= 64) {
          GemmArgs<float> g = nt_args(xc, codebook + l * K * dim, scores, rows, K, dim, dim, dim, K, 0, 0, 0, -2.f, 0.f);
          g.bias = nc + l * K;
          g.scratch = planes;
          const int r = gemm_umma_try(g, 1, 3, st);
          if (r < 0) return r;
          if (r == 0) { done = false; break; }          // declined: redo everything with the FFMA tiles below
          km_rowmin_kernel<<<(unsigned)ceil_div(rows * 32, 256), 256, 0, st>>>(scores, nx + l * B + r0, rows, K, best + l * B + r0);
          count_launch(1);
        } else {
          dim3 tail((unsigned)ceil_div(K, k_per_cta), (unsigned)ceil_div(rows, KM_T), 1);
          km_nearest_kernel<<<tail, 256, 0, st>>>(xc, codebook + l * K * dim, nx + l * B + r0, nc + l * K, rows, K, dim,
                                                   k_per_cta, best + l * B + r0);
        }
        OTK_LAUNCH_CHECK();
      }
    }
    if (!done) OTK_CUDA(cudaMemsetAsync(best, 0xff, (size_t)L * B * 8, st));   // partial results of a declined attempt
  }
  if (!done) {
    km_nearest_kernel<<<grid, 256, 0, st>>>(x, codebook, nx, nc, B, K, dim, k_per_cta, best);
    OTK_LAUNCH_CHECK();
  }
  if (samples_sum) {
    OTK_CUDA(cudaMemsetAsync(samples_sum, 0, (size_t)L * K * dim * dtype_size(buf_dtype), st));
    OTK_CUDA(cudaMemsetAsync(weights_sum, 0, (size_t)L * K * dtype_size(buf_dtype), st));
  }
  km_scatter_kernel<<<(unsigned)ceil_div(L * B * 32, 256), 256, 0, st>>>(x, best, L, B, K, dim, index, weights_sum, samples_sum,
                                                                        buf_dtype);
  OTK_LAUNCH_CHECK();
  return OTK_OK;
}
