// K1 on tcgen05: streaming sum (x-c)(x-c)^T and sum (x-c) with fp32-accurate 3xTF32 arithmetic.
// Reference: GaussianModel._stats (gaussian_model.py:144-157: einsum SYRK + column sum) and fid.py:103-104.
//
// The latents X [rows, dim] are row-major, so BOTH operands of X^T X are MN-major: one TMA-loaded tile of
// 32 latents x 128 features serves as the A operand of one output tile and the B operand of another.  MN-major TF32
// operands require the 32-byte-atom 128B swizzle (TMA SWIZZLE_128B_ATOM_32B / UMMA layout SWIZZLE_128B_BASE32B).
//
// Pivot shift: the converter warps subtract a per-feature pivot c (an estimate of the mean taken from the head of the
// batch) before the TF32 hi/lo split, so the fp32 tensor-core accumulation works on centred data and the
// cancellation in  cov = Sxx/n - mu mu^T  (ot/matrix_utils.py:155-157) is not amplified; the raw sums the reference
// keeps in its buffers are rebuilt exactly in fp64 by the merge kernel:
//     sum x = S' + n c ,   sum x x^T = P' + c S'^T + S' c^T + n c c^T .
//
// One CTA = one upper-triangular pair of 128-wide feature tiles (ti <= tj) x one contiguous range of rows.
// CTA = 576 threads: warp 0 TMA, warp 1 TMEM alloc + MMA issuer, warps 2-5 convert the i-tile (thread <-> feature
// column, which also yields the column sums for free), warps 6-9 convert the j-tile, warps 10-17 are the epilogue.
#include <cuda.h>

#include "otk_ptx.cuh"
#include "stats_umma.cuh"
#include "tensormap.cuh"

namespace otk {

constexpr int SU_T = 128, SU_BK = 32, SU_STAGES = 3, SU_ACC = 2;
constexpr int SU_THREADS = 64 + 256 + 256;            // TMA, MMA | 8 converter warps | 8 epilogue warps
constexpr int SU_TILE = SU_T * SU_BK * 4;            // 16 KiB: four 32-feature slabs of 32 rows x 128 B
constexpr int SU_STAGE = 4 * SU_TILE;                // i hi (raw in place), i lo, j hi, j lo
constexpr int SU_SMEM = SU_STAGES * SU_STAGE + 1024 + 256;
constexpr int SU_SUB = 1024;                         // rows accumulated in TMEM (fp32, truncating adder) per sub-chunk

// Persistent: CTA b works on unit (l, tile pair) = b / ctas_per_unit and the row range part = b % ctas_per_unit.
// The k-step ring (TMA -> converters -> MMA) streams over the whole range; every SU_SUB rows the TMEM accumulator is
// handed to the epilogue warps, which add it into fp32 REGISTER accumulators (round-to-nearest adds) while the next
// sub-chunk runs into the other TMEM buffer.  One fp64 atomic flush per CTA at the very end.
__global__ void __launch_bounds__(SU_THREADS, 1)
stats_umma_kernel(const __grid_constant__ CUtensorMap mapX, const float* __restrict__ pivot, int rows, int dim,
                  int rows_per_cta, int ctas_per_unit, int n_tiles, int n_pairs, double* __restrict__ ws_cov,
                  double* __restrict__ ws_sum) {
  using namespace ptx;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + SU_STAGES * SU_STAGE);
  uint64_t* ready = full + SU_STAGES;
  uint64_t* empty = ready + SU_STAGES;
  uint64_t* acc_full = empty + SU_STAGES;
  uint64_t* acc_empty = acc_full + SU_ACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + SU_ACC);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int unit = blockIdx.x / ctas_per_unit, part = blockIdx.x % ctas_per_unit;
  const int l = unit / n_pairs;
  int p = unit % n_pairs, ti = 0;
  while (p >= n_tiles - ti) { p -= n_tiles - ti; ++ti; }
  const int tj = ti + p;
  const bool diag = (ti == tj);
  const int r0 = part * rows_per_cta;
  const int r1 = min(rows, r0 + rows_per_cta);
  const int n_rows = max(0, r1 - r0);
  const int num_k = (n_rows + SU_BK - 1) / SU_BK;                 // k-steps of this CTA
  const int k_per_sub = SU_SUB / SU_BK;
  const int num_sub = (num_k + k_per_sub - 1) / k_per_sub;
  const int i0 = ti * SU_T, j0 = tj * SU_T;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapX);
    for (int s = 0; s < SU_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&ready[s], diag ? 128 : 256); mbar_init(&empty[s], 1); }
    for (int a = 0; a < SU_ACC; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 256); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, SU_ACC * SU_T); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int kt = 0; kt < num_k; ++kt) {
        const int s = kt % SU_STAGES;
        mbar_wait(&empty[s], ((kt / SU_STAGES) & 1) ^ 1);
        uint8_t* st = smem + s * SU_STAGE;
        mbar_arrive_expect_tx(&full[s], (diag ? 1u : 2u) * SU_TILE);
        const int k0 = r0 + kt * SU_BK;
#pragma unroll
        for (int sl = 0; sl < 4; ++sl) {
          tma_load_3d(st + sl * 4096, &mapX, i0 + 32 * sl, k0, l, &full[s]);
          if (!diag) tma_load_3d(st + 2 * SU_TILE + sl * 4096, &mapX, j0 + 32 * sl, k0, l, &full[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = idesc_tf32(SU_T, SU_T, 1, 1);   // both operands MN-major
      for (int sub = 0; sub < num_sub; ++sub) {
        const int a = sub % SU_ACC;
        mbar_wait(&acc_empty[a], ((sub / SU_ACC) & 1) ^ 1);
        tc_fence_after();
        const uint32_t acc = tmem_base + a * SU_T;
        const int kt_end = min(num_k, (sub + 1) * k_per_sub);
        for (int kt = sub * k_per_sub; kt < kt_end; ++kt) {
          const int s = kt % SU_STAGES;
          mbar_wait(&ready[s], (kt / SU_STAGES) & 1);
          tc_fence_after();
          const uint32_t base = smem_u32(smem + s * SU_STAGE);
          const uint32_t jbase = diag ? base : base + 2 * SU_TILE;
          const bool first = (kt == sub * k_per_sub);
#pragma unroll
          for (int kk = 0; kk < SU_BK / 8; ++kk) {
            const uint64_t a_hi = smem_desc_mn_tf32(base + kk * 1024, 4096);
            const uint64_t a_lo = smem_desc_mn_tf32(base + SU_TILE + kk * 1024, 4096);
            const uint64_t b_hi = smem_desc_mn_tf32(jbase + kk * 1024, 4096);
            const uint64_t b_lo = smem_desc_mn_tf32(jbase + SU_TILE + kk * 1024, 4096);
            umma_tf32(acc, a_lo, b_hi, idesc, !(first && kk == 0));
            umma_tf32(acc, a_hi, b_lo, idesc, 1);
            umma_tf32(acc, a_hi, b_hi, idesc, 1);
          }
          umma_commit(&empty[s]);
        }
        umma_commit(&acc_full[a]);
      }
    }
  } else if (warp < 10) {
    // ===== converters: thread <-> feature column t of its tile (i-tile: warps 2-5, j-tile: warps 6-9) =====
    const bool is_j = warp >= 6;
    const int t = (warp - (is_j ? 6 : 2)) * 32 + lane;
    if (!(is_j && diag)) {
      const int col = (is_j ? j0 : i0) + t;
      const float c = col < dim ? pivot[(int64_t)l * dim + col] : 0.f;
      const uint32_t slab_off = (uint32_t)(t / 32) * 4096 + (is_j ? 2u * SU_TILE : 0u);
      const uint32_t cc = (uint32_t)(t % 32) / 8, within = (uint32_t)(t % 8) * 4;
      double colsum = 0.0;
      for (int kt = 0; kt < num_k; ++kt) {
        const int s = kt % SU_STAGES;
        mbar_wait(&full[s], (kt / SU_STAGES) & 1);
        uint8_t* hi_t = smem + s * SU_STAGE + slab_off;
        uint8_t* lo_t = hi_t + SU_TILE;
        const int valid = min(SU_BK, r1 - (r0 + kt * SU_BK));   // rows past the range / batch end contribute nothing
        float part_sum = 0.f;
#pragma unroll 8
        for (int r = 0; r < SU_BK; ++r) {
          const uint32_t off = (uint32_t)r * 128 + ((cc ^ (uint32_t)(r & 3)) * 32) + within;
          float x = *reinterpret_cast<const float*>(hi_t + off);
          x = (r < valid && col < dim) ? x - c : 0.f;
          float h, lo;
          split_tf32(x, h, lo);
          *reinterpret_cast<float*>(hi_t + off) = h;
          *reinterpret_cast<float*>(lo_t + off) = lo;
          part_sum += x;
        }
        colsum += (double)part_sum;
        fence_proxy_async_smem();
        mbar_arrive(&ready[s]);
      }
      if (!is_j && diag && col < dim && num_k > 0) atomicAdd(&ws_sum[(int64_t)l * dim + col], colsum);
    }
  } else {
    // ===== epilogue: 8 warps; warp -> (TMEM lane quarter, column half); fp32 register accumulation across sub-chunks ====
    const int q = warp % 4, half = (warp - 10) / 4;
    const int gi = i0 + q * 32 + lane;
    float acc[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) acc[j] = 0.f;
    for (int sub = 0; sub < num_sub; ++sub) {
      const int a = sub % SU_ACC;
      mbar_wait(&acc_full[a], (sub / SU_ACC) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + a * SU_T + half * 64;
      {
        float v[32];
        tmem_ld32(taddr, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] += v[j];
        tmem_ld32(taddr + 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[32 + j] += v[j];
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[a]);
    }
    if (num_k > 0 && gi < dim) {
      double* cov = ws_cov + (int64_t)l * dim * dim;
#pragma unroll
      for (int j = 0; j < 64; ++j) {
        const int gj = j0 + half * 64 + j;
        // diagonal tile pairs: the merge kernel reads the (min, max) element, so only gi <= gj is needed
        if (gj < dim && (!diag || gi <= gj)) atomicAdd(&cov[(int64_t)gi * dim + gj], (double)acc[j]);
      }
    }
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, SU_ACC * SU_T); }
}

// pivot[l, :] = mean of the first min(rows, 64) latents of the batch
__global__ void pivot_kernel(const float* __restrict__ x, int64_t rows, int64_t dim, int64_t row_stride, int64_t batch_stride,
                             float* __restrict__ pivot) {
  const int64_t l = blockIdx.y;
  const int64_t col = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (col >= dim) return;
  const int64_t n = rows < 64 ? rows : 64;
  float acc = 0.f;
  for (int64_t r = 0; r < n; ++r) acc += x[l * batch_stride + r * row_stride + col];
  pivot[l * dim + col] = acc / (float)n;
}

// ws holds P' = sum (x-c)(x-c)^T (upper tile pairs) and S' = sum (x-c); rebuild the raw sums in place (fp64)
__global__ void unshift_kernel(double* __restrict__ ws_cov, double* __restrict__ ws_sum, const float* __restrict__ pivot,
                               int64_t L, int64_t dim, double rows) {
  const int64_t total = L * dim * dim;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t l = e / (dim * dim), r = e % (dim * dim), i = r / dim, j = r % dim;
    if (i > j) continue;  // the merge kernel only reads (min, max)
    const double ci = pivot[l * dim + i], cj = pivot[l * dim + j];
    const double si = ws_sum[l * dim + i], sj = ws_sum[l * dim + j];
    ws_cov[e] += ci * sj + si * cj + rows * ci * cj;
  }
}
__global__ void unshift_sum_kernel(double* __restrict__ ws_sum, const float* __restrict__ pivot, int64_t n, double rows) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e < n) ws_sum[e] += rows * (double)pivot[e];
}

size_t stats_umma_extra_workspace(int64_t L, int64_t dim) { return align_up((size_t)L * dim * 4, 256) + 256; }

int stats_umma_try(const float* x, int64_t L, int64_t rows, int64_t dim, int64_t row_stride, int64_t batch_stride,
                   double* ws_cov, double* ws_sum, Arena& ar, cudaStream_t st, int* tile) {
  if (dim < 64 || dim % 4 != 0 || row_stride % 4 != 0 || batch_stride % 4 != 0) return 0;
  if (rows < 1 || rows > INT32_MAX || dim > INT32_MAX || L > 65535) return 0;
  if (reinterpret_cast<uintptr_t>(x) & 15) return 0;
  if (!tensormap_encoder()) return 0;
  float* pivot = ar.take<float>((size_t)L * dim);
  if (!ar.ok()) return OTK_ERR_WORKSPACE;
  pivot_kernel<<<dim3((unsigned)ceil_div(dim, 128), (unsigned)L), 128, 0, st>>>(x, rows, dim, row_stride, batch_stride, pivot);
  OTK_LAUNCH_CHECK();
  CUtensorMap mX;
  if (!encode_map_f32_3d(&mX, x, dim, rows, L, row_stride, batch_stride, 32, SU_BK, /*atom32=*/true)) return 0;
  const int n_tiles = (int)ceil_div(dim, SU_T);
  const int64_t pairs = (int64_t)n_tiles * (n_tiles + 1) / 2;
  const int64_t units = pairs * L;
  // one CTA per SM when the units fit: each unit's rows are split over ctas_per_unit CTAs (multiples of 32 rows)
  int64_t cpu = units >= sm_count() ? 1 : sm_count() / units;
  int64_t rows_per_cta = ceil_div(ceil_div(rows, cpu), SU_BK) * SU_BK;
  if (rows_per_cta < 256) rows_per_cta = 256;                       // tiny batches: fewer, fuller CTAs
  cpu = ceil_div(rows, rows_per_cta);
  if (units * cpu > INT32_MAX) return 0;
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    OTK_CUDA(cudaFuncSetAttribute(stats_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SU_SMEM));
    attr_set[dev] = true;
  }
  stats_umma_kernel<<<(unsigned)(units * cpu), SU_THREADS, SU_SMEM, st>>>(mX, pivot, (int)rows, (int)dim, (int)rows_per_cta,
                                                                         (int)cpu, n_tiles, (int)pairs, ws_cov, ws_sum);
  OTK_LAUNCH_CHECK();
  int64_t blocks = ceil_div(L * dim * dim, 256);
  if (blocks > (int64_t)sm_count() * 16) blocks = (int64_t)sm_count() * 16;
  unshift_kernel<<<(unsigned)blocks, 256, 0, st>>>(ws_cov, ws_sum, pivot, L, dim, (double)rows);
  unshift_sum_kernel<<<(unsigned)ceil_div(L * dim, 256), 256, 0, st>>>(ws_sum, pivot, L * dim, (double)rows);
  count_launch(1);
  OTK_LAUNCH_CHECK();
  *tile = SU_T;
  return 1;
}

}  // namespace otk
