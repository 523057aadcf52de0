// K7 on tcgen05: Y = (X - 1 mean_s^T) T^T + 1 mean_t^T  with fp32-accurate 3xTF32 arithmetic.
// Reference: apply_transport, ot/w2_utils.py:517-520 (B broadcast fp64 mat-vecs) - here one streaming GEMM.
//
// X is streamed ONCE from HBM: TMA drops the raw fp32 tile (128 latents x 32 features, 128B swizzle) into shared
// memory, four converter warps centre it (x - mean_s) and split it in place into TF32 hi / lo planes, and one thread
// issues tcgen05.mma kind::tf32 (lo*hi' + hi*lo' + hi*hi') against the pre-split rows of T, accumulating a
// 128 x 128 fp32 tile in TMEM.  The converter warps then become the epilogue (TMEM -> + mean_t -> global).
//
// CTA = 192 threads: warp 0 TMA producer, warp 1 TMEM alloc + MMA issuer, warps 2-5 converter / epilogue.
#include <cuda.h>

#include "apply_umma.cuh"
#include "otk_ptx.cuh"
#include "tensormap.cuh"

namespace otk {

constexpr int AP_BM = 128, AP_BN = 128, AP_BK = 32, AP_STAGES = 3, AP_THREADS = 320, AP_ACC = 2;
constexpr int AP_TILE = AP_BM * AP_BK * 4;          // 16 KiB
constexpr int AP_STAGE = 4 * AP_TILE;               // X hi (raw in place), X lo, T hi, T lo
constexpr int AP_SMEM = AP_STAGES * AP_STAGE + 1024 + 256;

// Persistent: one CTA per SM walks the (row tile, column tile) list; the smem ring keeps streaming across tiles and the
// two TMEM accumulators let the epilogue of tile i overlap the main loop of tile i+1.
// warp 0 TMA producer | warp 1 TMEM alloc + MMA issuer | warps 2-5 converter | warps 6-9 epilogue
__global__ void __launch_bounds__(AP_THREADS, 1)
apply_umma_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapT_hi,
                  const __grid_constant__ CUtensorMap mapT_lo, const float* __restrict__ mean_s,
                  const float* __restrict__ mean_t, float* __restrict__ y, int rows, int dim, int m_tiles, int n_tiles,
                  int total_tiles) {
  using namespace ptx;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + AP_STAGES * AP_STAGE);   // TMA landed
  uint64_t* ready = full + AP_STAGES;                                           // converted, MMA may read
  uint64_t* empty = ready + AP_STAGES;                                          // MMA done with the stage
  uint64_t* acc_full = empty + AP_STAGES;                                       // accumulator complete
  uint64_t* acc_empty = acc_full + AP_ACC;                                      // accumulator drained by the epilogue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + AP_ACC);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int num_k = (dim + AP_BK - 1) / AP_BK;
  const int tiles_per_l = m_tiles * n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapX); tma_prefetch_desc(&mapT_hi); tma_prefetch_desc(&mapT_lo);
    for (int s = 0; s < AP_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&ready[s], 128); mbar_init(&empty[s], 1); }
    for (int a = 0; a < AP_ACC; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 128); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, AP_ACC * AP_BN); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int l = tile / tiles_per_l, rem = tile % tiles_per_l;
        const int m0 = (rem / n_tiles) * AP_BM, n0 = (rem % n_tiles) * AP_BN;
        for (int kt = 0; kt < num_k; ++kt, ++it) {
          const int s = it % AP_STAGES;
          mbar_wait(&empty[s], ((it / AP_STAGES) & 1) ^ 1);
          uint8_t* st = smem + s * AP_STAGE;
          mbar_arrive_expect_tx(&full[s], 3u * AP_TILE);
          tma_load_3d(st, &mapX, kt * AP_BK, m0, l, &full[s]);
          tma_load_3d(st + 2 * AP_TILE, &mapT_hi, kt * AP_BK, n0, l, &full[s]);
          tma_load_3d(st + 3 * AP_TILE, &mapT_lo, kt * AP_BK, n0, l, &full[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = idesc_tf32(AP_BM, AP_BN, 0, 0);
      int it = 0, ti = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti) {
        const int a = ti % AP_ACC;
        mbar_wait(&acc_empty[a], ((ti / AP_ACC) & 1) ^ 1);
        tc_fence_after();
        const uint32_t acc = tmem_base + a * AP_BN;
        for (int kt = 0; kt < num_k; ++kt, ++it) {
          const int s = it % AP_STAGES;
          mbar_wait(&ready[s], (it / AP_STAGES) & 1);
          tc_fence_after();
          const uint32_t base = smem_u32(smem + s * AP_STAGE);
#pragma unroll
          for (int kk = 0; kk < AP_BK / 8; ++kk) {
            const uint64_t x_hi = smem_desc_sw128(base + kk * 32, 16, 1024);
            const uint64_t x_lo = smem_desc_sw128(base + AP_TILE + kk * 32, 16, 1024);
            const uint64_t t_hi = smem_desc_sw128(base + 2 * AP_TILE + kk * 32, 16, 1024);
            const uint64_t t_lo = smem_desc_sw128(base + 3 * AP_TILE + kk * 32, 16, 1024);
            umma_tf32(acc, x_lo, t_hi, idesc, (kt | kk) != 0);
            umma_tf32(acc, x_hi, t_lo, idesc, 1);
            umma_tf32(acc, x_hi, t_hi, idesc, 1);
          }
          umma_commit(&empty[s]);
        }
        umma_commit(&acc_full[a]);
      }
    }
  } else if (warp < 6) {
    // ===== converter: thread r owns row r of the 128 x 32 tile (eight 16-byte chunks, XOR-swizzled by r % 8) =====
    const int r = (warp - 2) * 32 + lane;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int l = tile / tiles_per_l;
      const float* ms = mean_s + (int64_t)l * dim;
      for (int kt = 0; kt < num_k; ++kt, ++it) {
        const int s = it % AP_STAGES;
        mbar_wait(&full[s], (it / AP_STAGES) & 1);
        uint8_t* hi_row = smem + s * AP_STAGE + r * 128;
        uint8_t* lo_row = hi_row + AP_TILE;
        const int k0 = kt * AP_BK;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int phys = (c ^ (r & 7)) * 16;
          float4 x = *reinterpret_cast<const float4*>(hi_row + phys);
          const int k = k0 + c * 4;
          float4 mu = make_float4(0.f, 0.f, 0.f, 0.f);
          if (k + 3 < dim) mu = *reinterpret_cast<const float4*>(ms + k);
          else {
            if (k < dim) mu.x = ms[k];
            if (k + 1 < dim) mu.y = ms[k + 1];
            if (k + 2 < dim) mu.z = ms[k + 2];
          }
          float4 h, lo;
          split_tf32(x.x - mu.x, h.x, lo.x);
          split_tf32(x.y - mu.y, h.y, lo.y);
          split_tf32(x.z - mu.z, h.z, lo.z);
          split_tf32(x.w - mu.w, h.w, lo.w);
          *reinterpret_cast<float4*>(hi_row + phys) = h;
          *reinterpret_cast<float4*>(lo_row + phys) = lo;
        }
        fence_proxy_async_smem();   // make the generic-proxy writes visible to the tensor core (async proxy)
        mbar_arrive(&ready[s]);
      }
    }
  } else {
    // ===== epilogue: TMEM -> + mean_t -> global; overlaps the next tile's main loop =====
    const int q = warp % 4;
    const bool vec_ok = (dim % 4 == 0);
    int ti = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti) {
      const int l = tile / tiles_per_l, rem = tile % tiles_per_l;
      const int m0 = (rem / n_tiles) * AP_BM, n0 = (rem % n_tiles) * AP_BN;
      const int a = ti % AP_ACC;
      const int m = m0 + q * 32 + lane;
      const float* mt = mean_t + (int64_t)l * dim;
      float* dst = y + ((int64_t)l * rows + m) * dim;
      mbar_wait(&acc_full[a], (ti / AP_ACC) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < AP_BN; c0 += 32) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + a * AP_BN + c0, v);
        tmem_ld_wait();
        if (m < rows) {
#pragma unroll
          for (int j4 = 0; j4 < 32; j4 += 4) {
            const int n = n0 + c0 + j4;
            if (vec_ok && n + 3 < dim) {
              const float4 b = *reinterpret_cast<const float4*>(mt + n);
              *reinterpret_cast<float4*>(dst + n) = make_float4(v[j4] + b.x, v[j4 + 1] + b.y, v[j4 + 2] + b.z, v[j4 + 3] + b.w);
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (n + j < dim) dst[n + j] = v[j4 + j] + mt[n + j];
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[a]);
    }
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, AP_ACC * AP_BN); }
}

__global__ void split_matrix_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ hi, float* __restrict__ lo) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    float h, l;
    ptx::split_tf32(x[e], h, l);
    hi[e] = h;
    lo[e] = l;
  }
}

// Thi: in-place over T32 is not allowed (T32 stays the caller's); uses T32 -> (Tlo_scratch as lo, and the hi plane is
// written over ... ) - we need two planes: hi goes to `Thi_scratch`, lo to `Tlo_scratch`.
int apply_umma_try(const float* x, int64_t L, int64_t rows, int64_t dim, const float* ms32, const float* mt32,
                   const float* T32, float* Thi_scratch, float* Tlo_scratch, float* y, cudaStream_t st) {
  if (dim < 64 || dim % 4 != 0 || rows < 1 || L > 65535 || rows > INT32_MAX || dim > INT32_MAX) return 0;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & 15)) return 0;
  if (!tensormap_encoder()) return 0;
  const int64_t n = L * dim * dim;
  int64_t blocks = ceil_div(n, 256);
  if (blocks > (int64_t)sm_count() * 8) blocks = (int64_t)sm_count() * 8;
  split_matrix_kernel<<<(unsigned)blocks, 256, 0, st>>>(T32, n, Thi_scratch, Tlo_scratch);
  OTK_LAUNCH_CHECK();
  CUtensorMap mX, mTh, mTl;
  if (!encode_map_f32_3d(&mX, x, dim, rows, L, dim, rows * dim, 32, AP_BM)) return 0;
  if (!encode_map_f32_3d(&mTh, Thi_scratch, dim, dim, L, dim, dim * dim, 32, AP_BN)) return 0;
  if (!encode_map_f32_3d(&mTl, Tlo_scratch, dim, dim, L, dim, dim * dim, 32, AP_BN)) return 0;
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    OTK_CUDA(cudaFuncSetAttribute(apply_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AP_SMEM));
    attr_set[dev] = true;
  }
  const int64_t m_tiles = ceil_div(rows, AP_BM), n_tiles = ceil_div(dim, AP_BN);
  const int64_t total = L * m_tiles * n_tiles;
  if (total > INT32_MAX) return 0;
  const unsigned grid = (unsigned)(total < sm_count() ? total : sm_count());
  apply_umma_kernel<<<grid, AP_THREADS, AP_SMEM, st>>>(mX, mTh, mTl, ms32, mt32, y, (int)rows, (int)dim, (int)m_tiles,
                                                      (int)n_tiles, (int)total);
  OTK_LAUNCH_CHECK();
  return 1;
}

}  // namespace otk
