// K7 on tcgen05: Y = (X - 1 mean_s^T) T^T + 1 mean_t^T with fp32-accurate 3xTF32 arithmetic.
// Reference: apply_transport, ot/w2_utils.py:517-520 (B broadcast fp64 mat-vecs).
//
// Data flow of one k-step (32 features of 128 latents per CTA):
//   TMA            raw X tile [128 x 32] fp32 -> shared memory (128B swizzle), T_hi / T_lo tiles -> shared memory
//   converter warps  read the raw tile (thread <-> latent row), subtract mean_s, split into TF32 hi / lo and write both
//                    planes straight into TENSOR MEMORY with tcgen05.st  (the A operand never returns to shared memory)
//   MMA thread       three kind::tf32 MMAs per K=8 slice with A from TMEM and B (= T planes) from shared memory:
//                    lo*hi + hi*lo + hi*hi into one fp32 TMEM accumulator
//   epilogue warps   tcgen05.ld -> + mean_t -> swizzled staging tile -> TMA store (clips the ragged edges)
// Keeping A in TMEM halves the shared-memory traffic of the main loop: an SS-mode 128x128x8 TF32 MMA reads 8 KB of
// operands per 64 cycles, i.e. the SM's whole 128 B/clk, so the TMA fills and the converter used to starve it.
//
// Two instantiations:
//   <CG=1, BN=128>  one CTA per 128 x 128 tile, two accumulators (epilogue overlaps the next tile)      - dim < 256
//   <CG=2, BN=256>  a CTA pair per 256 x 256 tile (tcgen05 cta_group::2): each CTA converts its own 128 latents and
//                   stages half of the T tile, so T is fetched once per 256 latents and X once per 256 outputs - dim >= 256
#include <cuda.h>

#include "apply_umma.cuh"
#include "otk_ptx.cuh"
#include "tensormap.cuh"

namespace otk {

constexpr int AP_BM = 128, AP_BK = 32;
constexpr int AP_XS = 3, AP_TS = 3, AP_AS = 4;       // ring depths: raw X (smem), T planes (smem), converted A (TMEM)
constexpr int AP_XTILE = AP_BM * AP_BK * 4;          // 16 KiB raw X tile
constexpr int AP_TPLANE = 128 * AP_BK * 4;           // 16 KiB: the 128 rows of one T plane a CTA stages per k-step
constexpr int AP_TSTAGE = 2 * AP_TPLANE;             // hi + lo
constexpr int AP_OUT = 32 * 32 * 4;                  // 4 KiB staging tile per TMA store
constexpr int AP_THREADS = 18 * 32;                  // TMA, MMA | 8 converter warps | 8 epilogue warps
constexpr int AP_ACOL0 = 256;                        // TMEM columns [0,256): accumulators, [256,512): A ring
constexpr int AP_SMEM = AP_XS * AP_XTILE + AP_TS * AP_TSTAGE + 16 * AP_OUT + 1024 + 512;

template <int CG, int BN>
__global__ void __launch_bounds__(AP_THREADS, 1)
apply_ts_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapT_hi,
                const __grid_constant__ CUtensorMap mapT_lo, const __grid_constant__ CUtensorMap mapY,
                const float* __restrict__ mean_s, const float* __restrict__ mean_t, int rows, int dim, int m_tiles,
                int n_tiles, int total_tiles, int dbg, const int* __restrict__ run_flag) {
  using namespace ptx;
  // fallback launch behind the FP16-split kernel (apply_h.cu): nothing to do unless that kernel raised its overflow flag
  if (run_flag != nullptr && *run_flag == 0) return;
  constexpr int NACC = 256 / BN;                      // accumulator buffers
  static_assert(BN / CG == 128, "each CTA stages 128 rows of the T tile");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* xraw = smem;
  uint8_t* tpl = xraw + AP_XS * AP_XTILE;
  uint8_t* outb = tpl + AP_TS * AP_TSTAGE;
  uint64_t* full_x = reinterpret_cast<uint64_t*>(outb + 16 * AP_OUT);   // TMA landed the raw X tile
  uint64_t* empty_x = full_x + AP_XS;                                  // converters have read it
  uint64_t* full_t = empty_x + AP_XS;                                  // T planes landed (leader: both CTAs' halves)
  uint64_t* empty_t = full_t + AP_TS;                                  // MMAs reading them retired
  uint64_t* ready_a = empty_t + AP_TS;                                 // A planes in TMEM written (leader: both CTAs)
  uint64_t* empty_a = ready_a + AP_AS;                                 // MMAs reading them retired
  uint64_t* acc_full = empty_a + AP_AS;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;
  const int group = CG == 2 ? blockIdx.x / 2 : blockIdx.x;           // tile-processing unit (CTA or CTA pair)
  const int n_groups = CG == 2 ? gridDim.x / 2 : gridDim.x;
  const int num_k = (dim + AP_BK - 1) / AP_BK;
  const int tiles_per_l = m_tiles * n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapX); tma_prefetch_desc(&mapT_hi); tma_prefetch_desc(&mapT_lo); tma_prefetch_desc(&mapY);
    for (int s = 0; s < AP_XS; ++s) { mbar_init(&full_x[s], 1); mbar_init(&empty_x[s], 8); }
    for (int s = 0; s < AP_TS; ++s) { mbar_init(&full_t[s], 1); mbar_init(&empty_t[s], 1); }
    for (int s = 0; s < AP_AS; ++s) { mbar_init(&ready_a[s], 8 * CG); mbar_init(&empty_a[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 8 * CG); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc_cg<CG>(tmem_slot, 512); tmem_relinquish_cg<CG>(); }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer: raw X tile of this CTA's 128 latents, and this CTA's 128 rows of the T hi/lo tiles.
    // The whole warp runs the loop and one elected lane issues: with warp-uniform control flow the descriptors stay in
    // uniform registers (a lane-0-only branch makes the compiler wrap every UTMALDG / UTCHMMA in an elect+R2UR loop,
    // measured at ~56 clk per instruction instead of ~20).
    {
      const uint32_t full_t_leader0 = CG == 2 ? map_to_cta(smem_u32(&full_t[0]), 0) : smem_u32(&full_t[0]);
      int it = 0;
      for (int tile = group; tile < total_tiles; tile += n_groups) {
        const int l = tile / tiles_per_l, rem = tile % tiles_per_l;
        const int m0 = (rem / n_tiles) * (AP_BM * CG) + (int)rank * AP_BM;
        const int n0 = (rem % n_tiles) * BN + (int)rank * 128;
        for (int kt = 0; kt < num_k; ++kt, ++it) {
          const int sx = it % AP_XS, st = it % AP_TS;
          mbar_wait(&empty_x[sx], ((it / AP_XS) & 1) ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&full_x[sx], AP_XTILE);
            tma_load_3d(xraw + sx * AP_XTILE, &mapX, kt * AP_BK, m0, l, &full_x[sx]);
          }
          __syncwarp();
          mbar_wait(&empty_t[st], ((it / AP_TS) & 1) ^ 1);
          uint8_t* td = tpl + st * AP_TSTAGE;
          if (elect_one()) {
            if constexpr (CG == 1) {
              mbar_arrive_expect_tx(&full_t[st], AP_TSTAGE);
              tma_load_3d(td, &mapT_hi, kt * AP_BK, n0, l, &full_t[st]);
              tma_load_3d(td + AP_TPLANE, &mapT_lo, kt * AP_BK, n0, l, &full_t[st]);
            } else {
              if (rank == 0) mbar_arrive_expect_tx(&full_t[st], 2 * AP_TSTAGE);   // both CTAs' bytes land on the leader's barrier
              const uint32_t bar = full_t_leader0 + st * 8;
              tma_load_3d_cg2(td, &mapT_hi, kt * AP_BK, n0, l, bar);
              tma_load_3d_cg2(td + AP_TPLANE, &mapT_lo, kt * AP_BK, n0, l, bar);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA of the pair only): warp-uniform loop, one elected lane issues =====
    if (rank == 0) {
      const uint32_t idesc = idesc_tf32(AP_BM * CG, BN, 0, 0);
      int it = 0, ti = 0;
      for (int tile = group; tile < total_tiles; tile += n_groups, ++ti) {
        const int a = ti % NACC;
        mbar_wait(&acc_empty[a], ((ti / NACC) & 1) ^ 1);
        tc_fence_after();
        const uint32_t acc = tmem_base + a * BN;
        for (int kt = 0; kt < num_k; ++kt, ++it) {
          const int st = it % AP_TS, sa = it % AP_AS;
          mbar_wait(&full_t[st], (it / AP_TS) & 1);
          mbar_wait(&ready_a[sa], (it / AP_AS) & 1);
          tc_fence_after();
          const uint32_t tb = smem_u32(tpl + st * AP_TSTAGE);
          const uint32_t ab = tmem_base + AP_ACOL0 + sa * 64;
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < AP_BK / 8; ++kk) {
              const uint64_t t_hi = smem_desc_sw128(tb + kk * 32, 16, 1024);
              const uint64_t t_lo = smem_desc_sw128(tb + AP_TPLANE + kk * 32, 16, 1024);
              umma_tf32_ts<CG>(acc, ab + 32 + kk * 8, t_hi, idesc, (kt | kk) != 0);   // lo * hi
              umma_tf32_ts<CG>(acc, ab + kk * 8, t_lo, idesc, 1);                     // hi * lo
              umma_tf32_ts<CG>(acc, ab + kk * 8, t_hi, idesc, 1);                     // hi * hi
            }
            umma_commit_cg<CG>(&empty_t[st]);
            umma_commit_cg<CG>(&empty_a[sa]);
            if (kt == num_k - 1) umma_commit_cg<CG>(&acc_full[a]);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp < 10) {
    // ===== converters: thread <-> latent row r (TMEM lane r); warps 2-5 take features 0-15 of the k-step, 6-9 take 16-31
    const int q = warp % 4, half = (warp - 2) / 4;
    const int r = q * 32 + lane;
    const uint32_t row_off = (uint32_t)r * 128;
    const uint32_t ready_addr = CG == 2 ? map_to_cta(smem_u32(&ready_a[0]), 0) : smem_u32(&ready_a[0]);
    const uint32_t xbase = smem_u32(xraw);
    int it = 0;
    for (int tile = group; tile < total_tiles; tile += n_groups) {
      const int l = tile / tiles_per_l;
      const float* ms = mean_s + (int64_t)l * dim;
      for (int kt = 0; kt < num_k; ++kt, ++it) {
        const int sx = it % AP_XS, sa = it % AP_AS;
        mbar_wait(&full_x[sx], (it / AP_XS) & 1);
        float4 x[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t chunk = (uint32_t)(half * 4 + c);
          x[c] = lds128(xbase + sx * AP_XTILE + row_off + ((chunk ^ (uint32_t)(r & 7)) * 16));
        }
        float hi[16], lo[16];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int k = kt * AP_BK + half * 16 + c * 4;
          float4 mu = make_float4(0.f, 0.f, 0.f, 0.f);
          if (k < dim) mu = __ldg(reinterpret_cast<const float4*>(ms + k));   // dim % 4 == 0
          split_tf32_fast(x[c].x - mu.x, hi[4 * c + 0], lo[4 * c + 0]);
          split_tf32_fast(x[c].y - mu.y, hi[4 * c + 1], lo[4 * c + 1]);
          split_tf32_fast(x[c].z - mu.z, hi[4 * c + 2], lo[4 * c + 2]);
          split_tf32_fast(x[c].w - mu.w, hi[4 * c + 3], lo[4 * c + 3]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_x[sx]);                 // raw tile consumed (values are in registers)
        mbar_wait(&empty_a[sa], ((it / AP_AS) & 1) ^ 1);
        tc_fence_after();
        const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + AP_ACOL0 + sa * 64 + half * 16;
        tmem_st16(ta, hi);
        tmem_st16(ta + 32, lo);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(ready_addr + sa * 8);
      }
    }
  } else {
    // ===== epilogue: TMEM -> + mean_t -> 32x32 swizzled staging tile -> TMA store.  Eight warps (two per TMEM lane
    // quarter, half of the tile's columns each): with a single 256-column accumulator the MMA thread waits for the drain,
    // so the drain is kept short; with two accumulators it overlaps the next tile's main loop.
    const int q = warp % 4, half = (warp - 10) / 4;
    constexpr int HC = BN / 2;                                  // columns drained by this warp
    const uint32_t stage0 = smem_u32(outb) + (uint32_t)(warp - 10) * 2 * AP_OUT;
    const uint32_t acc_empty_addr = CG == 2 ? map_to_cta(smem_u32(&acc_empty[0]), 0) : smem_u32(&acc_empty[0]);
    int ti = 0, nstore = 0;
    for (int tile = group; tile < total_tiles; tile += n_groups, ++ti) {
      const int l = tile / tiles_per_l, rem = tile % tiles_per_l;
      const int m0 = (rem / n_tiles) * (AP_BM * CG) + (int)rank * AP_BM + q * 32;
      const int n0 = (rem % n_tiles) * BN + half * HC;
      const int a = ti % NACC;
      const float* mt = mean_t + (int64_t)l * dim;
      mbar_wait(&acc_full[a], (ti / NACC) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < HC; c0 += 32, ++nstore) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + a * BN + half * HC + c0, v);
        tmem_ld_wait();
        if (c0 + 32 == HC) {                                      // this warp's share is read: hand the accumulator back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(acc_empty_addr + a * 8);
        }
        if (dbg & 1) continue;
        const uint32_t buf = stage0 + (uint32_t)(nstore & 1) * AP_OUT;
        if (elect_one()) tma_store_wait_read<1>();                // the store issued two chunks ago has read this buffer
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int n = n0 + c0 + c * 4;
          float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
          if (n < dim) b = __ldg(reinterpret_cast<const float4*>(mt + n));
          sts128(buf + (uint32_t)lane * 128 + (((uint32_t)c ^ (uint32_t)(lane & 7)) * 16),
                 make_float4(v[4 * c] + b.x, v[4 * c + 1] + b.y, v[4 * c + 2] + b.z, v[4 * c + 3] + b.w));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (n0 + c0 < dim && m0 < rows) {
          if (elect_one()) {
            tma_store_3d(&mapY, buf, n0 + c0, m0, l);
            tma_store_commit();
          }
        }
      }
    }
    if (elect_one()) tma_store_wait_all<0>();
    __syncwarp();
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc_cg<CG>(tmem_base, 512); }
}

__global__ void split_matrix_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ hi, float* __restrict__ lo) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    float h, l;
    ptx::split_tf32(x[e], h, l);
    hi[e] = h;
    lo[e] = l;
  }
}

int g_apply_dbg = 0;        // tuning aid: bit0 = epilogue drains TMEM but stores nothing
template <int CG, int BN>
static int launch_apply(const CUtensorMap& mX, const CUtensorMap& mTh, const CUtensorMap& mTl, const CUtensorMap& mY,
                        const float* ms32, const float* mt32, int64_t L, int64_t rows, int64_t dim, cudaStream_t st,
                        const int* run_flag) {
  auto kern = apply_ts_kernel<CG, BN>;
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    OTK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, AP_SMEM));
    attr_set[dev] = true;
  }
  const int64_t m_tiles = ceil_div(rows, AP_BM * CG), n_tiles = ceil_div(dim, BN);
  const int64_t total = L * m_tiles * n_tiles;
  if (total > INT32_MAX) return 0;
  const int64_t max_groups = sm_count() / CG;
  const unsigned groups = (unsigned)(total < max_groups ? total : max_groups);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(groups * CG);
  cfg.blockDim = dim3(AP_THREADS);
  cfg.dynamicSmemBytes = AP_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  OTK_CUDA(cudaLaunchKernelEx(&cfg, kern, mX, mTh, mTl, mY, ms32, mt32, (int)rows, (int)dim, (int)m_tiles, (int)n_tiles,
                              (int)total, g_apply_dbg, run_flag));
  OTK_LAUNCH_CHECK();
  return 1;
}

int g_apply_force_cg = 0;   // tuning aid: 1 / 2 force the single-CTA / CTA-pair TF32 instantiation (and skip the FP16 kernel)

bool apply_umma_eligible(const float* x, const float* y, int64_t L, int64_t rows, int64_t dim) {
  if (dim < 64 || dim % 4 != 0 || rows < 1 || L > 65535 || rows > INT32_MAX || dim > INT32_MAX) return false;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & 15)) return false;
  return tensormap_encoder() != nullptr;
}
bool apply_umma_pair(int64_t rows, int64_t dim) { return g_apply_force_cg ? g_apply_force_cg == 2 : (dim >= 256 && rows > 128); }
void apply_umma_split(const float* T32, int64_t n, float* Thi, float* Tlo, cudaStream_t st) {
  int64_t blocks = ceil_div(n, 256);
  if (blocks > (int64_t)sm_count() * 8) blocks = (int64_t)sm_count() * 8;
  split_matrix_kernel<<<(unsigned)blocks, 256, 0, st>>>(T32, n, Thi, Tlo);
  count_launch(1);
}
int apply_umma_run_planes(const float* x, int64_t L, int64_t rows, int64_t dim, const float* ms32, const float* mt32,
                          const float* Thi, const float* Tlo, float* y, bool pair, const int* run_flag, cudaStream_t st,
                          int64_t x_row_stride, int64_t x_batch_stride) {
  CUtensorMap mX, mTh, mTl, mY;
  const int64_t xrs = x_row_stride > 0 ? x_row_stride : dim, xbs = x_batch_stride > 0 ? x_batch_stride : rows * dim;
  if (!encode_map_f32_3d(&mX, x, dim, rows, L, xrs, xbs, 32, AP_BM)) return 0;
  if (!encode_map_f32_3d(&mTh, Thi, dim, dim, L, dim, dim * dim, 32, 128)) return 0;
  if (!encode_map_f32_3d(&mTl, Tlo, dim, dim, L, dim, dim * dim, 32, 128)) return 0;
  if (!encode_map_f32_3d(&mY, y, dim, rows, L, dim, rows * dim, 32, 32)) return 0;
  if (pair) return launch_apply<2, 256>(mX, mTh, mTl, mY, ms32, mt32, L, rows, dim, st, run_flag);
  return launch_apply<1, 128>(mX, mTh, mTl, mY, ms32, mt32, L, rows, dim, st, run_flag);
}

int apply_umma_try(const float* x, int64_t L, int64_t rows, int64_t dim, const float* ms32, const float* mt32,
                   const float* T32, float* Thi_scratch, float* Tlo_scratch, float* y, Arena& ar, cudaStream_t st) {
  if (!apply_umma_eligible(x, y, L, rows, dim)) return 0;
  apply_umma_split(T32, L * dim * dim, Thi_scratch, Tlo_scratch, st);
  OTK_CUDA(cudaGetLastError());
  const bool pair = apply_umma_pair(rows, dim);
  // FP16-split kernel first (twice the MMA rate); the TF32 kernel behind it recomputes Y only if a value left the FP16 range
  int* flag = nullptr;
  if (!g_apply_force_cg) {
    int used = apply_h_try(x, L, rows, dim, ms32, mt32, T32, y, ar, pair, st, &flag);
    if (used < 0) return used;
    if (!used) flag = nullptr;
  }
  return apply_umma_run_planes(x, L, rows, dim, ms32, mt32, Thi_scratch, Tlo_scratch, y, pair, flag, st);
}

}  // namespace otk
