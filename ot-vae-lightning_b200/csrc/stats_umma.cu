// K1 on tcgen05: streaming sum (x-c)(x-c)^T and sum (x-c) with fp32-accurate 3xTF32 arithmetic.
// Reference: GaussianModel._stats (gaussian_model.py:144-157: einsum SYRK + column sum) and fid.py:103-104.
//
// X^T X with row-major latents X [rows, dim]: the reduction index K is the latent row.
//   A operand = (X^T) block, 128 features per CTA: TMA brings the raw [32 rows x 128 features] tile into shared memory,
//               converter warps (thread <-> feature = TMEM lane) subtract the pivot, split into TF32 hi / lo and write
//               both planes into TENSOR MEMORY (tcgen05.st); A never returns to shared memory.
//   B operand = X block, MN-major in shared memory (32-byte-atom 128B swizzle, the only layout tcgen05 takes for
//               MN-major TF32): converted in place (hi) plus a second plane (lo).
//   MMA       = lo*hi + hi*lo + hi*hi, kind::tf32, A from TMEM, fp32 accumulator in TMEM.
// With CG == 2 a CTA pair (tcgen05 cta_group::2) computes a 256 x 128 block: each CTA converts its own 128 A-features
// and stages 64 of the 128 B-features, which halves the shared-memory traffic per flop.
//
// Pivot shift: the converters subtract a per-feature pivot c (mean of the head of the batch) so that the fp32
// accumulation works on centred data and the cancellation in cov = Sxx/n - mu mu^T (ot/matrix_utils.py:155-157) is not
// amplified; the raw sums the reference keeps are rebuilt exactly in fp64 by the merge kernel (stats.cu):
//     sum x = S' + n c ,   sum x x^T = P' + c S'^T + S' c^T + n c c^T .
//
// TMEM accumulation truncates (error grows with the number of accumulation steps), so every SU_SUB rows the accumulator
// is handed to the epilogue warps, which add it into fp32 REGISTER accumulators (round-to-nearest) while the next
// sub-chunk runs into the other TMEM buffer; one fp64 atomic flush per work segment.
//
// Work split: the (unit, row) space - unit = (l, block row I, block column j) of the upper block triangle - is cut
// into equal contiguous ranges, one per CTA (pair); a range that crosses a unit boundary is processed as two segments.
#include <cuda.h>

#include "otk_ptx.cuh"
#include "stats_umma.cuh"
#include "tensormap.cuh"

namespace otk {

constexpr int SU_T = 128, SU_BK = 32;
constexpr int SU_AS = 4, SU_ACC = 2;                 // A-in-TMEM ring depth; accumulators
// smem ring depths.  The B ring is the deep one: a B stage is only released by the MMA that consumed it, so its refill
// (HBM/L2 latency + conversion) must be hidden behind the other stages' MMAs; the raw A ring is released by the converters.
template <int CG> __host__ __device__ constexpr int su_xs() { return CG == 1 ? 2 : 3; }
template <int CG> __host__ __device__ constexpr int su_bs() { return CG == 1 ? 4 : 7; }
constexpr int SU_THREADS = 22 * 32;                  // TMA, MMA | 2x4 A-converter | 2x4 B-converter | 4 epilogue warps
constexpr int SU_ATILE = SU_T * SU_BK * 4;           // 16 KiB: four 32-feature slabs of [32 rows x 128 B]
constexpr int SU_SLAB = 32 * SU_BK * 4;              // 4 KiB
constexpr int SU_SUB = 1024;                         // rows accumulated in TMEM per sub-chunk
constexpr int SU_ACOL0 = 256;                        // TMEM columns [0,256): two accumulators, [256,512): A ring
template <int CG> constexpr int su_bplane() { return (4 / CG) * SU_SLAB; }
constexpr int SU_SACC = SU_T * SU_T * 4;             // 64 KiB fp32 second-stage accumulators [column][row]
template <int CG> constexpr int su_smem() { return su_xs<CG>() * SU_ATILE + su_bs<CG>() * 2 * su_bplane<CG>() + SU_SACC + 1024 + 512; }

struct SegIter {
  int64_t cur, end;
  int rows;
  __device__ bool next(int& unit, int& r0, int& r1) {
    if (cur >= end) return false;
    unit = (int)(cur / rows);
    r0 = (int)(cur % rows);
    const int64_t left = end - cur;
    r1 = (int)(left < (int64_t)(rows - r0) ? r0 + left : rows);
    cur += r1 - r0;
    return true;
  }
};

// unit -> (l, I, j): block row I of 128*CG features, block column j of 128 features, j >= I*CG
template <int CG>
__device__ __forceinline__ void decode_unit(int unit, int upl, int nJ, int& l, int& I, int& j) {
  l = unit / upl;
  int w = unit % upl;
  I = 0;
  while (w >= nJ - I * CG) { w -= nJ - I * CG; ++I; }
  j = I * CG + w;
}

template <int CG>
__global__ void __launch_bounds__(SU_THREADS, 1)
stats_ts_kernel(const __grid_constant__ CUtensorMap mapX, const float* __restrict__ pivot, int rows, int dim, int upl,
                int nJ, long long range_len, long long flat_total, int parts, double* __restrict__ ws_cov,
                double* __restrict__ ws_sum, int dbg, const int* __restrict__ run_flag) {
  using namespace ptx;
  // fallback launch behind the FP16-split kernel (stats_h.cu): nothing to do unless that kernel raised its overflow flag
  pdl_launch_dependents();
  pdl_wait();      // programmatic launch: the predecessor's flag / staging area / pivot are visible from here
  if (run_flag != nullptr && *run_flag == 0) return;
  constexpr int SU_XS = su_xs<CG>(), SU_BS = su_bs<CG>();
  constexpr int BSL = 4 / CG;                          // 32-feature slabs of the B block staged by this CTA
  constexpr int BPLANE = BSL * SU_SLAB, BSTAGE = 2 * BPLANE;
  constexpr int BROWS = SU_BK / CG;                    // rows of a slab one B-converter warp handles
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* xa = smem;
  uint8_t* xb = xa + SU_XS * SU_ATILE;
  uint8_t* sacc = xb + SU_BS * BSTAGE;
  uint64_t* full_a = reinterpret_cast<uint64_t*>(sacc + SU_SACC);        // raw A tile landed
  uint64_t* empty_ra = full_a + SU_XS;                                   // A converters have read it
  uint64_t* full_b = empty_ra + SU_XS;                                   // raw B tile landed
  uint64_t* ready_b = full_b + SU_BS;                                    // B planes converted (leader: both CTAs)
  uint64_t* empty_b = ready_b + SU_BS;                                   // MMAs reading them retired
  uint64_t* ready_a = empty_b + SU_BS;                                   // A planes written to TMEM (leader: both CTAs)
  uint64_t* empty_a = ready_a + SU_AS;                                   // MMAs reading them retired
  uint64_t* acc_full = empty_a + SU_AS;
  uint64_t* acc_empty = acc_full + SU_ACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + SU_ACC);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;
  const int group = CG == 2 ? blockIdx.x / 2 : blockIdx.x;
  // Every unit of this launch (units unit_off ... ) is cut into the same `parts` row ranges: CTAs working on the same rows
  // of different units run in lockstep, so X is fetched from HBM once and re-read from L2, and a CTA (pair) owns exactly
  // one segment.  (`flat_total` carries unit_off; a launch covers at most #SMs / CG units, the host loops over the rest.)
  const int64_t ubase = (int64_t)((int)flat_total + group / parts) * rows;
  int64_t flat0 = ubase + (int64_t)(group % parts) * range_len;
  const int64_t flat1 = flat0 + range_len < ubase + rows ? flat0 + range_len : ubase + rows;
  if (flat0 > flat1) flat0 = flat1;
  const int k_per_sub = SU_SUB / SU_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapX);
    for (int s = 0; s < SU_XS; ++s) { mbar_init(&full_a[s], 1); mbar_init(&empty_ra[s], 4); }
    for (int s = 0; s < SU_BS; ++s) { mbar_init(&full_b[s], 1); mbar_init(&ready_b[s], 4 * CG); mbar_init(&empty_b[s], 1); }
    for (int s = 0; s < SU_AS; ++s) { mbar_init(&ready_a[s], 4 * CG); mbar_init(&empty_a[s], 1); }
    for (int a = 0; a < SU_ACC; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 4 * CG); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc_cg<CG>(tmem_slot, 512); tmem_relinquish_cg<CG>(); }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  SegIter seg{flat0, flat1, rows};
  int unit, r0, r1, l, I, j;

  if (warp == 0) {
    // ===== TMA producer: raw tile of this CTA's 128 A-features and of its 128/CG B-features, 32 latent rows per step.
    // Warp-uniform loop, one elected lane issues (keeps the descriptors in uniform registers, see apply_umma.cu).
    {
      int it = 0;
      while (seg.next(unit, r0, r1)) {
        decode_unit<CG>(unit, upl, nJ, l, I, j);
        const int i0 = I * SU_T * CG + (int)rank * SU_T;
        const int j0 = j * SU_T + (int)rank * (SU_T / CG);
        const int num_k = (r1 - r0 + SU_BK - 1) / SU_BK;
        for (int kt = 0; kt < num_k; ++kt, ++it) {
          const int sx = it % SU_XS, sb = it % SU_BS;
          const int k0 = r0 + kt * SU_BK;
          mbar_wait(&empty_ra[sx], ((it / SU_XS) & 1) ^ 1);
          if (dbg & 1) {
            if (elect_one()) mbar_arrive(&full_a[sx]);
            __syncwarp();
            mbar_wait(&empty_b[sb], ((it / SU_BS) & 1) ^ 1);
            if (elect_one()) mbar_arrive(&full_b[sb]);
            __syncwarp();
            continue;
          }
          if (elect_one()) {
            mbar_arrive_expect_tx(&full_a[sx], SU_ATILE);
#pragma unroll
            for (int sl = 0; sl < 4; ++sl) tma_load_3d(xa + sx * SU_ATILE + sl * SU_SLAB, &mapX, i0 + 32 * sl, k0, l, &full_a[sx]);
          }
          __syncwarp();
          mbar_wait(&empty_b[sb], ((it / SU_BS) & 1) ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&full_b[sb], BPLANE);
#pragma unroll
            for (int sl = 0; sl < BSL; ++sl) tma_load_3d(xb + sb * BSTAGE + sl * SU_SLAB, &mapX, j0 + 32 * sl, k0, l, &full_b[sb]);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA only): warp-uniform loop, one elected lane issues =====
    if (rank == 0) {
      const uint32_t idesc = idesc_tf32(SU_T * CG, SU_T, 0, 1);   // A: TMEM (K-major), B: MN-major
      int it = 0, subc = 0;
      while (seg.next(unit, r0, r1)) {
        const int num_k = (r1 - r0 + SU_BK - 1) / SU_BK;
        const int num_sub = (num_k + k_per_sub - 1) / k_per_sub;
        for (int sub = 0; sub < num_sub; ++sub, ++subc) {
          const int a = subc % SU_ACC;
          mbar_wait(&acc_empty[a], ((subc / SU_ACC) & 1) ^ 1);
          tc_fence_after();
          const uint32_t acc = tmem_base + a * SU_T;
          const int kt_end = min(num_k, (sub + 1) * k_per_sub);
          for (int kt = sub * k_per_sub; kt < kt_end; ++kt, ++it) {
            const int sb = it % SU_BS, sa = it % SU_AS;
            mbar_wait(&ready_b[sb], (it / SU_BS) & 1);
            mbar_wait(&ready_a[sa], (it / SU_AS) & 1);
            tc_fence_after();
            const uint32_t bb = smem_u32(xb + sb * BSTAGE);
            const uint32_t ab = tmem_base + SU_ACOL0 + sa * 64;
            const bool first = (kt == sub * k_per_sub);
            if (elect_one()) {
#pragma unroll
              for (int kk = 0; kk < SU_BK / 8; ++kk) {
                if (dbg & 8) break;
                const uint64_t b_hi = smem_desc_mn_tf32(bb + kk * 1024, SU_SLAB);
                const uint64_t b_lo = smem_desc_mn_tf32(bb + BPLANE + kk * 1024, SU_SLAB);
                umma_tf32_ts<CG>(acc, ab + 32 + kk * 8, b_hi, idesc, !(first && kk == 0));   // lo * hi
                umma_tf32_ts<CG>(acc, ab + kk * 8, b_lo, idesc, 1);                           // hi * lo
                umma_tf32_ts<CG>(acc, ab + kk * 8, b_hi, idesc, 1);                           // hi * hi
              }
              umma_commit_cg<CG>(&empty_b[sb]);
              umma_commit_cg<CG>(&empty_a[sa]);
              if (kt == kt_end - 1) umma_commit_cg<CG>(&acc_full[a]);
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp < 10) {
    // ===== A converters: thread <-> feature (TMEM lane q*32 + lane) of slab q; also yields the column sums.
    // Two sets of four warps take alternate k-steps, so one step's latency chain (smem read -> split -> tcgen05.st ->
    // wait) may span two MMA step times.
    const int q = warp % 4, cset = (warp - 2) / 4;
    const uint32_t ready_addr = CG == 2 ? map_to_cta(smem_u32(&ready_a[0]), 0) : smem_u32(&ready_a[0]);
    const uint32_t cc = (uint32_t)lane / 8, within = (uint32_t)(lane % 8) * 4;
    const uint32_t xbase = smem_u32(xa) + (uint32_t)q * SU_SLAB;
    int it = 0;
    while (seg.next(unit, r0, r1)) {
      decode_unit<CG>(unit, upl, nJ, l, I, j);
      const int col = I * SU_T * CG + (int)rank * SU_T + q * 32 + lane;
      const bool in = col < dim;
      const float c = in ? pivot[(int64_t)l * dim + col] : 0.f;
      const bool do_sum = (j == I * CG);                     // one unit per block row contributes the column sums
      const int num_k = (r1 - r0 + SU_BK - 1) / SU_BK;
      double colsum = 0.0;
      for (int kt = 0; kt < num_k; ++kt, ++it) {
        const int sx = it % SU_XS, sa = it % SU_AS;
        if ((it & 1) != cset) continue;
        const int valid = in ? min(SU_BK, r1 - (r0 + kt * SU_BK)) : 0;   // rows past the segment end contribute nothing
        mbar_wait(&full_a[sx], (it / SU_XS) & 1);
        if (dbg & 2) {
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty_ra[sx]);
          mbar_wait(&empty_a[sa], ((it / SU_AS) & 1) ^ 1);
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(ready_addr + sa * 8);
          continue;
        }
        float part_sum = 0.f;
        const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + SU_ACOL0 + sa * 64;
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {                        // two halves of 16 rows keep the register footprint low
          float hi[16], lo[16];
#pragma unroll
          for (int rr = 0; rr < 16; ++rr) {
            const int r = h2 * 16 + rr;
            float x = lds32(xbase + sx * SU_ATILE + (uint32_t)r * 128 + ((cc ^ (uint32_t)(r & 3)) * 32) + within);
            x = r < valid ? x - c : 0.f;
            split_tf32_fast(x, hi[rr], lo[rr]);
            part_sum += x;
          }
          if (h2 == 0) { mbar_wait(&empty_a[sa], ((it / SU_AS) & 1) ^ 1); tc_fence_after(); }
          tmem_st16(ta + h2 * 16, hi);
          tmem_st16(ta + 32 + h2 * 16, lo);
        }
        colsum += (double)part_sum;
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_ra[sx]);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(ready_addr + sa * 8);
      }
      if (do_sum && in && num_k > 0) atomicAdd(&ws_sum[(int64_t)l * dim + col], colsum);
    }
  } else if (warp < 18) {
    // ===== B converters: warp -> (slab, row range); thread <-> feature; hi in place, lo into the second plane;
    // two sets of four warps on alternate k-steps =====
    const int cw = (warp - 10) % 4, cset = (warp - 10) / 4;
    const int sl = cw / CG, rh = cw % CG;                    // CG == 1: four slabs x 32 rows; CG == 2: two slabs x 2 x 16 rows
    const uint32_t ready_addr = CG == 2 ? map_to_cta(smem_u32(&ready_b[0]), 0) : smem_u32(&ready_b[0]);
    const uint32_t cc = (uint32_t)lane / 8, within = (uint32_t)(lane % 8) * 4;
    const uint32_t bbase = smem_u32(xb) + (uint32_t)sl * SU_SLAB;
    int it = 0;
    while (seg.next(unit, r0, r1)) {
      decode_unit<CG>(unit, upl, nJ, l, I, j);
      const int col = j * SU_T + (int)rank * (SU_T / CG) + sl * 32 + lane;
      const bool in = col < dim;
      const float c = in ? pivot[(int64_t)l * dim + col] : 0.f;
      const int num_k = (r1 - r0 + SU_BK - 1) / SU_BK;
      for (int kt = 0; kt < num_k; ++kt, ++it) {
        const int sb = it % SU_BS;
        if ((it & 1) != cset) continue;
        const int valid = in ? min(SU_BK, r1 - (r0 + kt * SU_BK)) : 0;
        mbar_wait(&full_b[sb], (it / SU_BS) & 1);
        const uint32_t hb = bbase + sb * BSTAGE;
#pragma unroll
        for (int rr = 0; rr < ((dbg & 4) ? 0 : BROWS); ++rr) {
          const int r = rh * BROWS + rr;
          const uint32_t off = (uint32_t)r * 128 + ((cc ^ (uint32_t)(r & 3)) * 32) + within;
          float x = lds32(hb + off);
          x = r < valid ? x - c : 0.f;
          float h, lo;
          split_tf32_fast(x, h, lo);
          sts32(hb + off, h);
          sts32(hb + BPLANE + off, lo);
        }
        fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(ready_addr + sb * 8);
      }
    }
  } else {
    // ===== epilogue: 4 warps (warp <-> TMEM lane quarter); second-stage fp32 accumulation in shared memory =====
    // sacc[column][row]: the 32 lanes of a warp touch 32 consecutive floats, so the read-modify-write is conflict-free
    const int q = warp % 4;
    const uint32_t acc_empty_addr = CG == 2 ? map_to_cta(smem_u32(&acc_empty[0]), 0) : smem_u32(&acc_empty[0]);
    const uint32_t srow = smem_u32(sacc) + (uint32_t)(q * 32 + lane) * 4;
    int subc = 0;
    while (seg.next(unit, r0, r1)) {
      decode_unit<CG>(unit, upl, nJ, l, I, j);
      const int num_k = (r1 - r0 + SU_BK - 1) / SU_BK;
      const int num_sub = (num_k + k_per_sub - 1) / k_per_sub;
      for (int sub = 0; sub < num_sub; ++sub, ++subc) {
        const int a = subc % SU_ACC;
        mbar_wait(&acc_full[a], (subc / SU_ACC) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + a * SU_T;
#pragma unroll 1
        for (int c0 = 0; c0 < SU_T; c0 += 32) {
          float v[32];
          tmem_ld32(taddr + c0, v);
          tmem_ld_wait();
          if (c0 + 32 == SU_T) {                               // accumulator fully read: hand it back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(acc_empty_addr + a * 8);
          }
          if (sub == 0) {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) sts32(srow + (uint32_t)(c0 + jj) * (SU_T * 4), v[jj]);
          } else {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) {
              const uint32_t ad = srow + (uint32_t)(c0 + jj) * (SU_T * 4);
              sts32(ad, lds32(ad) + v[jj]);
            }
          }
        }
      }
      // flush: P'[gi][gj] for gi <= gj, stored TRANSPOSED (ws[gj][gi]) so that the 32 lanes of an atomic instruction
      // hit 32 consecutive doubles; the merge kernel reads the transposed position
      const int gi = I * SU_T * CG + (int)rank * SU_T + q * 32 + lane;
      if (num_k > 0 && gi < dim) {
        double* cov = ws_cov + (int64_t)l * dim * dim;
#pragma unroll 4
        for (int jj = 0; jj < SU_T; ++jj) {
          const int gj = j * SU_T + jj;
          if (gj < dim && gi <= gj) atomicAdd(&cov[(int64_t)gj * dim + gi], (double)lds32(srow + (uint32_t)jj * (SU_T * 4)));
        }
      }
    }
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc_cg<CG>(tmem_base, 512); }
}

// pivot[l, :] = mean of the first min(rows, 64) latents of the batch.  Block = 32 features x 8 row groups.
__global__ void pivot_kernel(const float* __restrict__ x, int64_t rows, int64_t dim, int64_t row_stride, int64_t batch_stride,
                             float* __restrict__ pivot) {
  __shared__ float part[8][33];
  const int64_t l = blockIdx.y;
  const int64_t col = blockIdx.x * 32 + threadIdx.x;
  const int64_t n = rows < 64 ? rows : 64;
  float acc = 0.f;
  if (col < dim)
    for (int64_t r = threadIdx.y; r < n; r += 8) acc += x[l * batch_stride + r * row_stride + col];
  part[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && col < dim) {
#pragma unroll
    for (int g = 1; g < 8; ++g) acc += part[g][threadIdx.x];
    pivot[l * dim + col] = acc / (float)n;
  }
}

size_t stats_umma_extra_workspace(int64_t L, int64_t dim) {
  return align_up((size_t)L * dim * 4, 256) + 256 + stats_h2_extra_workspace(L, dim) + 1024;
}

int g_stats_dbg = 0;        // tuning aid: bit0 no TMA loads, bit1 no A conversion, bit2 no B conversion, bit3 no MMAs
template <int CG>
static int launch_stats(const CUtensorMap& mX, const float* pivot, int64_t L, int64_t rows, int64_t dim, double* ws_cov,
                        double* ws_sum, cudaStream_t st, const int* run_flag = nullptr) {
  auto kern = stats_ts_kernel<CG>;
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    OTK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, su_smem<CG>()));
    attr_set[dev] = true;
  }
  const int nJ = (int)ceil_div(dim, SU_T), nI = (int)ceil_div(dim, SU_T * CG);
  int upl = 0;
  for (int I = 0; I < nI; ++I) upl += nJ - I * CG;
  // Aligned cuts only: each launch takes at most #SMs / CG units and gives every unit the same `parts` row ranges
  // (multiples of 32 rows, at least 256 rows each).  Wide / batched problems with more units than CTA groups take several
  // launches.  (An earlier "flat" cut, where a CTA's range crossed unit boundaries, produced wrong sums for some of its
  // multi-segment schedules and was removed.)
  const int64_t max_groups = sm_count() / CG;
  const int64_t n_units = L * upl;
  for (int64_t unit_off = 0; unit_off < n_units; unit_off += max_groups) {
  const int64_t nu = n_units - unit_off < max_groups ? n_units - unit_off : max_groups;
  int64_t p = max_groups / nu;
  int64_t range_len = ceil_div(ceil_div(rows, p), SU_BK) * SU_BK;
  if (range_len < 256) range_len = 256;
  p = ceil_div(rows, range_len);
  const int parts = (int)p;
  const int64_t groups = nu * p;
  const long long flat_total = unit_off;
  OTK_CUDA(launch_pdl_site(2, kern, dim3((unsigned)(groups * CG)), dim3(SU_THREADS), su_smem<CG>(), st, CG, mX, pivot, (int)rows, (int)dim,
                      upl, nJ, (long long)range_len, (long long)flat_total, parts, ws_cov, ws_sum, g_stats_dbg, run_flag));
  OTK_LAUNCH_CHECK();
  }
  return 1;
}

int g_stats_force_cg = 0;   // tuning aid: 1 / 2 force the single-CTA / CTA-pair instantiation

int stats_umma_try(const float* x, int64_t L, int64_t rows, int64_t dim, int64_t row_stride, int64_t batch_stride,
                   double* ws_cov, double* ws_sum, Arena& ar, cudaStream_t st, int* tile, const float** pivot_out,
                   const StatsRunning& run) {
  if (dim < 64 || dim % 4 != 0 || row_stride % 4 != 0 || batch_stride % 4 != 0) return 0;
  if (rows < 1 || rows > INT32_MAX || dim > 16384 || L > 65535) return 0;
  if (reinterpret_cast<uintptr_t>(x) & 15) return 0;
  if (!tensormap_encoder()) return 0;
  float* pivot = ar.take<float>((size_t)L * dim);
  if (!ar.ok()) return OTK_ERR_WORKSPACE;
  const bool narrow = stats_h_eligible(L, rows, dim), wide = stats_h2_eligible(L, rows, dim);
  if (!g_stats_force_cg && (narrow || wide)) {
    // FP16-split kernels; the TF32 kernel follows as a device-gated fallback that only runs if a value left the FP16
    // range (it then recomputes into the re-zeroed staging area with the same pivot)
    StatsHPlan plan{};
    int used = narrow ? stats_h_launch(x, L, rows, dim, row_stride, batch_stride, pivot, ws_cov, ws_sum, ar, st, &plan)
                      : stats_h2_launch(x, L, rows, dim, row_stride, batch_stride, pivot, ws_cov, ws_sum, ar, st, &plan);
    if (used < 0) return used;
    if (used == 1) {
      CUtensorMap mF;
      if (!encode_map_f32_3d(&mF, x, dim, rows, L, row_stride, batch_stride, 32, SU_BK, /*atom32=*/true)) return OTK_ERR_CUDA;
      // flag up: clear the staging area (P' and S') and let the TF32 kernel recompute the call into it
      const int64_t staged = (reinterpret_cast<char*>(ws_sum) - reinterpret_cast<char*>(ws_cov)) / 8 + L * dim;
      OTK_TRY(stats_zero_if(ws_cov, staged, plan.flag, st));
      used = dim >= 512 ? launch_stats<2>(mF, pivot, L, rows, dim, ws_cov, ws_sum, st, plan.flag)
                        : launch_stats<1>(mF, pivot, L, rows, dim, ws_cov, ws_sum, st, plan.flag);
      if (used <= 0) return used < 0 ? used : OTK_ERR_CUDA;
      *pivot_out = pivot;
      *tile = SU_T;
      if (plan.mode == 0) return 1;        // several super-chunks: the staging area holds the result, the caller merges
      OTK_TRY(stats_h_merge(plan, pivot, ws_cov, ws_sum, L, rows, dim, run, st));
      return 2;
    }
  }
  OTK_CUDA(cudaMemsetAsync(ws_cov, 0, (reinterpret_cast<char*>(ws_sum) - reinterpret_cast<char*>(ws_cov)) + (size_t)L * dim * 8, st));
  pivot_kernel<<<dim3((unsigned)ceil_div(dim, 32), (unsigned)L), dim3(32, 8), 0, st>>>(x, rows, dim, row_stride, batch_stride, pivot);
  OTK_LAUNCH_CHECK();
  CUtensorMap mX;
  if (!encode_map_f32_3d(&mX, x, dim, rows, L, row_stride, batch_stride, 32, SU_BK, /*atom32=*/true)) return 0;
  const bool pair = g_stats_force_cg ? g_stats_force_cg == 2 : dim >= 512;
  int used = pair ? launch_stats<2>(mX, pivot, L, rows, dim, ws_cov, ws_sum, st)
                  : launch_stats<1>(mX, pivot, L, rows, dim, ws_cov, ws_sum, st);
  if (used <= 0) return used;
  *pivot_out = pivot;     // the merge kernel rebuilds the raw sums from (P' transposed, S', pivot)
  *tile = SU_T;
  return 1;
}

}  // namespace otk
