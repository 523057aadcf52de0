// K1 for narrow latents (dim <= 128): streaming sum (x-c)(x-c)^T and sum (x-c) with an fp32-accurate FP16 hi/lo split.
// Reference: GaussianModel._stats (gaussian_model.py:144-157: einsum SYRK + column sum) and fid.py:103-104.
//
// At dim <= 128 one read of X (4 d bytes per latent) takes less time than the 3 x 2 d^2 TF32 flops per latent of
// stats_umma.cu, i.e. the 3xTF32 kernel is tensor-bound below the HBM roofline.  FP16 carries the same 11-bit significand
// as TF32 at twice the MMA rate and half the operand bytes, so the same three-product scheme
//     x y ~= lo_x hi_y + hi_x lo_y + hi_x hi_y ,   hi = fp16(x'), lo = fp16(x' - hi),  x' = (x - c) s
// runs at the HBM bound instead - provided x' stays inside the FP16 range.  The per-feature pivot c (mean of the head of
// the batch) and power-of-two scale s (head deviation mapped to [64, 128)) leave a factor 512 of headroom over the
// largest deviation seen in the head; a converter that meets a larger value raises a device flag and the caller's
// stream then runs the TF32 kernel on the same staging area (the launch is a no-op while the flag is clear).
//
// There is a single output unit (the d x d block), so ONE raw tile feeds both operands and is converted once:
//   TMA      raw [64 rows x 128 features] fp32 tile (four 32-feature slabs) into a 3-deep ring;
//   convert  12 warps in three sets (thread <-> feature): subtract pivot, scale, split, pack pairs along the row index;
//            A planes -> tensor memory (lane = feature, 32-bit column = two consecutive rows),
//            B planes -> shared memory, K-major [feature][64 rows] with the 128-byte swizzle (8 x st.shared.v4 / plane);
//   MMA      kind::f16, A from tensor memory, 128 x 128 fp32 accumulator in tensor memory, 12 instructions per tile;
//   epilogue every 1024 rows the accumulator is added into fp32 second-stage accumulators in shared memory (tensor-memory
//            accumulation truncates), one fp64 atomic flush per CTA with the scales divided out.
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include "role_timing.cuh"

#include "otk_ptx.cuh"
#include "stats_umma.cuh"
#include "tensormap.cuh"

namespace otk {

constexpr int SH_T = 128, SH_BK = 64;
// Ring depths.  All three rings are as deep as there are converter sets, so a raw stage, an A slot and a B stage belong
// to ONE set and their barriers advance by exactly one phase per tile of that set (a parity wait cannot tell "two phases
// behind" from "done"); the sets are fully independent pipelines that meet only in the MMA issue order.
constexpr int SH_SETS = 3;
constexpr int SH_XS = SH_SETS, SH_BS = SH_SETS, SH_AS = SH_SETS, SH_ACC = 2;
constexpr int SH_THREADS = (2 + 4 * SH_SETS + 4) * 32;   // TMA, MMA | SH_SETS x 4 converter warps | 4 epilogue warps
constexpr int SH_SLAB = 32 * SH_BK * 4;              // 8 KiB: [64 rows x 32 features] fp32
constexpr int SH_RAW = 4 * SH_SLAB;                  // 32 KiB
constexpr int SH_BPLANE = SH_T * SH_BK * 2;          // 16 KiB: [128 features x 64 rows] fp16
constexpr int SH_BSTAGE = 2 * SH_BPLANE;             // hi + lo
constexpr int SH_SACC = SH_T * (SH_T + 1) / 2 * 4;   // 32.25 KiB second-stage accumulators, packed upper triangle: element
                                                     // (row gi <= column gj) at gj (gj + 1) / 2 + gi  (the third B stage
                                                     // lives in the half this packing frees)
constexpr int SH_TRI = SH_T * (SH_T + 1) / 2;        // floats per record (packed upper triangle of the block)
constexpr int SH_SUB = 1024;                         // rows accumulated in tensor memory per sub-chunk
constexpr int SH_ACOL0 = 256;                        // TMEM columns [0,256): two accumulators, [256,448): A ring (3 x 64)
constexpr int SH_SMEM = SH_XS * SH_RAW + SH_BS * SH_BSTAGE + SH_SACC + 1024 + 512;

__device__ __forceinline__ void sts128u(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tmem_st16u(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc], kind::f16 (UMMA_K = 16)
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// Converter core, thread <-> feature: 32 rows of the raw slab -> 16 packed hi words + 16 packed lo words (two consecutive
// rows per word, even row in the low half).  v = x s - c s is one FFMA (exact in the same sense as (x - c) s: s is a power
// of two); `vsum` accumulates sum v (the caller divides by s); `chk` stays 0 unless a value left the FP16 range or was not
// finite: then hi = inf, the residual v - hi is -inf / NaN and 0 * residual poisons chk - detection costs FMA-pipe slots
// only.  FULL tiles need no row masking (features past `dim` are zero-filled by TMA and have c = 0).
// (r2 late) The converters are issue-bound (~8 instructions per element), so the check rides on ONE packed HFMA2 per pair
// (0 * lo of both halves at once) instead of two FFMAs, and SUM = false drops the two column-sum FADDs where the sums are
// not needed (B converters, A converters of off-diagonal blocks).
__device__ __forceinline__ bool chk_clean(__half2 chk) { return __low2float(chk) == 0.f && __high2float(chk) == 0.f; }
template <bool FULL, bool SUM = true>
__device__ __forceinline__ void split_rows32(uint32_t slab, int row0, int valid, uint32_t cc, uint32_t within, float s,
                                             float ncs, float& vsum, __half2& chk, uint32_t* hw, uint32_t* lw) {
  using namespace ptx;
#pragma unroll
  for (int p = 0; p < 16; ++p) {
    const int ra = row0 + 2 * p, rb = ra + 1;
    const float x0 = lds32(slab + (uint32_t)ra * 128 + ((cc ^ (uint32_t)(ra & 3)) * 32) + within);
    const float x1 = lds32(slab + (uint32_t)rb * 128 + ((cc ^ (uint32_t)(rb & 3)) * 32) + within);
    float v0 = fmaf(x0, s, ncs), v1 = fmaf(x1, s, ncs);
    if (!FULL) { v0 = ra < valid ? v0 : 0.f; v1 = rb < valid ? v1 : 0.f; }
    if (SUM) vsum += v0 + v1;
    const __half2 h = __floats2half2_rn(v0, v1);                  // .x (low half) = the even row
    const float2 hf = __half22float2(h);
    const float l0 = v0 - hf.x, l1 = v1 - hf.y;
    const __half2 lo = __floats2half2_rn(l0, l1);
    chk = __hfma2(lo, __float2half2_rn(0.f), chk);                            // +-inf / NaN residual -> NaN
    hw[p] = *reinterpret_cast<const uint32_t*>(&h);
    lw[p] = *reinterpret_cast<const uint32_t*>(&lo);
  }
}

__global__ void __launch_bounds__(SH_THREADS, 1)
stats_h_kernel(const __grid_constant__ CUtensorMap mapX, const float* __restrict__ pivot, const float* __restrict__ scale,
               int rows, int dim, int parts, int range_len, float* __restrict__ rec_cov, double* __restrict__ ws_sum,
               int* __restrict__ overflow) {
  using namespace ptx;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* xa = smem;
  uint8_t* xb = xa + SH_XS * SH_RAW;
  uint8_t* sacc = xb + SH_BS * SH_BSTAGE;
  uint64_t* full_a = reinterpret_cast<uint64_t*>(sacc + SH_SACC);        // raw tile landed
  uint64_t* empty_ra = full_a + SH_XS;                                   // the converter set has read it
  uint64_t* ready_b = empty_ra + SH_XS;                                  // B planes written
  uint64_t* empty_b = ready_b + SH_BS;                                   // MMAs reading them retired
  uint64_t* ready_a = empty_b + SH_BS;                                   // A planes written to tensor memory
  uint64_t* empty_a = ready_a + SH_AS;                                   // MMAs reading them retired
  uint64_t* acc_full = empty_a + SH_AS;
  uint64_t* acc_empty = acc_full + SH_ACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + SH_ACC);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
#ifdef OTK_SH_TIMING
  const long long t_kernel0 = clock64();
  unsigned long long g_kernel0;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g_kernel0));
#endif
  const int l = blockIdx.x / parts, part = blockIdx.x % parts;
  const int r0 = part * range_len, r1 = min(rows, r0 + range_len);
  const int num_k = r1 > r0 ? (r1 - r0 + SH_BK - 1) / SH_BK : 0;
  constexpr int k_per_sub = SH_SUB / SH_BK;
  const int num_sub = (num_k + k_per_sub - 1) / k_per_sub;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapX);
    for (int s = 0; s < SH_XS; ++s) { mbar_init(&full_a[s], 1); mbar_init(&empty_ra[s], 4); }
    for (int s = 0; s < SH_BS; ++s) { mbar_init(&ready_b[s], 4); mbar_init(&empty_b[s], 1); }
    for (int s = 0; s < SH_AS; ++s) { mbar_init(&ready_a[s], 4); mbar_init(&empty_a[s], 1); }
    for (int a = 0; a < SH_ACC; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 4); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Launched programmatically behind pivot_scale_kernel (which releases its dependents at once): the prologue above and
  // the TMA producer - X was final before pivot_scale started - run beside it; everyone else (pivot / scale readers, the
  // overflow flag, the column-sum atomics it zeroes) waits for its completion here.
  pdl_launch_dependents();
  if (warp != 0) pdl_wait();

  if (warp == 0) {
    // ===== TMA producer (warp-uniform loop, one elected lane issues) =====
    for (int it = 0; it < num_k; ++it) {
      const int sx = it % SH_XS;
      mbar_wait(&empty_ra[sx], ((it / SH_XS) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&full_a[sx], SH_RAW);
#pragma unroll
        for (int sl = 0; sl < 4; ++sl) tma_load_3d(xa + sx * SH_RAW + sl * SH_SLAB, &mapX, 32 * sl, r0 + it * SH_BK, l, &full_a[sx]);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t idesc = idesc_f16(SH_T, SH_T);
    int it = 0;
    for (int sub = 0; sub < num_sub; ++sub) {
      const int a = sub % SH_ACC;
      mbar_wait(&acc_empty[a], ((sub / SH_ACC) & 1) ^ 1);
      tc_fence_after();
      const uint32_t acc = tmem_base + a * SH_T;
      const int kt_end = min(num_k, (sub + 1) * k_per_sub);
      for (int kt = sub * k_per_sub; kt < kt_end; ++kt, ++it) {
        const int sb = it % SH_BS, sa = it % SH_AS;
        mbar_wait(&ready_b[sb], (it / SH_BS) & 1);
        mbar_wait(&ready_a[sa], (it / SH_AS) & 1);
        tc_fence_after();
        const uint32_t bb = smem_u32(xb + sb * SH_BSTAGE);
        const uint32_t ab = tmem_base + SH_ACOL0 + sa * 64;
        const bool first = (kt == sub * k_per_sub);
        if (elect_one()) {
#ifndef OTK_SH_NOMMA
#pragma unroll
          for (int kk = 0; kk < SH_BK / 16; ++kk) {
            const uint64_t b_hi = smem_desc_sw128(bb + kk * 32, 16, 1024);
            const uint64_t b_lo = smem_desc_sw128(bb + SH_BPLANE + kk * 32, 16, 1024);
            umma_f16_ts(acc, ab + 32 + kk * 8, b_hi, idesc, !(first && kk == 0));   // lo * hi
#ifndef OTK_SH_MMA1
            umma_f16_ts(acc, ab + kk * 8, b_lo, idesc, 1);                           // hi * lo
            umma_f16_ts(acc, ab + kk * 8, b_hi, idesc, 1);                           // hi * hi
#endif
          }
#endif
          umma_commit(&empty_b[sb]);
          umma_commit(&empty_a[sa]);
          if (kt == kt_end - 1) umma_commit(&acc_full[a]);
        }
        __syncwarp();
      }
    }
  } else if (warp < 2 + 4 * SH_SETS) {
    // ===== converters: thread <-> feature (TMEM lane q*32 + lane).  SH_SETS sets of four warps, tile it -> set it % SH_SETS.
    const int q = warp % 4, cset = (warp - 2) / 4;
    const int col = q * 32 + lane;
    const bool in = col < dim;
    const float c = in ? pivot[(int64_t)l * dim + col] : 0.f;
    const float s = in ? scale[(int64_t)l * dim + col] : 1.f;
    const uint32_t cc = (uint32_t)lane / 8, within = (uint32_t)(lane % 8) * 4;
    const uint32_t xbase = smem_u32(xa) + (uint32_t)q * SH_SLAB;
    const uint32_t brow = smem_u32(xb) + (uint32_t)col * 128;            // this feature's 128-byte row of a B plane
    const uint32_t sw = (uint32_t)(col & 7);
    const float ncs = -c * s;
    double colsum = 0.0;
    __half2 chk = __float2half2_rn(0.f);
#ifdef OTK_SH_TIMING
    long long tw_a = 0, tw_x = 0, tw_b = 0, t_cv = 0, t_st = 0, t_all = clock64();
#define SH_TICK(acc) { const long long t_now = clock64(); acc += t_now - t_prev; t_prev = t_now; }
#else
#define SH_TICK(acc)
#endif
    for (int it = cset; it < num_k; it += SH_SETS) {
      const int sx = it % SH_XS, sa = it % SH_AS, sb = it % SH_BS;
      const int valid = min(SH_BK, r1 - (r0 + it * SH_BK));              // rows past the range end contribute nothing
#ifdef OTK_SH_TIMING
      long long t_prev = clock64();
#endif
      // the A slot, the raw stage and the B stage of tile `it` were last used by this set's previous tile (it - SH_SETS)
      mbar_wait(&empty_a[sa], ((it / SH_AS) & 1) ^ 1);
      tc_fence_after();
      SH_TICK(tw_a)
      mbar_wait(&full_a[sx], (it / SH_XS) & 1);
      SH_TICK(tw_x)
      mbar_wait(&empty_b[sb], ((it / SH_BS) & 1) ^ 1);
      SH_TICK(tw_b)
      float vsum = 0.f;
      const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + SH_ACOL0 + sa * 64;
      const uint32_t hb = brow + sb * SH_BSTAGE;
      const uint32_t slab = xbase + sx * SH_RAW;
#ifndef OTK_SH_NOCONV
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {                                   // two halves of 32 rows
        uint32_t hw[16], lw[16];
        if (valid == SH_BK) split_rows32<true>(slab, h2 * 32, valid, cc, within, s, ncs, vsum, chk, hw, lw);
        else split_rows32<false>(slab, h2 * 32, valid, cc, within, s, ncs, vsum, chk, hw, lw);
        tmem_st16u(ta + h2 * 16, hw);
        tmem_st16u(ta + 32 + h2 * 16, lw);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {                                 // 16-byte chunks (8 rows each) of this half
          const uint32_t off = (((uint32_t)(h2 * 4 + ch)) ^ sw) * 16;
          sts128u(hb + off, hw[4 * ch], hw[4 * ch + 1], hw[4 * ch + 2], hw[4 * ch + 3]);
          sts128u(hb + SH_BPLANE + off, lw[4 * ch], lw[4 * ch + 1], lw[4 * ch + 2], lw[4 * ch + 3]);
        }
      }
#endif
      colsum += (double)vsum;
      fence_proxy_async_smem();   // generic-proxy writes of the B planes -> visible to the tensor core
      __syncwarp();
      SH_TICK(t_cv)
      if (lane == 0) { mbar_arrive(&empty_ra[sx]); mbar_arrive(&ready_b[sb]); }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ready_a[sa]);
      SH_TICK(t_st)
    }
#ifdef OTK_SH_TIMING
    if (lane == 0 && q == 0 && (blockIdx.x == 0 || blockIdx.x == 77) && num_k > 8)
      printf("cta %d set %d tiles %d: total %lld | wait empty_a %lld full_a %lld empty_b %lld | convert %lld | st-wait %lld (cycles per tile of this set)\n",
             blockIdx.x, cset, (num_k - cset + SH_SETS - 1) / SH_SETS, (clock64() - t_all) / ((num_k - cset + SH_SETS - 1) / SH_SETS),
             tw_a / ((num_k - cset + SH_SETS - 1) / SH_SETS), tw_x / ((num_k - cset + SH_SETS - 1) / SH_SETS),
             tw_b / ((num_k - cset + SH_SETS - 1) / SH_SETS), t_cv / ((num_k - cset + SH_SETS - 1) / SH_SETS),
             t_st / ((num_k - cset + SH_SETS - 1) / SH_SETS));
#endif
    // the set's raw stage is idle from here on (its last tile has been read): park the column sums there, the CTA adds
    // the three sets up after the final barrier - one atomic per feature and CTA
    *reinterpret_cast<double*>(xa + cset * SH_RAW + col * 8) = colsum / (double)s;
    if (!chk_clean(chk)) atomicOr(overflow, 1);                           // a value left the FP16 window, or NaN / inf input
  } else {
    // ===== epilogue: 4 warps (warp <-> TMEM lane quarter); second-stage fp32 accumulation in shared memory =====
    const int q = warp % 4;
    const int gi = q * 32 + lane;                                        // TMEM lane = row of the block
    const uint32_t srow = smem_u32(sacc) + (uint32_t)gi * 4;
    for (int sub = 0; sub < num_sub; ++sub) {
      const int a = sub % SH_ACC;
      mbar_wait(&acc_full[a], (sub / SH_ACC) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + a * SH_T;
#pragma unroll 1
      for (int c0 = q * 32; c0 < SH_T; c0 += 32) {                       // columns left of the warp's rows are never kept
        float v[32];
        tmem_ld32(taddr + c0, v);
        tmem_ld_wait();
        if (c0 + 32 == SH_T) {                                           // accumulator fully read: hand it back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[a]);
        }
        const uint32_t tri0 = (uint32_t)(c0 * (c0 + 1) / 2);
        if (sub == 0) {
#pragma unroll
          for (int jj = 0; jj < 32; ++jj)
            if (gi <= c0 + jj) sts32(srow + (tri0 + (uint32_t)(c0 * jj + jj * (jj + 1) / 2)) * 4, v[jj]);
        } else {
#pragma unroll
          for (int jj = 0; jj < 32; ++jj)
            if (gi <= c0 + jj) {
              const uint32_t ad = srow + (tri0 + (uint32_t)(c0 * jj + jj * (jj + 1) / 2)) * 4;
              sts32(ad, lds32(ad) + v[jj]);
            }
        }
      }
    }
#ifdef OTK_SH_TIMING
    const long long t_flush0 = clock64();
#endif
    // flush: the CTA's packed triangle goes to its own record with plain stores (a warp stores the 32 consecutive floats
    // it owns of a column); stats_h_merge_kernel sums the records in fp64 and divides the power-of-two scales out.
    // (148 CTAs x 8256 fp64 atomics on the same addresses cost 26 us here - a quarter of the kernel.)
    {
      float* rec = rec_cov + (int64_t)blockIdx.x * SH_TRI;
      for (int gj = q * 32; gj < SH_T; ++gj)
        if (gi <= gj) {
          const uint32_t idx = (uint32_t)(gj * (gj + 1) / 2);
          rec[idx + gi] = num_k > 0 ? lds32(srow + idx * 4) : 0.f;
        }
    }
#ifdef OTK_SH_TIMING
    if (lane == 0 && (blockIdx.x == 0 || blockIdx.x == 77 || blockIdx.x == 140)) {
      unsigned long long g1;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g1));
      printf("cta %d epi warp %d: loop end at %lld cycles after kernel start, flush %lld cycles; kernel start->flush end %llu ns (start stamp %llu)\n",
             blockIdx.x, q, t_flush0 - t_kernel0, clock64() - t_flush0, g1 - g_kernel0, g_kernel0 % 10000000ull);
    }
#endif
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
  if ((int)threadIdx.x < dim && num_k > 0) {
    double sv = 0.0;
#pragma unroll
    for (int cs = 0; cs < SH_SETS; ++cs) sv += *reinterpret_cast<const double*>(xa + cs * SH_RAW + threadIdx.x * 8);
    atomicAdd(&ws_sum[(int64_t)l * dim + threadIdx.x], sv);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Wide latents (dim > 128): the same FP16 hi/lo split on the 128 x 128 blocks (bi <= bj) of the upper block triangle.
// Work item = (row segment, block); items are dealt round-robin to a persistent grid, segment-major, so the CTAs that run
// concurrently read the same rows and X comes from HBM once (re-reads are L2 hits).  The A block (-> tensor memory) and the
// B block (-> K-major swizzled planes in shared memory) have their own raw rings and eight converter warps each (TMEM lane
// quarter x 32-row half of the tile).  Every converter warp visits every tile in order, so no barrier can run two phases
// ahead of a waiter and the output rings may be deep (4 A slots in tensor memory, 3 B stages): the converters run up to
// three tiles ahead of the MMAs instead of ping-ponging with them.
// A segment is short enough (<= 2048 rows) for the truncating tensor-memory accumulation; its 128 x 128 fp32 tile is
// written with plain stores to a per-item slot and the partial tiles are summed in fp64 by stats_h2_reduce_kernel (no
// atomics on the covariance, no second-stage accumulators in shared memory).
constexpr int S2_XS = 3, S2_PB = 4, S2_AS = 4;                        // raw A ring, B ring (raw tile -> planes, in place), A slots
constexpr int S2_THREADS = (2 + 8 + 8 + 4 + 1) * 32;                   // TMA (A), MMA | 8 A conv. | 8 B conv. | 4 epilogue | TMA (B)
constexpr int S2_SMEM = S2_XS * SH_RAW + S2_PB * SH_BSTAGE + 1024 + 512;
static_assert(SH_RAW == SH_BSTAGE, "the B planes overwrite the raw tile they were converted from");
constexpr int S2_TILE = SH_T * SH_T;                                    // floats per 128 x 128 partial tile
constexpr int S2_MAX_ITEMS = 2048;                                      // 128 x 128 slots per launch (128 MiB); a pair item takes 4

template <int CG>
__device__ __forceinline__ void h2_umma(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (CG == 1)
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

struct S2Item { int l, bi, bj, r0, r1; };
__device__ __forceinline__ S2Item s2_decode(int item, int n_units, int upl, int nB, int seg_len, int row_lo, int row_hi) {
  S2Item t;
  const int seg = item / n_units, u = item % n_units;
  t.l = u / upl;
  int w = u % upl;
  t.bi = 0;
  while (w >= nB - t.bi) { w -= nB - t.bi; ++t.bi; }
  t.bj = t.bi + w;
  t.r0 = row_lo + seg * seg_len;
  t.r1 = min(row_hi, t.r0 + seg_len);
  return t;
}

// CG == 2: a CTA pair (tcgen05 cta_group::2) owns a 256 x 256 block.  Each CTA still loads and converts ONE raw A block
// (its 128 of the 256 A features -> its own tensor memory) and ONE raw B block (its 128 of the 256 B features -> its own
// shared memory) per 64-row tile, exactly the per-CTA work of CG == 1, but the pair's MMAs (M = N = 256, issued by the
// leader) do twice the flops per CTA with it: L2 -> SM bytes, shared-memory traffic and converter instructions per flop
// all halve - the three things the single-CTA kernel is bound by (53 us of its 87 us per 65536 x 512 chunk is raw-tile
// delivery alone).  Converters of both CTAs arrive on the leader's ready barriers, commits are multicast to both CTAs, each
// CTA drains its own 128 x 256 accumulator (one accumulator: 256 of the 512 tensor-memory columns, the A ring has the rest).
template <int CG>
__global__ void __launch_bounds__(S2_THREADS, 1)
stats_h2_kernel(const __grid_constant__ CUtensorMap mapX, const float* __restrict__ pivot, const float* __restrict__ scale,
                int dim, int nB, int upl, int n_units, int n_items, int seg_len, int row_lo, int row_hi,
                float* __restrict__ parts, double* __restrict__ ws_sum, int* __restrict__ overflow) {
  using namespace ptx;
  constexpr int S2_ACC = CG == 2 ? 1 : 2;                 // accumulators of 128 * CG columns in tensor-memory columns [0, 256)
  constexpr int BW = SH_T * CG;                           // block width = accumulator columns = row length of a partial tile
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* xa = smem;                                     // raw A ring
  uint8_t* xb = xa + S2_XS * SH_RAW;                      // B ring: a stage is a raw tile, then (in place) its hi / lo planes
  uint64_t* full_a = reinterpret_cast<uint64_t*>(xb + S2_PB * SH_BSTAGE);
  uint64_t* empty_ra = full_a + S2_XS;
  uint64_t* full_b = empty_ra + S2_XS;
  uint64_t* ready_a = full_b + S2_PB;
  uint64_t* empty_a = ready_a + S2_AS;
  uint64_t* ready_b = empty_a + S2_AS;
  uint64_t* empty_b = ready_b + S2_PB;
  uint64_t* acc_full = empty_b + S2_PB;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;
  const int group = CG == 2 ? blockIdx.x / 2 : blockIdx.x;           // item-processing unit (CTA or CTA pair)
  const int n_groups = CG == 2 ? gridDim.x / 2 : gridDim.x;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapX);
    for (int s = 0; s < S2_XS; ++s) { mbar_init(&full_a[s], 1); mbar_init(&empty_ra[s], 16); }   // 8 A + 8 B converter warps
    for (int s = 0; s < S2_PB; ++s) mbar_init(&full_b[s], 1);
    for (int s = 0; s < S2_AS; ++s) { mbar_init(&ready_a[s], 8 * CG); mbar_init(&empty_a[s], 1); }
    for (int s = 0; s < S2_PB; ++s) { mbar_init(&ready_b[s], 8 * CG); mbar_init(&empty_b[s], 1); }
    for (int a = 0; a < S2_ACC; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 4 * CG); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc_cg<CG>(tmem_slot, 512); tmem_relinquish_cg<CG>(); }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  if (warp != 0 && warp != 22) pdl_wait();      // as in stats_h_kernel: only the two TMA producer warps run ahead of pivot_scale

  if (warp == 0 || warp == 22) {
    // ===== TMA producers: warp 0 streams the A blocks, warp 22 the B blocks (independent rings: a slow side must not
    // hold back the other side's loads) =====
    const bool side_b = warp == 22;
    // (a B stage is free again when the MMAs that read its planes have retired: empty_b, armed by tcgen05.commit)
    uint8_t* ring = side_b ? xb : xa;
    uint64_t* full = side_b ? full_b : full_a;
    uint64_t* empty = side_b ? empty_b : empty_ra;
    const int depth = side_b ? S2_PB : S2_XS;
    int it = 0;
    S2_T0
    for (int item = group; item < n_items; item += n_groups) {
      const S2Item t = s2_decode(item, n_units, upl, nB, seg_len, row_lo, row_hi);
      const int num_k = (t.r1 - t.r0 + SH_BK - 1) / SH_BK;
      const int f0 = (side_b ? t.bj : t.bi) * BW + (int)rank * SH_T;
      // Diagonal block (bi == bj): the B block IS the A block.  It is loaded once, into the A ring, and the B converters
      // read it there; the B stage still cycles (its planes are written as usual), so its full barrier gets a plain
      // arrival.  At d = 512 that is 4 instead of 6 raw blocks per row tile: the kernel is bound by raw-tile delivery
      // (ablation builds: loads alone 44 us of the 71 us per 65536-row chunk).
      const bool skip_load = side_b && t.bi == t.bj;
      for (int kt = 0; kt < num_k; ++kt, ++it) {
        const int sx = it % depth;
        S2_TICK(s2_b)
        mbar_wait(&empty[sx], ((it / depth) & 1) ^ 1);
        S2_TICK(s2_a)
        S2_COUNT
        if (elect_one()) {
          if (skip_load) {
            mbar_arrive(&full[sx]);
          } else {
            mbar_arrive_expect_tx(&full[sx], SH_RAW);
#pragma unroll
            for (int sl = 0; sl < 4; ++sl)
              tma_load_3d(ring + sx * SH_RAW + sl * SH_SLAB, &mapX, f0 + 32 * sl, t.r0 + kt * SH_BK, t.l, &full[sx]);
          }
        }
        __syncwarp();
      }
    }
    S2_REPORT(side_b ? "tma B" : "tma A", "wait empty", "issue", "-", "-", "-")
  } else if (warp == 1) {
    // ===== MMA issuer (the leader CTA of a pair issues for both) =====
    const uint32_t idesc = idesc_f16(BW, BW);
    int it = 0, n = 0;
    S2_T0
    if (rank == 0)
    for (int item = group; item < n_items; item += n_groups, ++n) {
      const S2Item t = s2_decode(item, n_units, upl, nB, seg_len, row_lo, row_hi);
      const int num_k = (t.r1 - t.r0 + SH_BK - 1) / SH_BK;
      const int a = n % S2_ACC;
      S2_TICK(s2_d)
      mbar_wait(&acc_empty[a], ((n / S2_ACC) & 1) ^ 1);
      S2_TICK(s2_a)
      tc_fence_after();
      const uint32_t acc = tmem_base + a * BW;
      for (int kt = 0; kt < num_k; ++kt, ++it) {
        const int sp = it % S2_PB, sa = it % S2_AS;
        S2_TICK(s2_d)
        mbar_wait(&ready_b[sp], (it / S2_PB) & 1);
        S2_TICK(s2_b)
        mbar_wait(&ready_a[sa], (it / S2_AS) & 1);
        S2_TICK(s2_c)
        S2_COUNT
        tc_fence_after();
        const uint32_t bb = smem_u32(xb + sp * SH_BSTAGE);
        const uint32_t ab = tmem_base + SH_ACOL0 + sa * 64;
        if (elect_one()) {
#ifndef OTK_SH_NOMMA
#pragma unroll
          for (int kk = 0; kk < SH_BK / 16; ++kk) {
            const uint64_t b_hi = smem_desc_sw128(bb + kk * 32, 16, 1024);
            const uint64_t b_lo = smem_desc_sw128(bb + SH_BPLANE + kk * 32, 16, 1024);
            h2_umma<CG>(acc, ab + 32 + kk * 8, b_hi, idesc, !(kt == 0 && kk == 0));   // lo * hi
            h2_umma<CG>(acc, ab + kk * 8, b_lo, idesc, 1);                             // hi * lo
            h2_umma<CG>(acc, ab + kk * 8, b_hi, idesc, 1);                             // hi * hi
          }
#endif
          umma_commit_cg<CG>(&empty_b[sp]);
          umma_commit_cg<CG>(&empty_a[sa]);
          if (kt == num_k - 1) umma_commit_cg<CG>(&acc_full[a]);
        }
        __syncwarp();
      }
    }
    S2_REPORT("mma", "wait acc_empty", "wait ready_b", "wait ready_a", "issue", "-")
  } else if (warp < 10) {
    // ===== A converters: thread <-> feature of block bi (TMEM lane q*32 + lane), warp <-> (lane quarter, 32-row half) =====
    const int q = warp % 4, h2 = (warp - 2) / 4;
    const uint32_t cc = (uint32_t)lane / 8, within = (uint32_t)(lane % 8) * 4;
    const uint32_t slab0 = smem_u32(xa) + (uint32_t)q * SH_SLAB;
    const uint32_t ta0 = tmem_base + ((uint32_t)(q * 32) << 16) + SH_ACOL0 + h2 * 16;
    const uint32_t ready_a_addr = CG == 2 ? map_to_cta(smem_u32(&ready_a[0]), 0) : smem_u32(&ready_a[0]);
    __half2 chk = __float2half2_rn(0.f);
    int it = 0;
    S2_T0
    for (int item = group; item < n_items; item += n_groups) {
      const S2Item t = s2_decode(item, n_units, upl, nB, seg_len, row_lo, row_hi);
      const int num_k = (t.r1 - t.r0 + SH_BK - 1) / SH_BK;
      const int col = t.bi * BW + (int)rank * SH_T + q * 32 + lane;
      const bool in = col < dim;
      const float c = in ? pivot[(int64_t)t.l * dim + col] : 0.f;
      const float s = in ? scale[(int64_t)t.l * dim + col] : 1.f;
      const float ncs = -c * s;
      const bool diag = t.bi == t.bj;                     // the column sums are taken once per feature: on the diagonal blocks
      double colsum = 0.0;
      float vsum = 0.f;                                   // fp32 over at most four tiles (128 values), then fp64
      for (int kt = 0; kt < num_k; ++kt, ++it) {
        const int sx = it % S2_XS, sa = it % S2_AS;
        const int valid = min(SH_BK, t.r1 - (t.r0 + kt * SH_BK));
        S2_TICK(s2_e)
        mbar_wait(&empty_a[sa], ((it / S2_AS) & 1) ^ 1);
        S2_TICK(s2_a)
        tc_fence_after();
        mbar_wait(&full_a[sx], (it / S2_XS) & 1);
        S2_TICK(s2_b)
        S2_COUNT
#ifndef OTK_SH_NOCONV
        uint32_t hw[16], lw[16];
        if (valid == SH_BK) {
          if (diag) split_rows32<true, true>(slab0 + sx * SH_RAW, h2 * 32, valid, cc, within, s, ncs, vsum, chk, hw, lw);
          else split_rows32<true, false>(slab0 + sx * SH_RAW, h2 * 32, valid, cc, within, s, ncs, vsum, chk, hw, lw);
        } else {
          split_rows32<false, true>(slab0 + sx * SH_RAW, h2 * 32, valid, cc, within, s, ncs, vsum, chk, hw, lw);
        }
        tmem_st16u(ta0 + sa * 64, hw);
        tmem_st16u(ta0 + sa * 64 + 32, lw);
#endif
        if (diag && ((kt & 3) == 3 || kt == num_k - 1)) { colsum += (double)vsum; vsum = 0.f; }   // DADD: 6 % of the kernel's stall samples when unconditional
        __syncwarp();
        S2_TICK(s2_c)
        if (lane == 0) mbar_arrive(&empty_ra[sx]);
        tmem_st_wait();
        S2_TICK(s2_d)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(ready_a_addr + sa * 8);
      }
      if (diag && in && num_k > 0) atomicAdd(&ws_sum[(int64_t)t.l * dim + col], colsum / (double)s);   // each feature once
    }
    if (!chk_clean(chk)) atomicOr(overflow, 1);
    if (q == 0 && h2 == 0) { S2_REPORT("conv A", "wait empty_a", "wait full_a", "convert", "tmem st-wait", "arrive+loop") }
  } else if (warp < 18) {
    // ===== B converters: thread <-> feature of block bj (row of the K-major planes), warp <-> (quarter, 32-row half) =====
    const int q = warp % 4, h2 = (warp - 10) / 4;
    const int nloc = q * 32 + lane;
    const uint32_t cc = (uint32_t)lane / 8, within = (uint32_t)(lane % 8) * 4;
    const uint32_t slab0 = smem_u32(xb) + (uint32_t)q * SH_SLAB;
    const uint32_t slab0_a = smem_u32(xa) + (uint32_t)q * SH_SLAB;      // diagonal blocks: the raw tile sits in the A ring
    const uint32_t hb0 = smem_u32(xb) + (uint32_t)nloc * 128;
    const uint32_t sw = (uint32_t)(nloc & 7);
    const uint32_t ready_b_addr = CG == 2 ? map_to_cta(smem_u32(&ready_b[0]), 0) : smem_u32(&ready_b[0]);
    __half2 chk = __float2half2_rn(0.f);
    float vsum = 0.f;                                     // unused (SUM = false)
    int it = 0;
    S2_T0
    for (int item = group; item < n_items; item += n_groups) {
      const S2Item t = s2_decode(item, n_units, upl, nB, seg_len, row_lo, row_hi);
      const int num_k = (t.r1 - t.r0 + SH_BK - 1) / SH_BK;
      const int col = t.bj * BW + (int)rank * SH_T + nloc;
      const bool in = col < dim;
      const float c = in ? pivot[(int64_t)t.l * dim + col] : 0.f;
      const float s = in ? scale[(int64_t)t.l * dim + col] : 1.f;
      const float ncs = -c * s;
      const bool diag = t.bi == t.bj;
      for (int kt = 0; kt < num_k; ++kt, ++it) {
        const int sp = it % S2_PB, sx = it % S2_XS;
        const int valid = min(SH_BK, t.r1 - (t.r0 + kt * SH_BK));
        S2_TICK(s2_e)
        mbar_wait(&full_b[sp], (it / S2_PB) & 1);
        S2_TICK(s2_a)
        S2_COUNT
        // The A ring's stage of this tile is released by the A AND the B converters (its barrier counts all 16 warps, so
        // every tile needs both arrivals): wait until the tile has landed there - that also pins the barrier's phase.
        mbar_wait(&full_a[sx], (it / S2_XS) & 1);
        S2_TICK(s2_b)
        const uint32_t src = diag ? slab0_a + sx * SH_RAW : slab0 + sp * SH_RAW;
#ifndef OTK_SH_NOCONV
        uint32_t hw[16], lw[16];
        if (valid == SH_BK) split_rows32<true, false>(src, h2 * 32, valid, cc, within, s, ncs, vsum, chk, hw, lw);
        else split_rows32<false, false>(src, h2 * 32, valid, cc, within, s, ncs, vsum, chk, hw, lw);
#endif
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_ra[sx]);                 // values are in registers
        S2_TICK(s2_c)
#ifndef OTK_SH_NOCONV
        // the planes overwrite the raw tile: every B converter must have read its share first
        asm volatile("bar.sync 1, 256;" ::: "memory");
        S2_TICK(s2_d)
        const uint32_t hb = hb0 + sp * SH_BSTAGE;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const uint32_t off = (((uint32_t)(h2 * 4 + ch)) ^ sw) * 16;
          sts128u(hb + off, hw[4 * ch], hw[4 * ch + 1], hw[4 * ch + 2], hw[4 * ch + 3]);
          sts128u(hb + SH_BPLANE + off, lw[4 * ch], lw[4 * ch + 1], lw[4 * ch + 2], lw[4 * ch + 3]);
        }
#endif
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(ready_b_addr + sp * 8);
      }
    }
    if (!chk_clean(chk)) atomicOr(overflow, 1);
    if (q == 0 && h2 == 0) { S2_REPORT("conv B", "wait full_b", "wait full_a", "convert", "bar.sync", "sts+arrive") }
  } else if (warp < 22) {
    // ===== epilogue: TMEM accumulator of an item -> its partial-tile slot (thread <-> row of the block) =====
    const int q = warp % 4;
    const uint32_t acc_empty_addr = CG == 2 ? map_to_cta(smem_u32(&acc_empty[0]), 0) : smem_u32(&acc_empty[0]);
    int n = 0;
    for (int item = group; item < n_items; item += n_groups, ++n) {
      const int a = n % S2_ACC;
      mbar_wait(&acc_full[a], (n / S2_ACC) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + a * BW;
      float4* out = reinterpret_cast<float4*>(parts + (int64_t)item * (BW * BW) + ((int)rank * SH_T + q * 32 + lane) * BW);
#pragma unroll 1
      for (int c0 = 0; c0 < BW; c0 += 32) {
        float v[32];
        tmem_ld32(taddr + c0, v);
        tmem_ld_wait();
        if (c0 + 32 == BW) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(acc_empty_addr + a * 8);
        }
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) out[c0 / 4 + j4] = make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
      }
    }
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc_cg<CG>(tmem_base, 512); }
}

// ws_cov[l][gj][gi] (the transposed position the merge kernel reads, gi <= gj) += sum over the segments of the partial
// tiles, in fp64, with the power-of-two scales divided out.  Once the overflow flag is up the staging area is cleared
// instead (covariance and column sums): the TF32 kernel launched behind recomputes the whole call into it.
__global__ void stats_h2_reduce_kernel(const float* __restrict__ parts, const float* __restrict__ scale, int64_t L, int dim,
                                       int nB, int upl, int n_units, int n_seg, double* __restrict__ ws_cov,
                                       double* __restrict__ ws_sum, const int* __restrict__ overflow) {
  const int64_t total = L * (int64_t)dim * dim;
  if (*overflow) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
      ws_cov[e] = 0.0;
      if (e < L * dim) ws_sum[e] = 0.0;
    }
    return;
  }
  // one thread per 4 consecutive columns gj of a row gi (dim % 4 == 0; a group never straddles a 128-column block);
  // groups entirely below the diagonal are skipped; the segment loop keeps four independent 16-byte loads in flight
  const int dq = dim / 4;
  const int64_t groups = L * (int64_t)dim * dq;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < groups; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t l = e / ((int64_t)dim * dq);
    const int r = (int)(e % ((int64_t)dim * dq)), gi = r / dq, gj = (r % dq) * 4;
    if (gi > gj + 3) continue;
    const int bi = gi / SH_T, bj = gj / SH_T;
    const int u = (int)l * upl + bi * nB - bi * (bi - 1) / 2 + (bj - bi);
    const float4* p = reinterpret_cast<const float4*>(parts + (int64_t)u * S2_TILE + (gi % SH_T) * SH_T + (gj % SH_T));
    const int64_t stride = (int64_t)n_units * S2_TILE / 4;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int sg = 0;
    for (; sg + 4 <= n_seg; sg += 4) {
      const float4 v0 = p[(int64_t)sg * stride], v1 = p[(int64_t)(sg + 1) * stride], v2 = p[(int64_t)(sg + 2) * stride],
                   v3 = p[(int64_t)(sg + 3) * stride];
      a0 += ((double)v0.x + (double)v1.x) + ((double)v2.x + (double)v3.x);
      a1 += ((double)v0.y + (double)v1.y) + ((double)v2.y + (double)v3.y);
      a2 += ((double)v0.z + (double)v1.z) + ((double)v2.z + (double)v3.z);
      a3 += ((double)v0.w + (double)v1.w) + ((double)v2.w + (double)v3.w);
    }
    for (; sg < n_seg; ++sg) {
      const float4 v = p[(int64_t)sg * stride];
      a0 += (double)v.x; a1 += (double)v.y; a2 += (double)v.z; a3 += (double)v.w;
    }
    const float* sc = scale + l * dim;
    const double inv_i = 1.0 / (double)sc[gi];
    double* out = ws_cov + l * dim * dim + gi;                      // transposed position: [gj][gi]
    if (gi <= gj) out[(int64_t)gj * dim] += a0 * (inv_i / (double)sc[gj]);
    if (gi <= gj + 1) out[(int64_t)(gj + 1) * dim] += a1 * (inv_i / (double)sc[gj + 1]);
    if (gi <= gj + 2) out[(int64_t)(gj + 2) * dim] += a2 * (inv_i / (double)sc[gj + 2]);
    out[(int64_t)(gj + 3) * dim] += a3 * (inv_i / (double)sc[gj + 3]);
  }
}

// pivot[l, f] = mean of the first min(rows, 64) latents; scale[l, f] = power of two mapping the largest deviation from
// the pivot seen in those rows into [64, 128)  (1 if the feature is constant there).  Block = 32 features x 8 row groups.
// Also resets the overflow flag and the S' staging vector of the launch that follows (saves two memset nodes).
__global__ void pivot_scale_kernel(const float* __restrict__ x, int64_t rows, int64_t dim, int64_t row_stride,
                                   int64_t batch_stride, float* __restrict__ pivot, float* __restrict__ scale,
                                   int* __restrict__ flag, double* __restrict__ ws_sum) {
  __shared__ float part[8][33];
  __shared__ float piv[32];
  const int64_t l = blockIdx.y;
  const int64_t col = blockIdx.x * 32 + threadIdx.x;
  ptx::pdl_launch_dependents();      // the statistics kernel behind may start its prologue and its raw-tile loads now
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && threadIdx.y == 0) *flag = 0;
  if (threadIdx.y == 1 && col < dim) ws_sum[l * dim + col] = 0.0;
  const int64_t n = rows < 64 ? rows : 64;
  const float* base = x + l * batch_stride + col;
  float acc = 0.f;
  if (col < dim)
    for (int64_t r = threadIdx.y; r < n; r += 8) acc += base[r * row_stride];
  part[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0) {
#pragma unroll
    for (int g = 1; g < 8; ++g) acc += part[g][threadIdx.x];
    piv[threadIdx.x] = acc / (float)n;
  }
  __syncthreads();
  const float c = piv[threadIdx.x];
  float dev = 0.f;
  if (col < dim)
    for (int64_t r = threadIdx.y; r < n; r += 8) dev = fmaxf(dev, fabsf(base[r * row_stride] - c));
  part[threadIdx.y][threadIdx.x] = dev;
  __syncthreads();
  if (threadIdx.y == 0 && col < dim) {
#pragma unroll
    for (int g = 1; g < 8; ++g) dev = fmaxf(dev, part[g][threadIdx.x]);
    float s = 1.f;
    if (dev > 0.f && dev < 3.0e38f) {
      int e;
      frexpf(dev, &e);                       // dev = m 2^e, m in [0.5, 1)  ->  dev * 2^(7 - e) in [64, 128)
      e = 7 - e;
      e = e < -100 ? -100 : (e > 100 ? 100 : e);
      s = ldexpf(1.f, e);
    }
    pivot[l * dim + col] = c;
    scale[l * dim + col] = s;
  }
}

__global__ void zero_if_kernel(double* __restrict__ p, int64_t n, const int* __restrict__ flag) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  if (*flag == 0) return;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) p[e] = 0.0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Merge kernels: (records | partial tiles) -> running buffers, in one pass.  They replace "reduce into the fp64 staging
// area, then merge": P'[i][j] (i <= j) is summed over the partials in fp64, the power-of-two scales are divided out, the
// raw sums are rebuilt from the pivot-shifted ones ( sum x x^T = P' + c S'^T + S' c^T + n c c^T ,  sum x = S' + n c )
// and the accumulate / EMA rule (gaussian_model.py:104-108, utils/__init__.py:204-206) is applied to [i][j] and [j][i].
// If the overflow flag is up, the TF32 fallback has refilled the staging area (P' transposed at [j][i], S') and that is
// what gets merged instead - same kernel, no extra gated launch.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int HM_OUT = 64, HM_GROUPS = 4;

__device__ __forceinline__ void hm_store(const StatsRunning& run, int64_t l, int dim, int gi, int gj, double pp, double rows,
                                         const float* __restrict__ pivot, const double* __restrict__ ws_sum) {
  const double keep = run.decay < 0 ? 1.0 : run.decay, gain = run.decay < 0 ? 1.0 : 1.0 - run.decay;
  const double ci = pivot[l * dim + gi], cj = pivot[l * dim + gj];
  const double si = ws_sum[l * dim + gi], sj = ws_sum[l * dim + gj];
  const double v = pp + ci * sj + si * cj + rows * ci * cj;
  const int64_t e = l * dim * dim + (int64_t)gi * dim + gj;
  store_real(run.sum_cov, e, run.buf_dtype, load_real(run.sum_cov, e, run.buf_dtype) * keep + v * gain);
  if (gi != gj) {
    const int64_t e2 = l * dim * dim + (int64_t)gj * dim + gi;
    store_real(run.sum_cov, e2, run.buf_dtype, load_real(run.sum_cov, e2, run.buf_dtype) * keep + v * gain);
  } else {
    const int64_t f = l * dim + gj;
    store_real(run.sum, f, run.buf_dtype, load_real(run.sum, f, run.buf_dtype) * keep + (sj + rows * cj) * gain);
  }
}

// narrow: thread <-> packed index t = gj (gj + 1) / 2 + gi of the triangle (consecutive threads read consecutive floats of
// a record and write consecutive elements of row gj); HM_GROUPS thread groups share the records of an output.
__global__ void __launch_bounds__(HM_OUT * HM_GROUPS)
stats_h_merge_kernel(const float* __restrict__ rec_cov, int parts, const float* __restrict__ scale,
                     const float* __restrict__ pivot, const int* __restrict__ flag, const double* __restrict__ ws_cov,
                     const double* __restrict__ ws_sum, int dim, double rows, StatsRunning run) {
  __shared__ double red[HM_GROUPS][HM_OUT];
  ptx::pdl_wait();
  const int tri = dim * (dim + 1) / 2;
  const int blocks_per_l = (tri + HM_OUT - 1) / HM_OUT;
  const int64_t l = blockIdx.x / blocks_per_l;
  const int o = threadIdx.x % HM_OUT, g = threadIdx.x / HM_OUT;
  const int t = (blockIdx.x % blocks_per_l) * HM_OUT + o;
  const bool fallback = *flag != 0;
  int gj = 0, gi = 0;
  if (t < tri) {
    gj = (int)((sqrtf(8.f * (float)t + 1.f) - 1.f) * 0.5f);
    while ((gj + 1) * (gj + 2) / 2 <= t) ++gj;
    while (gj * (gj + 1) / 2 > t) --gj;
    gi = t - gj * (gj + 1) / 2;
  }
  double acc = 0.0;
  if (t < tri && !fallback) {
    const float* p = rec_cov + (int64_t)l * parts * SH_TRI + t;
    double a0 = 0.0, a1 = 0.0;
    int r = g;
    for (; r + HM_GROUPS < parts; r += 2 * HM_GROUPS) {
      a0 += (double)p[(int64_t)r * SH_TRI];
      a1 += (double)p[(int64_t)(r + HM_GROUPS) * SH_TRI];
    }
    if (r < parts) a0 += (double)p[(int64_t)r * SH_TRI];
    acc = a0 + a1;
  }
  red[g][o] = acc;
  __syncthreads();
  if (g != 0 || t >= tri) return;
  double pp;
  if (fallback) {
    pp = ws_cov[l * dim * dim + (int64_t)gj * dim + gi];
  } else {
#pragma unroll
    for (int k = 1; k < HM_GROUPS; ++k) acc += red[k][o];
    pp = acc * (1.0 / ((double)scale[l * dim + gi] * (double)scale[l * dim + gj]));
  }
  hm_store(run, l, dim, gi, gj, pp, rows, pivot, ws_sum);
  if (t == 0) {
    const double keep = run.decay < 0 ? 1.0 : run.decay, gain = run.decay < 0 ? 1.0 : 1.0 - run.decay;
    store_real(run.n_obs, l, run.n_dtype, load_real(run.n_obs, l, run.n_dtype) * keep + rows * gain);
  }
}

// wide: block <-> one row of one BW x BW unit (BW = 128: single-CTA kernel, 256: CTA-pair kernel); BW / 4 float4 columns x
// 1024 / BW segment groups, then one thread per column finalises
template <int BW>
__global__ void __launch_bounds__(256)
stats_h2_merge_kernel(const float* __restrict__ parts, int n_seg, int n_units, int upl, int nB,
                      const float* __restrict__ scale, const float* __restrict__ pivot, const int* __restrict__ flag,
                      const double* __restrict__ ws_cov, const double* __restrict__ ws_sum, int dim, double rows,
                      StatsRunning run) {
  constexpr int C4 = BW / 4, G = 256 / C4;
  __shared__ double red[G][C4][4];
  ptx::pdl_wait();
  const int u = blockIdx.x / BW, r = blockIdx.x % BW;
  const int64_t l = u / upl;
  int w = u % upl, bi = 0;
  while (w >= nB - bi) { w -= nB - bi; ++bi; }
  const int bj = bi + w;
  const int c4 = threadIdx.x % C4, g = threadIdx.x / C4;
  const int gi = bi * BW + r, gj0 = bj * BW + c4 * 4;
  const bool fallback = *flag != 0;
  const bool live = gi < dim && gj0 < dim && gi <= gj0 + 3;            // dim % 4 == 0: a group is inside or outside
  double a[4] = {0.0, 0.0, 0.0, 0.0};
  if (live && !fallback) {
    const float4* p = reinterpret_cast<const float4*>(parts + (int64_t)u * (BW * BW) + r * BW + c4 * 4);
    const int64_t stride = (int64_t)n_units * (BW * BW) / 4;
    for (int sg = g; sg < n_seg; sg += G) {
      const float4 v = p[(int64_t)sg * stride];
      a[0] += (double)v.x; a[1] += (double)v.y; a[2] += (double)v.z; a[3] += (double)v.w;
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) red[g][c4][k] = a[k];
  __syncthreads();
  // finalise with one thread per column (the read-modify-write latencies of the running buffers overlap)
  if (threadIdx.x < BW) {
    const int c = threadIdx.x, gj = bj * BW + c;
    if (gi < dim && gj < dim && gi <= gj) {
      double pp;
      if (fallback) {
        pp = ws_cov[l * dim * dim + (int64_t)gj * dim + gi];
      } else {
        pp = 0.0;
#pragma unroll
        for (int q = 0; q < G; ++q) pp += red[q][c / 4][c % 4];
        pp *= 1.0 / ((double)scale[l * dim + gi] * (double)scale[l * dim + gj]);
      }
      hm_store(run, l, dim, gi, gj, pp, rows, pivot, ws_sum);
    }
  }
  if (u % upl == 0 && r == 0 && threadIdx.x == 0) {
    const double keep = run.decay < 0 ? 1.0 : run.decay, gain = run.decay < 0 ? 1.0 : 1.0 - run.decay;
    store_real(run.n_obs, l, run.n_dtype, load_real(run.n_obs, l, run.n_dtype) * keep + rows * gain);
  }
}

int stats_h_merge(const StatsHPlan& plan, const float* pivot, const double* ws_cov, const double* ws_sum, int64_t L,
                  int64_t rows, int64_t dim, const StatsRunning& run, cudaStream_t st) {
  if (plan.mode == 1) {
    const int64_t tri = dim * (dim + 1) / 2, blocks = L * ceil_div(tri, HM_OUT);
    OTK_CUDA(launch_pdl_site(3, stats_h_merge_kernel, dim3((unsigned)blocks), dim3(HM_OUT * HM_GROUPS), 0, st, 1, plan.parts, plan.n_parts,
                        plan.scale, pivot, plan.flag, ws_cov, ws_sum, (int)dim, (double)rows, run));
  } else if (plan.mode == 2) {
    OTK_CUDA(launch_pdl_site(3, stats_h2_merge_kernel<128>, dim3((unsigned)(plan.n_units * 128)), dim3(256), 0, st, 1, plan.parts,
                        plan.n_parts, plan.n_units, plan.upl, plan.nB, plan.scale, pivot, plan.flag, ws_cov, ws_sum, (int)dim,
                        (double)rows, run));
  } else {   // mode 3: 256 x 256 units of the CTA-pair kernel
    OTK_CUDA(launch_pdl_site(3, stats_h2_merge_kernel<256>, dim3((unsigned)(plan.n_units * 256)), dim3(256), 0, st, 1, plan.parts,
                        plan.n_parts, plan.n_units, plan.upl, plan.nB, plan.scale, pivot, plan.flag, ws_cov, ws_sum, (int)dim,
                        (double)rows, run));
  }
  OTK_LAUNCH_CHECK();
  return OTK_OK;
}

constexpr int64_t SH_MAX_L = 1024;     // leading indices the narrow kernel takes (bounds the record area: 33 KB per CTA)
static int64_t stats_h_max_records(int64_t L) { const int64_t sms = sm_count(); return L > sms ? L : sms; }

size_t stats_h_extra_workspace(int64_t L, int64_t dim) {
  const size_t recs = (dim <= SH_T && L <= SH_MAX_L) ? align_up((size_t)stats_h_max_records(L) * SH_TRI * 4, 256) : 0;
  return align_up((size_t)L * dim * 4, 256) + 256 + recs;
}

bool stats_h_eligible(int64_t L, int64_t rows, int64_t dim) { return dim <= SH_T && rows >= 1 && L >= 1 && L <= SH_MAX_L; }

// Launches pivot/scale + the FP16-split kernel.  *plan->flag (device int) is non-zero afterwards iff a value left the FP16
// range, in which case the records are invalid and the caller's stream re-runs the call with the TF32 kernel
// (stats_umma.cu) into the staging area before stats_h_merge.
int stats_h_launch(const float* x, int64_t L, int64_t rows, int64_t dim, int64_t row_stride, int64_t batch_stride,
                   float* pivot, double* ws_cov, double* ws_sum, Arena& ar, cudaStream_t st, StatsHPlan* plan) {
  (void)ws_cov;
  float* scale = ar.take<float>((size_t)L * dim);
  int* flag = ar.take<int>(16);
  float* rec = ar.take<float>((size_t)stats_h_max_records(L) * SH_TRI);
  if (!ar.ok()) return OTK_ERR_WORKSPACE;
  CUtensorMap mX;
  if (!encode_map_f32_3d(&mX, x, dim, rows, L, row_stride, batch_stride, 32, SH_BK, /*atom32=*/true)) return 0;
  pivot_scale_kernel<<<dim3((unsigned)ceil_div(dim, 32), (unsigned)L), dim3(32, 8), 0, st>>>(x, rows, dim, row_stride, batch_stride,
                                                                                             pivot, scale, flag, ws_sum);
  OTK_LAUNCH_CHECK();
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    OTK_CUDA(cudaFuncSetAttribute(stats_h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SH_SMEM));
    attr_set[dev] = true;
  }
  // every leading index is cut into the same `parts` row ranges (multiples of 64 rows, at least 256 rows each)
  int64_t parts = sm_count() / L;
  if (parts < 1) parts = 1;
  int64_t range_len = ceil_div(ceil_div(rows, parts), SH_BK) * SH_BK;
  if (range_len < 256) range_len = 256;
  parts = ceil_div(rows, range_len);
  if (L * parts > stats_h_max_records(L)) return OTK_ERR_WORKSPACE;
  OTK_CUDA(launch_pdl(stats_h_kernel, dim3((unsigned)(L * parts)), dim3(SH_THREADS), SH_SMEM, st, 1, mX, (const float*)pivot,
                      (const float*)scale, (int)rows, (int)dim, (int)parts, (int)range_len, rec, ws_sum, flag));
  OTK_LAUNCH_CHECK();
  *plan = StatsHPlan{1, rec, (int)parts, 0, 0, 0, scale, flag};
  return 1;
}

size_t stats_h2_extra_workspace(int64_t L, int64_t dim) {
  const int64_t nB = ceil_div(dim, SH_T), units = L * nB * (nB + 1) / 2;
  const int64_t slots = units < S2_MAX_ITEMS ? S2_MAX_ITEMS : 0;      // not eligible beyond: no partial-tile area needed
  return align_up((size_t)slots * S2_TILE * 4, 256) + stats_h_extra_workspace(L, dim);
}

bool stats_h2_eligible(int64_t L, int64_t rows, int64_t dim) {
  const int64_t nB = ceil_div(dim, SH_T);
  return dim > SH_T && rows >= 1 && L * nB * (nB + 1) / 2 <= S2_MAX_ITEMS && rows <= INT32_MAX;
}

// FP16-split kernel for dim > 128: pivot/scale, then one launch + one partial-tile reduction per super-chunk of rows.
int stats_h2_launch(const float* x, int64_t L, int64_t rows, int64_t dim, int64_t row_stride, int64_t batch_stride,
                    float* pivot, double* ws_cov, double* ws_sum, Arena& ar, cudaStream_t st, StatsHPlan* plan) {
  const int64_t nB = ceil_div(dim, SH_T), upl = nB * (nB + 1) / 2, n_units = L * upl;
  float* scale = ar.take<float>((size_t)L * dim);
  int* flag = ar.take<int>(16);
  float* parts = ar.take<float>((size_t)S2_MAX_ITEMS * S2_TILE);
  if (!ar.ok()) return OTK_ERR_WORKSPACE;
  CUtensorMap mX;
  if (!encode_map_f32_3d(&mX, x, dim, rows, L, row_stride, batch_stride, 32, SH_BK, /*atom32=*/true)) return 0;
  pivot_scale_kernel<<<dim3((unsigned)ceil_div(dim, 32), (unsigned)L), dim3(32, 8), 0, st>>>(x, rows, dim, row_stride, batch_stride,
                                                                                             pivot, scale, flag, ws_sum);
  OTK_LAUNCH_CHECK();
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    OTK_CUDA(cudaFuncSetAttribute(stats_h2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, S2_SMEM));
    attr_set[dev] = true;
  }
  const int64_t sms = sm_count();
  // CTA-pair kernel (dim >= 512): 256 x 256 units, one super-chunk only (a pair item takes four 128 x 128 slots)
  {
    const int64_t nB2 = ceil_div(dim, 2 * SH_T), upl2 = nB2 * (nB2 + 1) / 2, n_units2 = L * upl2, pairs = sms / 2;
    const int64_t max_seg2 = (S2_MAX_ITEMS / 4) / (n_units2 > 0 ? n_units2 : 1);
    static const bool pair_ok = [] { const char* e = getenv("OTK_STATS_H2_PAIR"); return !(e && e[0] == '0'); }();   // tuning aid
    static const int64_t pair_min_dim = [] { const char* e = getenv("OTK_STATS_H2_PAIR_MIN_DIM"); return e ? (int64_t)atoi(e) : (int64_t)512; }();
    if (dim >= pair_min_dim && pair_ok && pairs >= 1 && max_seg2 >= 1 && rows <= max_seg2 * 2048) {
      static bool attr2_set[64] = {false};
      if (dev >= 0 && dev < 64 && !attr2_set[dev]) {
        OTK_CUDA(cudaFuncSetAttribute(stats_h2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, S2_SMEM));
        attr2_set[dev] = true;
      }
      int64_t seg_len = 0;
      for (int64_t k = 1;; ++k) {
        int64_t sgs = pairs * k / n_units2;
        if (sgs < 1) continue;
        if (sgs > max_seg2) sgs = max_seg2;
        seg_len = ceil_div(ceil_div(rows, sgs), SH_BK) * SH_BK;
        if (seg_len < 256) seg_len = 256;
        if (seg_len <= 2048 || sgs == max_seg2) break;
      }
      if (seg_len > 2048) seg_len = 2048;
      const int64_t n_seg = ceil_div(rows, seg_len), n_items = n_units2 * n_seg;
      if (n_seg <= max_seg2) {
        const unsigned groups = (unsigned)(n_items < pairs ? n_items : pairs);
        OTK_CUDA(launch_pdl(stats_h2_kernel<2>, dim3(groups * 2), dim3(S2_THREADS), S2_SMEM, st, 2, mX, (const float*)pivot,
                            (const float*)scale, (int)dim, (int)nB2, (int)upl2, (int)n_units2, (int)n_items, (int)seg_len, 0,
                            (int)rows, parts, ws_sum, flag));
        count_launch(1);
        *plan = StatsHPlan{3, parts, (int)n_seg, (int)n_units2, (int)upl2, (int)nB2, scale, flag};
        return 1;
      }
    }
  }
  const int64_t max_seg = S2_MAX_ITEMS / n_units;                       // segments per launch (>= 1 by eligibility)
  // One super-chunk (the streaming case: max_seg * 2048 rows = 209 k rows at d = 512): the partial tiles are merged
  // straight into the running buffers by stats_h2_merge_kernel.  Longer calls reduce every super-chunk into the fp64
  // staging area (cleared here) and leave the merge to the caller.
  const bool single = rows <= max_seg * 2048;
  if (!single) OTK_CUDA(cudaMemsetAsync(ws_cov, 0, (size_t)L * dim * dim * 8, st));
  for (int64_t row_lo = 0; row_lo < rows;) {
    // segment length: the smallest number of "rounds" k of the persistent grid whose segments are <= 2048 rows
    // (tensor-memory accumulation truncates) - items = n_units * segments just below k * #SMs
    int64_t left = rows - row_lo, seg_len = 0, n_seg = 0;
    for (int64_t k = 1;; ++k) {
      int64_t sgs = sms * k / n_units;
      if (sgs < 1) continue;
      if (sgs > max_seg) sgs = max_seg;
      seg_len = ceil_div(ceil_div(left, sgs), SH_BK) * SH_BK;
      if (seg_len < 256) seg_len = 256;
      if (seg_len <= 2048 || sgs == max_seg) break;
    }
    if (seg_len > 2048) seg_len = 2048;
    n_seg = ceil_div(left, seg_len);
    if (n_seg > max_seg) n_seg = max_seg;
    const int64_t row_hi = row_lo + n_seg * seg_len < rows ? row_lo + n_seg * seg_len : rows;
    const int64_t n_items = n_units * n_seg;
    const unsigned grid = (unsigned)(n_items < sms ? n_items : sms);
    stats_h2_kernel<1><<<grid, S2_THREADS, S2_SMEM, st>>>(mX, pivot, scale, (int)dim, (int)nB, (int)upl, (int)n_units, (int)n_items,
                                                      (int)seg_len, (int)row_lo, (int)row_hi, parts, ws_sum, flag);
    OTK_LAUNCH_CHECK();
    if (single) {
      if (row_hi != rows) return OTK_ERR_CUDA;   // cannot happen: `single` bounds the segment count
      *plan = StatsHPlan{2, parts, (int)n_seg, (int)n_units, (int)upl, (int)nB, scale, flag};
      return 1;
    }
    int64_t blocks = ceil_div(L * dim * dim / 4, 256);
    if (blocks > sms * 16) blocks = sms * 16;
    stats_h2_reduce_kernel<<<(unsigned)blocks, 256, 0, st>>>(parts, scale, L, (int)dim, (int)nB, (int)upl, (int)n_units, (int)n_seg,
                                                            ws_cov, ws_sum, flag);
    OTK_LAUNCH_CHECK();
    row_lo = row_hi;
  }
  *plan = StatsHPlan{0, parts, 0, (int)n_units, (int)upl, (int)nB, scale, flag};
  return 1;
}

int stats_zero_if(double* p, int64_t n, const int* flag, cudaStream_t st) {
  int64_t blocks = ceil_div(n, 256);
  if (blocks > 1024) blocks = 1024;
  OTK_CUDA(launch_pdl_site(1, zero_if_kernel, dim3((unsigned)blocks), dim3(256), 0, st, 1, p, n, flag));
  OTK_LAUNCH_CHECK();
  return OTK_OK;
}

}  // namespace otk
