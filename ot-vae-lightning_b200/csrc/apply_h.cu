// K7 with the fp32-accurate FP16 hi/lo split: Y = (X - 1 mean_s^T) T^T + 1 mean_t^T at twice the MMA rate of apply_umma.cu.
// Reference: apply_transport, ot/w2_utils.py:517-520 (B broadcast fp64 mat-vecs).
//
// FP16 has the 11-bit significand of TF32, so hi = fp16(v), lo = fp16(v - hi) with three kind::f16 MMAs per product
// (lo*hi + hi*lo + hi*hi) is as accurate as 3xTF32 - inside the FP16 range.  Both operands are therefore balanced with
// exact power-of-two scales:
//     x'_k    = (x_k - mean_s_k) s_k         s_k: the largest deviation in the first <= 64 latents mapped into [64, 128)
//     T'_{jk} = T_{jk} g_j / s_k             g_j: the largest |T_{jk} / s_k| of output row j mapped into [512, 1024)
//     y_j     = (sum_k T'_{jk} x'_k) / g_j + mean_t_j
// A converter that meets a value outside the FP16 range (or a non-finite one) raises a device flag; the TF32 kernel is
// enqueued right behind and recomputes Y only if the flag is up (apply_umma.cu, `run_flag`).
//
// Same structure as apply_umma.cu (persistent CTA pairs, 256 x 256 tiles; raw X tile by TMA, centred / scaled / split /
// packed by eight converter warps straight into tensor memory as the A operand; TMA-store epilogue), except that a stage
// of T planes holds 64 features (128-byte rows of FP16) and serves two 32-feature steps of X.
#include <cuda.h>
#include <cuda_fp16.h>

#include "apply_umma.cuh"
#include "role_timing.cuh"
#include "otk_ptx.cuh"
#include "tensormap.cuh"

namespace otk {

constexpr int HP_BM = 128, HP_BK = 32, HP_TK = 64;
constexpr int HP_XS = 4, HP_TS = 3, HP_AS = 8;       // ring depths: raw X (smem), T planes (smem), converted A (TMEM)
constexpr int HP_XTILE = HP_BM * HP_BK * 4;          // 16 KiB raw X tile
constexpr int HP_TPLANE = 128 * HP_TK * 2;           // 16 KiB: 128 rows x 64 features of one FP16 plane
constexpr int HP_TSTAGE = 2 * HP_TPLANE;             // hi + lo
constexpr int HP_OUT = 32 * 32 * 4;                  // 4 KiB staging tile per TMA store
constexpr int HP_THREADS = 18 * 32;                  // TMA, MMA | 8 converter warps | 8 epilogue warps
constexpr int HP_ACOL0 = 256;                        // TMEM columns [0,256): accumulators, [256,512): A ring (8 x 32)
constexpr int HP_SMEM = HP_XS * HP_XTILE + HP_TS * HP_TSTAGE + 16 * HP_OUT + 1024 + 512;

__device__ __forceinline__ void tmem_st8u(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
template <int CG>
__device__ __forceinline__ void umma_f16_ts_cg(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
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

// one pair of features of one latent: centre + scale (one FFMA each), split, pack; `chk` is poisoned (NaN) by any value
// outside the FP16 range or non-finite
__device__ __forceinline__ void split_pair(float x0, float x1, float s0, float s1, float n0, float n1, __half2& chk,
                                           uint32_t& hw, uint32_t& lw) {
  const float v0 = fmaf(x0, s0, n0), v1 = fmaf(x1, s1, n1);
  const __half2 h = __floats2half2_rn(v0, v1);                  // .x (low half) = the even feature
  const float2 hf = __half22float2(h);
  const float l0 = v0 - hf.x, l1 = v1 - hf.y;
  const __half2 lo = __floats2half2_rn(l0, l1);
  chk = __hfma2(lo, __float2half2_rn(0.f), chk);                            // one packed 0 * residual for the pair: +-inf / NaN -> NaN
  hw = *reinterpret_cast<const uint32_t*>(&h);
  lw = *reinterpret_cast<const uint32_t*>(&lo);
}

template <int CG, int BN>
__global__ void __launch_bounds__(HP_THREADS, 1)
apply_h_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapT_hi,
               const __grid_constant__ CUtensorMap mapT_lo, const __grid_constant__ CUtensorMap mapY,
               const float* __restrict__ sk, const float* __restrict__ nmsk, const float* __restrict__ inv_g,
               const float* __restrict__ mean_t, int rows, int dim, int m_tiles, int n_tiles, int total_tiles,
               int* __restrict__ overflow) {
  using namespace ptx;
  constexpr int NACC = 256 / BN;                      // accumulator buffers
  static_assert(BN / CG == 128, "each CTA stages 128 rows of the T tile");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* xraw = smem;
  uint8_t* tpl = xraw + HP_XS * HP_XTILE;
  uint8_t* outb = tpl + HP_TS * HP_TSTAGE;
  uint64_t* full_x = reinterpret_cast<uint64_t*>(outb + 16 * HP_OUT);   // TMA landed the raw X tile
  uint64_t* empty_x = full_x + HP_XS;                                  // converters have read it
  uint64_t* full_t = empty_x + HP_XS;                                  // T planes landed (leader: both CTAs' halves)
  uint64_t* empty_t = full_t + HP_TS;                                  // MMAs reading them retired
  uint64_t* ready_a = empty_t + HP_TS;                                 // A planes in TMEM written (leader: both CTAs)
  uint64_t* empty_a = ready_a + HP_AS;                                 // MMAs reading them retired
  uint64_t* acc_full = empty_a + HP_AS;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;
  const int group = CG == 2 ? blockIdx.x / 2 : blockIdx.x;           // tile-processing unit (CTA or CTA pair)
  const int n_groups = CG == 2 ? gridDim.x / 2 : gridDim.x;
  const int num_k = (dim + HP_BK - 1) / HP_BK;                        // 32-feature steps of X
  const int num_t = (num_k + 1) / 2;                                  // 64-feature stages of T per tile
  const int tiles_per_l = m_tiles * n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapX); tma_prefetch_desc(&mapT_hi); tma_prefetch_desc(&mapT_lo); tma_prefetch_desc(&mapY);
    for (int s = 0; s < HP_XS; ++s) { mbar_init(&full_x[s], 1); mbar_init(&empty_x[s], 8); }
    for (int s = 0; s < HP_TS; ++s) { mbar_init(&full_t[s], 1); mbar_init(&empty_t[s], 1); }
    for (int s = 0; s < HP_AS; ++s) { mbar_init(&ready_a[s], 8 * CG); mbar_init(&empty_a[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 8 * CG); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc_cg<CG>(tmem_slot, 512); tmem_relinquish_cg<CG>(); }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer: raw X tile of this CTA's 128 latents every step, this CTA's 128 rows of the T planes every
    // second step (warp-uniform loop, one elected lane issues) =====
    const uint32_t full_t_leader0 = CG == 2 ? map_to_cta(smem_u32(&full_t[0]), 0) : smem_u32(&full_t[0]);
    int it = 0, jt = 0;
    S2_T0
    for (int tile = group; tile < total_tiles; tile += n_groups) {
      const int l = tile / tiles_per_l, rem = tile % tiles_per_l;
      const int m0 = (rem / n_tiles) * (HP_BM * CG) + (int)rank * HP_BM;
      const int n0 = (rem % n_tiles) * BN + (int)rank * 128;
      for (int kt = 0; kt < num_k; ++kt, ++it) {
        S2_COUNT
        S2_TICK(s2_c)
        if ((kt & 1) == 0) {
          const int st = jt % HP_TS;
          mbar_wait(&empty_t[st], ((jt / HP_TS) & 1) ^ 1);
          S2_TICK(s2_a)
          uint8_t* td = tpl + st * HP_TSTAGE;
          if (elect_one()) {
            if constexpr (CG == 1) {
              mbar_arrive_expect_tx(&full_t[st], HP_TSTAGE);
              tma_load_3d(td, &mapT_hi, kt * HP_BK, n0, l, &full_t[st]);
              tma_load_3d(td + HP_TPLANE, &mapT_lo, kt * HP_BK, n0, l, &full_t[st]);
            } else {
              if (rank == 0) mbar_arrive_expect_tx(&full_t[st], 2 * HP_TSTAGE);   // both CTAs' bytes land on the leader's barrier
              const uint32_t bar = full_t_leader0 + st * 8;
              tma_load_3d_cg2(td, &mapT_hi, kt * HP_BK, n0, l, bar);
              tma_load_3d_cg2(td + HP_TPLANE, &mapT_lo, kt * HP_BK, n0, l, bar);
            }
          }
          __syncwarp();
          ++jt;
        }
        const int sx = it % HP_XS;
        S2_TICK(s2_c)
        mbar_wait(&empty_x[sx], ((it / HP_XS) & 1) ^ 1);
        S2_TICK(s2_b)
        if (elect_one()) {
          mbar_arrive_expect_tx(&full_x[sx], HP_XTILE);
          tma_load_3d(xraw + sx * HP_XTILE, &mapX, kt * HP_BK, m0, l, &full_x[sx]);
        }
        __syncwarp();
      }
    }
    S2_REPORT("tma", "wait empty_t", "wait empty_x", "issue", "-", "-")
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA of the pair only): warp-uniform loop, one elected lane issues =====
    if (rank == 0) {
      const uint32_t idesc = idesc_f16(HP_BM * CG, BN);
      int it = 0, ti = 0, jt = 0;
      S2_T0
      for (int tile = group; tile < total_tiles; tile += n_groups, ++ti) {
        const int a = ti % NACC;
        S2_TICK(s2_d)
        mbar_wait(&acc_empty[a], ((ti / NACC) & 1) ^ 1);
        S2_TICK(s2_a)
        tc_fence_after();
        const uint32_t acc = tmem_base + a * BN;
        for (int kt = 0; kt < num_k; ++kt, ++it) {
          const int js = jt + (kt >> 1), st = js % HP_TS, sa = it % HP_AS;
          S2_COUNT
          S2_TICK(s2_d)
          if ((kt & 1) == 0) mbar_wait(&full_t[st], (js / HP_TS) & 1);
          S2_TICK(s2_b)
          mbar_wait(&ready_a[sa], (it / HP_AS) & 1);
          S2_TICK(s2_c)
          tc_fence_after();
          const uint32_t tb = smem_u32(tpl + st * HP_TSTAGE) + (uint32_t)(kt & 1) * 64;
          const uint32_t ab = tmem_base + HP_ACOL0 + sa * 32;
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < HP_BK / 16; ++kk) {
              const uint64_t t_hi = smem_desc_sw128(tb + kk * 32, 16, 1024);
              const uint64_t t_lo = smem_desc_sw128(tb + HP_TPLANE + kk * 32, 16, 1024);
              umma_f16_ts_cg<CG>(acc, ab + 16 + kk * 8, t_hi, idesc, (kt | kk) != 0);   // lo * hi
              umma_f16_ts_cg<CG>(acc, ab + kk * 8, t_lo, idesc, 1);                     // hi * lo
              umma_f16_ts_cg<CG>(acc, ab + kk * 8, t_hi, idesc, 1);                     // hi * hi
            }
            if ((kt & 1) || kt == num_k - 1) umma_commit_cg<CG>(&empty_t[st]);
            umma_commit_cg<CG>(&empty_a[sa]);
            if (kt == num_k - 1) umma_commit_cg<CG>(&acc_full[a]);
          }
          __syncwarp();
        }
        jt += num_t;
      }
      S2_REPORT("mma", "wait acc_empty", "wait full_t", "wait ready_a", "issue", "-")
    }
  } else if (warp < 10) {
    // ===== converters: thread <-> latent row r (TMEM lane r); warps 2-5 take features 0-15 of the step, 6-9 take 16-31
    const int q = warp % 4, half = (warp - 2) / 4;
    const int r = q * 32 + lane;
    const uint32_t row_off = (uint32_t)r * 128;
    const uint32_t ready_addr = CG == 2 ? map_to_cta(smem_u32(&ready_a[0]), 0) : smem_u32(&ready_a[0]);
    const uint32_t xbase = smem_u32(xraw);
    __half2 chk = __float2half2_rn(0.f);
    int it = 0;
    S2_T0
    for (int tile = group; tile < total_tiles; tile += n_groups) {
      const int l = tile / tiles_per_l, rem = tile % tiles_per_l;
      const int m0 = (rem / n_tiles) * (HP_BM * CG) + (int)rank * HP_BM;
      const bool live = m0 + r < rows;                             // rows past the end (zero-filled by TMA) get s = -mean s = 0
      const float* sl = sk + (int64_t)l * dim;
      const float* nl = nmsk + (int64_t)l * dim;
      for (int kt = 0; kt < num_k; ++kt, ++it) {
        const int sx = it % HP_XS, sa = it % HP_AS;
        S2_COUNT
        S2_TICK(s2_e)
        mbar_wait(&full_x[sx], (it / HP_XS) & 1);
        S2_TICK(s2_a)
        float4 x[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t chunk = (uint32_t)(half * 4 + c);
          x[c] = lds128(xbase + sx * HP_XTILE + row_off + ((chunk ^ (uint32_t)(r & 7)) * 16));
        }
        uint32_t hw[8], lw[8];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int k = kt * HP_BK + half * 16 + c * 4;
          float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f), n4 = s4;
          if (k < dim && live) { s4 = __ldg(reinterpret_cast<const float4*>(sl + k)); n4 = __ldg(reinterpret_cast<const float4*>(nl + k)); }
          split_pair(x[c].x, x[c].y, s4.x, s4.y, n4.x, n4.y, chk, hw[2 * c], lw[2 * c]);
          split_pair(x[c].z, x[c].w, s4.z, s4.w, n4.z, n4.w, chk, hw[2 * c + 1], lw[2 * c + 1]);
        }
        __syncwarp();
        S2_TICK(s2_b)
        if (lane == 0) mbar_arrive(&empty_x[sx]);                 // raw tile consumed (values are in registers)
        mbar_wait(&empty_a[sa], ((it / HP_AS) & 1) ^ 1);
        S2_TICK(s2_c)
        tc_fence_after();
        const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + HP_ACOL0 + sa * 32 + half * 8;
        tmem_st8u(ta, hw);
        tmem_st8u(ta + 16, lw);
        tmem_st_wait();
        S2_TICK(s2_d)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(ready_addr + sa * 8);
      }
    }
    if (warp == 2) { S2_REPORT("convert", "wait full_x", "convert", "wait empty_a", "tmem st", "arrive+loop") }
    if (!(__low2float(chk) == 0.f && __high2float(chk) == 0.f)) atomicOr(overflow, 1);
  } else {
    // ===== epilogue: TMEM -> / g_j + mean_t -> 32x32 swizzled staging tile -> TMA store (eight warps: two per TMEM lane
    // quarter, half of the tile's columns each) =====
    const int q = warp % 4, half = (warp - 10) / 4;
    constexpr int HC = BN / 2;                                  // columns drained by this warp
    const uint32_t stage0 = smem_u32(outb) + (uint32_t)(warp - 10) * 2 * HP_OUT;
    const uint32_t acc_empty_addr = CG == 2 ? map_to_cta(smem_u32(&acc_empty[0]), 0) : smem_u32(&acc_empty[0]);
    int ti = 0, nstore = 0;
    S2_T0
    for (int tile = group; tile < total_tiles; tile += n_groups, ++ti) {
      const int l = tile / tiles_per_l, rem = tile % tiles_per_l;
      const int m0 = (rem / n_tiles) * (HP_BM * CG) + (int)rank * HP_BM + q * 32;
      const int n0 = (rem % n_tiles) * BN + half * HC;
      const int a = ti % NACC;
      const float* mt = mean_t + (int64_t)l * dim;
      const float* ig = inv_g + (int64_t)l * dim;
      S2_COUNT
      S2_TICK(s2_e)
      mbar_wait(&acc_full[a], (ti / NACC) & 1);
      S2_TICK(s2_a)
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < HC; c0 += 32, ++nstore) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + a * BN + half * HC + c0, v);
        tmem_ld_wait();
        S2_TICK(s2_b)
        if (c0 + 32 == HC) {                                      // this warp's share is read: hand the accumulator back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(acc_empty_addr + a * 8);
        }
        const uint32_t buf = stage0 + (uint32_t)(nstore & 1) * HP_OUT;
        if (elect_one()) tma_store_wait_read<1>();                // the store issued two chunks ago has read this buffer
        __syncwarp();
        S2_TICK(s2_c)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int n = n0 + c0 + c * 4;
          float4 b = make_float4(0.f, 0.f, 0.f, 0.f), g = b;
          if (n < dim) { b = __ldg(reinterpret_cast<const float4*>(mt + n)); g = __ldg(reinterpret_cast<const float4*>(ig + n)); }
          sts128(buf + (uint32_t)lane * 128 + (((uint32_t)c ^ (uint32_t)(lane & 7)) * 16),
                 make_float4(fmaf(v[4 * c], g.x, b.x), fmaf(v[4 * c + 1], g.y, b.y), fmaf(v[4 * c + 2], g.z, b.z),
                             fmaf(v[4 * c + 3], g.w, b.w)));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (n0 + c0 < dim && m0 < rows) {
          if (elect_one()) {
            tma_store_3d(&mapY, buf, n0 + c0, m0, l);
            tma_store_commit();
          }
        }
        S2_TICK(s2_d)
      }
    }
    if (warp == 10) { S2_REPORT("epilogue", "wait acc_full", "tmem ld", "store-buffer wait", "scale+sts+store", "loop") }
    if (elect_one()) tma_store_wait_all<0>();
    __syncwarp();
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc_cg<CG>(tmem_base, 512); }
}

// per input feature k: s_k = power of two mapping the largest |x_k - mean_s_k| of the first min(rows, 64) latents into
// [64, 128) (1 if there is no deviation), nms_k = -mean_s_k s_k.  Block = 32 features x 8 row groups; also clears the flag.
__global__ void apply_scale_kernel(const float* __restrict__ x, int64_t rows, int64_t dim, const float* __restrict__ mean_s,
                                   float* __restrict__ sk, float* __restrict__ nmsk, int* __restrict__ overflow) {
  __shared__ float part[8][33];
  const int64_t l = blockIdx.y;
  const int64_t col = blockIdx.x * 32 + threadIdx.x;
  const int64_t n = rows < 64 ? rows : 64;
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && threadIdx.y == 0) *overflow = 0;
  const float c = col < dim ? mean_s[l * dim + col] : 0.f;
  float dev = 0.f;
  if (col < dim)
    for (int64_t r = threadIdx.y; r < n; r += 8) dev = fmaxf(dev, fabsf(x[(l * rows + r) * dim + col] - c));
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
    sk[l * dim + col] = s;
    nmsk[l * dim + col] = -c * s;
  }
}

// prepared form: s_k from the source standard deviation (8 sigma_k mapped into [64, 128), i.e. +-4000 sigma fit the FP16
// range), nms_k = -mean_s_k s_k
__global__ void apply_scale_from_var_kernel(const float* __restrict__ var_s, const float* __restrict__ mean_s, int64_t n,
                                            float* __restrict__ sk, float* __restrict__ nmsk) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= n) return;
  const float dev = 8.f * sqrtf(fmaxf(var_s[e], 0.f));
  float s = 1.f;
  if (dev > 0.f && dev < 3.0e38f) {
    int ex;
    frexpf(dev, &ex);
    ex = 7 - ex;
    ex = ex < -100 ? -100 : (ex > 100 ? 100 : ex);
    s = ldexpf(1.f, ex);
  }
  sk[e] = s;
  nmsk[e] = -mean_s[e] * s;
}
__global__ void clear_flag_kernel(int* flag) { *flag = 0; }

// one warp per output row j of T: g_j = power of two mapping max_k |T_jk / s_k| into [512, 1024); FP16 hi / lo planes of
// T_jk g_j / s_k; inv_g[j] = 1 / g_j
__global__ void apply_split_t_kernel(const float* __restrict__ T, const float* __restrict__ sk, int64_t L, int64_t dim,
                                     __half* __restrict__ hi, __half* __restrict__ lo, float* __restrict__ inv_g) {
  const int lane = threadIdx.x % 32;
  const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / 32;
  if (row >= L * dim) return;
  const int64_t l = row / dim;
  const float* t = T + row * dim;
  const float* s = sk + l * dim;
  float mx = 0.f;
  for (int64_t k = lane; k < dim; k += 32) mx = fmaxf(mx, fabsf(t[k] / s[k]));
  mx = warp_max(mx);
  float g = 1.f;
  if (mx > 0.f && mx < 3.0e38f) {
    int e;
    frexpf(mx, &e);
    e = 10 - e;
    e = e < -100 ? -100 : (e > 100 ? 100 : e);
    g = ldexpf(1.f, e);
  }
  for (int64_t k = lane; k < dim; k += 32) {
    const float v = t[k] / s[k] * g;                       // exact: s, g are powers of two
    const __half h = __float2half_rn(v);
    hi[row * dim + k] = h;
    lo[row * dim + k] = __float2half_rn(v - __half2float(h));
  }
  if (lane == 0) inv_g[row] = 1.f / g;
}

template <int CG, int BN>
static int launch_apply_h(const CUtensorMap& mX, const CUtensorMap& mTh, const CUtensorMap& mTl, const CUtensorMap& mY,
                          const float* sk, const float* nmsk, const float* inv_g, const float* mt32, int64_t L, int64_t rows,
                          int64_t dim, int* flag, cudaStream_t st) {
  auto kern = apply_h_kernel<CG, BN>;
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    OTK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, HP_SMEM));
    attr_set[dev] = true;
  }
  const int64_t m_tiles = ceil_div(rows, HP_BM * CG), n_tiles = ceil_div(dim, BN);
  const int64_t total = L * m_tiles * n_tiles;
  if (total > INT32_MAX) return 0;
  const int64_t max_groups = sm_count() / CG;
  const unsigned groups = (unsigned)(total < max_groups ? total : max_groups);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(groups * CG);
  cfg.blockDim = dim3(HP_THREADS);
  cfg.dynamicSmemBytes = HP_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  OTK_CUDA(cudaLaunchKernelEx(&cfg, kern, mX, mTh, mTl, mY, sk, nmsk, inv_g, mt32, (int)rows, (int)dim, (int)m_tiles,
                              (int)n_tiles, (int)total, flag));
  OTK_LAUNCH_CHECK();
  return 1;
}

size_t apply_h_workspace_bytes(int64_t L, int64_t dim) {
  return 2 * align_up((size_t)L * dim * dim * 2, 256) + 3 * align_up((size_t)L * dim * 4, 256) + 512;
}

bool apply_h_eligible(int64_t L, int64_t rows, int64_t dim) {
  return dim >= 64 && dim % 8 == 0 && rows >= 1 && L <= 65535 && rows <= INT32_MAX;
}

struct ApplyHPlanes { __half *Thi, *Tlo; float *sk, *nmsk, *inv_g; int* flag; };
static bool carve_apply_h(Arena& ar, int64_t L, int64_t dim, ApplyHPlanes* p) {
  p->Thi = ar.take<__half>((size_t)L * dim * dim);
  p->Tlo = ar.take<__half>((size_t)L * dim * dim);
  p->sk = ar.take<float>((size_t)L * dim);
  p->nmsk = ar.take<float>((size_t)L * dim);
  p->inv_g = ar.take<float>((size_t)L * dim);
  p->flag = ar.take<int>(16);
  return ar.ok();
}
static int run_apply_h(const float* x, int64_t L, int64_t rows, int64_t dim, const float* mt32, const ApplyHPlanes& p, float* y,
                       bool pair, cudaStream_t st, int64_t x_row_stride = 0, int64_t x_batch_stride = 0) {
  CUtensorMap mX, mTh, mTl, mY;
  const int64_t xrs = x_row_stride > 0 ? x_row_stride : dim, xbs = x_batch_stride > 0 ? x_batch_stride : rows * dim;
  if (!encode_map_f32_3d(&mX, x, dim, rows, L, xrs, xbs, 32, HP_BM)) return 0;
  if (!encode_map_f16_3d(&mTh, p.Thi, dim, dim, L, dim, dim * dim, HP_TK, 128)) return 0;
  if (!encode_map_f16_3d(&mTl, p.Tlo, dim, dim, L, dim, dim * dim, HP_TK, 128)) return 0;
  if (!encode_map_f32_3d(&mY, y, dim, rows, L, dim, rows * dim, 32, 32)) return 0;
  return pair ? launch_apply_h<2, 256>(mX, mTh, mTl, mY, p.sk, p.nmsk, p.inv_g, mt32, L, rows, dim, p.flag, st)
              : launch_apply_h<1, 128>(mX, mTh, mTl, mY, p.sk, p.nmsk, p.inv_g, mt32, L, rows, dim, p.flag, st);
}

// FP16-split transport.  Returns 1 if launched (*flag_out: device int, non-zero afterwards iff a value left the FP16 range
// and Y must be recomputed by the TF32 kernel), 0 if not eligible, < 0 on error.
int apply_h_try(const float* x, int64_t L, int64_t rows, int64_t dim, const float* ms32, const float* mt32, const float* T32,
                float* y, Arena& ar, bool pair, cudaStream_t st, int** flag_out) {
  if (!apply_h_eligible(L, rows, dim) || !tensormap_encoder()) return 0;
  ApplyHPlanes p;
  if (!carve_apply_h(ar, L, dim, &p)) return OTK_ERR_WORKSPACE;
  apply_scale_kernel<<<dim3((unsigned)ceil_div(dim, 32), (unsigned)L), dim3(32, 8), 0, st>>>(x, rows, dim, ms32, p.sk, p.nmsk, p.flag);
  OTK_LAUNCH_CHECK();
  apply_split_t_kernel<<<(unsigned)ceil_div(L * dim * 32, 256), 256, 0, st>>>(T32, p.sk, L, dim, p.Thi, p.Tlo, p.inv_g);
  OTK_LAUNCH_CHECK();
  int used = run_apply_h(x, L, rows, dim, mt32, p, y, pair, st);
  if (used == 1) *flag_out = p.flag;
  return used;
}

// prepared form: the planes are built once from (mean_s, T, var_s); the arena must be carved identically in both calls
int apply_h_prepare(int64_t L, int64_t dim, const float* ms32, const float* T32, const float* var32, Arena& ar, cudaStream_t st) {
  ApplyHPlanes p;
  if (!carve_apply_h(ar, L, dim, &p)) return OTK_ERR_WORKSPACE;
  if (!apply_h_eligible(L, 1, dim)) return 0;
  apply_scale_from_var_kernel<<<(unsigned)ceil_div(L * dim, 256), 256, 0, st>>>(var32, ms32, L * dim, p.sk, p.nmsk);
  OTK_LAUNCH_CHECK();
  apply_split_t_kernel<<<(unsigned)ceil_div(L * dim * 32, 256), 256, 0, st>>>(T32, p.sk, L, dim, p.Thi, p.Tlo, p.inv_g);
  OTK_LAUNCH_CHECK();
  return 1;
}
int apply_h_run_prepared(const float* x, int64_t L, int64_t rows, int64_t dim, const float* mt32, float* y, Arena& ar, bool pair,
                         cudaStream_t st, int** flag_out, int64_t x_row_stride, int64_t x_batch_stride) {
  ApplyHPlanes p;
  if (!carve_apply_h(ar, L, dim, &p)) return OTK_ERR_WORKSPACE;
  if (!apply_h_eligible(L, rows, dim) || !tensormap_encoder()) return 0;
  clear_flag_kernel<<<1, 1, 0, st>>>(p.flag);
  OTK_LAUNCH_CHECK();
  int used = run_apply_h(x, L, rows, dim, mt32, p, y, pair, st, x_row_stride, x_batch_stride);
  if (used == 1) *flag_out = p.flag;
  return used;
}

}  // namespace otk
