// K8/K9 fused: log-domain Sinkhorn between point clouds with the cost tile recomputed on the tensor cores.
// Reference recurrences: sinkhorn_log, ot/w2_utils.py:301-319 ; cost: |x-y|^2 (w2_utils.py:121-125) * scale.
//
// Never materialises the N x M cost / kernel matrix.  With  Cr_ij = -scale |x_i - y_j|^2 / reg  one half-step is
//     pot_r = log(marg_r + 1e-8) + nrm_r - LSE_q( bias_q + gamma p_r.q_q ),   nrm = scale |p|^2 / reg,
//     bias_q = pot_q - scale |q_q|^2 / reg,  gamma = 2 scale / reg,
// i.e. a FlashAttention-shaped pass: S = P Q^T on tcgen05 (kind::tf32, operands rounded once to TF32 by the prep
// kernel) into a ring of four 128x128 fp32 TMEM buffers, then an online log-sum-exp straight out of TMEM
// (thread <-> row, base-2 exponentials on the MUFU pipe), no shared-memory round trip for S.
//
// CTA = 320 threads: warp 0 TMA producer (P block once, Q tiles through a 2-stage ring), warp 1 TMEM alloc + MMA
// issuer, warps 2-5 / 6-9 two softmax warpgroups that alternate over the Q tiles.  One CTA = one block of 128 P rows
// x one contiguous range of Q tiles; the (max, sum) partials of the ranges - and, in the row-sharded multi-GPU path,
// of the ranks - are merged by sk_finalize_kernel.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cfloat>

#include "otk_ptx.cuh"
#include "sinkhorn_umma.cuh"
#include "tensormap.cuh"

namespace otk {

constexpr int FS_SUB = 128, FS_BM = 2 * FS_SUB, FS_BN = 128, FS_SLAB = 64 /*fp16 per 128-byte row*/;
constexpr int FS_SLAB_BYTES = 128 * 128, FS_MAX_SLABS = 2, FS_QSTAGES = 3, FS_SBUF = 4;
constexpr int FS_THREADS = 64 + 4 * 128;   // TMA warp, MMA warp, four softmax warpgroups
constexpr int FS_P_BYTES = 2 * FS_MAX_SLABS * FS_SLAB_BYTES;       // 64 KiB: two 128-row sub-blocks
constexpr int FS_Q_BYTES = FS_MAX_SLABS * FS_SLAB_BYTES;            // 32 KiB per stage
constexpr int FS_SMEM = FS_P_BYTES + FS_QSTAGES * FS_Q_BYTES + 1024 /*align*/ + 4096 /*bias staging*/ + 512 /*barriers*/;
constexpr float FS_NEG = -1.0e30f;
constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;

struct FsState { int done; int iters; };

__device__ __forceinline__ float ex2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// packed fp32 pairs (FFMA2 / FADD2: one issue slot for two lanes of work) and the 3-input maximum (FMNMX3)
typedef unsigned long long f2;
__device__ __forceinline__ f2 pack2(float a, float b) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(f2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float max3(float a, float b, float c) { float r; asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ void named_bar(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// Operands are FP16 planes of the points divided by sigma = max |coordinate| (10-bit mantissa, the same rounding as
// TF32, at twice the tensor rate and half the bytes); *sig2 = sigma^2 rescales the dot products.
// out: part_m / part_l [split][Np] (base-2 max and sum of 2^(t - max)); COST also part_c = sum 2^(t-max) (nq_q - 2 p.q)
template <bool COST>
__global__ void __launch_bounds__(FS_THREADS, 1)
fused_lse_kernel(const __grid_constant__ CUtensorMap mapP, const __grid_constant__ CUtensorMap mapQ,
                 const float* __restrict__ bias2, const float* __restrict__ nq, float g2_unit,
                 const float* __restrict__ sig2, int Np, int Nq, int dim, int tiles_per_split, float* __restrict__ part_m,
                 float* __restrict__ part_l, float* __restrict__ part_c, const FsState* __restrict__ state,
                 const float* __restrict__ lse_prev, const float* __restrict__ dmax) {
  using namespace ptx;
  if (state && state->done) return;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sP = smem;                                   // [sub][slab][128 rows x 128 B]
  uint8_t* sQ = smem + FS_P_BYTES;                      // [stage][slab][128 rows x 128 B]
  float* s_bias = reinterpret_cast<float*>(smem + FS_P_BYTES + FS_QSTAGES * FS_Q_BYTES);   // [2 buffers][4 groups][64]
  float* s_nq = s_bias + 512;                                                                // [2 buffers][4 groups][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_nq + 512);
  uint64_t* p_full = bars;
  uint64_t* q_full = bars + 1;
  uint64_t* q_empty = q_full + FS_QSTAGES;
  uint64_t* s_full = q_empty + FS_QSTAGES;
  uint64_t* s_empty = s_full + FS_SBUF;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_empty + FS_SBUF);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int pb = blockIdx.x, split = blockIdx.y;
  const int total_tiles = (Nq + FS_BN - 1) / FS_BN;
  const int jt0 = split * tiles_per_split;
  const int jt1 = min(total_tiles, jt0 + tiles_per_split);
  const int n_tiles = max(0, jt1 - jt0);
  const int nks = (dim + FS_SLAB - 1) / FS_SLAB;
  const float g2 = g2_unit * (*sig2);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapP); tma_prefetch_desc(&mapQ);
    mbar_init(p_full, 1);
    for (int s = 0; s < FS_QSTAGES; ++s) { mbar_init(&q_full[s], 1); mbar_init(&q_empty[s], 1); }
    for (int b = 0; b < FS_SBUF; ++b) { mbar_init(&s_full[b], 1); mbar_init(&s_empty[b], 256); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(p_full, (uint32_t)(2 * nks) * FS_SLAB_BYTES);
      for (int sub = 0; sub < 2; ++sub)
        for (int sl = 0; sl < nks; ++sl)
          tma_load_2d(sP + (sub * FS_MAX_SLABS + sl) * FS_SLAB_BYTES, &mapP, sl * FS_SLAB, pb * FS_BM + sub * FS_SUB, p_full);
      for (int idx = 0; idx < n_tiles; ++idx) {
        const int s = idx % FS_QSTAGES, it = idx / FS_QSTAGES;
        mbar_wait(&q_empty[s], (it & 1) ^ 1);
        mbar_arrive_expect_tx(&q_full[s], (uint32_t)nks * FS_SLAB_BYTES);
        for (int sl = 0; sl < nks; ++sl)
          tma_load_2d(sQ + s * FS_Q_BYTES + sl * FS_SLAB_BYTES, &mapQ, sl * FS_SLAB, (jt0 + idx) * FS_BN, &q_full[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = idesc_f16(FS_SUB, FS_BN);
      mbar_wait(p_full, 0);
      for (int idx = 0; idx < n_tiles; ++idx) {
        const int s = idx % FS_QSTAGES, pair = idx % 2;
        mbar_wait(&q_full[s], (idx / FS_QSTAGES) & 1);
        const uint32_t qbase = smem_u32(sQ + s * FS_Q_BYTES);
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
          const int b = 2 * pair + sub;
          mbar_wait(&s_empty[b], ((idx / 2) & 1) ^ 1);
          tc_fence_after();
          const uint32_t pbase = smem_u32(sP + sub * FS_MAX_SLABS * FS_SLAB_BYTES);
          for (int sl = 0; sl < nks; ++sl) {
            const int ksteps = min(4, (dim - sl * FS_SLAB + 15) / 16);
#pragma unroll 4
            for (int kk = 0; kk < ksteps; ++kk) {
              const uint64_t pd = smem_desc_sw128(pbase + sl * FS_SLAB_BYTES + kk * 32, 16, 1024);
              const uint64_t qd = smem_desc_sw128(qbase + sl * FS_SLAB_BYTES + kk * 32, 16, 1024);
              umma_f16(tmem_base + b * FS_BN, pd, qd, idesc, (sl | kk) != 0);
            }
          }
          umma_commit(&s_full[b]);
        }
        umma_commit(&q_empty[s]);
      }
    }
  } else {
    // ===== online log-sum-exp: four warpgroups = (P sub-block) x (column half of the Q tile); thread <-> row.
    // Four softmax warps per SM sub-partition keep the MUFU pipe fed while the others wait on TMEM / the FMA pipe.
    const int grp = (warp - 2) / 4;                      // 0..3
    const int sub = grp >> 1, half = grp & 1;
    const int tid = (warp - 2) % 4 * 32 + lane;          // 0..127 inside the warpgroup
    const int q4 = warp % 4;                             // TMEM lane quarter of this warp
    const float k2 = COST ? -2.f / g2_unit : 0.f;       // t - bias2 = g2_unit * (x.y)  ->  -2 x.y
    float m = -3.0e38f, l = 0.f, lc = 0.f;
    // Bounded-shift mode (Sinkhorn iterations >= 2): every row's LSE of the previous iteration is known and the biases
    // moved by at most *dmax since, so U = lse_prev + dmax + 1 bounds every exponent of the row from above and
    // overestimates the new LSE by < 2 dmax + 2.  With dmax < 32 the sum of 2^(t - U) can neither overflow nor
    // vanish, and the running maximum (a second sweep over the tile plus a rescale) is not needed at all.
    bool fast = false;
    if constexpr (!COST) {
      if (lse_prev != nullptr) {
        const float dm = *dmax;
        if (dm < 32.f) {
          fast = true;
          const int r = min(pb * FS_BM + sub * FS_SUB + q4 * 32 + lane, Np - 1);
          m = lse_prev[r] + dm + 1.f;
        }
      }
    }
    const f2 NU = pack2(-m, -m);
    // bias / |q|^2 of a Q tile are staged in shared memory by the first 64 threads of the warpgroup, double-buffered and
    // loaded one tile ahead, so a tile costs one named barrier and no exposed global-load latency
    float nx_bias = FS_NEG, nx_nq = 0.f;
    {
      const int qcol = jt0 * FS_BN + half * 64 + tid;
      if (tid < 64 && n_tiles > 0 && qcol < Nq) { nx_bias = bias2[qcol]; if (COST) nx_nq = nq[qcol]; }
    }
    for (int idx = 0; idx < n_tiles; ++idx) {
      const int b = 2 * (idx % 2) + sub;
      float* sb = s_bias + (idx & 1) * 256 + grp * 64;
      float* sn = s_nq + (idx & 1) * 256 + grp * 64;
      if (tid < 64) {
        sb[tid] = nx_bias;
        if (COST) sn[tid] = nx_nq;
        const int qcol = (jt0 + idx + 1) * FS_BN + half * 64 + tid;
        nx_bias = FS_NEG; nx_nq = 0.f;
        if (idx + 1 < n_tiles && qcol < Nq) { nx_bias = bias2[qcol]; if (COST) nx_nq = nq[qcol]; }
      }
      named_bar(1 + grp, 128);                           // staging of this tile visible; buffer of tile idx-1 is free
      mbar_wait(&s_full[b], (idx / 2) & 1);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(q4 * 32) << 16) + b * FS_BN + half * 64;
      float v[64];
      tmem_ld32(trow, v);
      tmem_ld32(trow + 32, v + 32);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&s_empty[b]);                          // the accumulator buffer can be refilled already
      if constexpr (!COST) {
        if (fast) {
          const f2 G = pack2(g2, g2);
          f2 acc0 = pack2(0.f, 0.f), acc1 = acc0, acc2 = acc0, acc3 = acc0;
#pragma unroll
          for (int j = 0; j < 64; j += 8) {
            const float4 b0 = *reinterpret_cast<const float4*>(sb + j), b1 = *reinterpret_cast<const float4*>(sb + j + 4);
            const f2 xa = add2(fma2(pack2(v[j], v[j + 1]), G, pack2(b0.x, b0.y)), NU);
            const f2 xb = add2(fma2(pack2(v[j + 2], v[j + 3]), G, pack2(b0.z, b0.w)), NU);
            const f2 xc = add2(fma2(pack2(v[j + 4], v[j + 5]), G, pack2(b1.x, b1.y)), NU);
            const f2 xd = add2(fma2(pack2(v[j + 6], v[j + 7]), G, pack2(b1.z, b1.w)), NU);
            float x0, x1;
            f2 ea, eb, ec, ed;
            unpack2(xa, x0, x1); ea = pack2(ex2(x0), ex2(x1));
            unpack2(xb, x0, x1); eb = pack2(ex2(x0), ex2(x1));
            unpack2(xc, x0, x1); ec = pack2(ex2(x0), ex2(x1));
            unpack2(xd, x0, x1); ed = pack2(ex2(x0), ex2(x1));
            acc0 = add2(acc0, ea); acc1 = add2(acc1, eb); acc2 = add2(acc2, ec); acc3 = add2(acc3, ed);
          }
          float s0, s1;
          unpack2(add2(add2(acc0, acc1), add2(acc2, acc3)), s0, s1);
          l += s0 + s1;
          continue;
        }
        // packed path: t = g2 s + bias (FFMA2), running maximum (FMNMX3), 2^(t - max) on the MUFU pipe for
        // 2^(t - max) on the MUFU pipe, packed accumulation (FADD2).  (A degree-4 polynomial 2^x on the FMA pipe for a
        // fraction of the pairs was measured SLOWER on B200 - the kernel runs at the 1000 W power cap - and was removed.)
        f2 t[32];
        const f2 G = pack2(g2, g2);
        float cm0 = -3.0e38f, cm1 = -3.0e38f, cm2 = -3.0e38f, cm3 = -3.0e38f;
#pragma unroll
        for (int j = 0; j < 64; j += 8) {
          const float4 b0 = *reinterpret_cast<const float4*>(sb + j), b1 = *reinterpret_cast<const float4*>(sb + j + 4);
          t[j / 2] = fma2(pack2(v[j], v[j + 1]), G, pack2(b0.x, b0.y));
          t[j / 2 + 1] = fma2(pack2(v[j + 2], v[j + 3]), G, pack2(b0.z, b0.w));
          t[j / 2 + 2] = fma2(pack2(v[j + 4], v[j + 5]), G, pack2(b1.x, b1.y));
          t[j / 2 + 3] = fma2(pack2(v[j + 6], v[j + 7]), G, pack2(b1.z, b1.w));
          float x0, x1;
          unpack2(t[j / 2], x0, x1); cm0 = max3(cm0, x0, x1);
          unpack2(t[j / 2 + 1], x0, x1); cm1 = max3(cm1, x0, x1);
          unpack2(t[j / 2 + 2], x0, x1); cm2 = max3(cm2, x0, x1);
          unpack2(t[j / 2 + 3], x0, x1); cm3 = max3(cm3, x0, x1);
        }
        const float m_new = fmaxf(max3(m, cm0, cm1), fmaxf(cm2, cm3));
        l *= ex2(m - m_new);
        const f2 NM = pack2(-m_new, -m_new);
        f2 acc0 = pack2(0.f, 0.f), acc1 = acc0;
#pragma unroll
        for (int q = 0; q < 32; ++q) {
          const f2 x = add2(t[q], NM);
          float x0, x1;
          unpack2(x, x0, x1);
          const f2 e = pack2(ex2(x0), ex2(x1));
          if (q & 1) acc1 = add2(acc1, e); else acc0 = add2(acc0, e);
        }
        float s0, s1;
        unpack2(add2(acc0, acc1), s0, s1);
        l += s0 + s1;
        m = m_new;
        continue;
      }
      float cm0 = -3.0e38f, cm1 = -3.0e38f, cm2 = -3.0e38f, cm3 = -3.0e38f;
#pragma unroll
      for (int j = 0; j < 64; j += 4) {
        const float4 bb = *reinterpret_cast<const float4*>(sb + j);
        v[j] = fmaf(v[j], g2, bb.x); v[j + 1] = fmaf(v[j + 1], g2, bb.y);
        v[j + 2] = fmaf(v[j + 2], g2, bb.z); v[j + 3] = fmaf(v[j + 3], g2, bb.w);
        cm0 = fmaxf(cm0, v[j]); cm1 = fmaxf(cm1, v[j + 1]); cm2 = fmaxf(cm2, v[j + 2]); cm3 = fmaxf(cm3, v[j + 3]);
      }
      const float m_new = fmaxf(fmaxf(m, fmaxf(cm0, cm1)), fmaxf(cm2, cm3));
      const float rescale = ex2(m - m_new);
      l *= rescale;
      if (COST) lc *= rescale;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
      for (int j = 0; j < 64; j += 4) {
        const float e0 = ex2(v[j] - m_new), e1 = ex2(v[j + 1] - m_new), e2 = ex2(v[j + 2] - m_new), e3 = ex2(v[j + 3] - m_new);
        a0 += e0; a1 += e1; a2 += e2; a3 += e3;
        if (COST) {
          const float4 nn = *reinterpret_cast<const float4*>(sn + j);
          const float4 bb = *reinterpret_cast<const float4*>(sb + j);
          lc += e0 * fmaf(v[j] - bb.x, k2, nn.x) + e1 * fmaf(v[j + 1] - bb.y, k2, nn.y) +
                e2 * fmaf(v[j + 2] - bb.z, k2, nn.z) + e3 * fmaf(v[j + 3] - bb.w, k2, nn.w);
        }
      }
      l += (a0 + a1) + (a2 + a3);
      m = m_new;
    }
    const int row = pb * FS_BM + sub * FS_SUB + q4 * 32 + lane;
    if (row < Np) {
      const int64_t part = (int64_t)split * 2 + half;
      part_m[part * Np + row] = m;
      part_l[part * Np + row] = l;
      if (COST) part_c[part * Np + row] = lc;
    }
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ---- small kernels ----------------------------------------------------------------------------------------------
// sigma = max |coordinate| over a cloud (atomicMax on the bit pattern of a non-negative float)
__global__ void fs_absmax_kernel(const float* __restrict__ x, int64_t n, float* out) {
  float mx = 0.f;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) mx = fmaxf(mx, fabsf(x[e]));
  mx = warp_max(mx);
  if (threadIdx.x % 32 == 0) atomicMax(reinterpret_cast<unsigned*>(out), __float_as_uint(mx));
}
__global__ void fs_sigma_kernel(float* sig) {   // sig[0] = sigma (>= tiny), sig[1] = sigma^2, sig[2] = 1/sigma
  const float s = fmaxf(sig[0], 1e-30f);
  sig[0] = s; sig[1] = s * s; sig[2] = 1.f / s;
}
// FP16 operand plane x / sigma (the tensor-core operands) and squared norms of the ORIGINAL points (exact norms keep
// the rounding of the cross term zero-mean along both axes, so it averages out of the marginals); one warp per point
__global__ void fs_prep_kernel(const float* __restrict__ x, int64_t n, int64_t d, const float* __restrict__ sig,
                               __half* __restrict__ hi, float* __restrict__ sq) {
  const int lane = threadIdx.x % 32;
  const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / 32;
  if (row >= n) return;
  const float inv = sig[2];
  float acc = 0.f;
  for (int64_t k = lane; k < d; k += 32) {
    const float xv = x[row * d + k];
    hi[row * d + k] = __float2half_rn(xv * inv);
    acc = fmaf(xv, xv, acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) sq[row] = acc;
}

// merge `parts` base-2 partials per row -> potentials.
//   mode 0 (Sinkhorn half-step): pot = logm + nrm - L ; bias2_out = (logm - L) log2e ; diff += |pot - pot_old|
//   mode 1 (partials out, natural log): out_m = m ln2, out_l = l            (row-sharded column step)
//   mode 2 (row max of the cost): out_m = sq + m                            (scale = 1/max)
__global__ void fs_finalize_kernel(const float* __restrict__ pm, const float* __restrict__ pl, int parts, int64_t n, int mode,
                                   const float* __restrict__ marg, const float* __restrict__ sq, float nrm_scale,
                                   float* __restrict__ pot, float* __restrict__ bias2_out, float* __restrict__ out_m,
                                   float* __restrict__ out_l, float* __restrict__ diff, const FsState* state,
                                   float* __restrict__ lse_out = nullptr, float* __restrict__ dmax_out = nullptr) {
  if (state && state->done) return;
  __shared__ float red[32];
  float acc = 0.f, dmx = 0.f;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    float m = pm[r], l = pl[r];
    for (int p = 1; p < parts; ++p) {
      const float m2 = pm[(int64_t)p * n + r], l2 = pl[(int64_t)p * n + r];
      const float mm = fmaxf(m, m2);
      l = l * ex2(m - mm) + l2 * ex2(m2 - mm);
      m = mm;
    }
    if (mode == 0) {
      const float L = (m + log2f(l)) * LN2;
      const float bnat = logf(marg[r] + 1e-8f) - L;
      const float pn = bnat + sq[r] * nrm_scale;
      acc += fabsf(pn - pot[r]);
      dmx = fmaxf(dmx, fabsf(pn - pot[r]));
      pot[r] = pn;
      bias2_out[r] = bnat * LOG2E;
      if (lse_out) lse_out[r] = m + log2f(l);       // base-2 LSE of this row: next iteration's shift
    } else if (mode == 1) {
      out_m[r] = m * LN2;
      out_l[r] = l;
      if (lse_out) lse_out[r] = m + log2f(l);       // partial LSE over the local rows: next call's shift
    } else {
      out_m[r] = sq[r] + m;
    }
  }
  if (mode == 0 && dmax_out) {                     // largest move of a bias of this side, in base-2 units
    dmx = warp_max(dmx);
    if (threadIdx.x % 32 == 0) atomicMax(reinterpret_cast<unsigned*>(dmax_out), __float_as_uint(dmx * LOG2E));
  }
  if (mode == 0 && diff) {
    acc = warp_sum(acc);
    if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = acc;
    __syncthreads();
    if (threadIdx.x == 0) { float t = 0; for (int w = 0; w < blockDim.x / 32; ++w) t += red[w]; atomicAdd(diff, t); }
  }
}

// bias2[r] = (pot[r] - sq[r] * nrm_scale) * log2e   (bias of a side from its potentials)
// dmax_out (optional, zeroed by the caller): max |new bias2 - bias2 of the previous call|, the bound the bounded-shift
// mode of the next pass needs (row-sharded half-steps: the previous bias is still in the workspace)
__global__ void fs_bias_kernel(const float* __restrict__ pot, const float* __restrict__ sq, float nrm_scale, int64_t n,
                               float* __restrict__ bias2, float* __restrict__ dmax_out = nullptr) {
  float dmx = 0.f;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    const float nb = (pot ? pot[r] - sq[r] * nrm_scale : -sq[r] * nrm_scale) * LOG2E;
    if (dmax_out) dmx = fmaxf(dmx, fabsf(nb - bias2[r]));
    bias2[r] = nb;
  }
  if (dmax_out) {
    dmx = warp_max(dmx);
    if (threadIdx.x % 32 == 0) atomicMax(reinterpret_cast<unsigned*>(dmax_out), __float_as_uint(dmx));
  }
}

__global__ void fs_check_kernel(float* diff, double threshold, FsState* state, float* dmax_next_x = nullptr,
                                float* dmax_next_y = nullptr) {
  if (state->done) return;
  if (dmax_next_x) { *dmax_next_x = 0.f; *dmax_next_y = 0.f; }
  const double d = (double)diff[0] + (double)diff[1];
  diff[0] = 0.f; diff[1] = 0.f;
  state->iters += 1;
  if (d < threshold) state->done = 1;
}
// diff[0..1]: sum |delta potential| of the two sides; diff[16..19]: max |delta bias| slots, [side][iteration parity]
__global__ void fs_init_state_kernel(FsState* st, float* diff) {
  st->done = 0; st->iters = 0; diff[0] = 0.f; diff[1] = 0.f;
  diff[16] = 0.f; diff[17] = 0.f; diff[18] = 0.f; diff[19] = 0.f;
}

__global__ void fs_max_reduce_kernel(const float* __restrict__ rowmax, int64_t n, float* out) {
  float mx = 0.f;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) mx = fmaxf(mx, rowmax[r]);
  mx = warp_max(mx);
  if (threadIdx.x % 32 == 0) atomicMax(reinterpret_cast<unsigned*>(out), __float_as_uint(fmaxf(mx, 0.f)));
}

// summary[0] += sum_r rowsum_r * cost part, [1] += mass, [2] = max |rowsum - marg|  (row side, with cost partials)
// or (cost partials null) only [3] = max |colsum - marg|
__global__ void fs_summary_kernel(const float* __restrict__ pm, const float* __restrict__ pl, const float* __restrict__ pc,
                                  int parts, int64_t n, const float* __restrict__ pot, const float* __restrict__ sq,
                                  float nrm_scale, float scale, const float* __restrict__ marg, double* summary, int slot,
                                  float* __restrict__ marg_out) {
  double cost = 0, mass = 0, err = 0;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    float m = pm[r], l = pl[r], c = pc ? pc[r] : 0.f;
    for (int p = 1; p < parts; ++p) {
      const float m2 = pm[(int64_t)p * n + r], l2 = pl[(int64_t)p * n + r];
      const float mm = fmaxf(m, m2), w0 = ex2(m - mm), w1 = ex2(m2 - mm);
      l = l * w0 + l2 * w1;
      if (pc) c = c * w0 + pc[(int64_t)p * n + r] * w1;
      m = mm;
    }
    // row sum of the plan: exp(pot - nrm) * 2^m * l
    const double pref = exp((double)pot[r] - (double)sq[r] * nrm_scale + (double)m * (double)LN2);
    const double rs = pref * l;
    if (marg_out) marg_out[r] = (float)rs;
    mass += rs;
    err = fmax(err, fabs(rs - (double)marg[r]));
    if (pc) cost += pref * (double)scale * ((double)sq[r] * l + (double)c);
  }
  cost = warp_sum(cost); mass = warp_sum(mass); err = warp_max(err);
  if (threadIdx.x % 32 == 0) {
    if (pc) { atomicAdd(&summary[0], cost); atomicAdd(&summary[1], mass); }
    atomicMax(reinterpret_cast<unsigned long long*>(&summary[slot]), (unsigned long long)__double_as_longlong(err));
  }
}

// ---- host side --------------------------------------------------------------------------------------------------
bool sk_umma_eligible(int64_t N, int64_t M, int64_t dim, int cost_kind) {
  return cost_kind == OTK_COST_SQEUCLIDEAN && dim >= 8 && dim <= 128 && dim % 8 == 0 && N >= 1 && M >= 1 &&
         N < (1ll << 30) && M < (1ll << 30) && tensormap_encoder() != nullptr;
}

static unsigned fs_grid(int64_t n) {
  int64_t b = ceil_div(n, 256), cap = (int64_t)sm_count() * 4;
  return (unsigned)(b < cap ? (b ? b : 1) : cap);
}

// Column splits of a pass.  The grid is (P blocks) x (splits) CTAs, each sweeping ceil(q tiles / splits) tiles after a fixed
// prologue (TMEM allocation, P-block load, first Q stage: about FS_CTA_OVERHEAD_TILES tiles' worth); CTAs are dealt to the
// SMs in waves, so the pass takes   waves(splits) x (tiles per CTA + overhead)   tile times.  Minimise that over the split
// count: the old rule (~6 waves) left a row shard of 64 P blocks with 14 splits = 6.05 waves, i.e. a seventh, almost empty
// wave - 27 % above the work bound; 9 splits (3.9 waves, longer CTAs) are within 9 %.
constexpr int FS_CTA_OVERHEAD_TILES = 3;
static int fs_splits(int64_t p_rows, int64_t q_rows) {
  const int64_t pblocks = ceil_div(p_rows, FS_BM), qtiles = ceil_div(q_rows, FS_BN), sms = sm_count();
  int best = 1;
  int64_t best_cost = INT64_MAX;
  for (int64_t s = 1; s <= 32 && s <= qtiles; ++s) {
    const int64_t waves = ceil_div(pblocks * s, sms);
    const int64_t cost = waves * (ceil_div(qtiles, s) + FS_CTA_OVERHEAD_TILES);
    if (cost < best_cost) { best_cost = cost; best = (int)s; }
  }
  return best;
}

struct FsSide {           // one point cloud, prepared
  const float* raw; __half* hi; float* sq; int64_t n;
  CUtensorMap map;
};

struct FsWork {
  FsSide X, Y;
  float *biasX2, *biasY2, *pm, *pl, *pc, *diff, *scratch_n, *sig, *lseX, *lseY;
  FsState* state;
  int max_parts;
};

size_t sk_umma_workspace_bytes(int64_t N, int64_t M, int64_t dim) {
  const int64_t mx = N > M ? N : M;
  return align_up((size_t)N * dim * 2, 256) + align_up((size_t)M * dim * 2, 256) + 8 * align_up((size_t)mx * 4, 256) +
         3 * align_up((size_t)64 * mx * 4, 256) + 8192;
}

static int fs_carve(FsWork& w, const float* x, const float* y, int64_t N, int64_t M, int64_t dim, void* workspace,
                    size_t workspace_bytes, cudaStream_t st, bool prepare = true) {
  Arena ar(workspace, workspace_bytes);
  const int64_t mx = N > M ? N : M;
  w.X = FsSide{x, ar.take<__half>((size_t)N * dim), ar.take<float>((size_t)N), N, {}};
  w.Y = FsSide{y, ar.take<__half>((size_t)M * dim), ar.take<float>((size_t)M), M, {}};
  w.biasX2 = ar.take<float>((size_t)N);
  w.biasY2 = ar.take<float>((size_t)M);
  w.scratch_n = ar.take<float>((size_t)mx);
  w.lseX = ar.take<float>((size_t)N);
  w.lseY = ar.take<float>((size_t)M);
  w.max_parts = 64;
  w.pm = ar.take<float>((size_t)64 * mx);
  w.pl = ar.take<float>((size_t)64 * mx);
  w.pc = ar.take<float>((size_t)64 * mx);
  w.diff = ar.take<float>(64);
  w.sig = ar.take<float>(64);
  w.state = ar.take<FsState>(1);
  if (!ar.ok()) return OTK_ERR_WORKSPACE;
  if (prepare) {       // otherwise the planes / norms / sigma of an earlier call on the same operands are still in place
    OTK_CUDA(cudaMemsetAsync(w.sig, 0, 16, st));
    fs_absmax_kernel<<<fs_grid(N * dim), 256, 0, st>>>(x, N * dim, w.sig);
    fs_absmax_kernel<<<fs_grid(M * dim), 256, 0, st>>>(y, M * dim, w.sig);
    fs_sigma_kernel<<<1, 1, 0, st>>>(w.sig);
    fs_prep_kernel<<<(unsigned)ceil_div(N * 32, 256), 256, 0, st>>>(x, N, dim, w.sig, w.X.hi, w.X.sq);
    fs_prep_kernel<<<(unsigned)ceil_div(M * 32, 256), 256, 0, st>>>(y, M, dim, w.sig, w.Y.hi, w.Y.sq);
    count_launch(4);
    OTK_LAUNCH_CHECK();
  }
  // 2-D fp16 maps [rows, dim], box 64 x 128, 128B swizzle (K-major operands)
  if (!encode_map_f16_2d(&w.X.map, w.X.hi, dim, N, dim, FS_SLAB, FS_SUB)) return OTK_ERR_CUDA;
  if (!encode_map_f16_2d(&w.Y.map, w.Y.hi, dim, M, dim, FS_SLAB, FS_SUB)) return OTK_ERR_CUDA;
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    OTK_CUDA(cudaFuncSetAttribute(fused_lse_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FS_SMEM));
    OTK_CUDA(cudaFuncSetAttribute(fused_lse_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FS_SMEM));
    attr_set[dev] = true;
  }
  return OTK_OK;
}

int g_fs_fast = 1;    // tuning aid: 0 disables the bounded-shift mode
// one pass: partials of LSE_q(bias2_q + g2 * p.q) for every row of P; returns the number of parts written
template <bool COST>
static int fs_pass(const FsSide& P, const FsSide& Q, const float* biasQ2, float g2, const float* sig2, int64_t dim, float* pm,
                   float* pl, float* pc, const FsState* state, int* parts_out, cudaStream_t st,
                   const float* lse_prev = nullptr, const float* dmax = nullptr) {
  const int splits = fs_splits(P.n, Q.n);
  const int64_t qtiles = ceil_div(Q.n, FS_BN);
  const int tps = (int)ceil_div(qtiles, splits);
  const int parts = (int)ceil_div(qtiles, tps);
  dim3 grid((unsigned)ceil_div(P.n, FS_BM), (unsigned)parts);
  fused_lse_kernel<COST><<<grid, FS_THREADS, FS_SMEM, st>>>(P.map, Q.map, biasQ2, Q.sq, g2, sig2, (int)P.n, (int)Q.n, (int)dim, tps, pm, pl, pc,
                                          state, g_fs_fast ? lse_prev : nullptr, dmax);
  OTK_LAUNCH_CHECK();
  *parts_out = 2 * parts;   // two column halves per split
  return OTK_OK;
}

// max_ij |x_i - y_j|^2 on the device -> *out_dev (one fused pass with g2 = -2, bias = |y|^2: the running max is the answer)
static int fs_cost_max(FsWork& w, int64_t dim, float* out_dev, cudaStream_t st) {
  int parts = 0;
  fs_bias_kernel<<<fs_grid(w.Y.n), 256, 0, st>>>(nullptr, w.Y.sq, -1.f / LOG2E, w.Y.n, w.biasY2);  // bias2 = +|y|^2
  // the LSE pass maximises bias2 + g2 * s ; we need max(|y|^2 - 2 x.y): g2 = -2
  OTK_TRY(fs_pass<false>(w.X, w.Y, w.biasY2, -2.f, w.sig + 1, dim, w.pm, w.pl, w.pc, nullptr, &parts, st));
  fs_finalize_kernel<<<fs_grid(w.X.n), 256, 0, st>>>(w.pm, w.pl, parts, w.X.n, 2, nullptr, w.X.sq, 0.f, nullptr, nullptr,
                                                    w.scratch_n, nullptr, nullptr, nullptr);
  OTK_CUDA(cudaMemsetAsync(out_dev, 0, 4, st));
  fs_max_reduce_kernel<<<fs_grid(w.X.n), 256, 0, st>>>(w.scratch_n, w.X.n, out_dev);
  count_launch(2);
  OTK_LAUNCH_CHECK();
  return OTK_OK;
}

int sk_umma_cost_max(const float* x, const float* y, int64_t N, int64_t M, int64_t dim, float* out, void* workspace,
                     size_t workspace_bytes, cudaStream_t st) {
  FsWork w;
  OTK_TRY(fs_carve(w, x, y, N, M, dim, workspace, workspace_bytes, st));
  return fs_cost_max(w, dim, out, st);
}

int sk_umma_solve(const float* x, const float* y, int64_t N, int64_t M, int64_t dim, const float* a, const float* b,
                  double scale, int scale_inv_max, double reg, int max_iter, double threshold, int poll_every, int precision,
                  float* u, float* v, double* summary, float* row_marginal, float* col_marginal, int* iters_done_host,
                  void* workspace, size_t workspace_bytes, cudaStream_t st) {
  (void)precision;
  FsWork w;
  OTK_TRY(fs_carve(w, x, y, N, M, dim, workspace, workspace_bytes, st));
  if (scale_inv_max) {
    float host_max = 0.f;
    OTK_TRY(fs_cost_max(w, dim, w.diff + 8, st));
    OTK_CUDA(cudaMemcpyAsync(&host_max, w.diff + 8, 4, cudaMemcpyDeviceToHost, st));
    OTK_CUDA(cudaStreamSynchronize(st));
    OTK_REQUIRE(host_max > 0.f, "sinkhorn_points: degenerate cost (max = 0)");
    scale = 1.0 / (double)host_max;
  }
  const float nrm_scale = (float)(scale / reg);                    // nrm_r = |p_r|^2 * scale / reg
  const float g2 = (float)(2.0 * scale / reg) * LOG2E;
  fs_init_state_kernel<<<1, 1, 0, st>>>(w.state, w.diff);
  OTK_CUDA(cudaMemsetAsync(u, 0, (size_t)N * 4, st));
  OTK_CUDA(cudaMemsetAsync(v, 0, (size_t)M * 4, st));
  fs_bias_kernel<<<fs_grid(N), 256, 0, st>>>(u, w.X.sq, nrm_scale, N, w.biasX2);
  count_launch(1);
  OTK_LAUNCH_CHECK();
  if (poll_every <= 0) poll_every = 16;
  FsState host_state{0, 0};
  int parts = 0;
  for (int it = 0; it < max_iter; ++it) {
    // v first: rows = Y, reduce over X (bias from u) ; then u: rows = X, reduce over Y (bias from the new v)
    // max |delta bias| slots: dX = diff[16 + parity], dY = diff[18 + parity]; the first iteration has no previous LSE
    float* dX_prev = w.diff + 16 + ((it + 1) & 1);
    float* dX_cur = w.diff + 16 + (it & 1);
    float* dY_cur = w.diff + 18 + (it & 1);
    float* dY_next = w.diff + 18 + ((it + 1) & 1);
    OTK_TRY(fs_pass<false>(w.Y, w.X, w.biasX2, g2, w.sig + 1, dim, w.pm, w.pl, w.pc, w.state, &parts, st,
                           it > 0 ? w.lseY : nullptr, dX_prev));
    fs_finalize_kernel<<<fs_grid(M), 256, 0, st>>>(w.pm, w.pl, parts, M, 0, b, w.Y.sq, nrm_scale, v, w.biasY2, nullptr, nullptr,
                                                  w.diff + 1, w.state, w.lseY, dY_cur);
    OTK_TRY(fs_pass<false>(w.X, w.Y, w.biasY2, g2, w.sig + 1, dim, w.pm, w.pl, w.pc, w.state, &parts, st,
                           it > 0 ? w.lseX : nullptr, dY_cur));
    fs_finalize_kernel<<<fs_grid(N), 256, 0, st>>>(w.pm, w.pl, parts, N, 0, a, w.X.sq, nrm_scale, u, w.biasX2, nullptr, nullptr,
                                                  w.diff, w.state, w.lseX, dX_cur);
    fs_check_kernel<<<1, 1, 0, st>>>(w.diff, threshold, w.state, dX_prev, dY_next);
    count_launch(2);
    OTK_LAUNCH_CHECK();
    if (threshold > 0 && (it + 1) % poll_every == 0 && it + 1 < max_iter) {
      OTK_CUDA(cudaMemcpyAsync(&host_state, w.state, sizeof(FsState), cudaMemcpyDeviceToHost, st));
      OTK_CUDA(cudaStreamSynchronize(st));
      if (host_state.done) break;
    }
  }
  if (summary) {
    OTK_CUDA(cudaMemsetAsync(summary, 0, 4 * sizeof(double), st));
    OTK_TRY(fs_pass<true>(w.X, w.Y, w.biasY2, g2, w.sig + 1, dim, w.pm, w.pl, w.pc, nullptr, &parts, st));
    fs_summary_kernel<<<fs_grid(N), 256, 0, st>>>(w.pm, w.pl, w.pc, parts, N, u, w.X.sq, nrm_scale, (float)scale, a, summary, 2, row_marginal);
    OTK_TRY(fs_pass<false>(w.Y, w.X, w.biasX2, g2, w.sig + 1, dim, w.pm, w.pl, w.pc, nullptr, &parts, st));
    fs_summary_kernel<<<fs_grid(M), 256, 0, st>>>(w.pm, w.pl, nullptr, parts, M, v, w.Y.sq, nrm_scale, (float)scale, b, summary, 3, col_marginal);
    count_launch(1);
    OTK_LAUNCH_CHECK();
  }
  if (iters_done_host) {
    OTK_CUDA(cudaMemcpyAsync(&host_state, w.state, sizeof(FsState), cudaMemcpyDeviceToHost, st));
    OTK_CUDA(cudaStreamSynchronize(st));
    *iters_done_host = host_state.iters;
  }
  return OTK_OK;
}

__global__ void fs_fold_kernel(float* m, const float* sq, float nrm_scale, int64_t n) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x)
    m[r] -= sq[r] * nrm_scale;
}

// row-sharded half-steps; the prepared operands live in the caller's workspace and are reused when the caller says so
// reuse_prepared: 0 = prepare the operands; 1 = operands in place; 2 = operands in place AND the state of the previous
// call on this workspace (biases, partial LSEs) is that of the previous Sinkhorn iteration: bounded-shift mode - the partial
// LSE over the local rows moved by at most max |delta bias| of the local rows, so one sweep per tile is enough
int sk_umma_colstep(const float* x_local, const float* y, int64_t n_local, int64_t M, int64_t dim, const float* u_local,
                    double scale, double reg, int reuse_prepared, float* col_max, float* col_sum, void* workspace,
                    size_t workspace_bytes, cudaStream_t st) {
  FsWork w;
  OTK_TRY(fs_carve(w, x_local, y, n_local, M, dim, workspace, workspace_bytes, st, !reuse_prepared));
  const float nrm_scale = (float)(scale / reg), g2 = (float)(2.0 * scale / reg) * LOG2E;
  const bool bounded = reuse_prepared >= 2;
  float* dmax = w.diff + 16;
  if (bounded) OTK_CUDA(cudaMemsetAsync(dmax, 0, 4, st));
  fs_bias_kernel<<<fs_grid(n_local), 256, 0, st>>>(u_local, w.X.sq, nrm_scale, n_local, w.biasX2, bounded ? dmax : nullptr);
  int parts = 0;
  OTK_TRY(fs_pass<false>(w.Y, w.X, w.biasX2, g2, w.sig + 1, dim, w.pm, w.pl, w.pc, nullptr, &parts, st,
                         bounded ? w.lseY : nullptr, dmax));
  // partial over the LOCAL rows of LSE_i(u_i + Cr_ij) = -nrm_j + LSE_i(bias_i + gamma x_i.y_j), natural log
  fs_finalize_kernel<<<fs_grid(M), 256, 0, st>>>(w.pm, w.pl, parts, M, 1, nullptr, nullptr, 0.f, nullptr, nullptr, col_max, col_sum,
                                                nullptr, nullptr, w.lseY);
  fs_fold_kernel<<<fs_grid(M), 256, 0, st>>>(col_max, w.Y.sq, nrm_scale, M);
  count_launch(2);
  OTK_LAUNCH_CHECK();
  return OTK_OK;
}

__global__ void fs_rowstep_finish_kernel(const float* __restrict__ pm, const float* __restrict__ pl, int parts, int64_t n,
                                         const float* __restrict__ marg, const float* __restrict__ sq, float nrm_scale,
                                         float* __restrict__ pot, float* diff, float* __restrict__ lse_out) {
  __shared__ float red[32];
  float acc = 0.f;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    float m = pm[r], l = pl[r];
    for (int p = 1; p < parts; ++p) {
      const float m2 = pm[(int64_t)p * n + r], l2 = pl[(int64_t)p * n + r];
      const float mm = fmaxf(m, m2);
      l = l * ex2(m - mm) + l2 * ex2(m2 - mm);
      m = mm;
    }
    const float lse2 = m + log2f(l);
    const float pn = logf(marg[r] + 1e-8f) - lse2 * LN2 + sq[r] * nrm_scale;
    acc += fabsf(pn - pot[r]);
    pot[r] = pn;
    lse_out[r] = lse2;
  }
  acc = warp_sum(acc);
  if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = acc;
  __syncthreads();
  if (threadIdx.x == 0 && diff) { float t = 0; for (int w = 0; w < blockDim.x / 32; ++w) t += red[w]; atomicAdd(diff, t); }
}

int sk_umma_rowstep(const float* x_local, const float* y, int64_t n_local, int64_t M, int64_t dim, const float* a_local,
                    const float* v, double scale, double reg, int reuse_prepared, float* u_local, float* diff, void* workspace,
                    size_t workspace_bytes, cudaStream_t st) {
  FsWork w;
  OTK_TRY(fs_carve(w, x_local, y, n_local, M, dim, workspace, workspace_bytes, st, !reuse_prepared));
  const float nrm_scale = (float)(scale / reg), g2 = (float)(2.0 * scale / reg) * LOG2E;
  const bool bounded = reuse_prepared >= 2;          // see sk_umma_colstep
  float* dmax = w.diff + 18;
  if (bounded) OTK_CUDA(cudaMemsetAsync(dmax, 0, 4, st));
  fs_bias_kernel<<<fs_grid(M), 256, 0, st>>>(v, w.Y.sq, nrm_scale, M, w.biasY2, bounded ? dmax : nullptr);
  int parts = 0;
  OTK_TRY(fs_pass<false>(w.X, w.Y, w.biasY2, g2, w.sig + 1, dim, w.pm, w.pl, w.pc, nullptr, &parts, st,
                         bounded ? w.lseX : nullptr, dmax));
  fs_rowstep_finish_kernel<<<fs_grid(n_local), 256, 0, st>>>(w.pm, w.pl, parts, n_local, a_local, w.X.sq, nrm_scale, u_local, diff,
                                                            w.lseX);
  count_launch(1);
  OTK_LAUNCH_CHECK();
  return OTK_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Peer-memory exchange of the column partials (row-sharded multi-GPU path).  Every rank owns one exchange buffer in
// symmetric memory, mapped into all peers (torch.distributed._symmetric_memory on the host side):
//     float    data[2][world][2][M]     slot (iteration parity), source rank, {max (natural log), sum}, column
//     uint32_t flag[2][world]           flag[slot][src] = iteration + 1 once src's partials of that iteration have landed
// The column half-step finishes with ONE kernel that merges the split partials of the pass, folds the column norms in and
// stores the result straight into every peer's buffer over NVLink (coalesced 4-byte stores, 2 x M x 4 bytes per peer),
// then the last CTA releases the flags; the combine kernel acquires the flags of all ranks and reduces the partials it
// finds in its own memory.  No collective call, no host involvement, and the iteration stays capturable in a CUDA graph.
// Double buffering by iteration parity is enough: a rank can only be one iteration ahead of the slowest peer, because its
// combine of iteration k needs every peer's push of iteration k, which follows that peer's combine of iteration k - 1.
// ---------------------------------------------------------------------------------------------------------------------
struct XchgView { float* data; uint32_t* flag; };
__host__ __device__ inline size_t xchg_data_floats(int world, int64_t M) { return (size_t)2 * world * 2 * M; }
__device__ __forceinline__ XchgView xchg_view(void* base, int world, int64_t M) {
  float* d = static_cast<float*>(base);
  return XchgView{d, reinterpret_cast<uint32_t*>(d + xchg_data_floats(world, M))};
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ctrl (local device ints): [0] iteration counter (advanced by the combine kernel), [1] ticket of the push kernel,
// [2] set if a wait timed out, [3] ticket of the combine kernel
__global__ void fs_finalize_push_kernel(const float* __restrict__ pm, const float* __restrict__ pl, int parts, int64_t M,
                                        const float* __restrict__ sq, float nrm_scale, float* __restrict__ lse_out,
                                        void* const* __restrict__ peers, int world, int rank, int* ctrl) {
  const uint32_t it = (uint32_t)ctrl[0];
  const int slot = it & 1;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < M; r += (int64_t)gridDim.x * blockDim.x) {
    float m = pm[r], l = pl[r];
    for (int p = 1; p < parts; ++p) {
      const float m2 = pm[(int64_t)p * M + r], l2 = pl[(int64_t)p * M + r];
      const float mm = fmaxf(m, m2);
      l = l * ex2(m - mm) + l2 * ex2(m2 - mm);
      m = mm;
    }
    lse_out[r] = m + log2f(l);                                  // partial LSE over the local rows: next call's shift
    const float mn = m * LN2 - sq[r] * nrm_scale;               // natural log, column norm folded in
    for (int p = 0; p < world; ++p) {
      float* d = xchg_view(peers[p], world, M).data + ((size_t)(slot * world + rank) * 2) * M;
      d[r] = mn;
      d[M + r] = l;
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int t = atomicAdd(&ctrl[1], 1);
    if (t == (int)gridDim.x - 1) {                              // every CTA's stores are fenced: publish
      ctrl[1] = 0;
      __threadfence_system();
      for (int p = 0; p < world; ++p) st_release_sys(xchg_view(peers[p], world, M).flag + slot * world + rank, it + 1);
    }
  }
}

__global__ void fs_combine_wait_kernel(void* xchg, int world, int64_t M, const float* __restrict__ b, float* __restrict__ v,
                                       float* diff, int* ctrl) {
  __shared__ float red[32];
  const uint32_t it = (uint32_t)ctrl[0];
  const int slot = it & 1;
  const XchgView x = xchg_view(xchg, world, M);
  if ((int)threadIdx.x < world) {
    const uint32_t* f = x.flag + slot * world + threadIdx.x;
    const long long t0 = clock64();
    while (ld_acquire_sys(f) < it + 1) {
      if (clock64() - t0 > (1ll << 32)) { ctrl[2] = 1; break; }     // ~2 s: a peer is gone - flag the error, do not hang
    }
  }
  __syncthreads();
  const float* d = x.data + (size_t)slot * world * 2 * M;
  float acc = 0.f;
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < M; j += (int64_t)gridDim.x * blockDim.x) {
    float mm = d[j], ss = d[M + j];
    for (int p = 1; p < world; ++p) {
      const float m2 = d[(size_t)p * 2 * M + j], s2 = d[(size_t)p * 2 * M + M + j];
      if (m2 > mm) { ss = ss * __expf(mm - m2) + s2; mm = m2; } else ss += s2 * __expf(m2 - mm);
    }
    const float vn = logf(b[j] + 1e-8f) - (mm + logf(ss));
    acc += fabsf(vn - v[j]);
    v[j] = vn;
  }
  acc = warp_sum(acc);
  if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0;
    for (int w = 0; w < (int)blockDim.x / 32; ++w) t += red[w];
    if (diff) atomicAdd(diff, t);
    if (atomicAdd(&ctrl[3], 1) == (int)gridDim.x - 1) { ctrl[3] = 0; ctrl[0] = (int)(it + 1); }   // all CTAs have read `it`
  }
}

size_t sk_umma_exchange_bytes(int world, int64_t M) { return xchg_data_floats(world, M) * 4 + align_up((size_t)2 * world * 4, 256); }

int sk_umma_colstep_push(const float* x_local, const float* y, int64_t n_local, int64_t M, int64_t dim, const float* u_local,
                         double scale, double reg, int reuse_prepared, void* const* peers_dev, int world, int rank, int* ctrl,
                         void* workspace, size_t workspace_bytes, cudaStream_t st) {
  FsWork w;
  OTK_TRY(fs_carve(w, x_local, y, n_local, M, dim, workspace, workspace_bytes, st, !reuse_prepared));
  const float nrm_scale = (float)(scale / reg), g2 = (float)(2.0 * scale / reg) * LOG2E;
  const bool bounded = reuse_prepared >= 2;
  float* dmax = w.diff + 16;
  if (bounded) OTK_CUDA(cudaMemsetAsync(dmax, 0, 4, st));
  fs_bias_kernel<<<fs_grid(n_local), 256, 0, st>>>(u_local, w.X.sq, nrm_scale, n_local, w.biasX2, bounded ? dmax : nullptr);
  int parts = 0;
  OTK_TRY(fs_pass<false>(w.Y, w.X, w.biasX2, g2, w.sig + 1, dim, w.pm, w.pl, w.pc, nullptr, &parts, st,
                         bounded ? w.lseY : nullptr, dmax));
  fs_finalize_push_kernel<<<fs_grid(M), 256, 0, st>>>(w.pm, w.pl, parts, M, w.Y.sq, nrm_scale, w.lseY, peers_dev, world, rank, ctrl);
  count_launch(1);
  OTK_LAUNCH_CHECK();
  return OTK_OK;
}

int sk_combine_wait(void* xchg, int world, int64_t M, const float* b, float* v, float* diff, int* ctrl, cudaStream_t st) {
  fs_combine_wait_kernel<<<fs_grid(M), 256, 0, st>>>(xchg, world, M, b, v, diff, ctrl);
  OTK_LAUNCH_CHECK();
  return OTK_OK;
}

// ---- one whole row-sharded iteration in five launches (peer-memory exchange inside) ------------------------------------
// Every kernel that finishes a half-step also produces what the NEXT pass needs - its operand bias, the bound max |delta
// bias| of the bounded-shift mode, the partial log-sum-exps - so the iteration is
//     pass (columns x local rows) -> finalize + push -> wait + combine -> pass (local rows x columns) -> finish
// with no memset / bias / fold launches in between.  Slots: diff[16] = max |delta bias X| (written by `finish`, read by the
// column pass, cleared by `combine`), diff[18] = max |delta bias Y| (written by `combine`, read by the row pass, cleared by
// `finish`); `diffs` = {sum |du| (local rows), sum |dv|}, cleared by `push`.
__global__ void fs_push_step_kernel(const float* __restrict__ pm, const float* __restrict__ pl, int parts, int64_t M,
                                    const float* __restrict__ sq, float nrm_scale, float* __restrict__ lse_out,
                                    void* const* __restrict__ peers, int world, int rank, int* ctrl, float* diffs) {
  if (blockIdx.x == 0 && threadIdx.x == 0 && diffs) { diffs[0] = 0.f; diffs[1] = 0.f; }
  const uint32_t it = (uint32_t)ctrl[0];
  const int slot = it & 1;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < M; r += (int64_t)gridDim.x * blockDim.x) {
    float m = pm[r], l = pl[r];
    for (int p = 1; p < parts; ++p) {
      const float m2 = pm[(int64_t)p * M + r], l2 = pl[(int64_t)p * M + r];
      const float mm = fmaxf(m, m2);
      l = l * ex2(m - mm) + l2 * ex2(m2 - mm);
      m = mm;
    }
    lse_out[r] = m + log2f(l);
    const float mn = m * LN2 - sq[r] * nrm_scale;
    for (int p = 0; p < world; ++p) {
      float* d = xchg_view(peers[p], world, M).data + ((size_t)(slot * world + rank) * 2) * M;
      d[r] = mn;
      d[M + r] = l;
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int t = atomicAdd(&ctrl[1], 1);
    if (t == (int)gridDim.x - 1) {
      ctrl[1] = 0;
      __threadfence_system();
      for (int p = 0; p < world; ++p) st_release_sys(xchg_view(peers[p], world, M).flag + slot * world + rank, it + 1);
    }
  }
}

__global__ void fs_combine_step_kernel(void* xchg, int world, int64_t M, const float* __restrict__ b, float* __restrict__ v,
                                       float* diffs, int* ctrl, const float* __restrict__ sq, float nrm_scale,
                                       float* __restrict__ bias2, float* dmax_y, float* dmax_x_clear) {
  __shared__ float red[32];
  const uint32_t it = (uint32_t)ctrl[0];
  const int slot = it & 1;
  const XchgView x = xchg_view(xchg, world, M);
  if (blockIdx.x == 0 && threadIdx.x == 32) *dmax_x_clear = 0.f;       // the column pass has consumed it
  if ((int)threadIdx.x < world) {
    const uint32_t* f = x.flag + slot * world + threadIdx.x;
    const long long t0 = clock64();
    while (ld_acquire_sys(f) < it + 1) {
      if (clock64() - t0 > (1ll << 32)) { ctrl[2] = 1; break; }
    }
  }
  __syncthreads();
  const float* d = x.data + (size_t)slot * world * 2 * M;
  float acc = 0.f, dmx = 0.f;
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < M; j += (int64_t)gridDim.x * blockDim.x) {
    float mm = d[j], ss = d[M + j];
    for (int p = 1; p < world; ++p) {
      const float m2 = d[(size_t)p * 2 * M + j], s2 = d[(size_t)p * 2 * M + M + j];
      if (m2 > mm) { ss = ss * __expf(mm - m2) + s2; mm = m2; } else ss += s2 * __expf(m2 - mm);
    }
    const float vn = logf(b[j] + 1e-8f) - (mm + logf(ss));
    acc += fabsf(vn - v[j]);
    v[j] = vn;
    const float nb = (vn - sq[j] * nrm_scale) * LOG2E;                  // operand bias of the row pass
    dmx = fmaxf(dmx, fabsf(nb - bias2[j]));
    bias2[j] = nb;
  }
  dmx = warp_max(dmx);
  if (threadIdx.x % 32 == 0 && dmx == dmx) atomicMax(reinterpret_cast<unsigned*>(dmax_y), __float_as_uint(dmx));
  acc = warp_sum(acc);
  if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0;
    for (int w = 0; w < (int)blockDim.x / 32; ++w) t += red[w];
    if (diffs) atomicAdd(&diffs[1], t);
    if (atomicAdd(&ctrl[3], 1) == (int)gridDim.x - 1) { ctrl[3] = 0; ctrl[0] = (int)(it + 1); }
  }
}

__global__ void fs_finish_step_kernel(const float* __restrict__ pm, const float* __restrict__ pl, int parts, int64_t n,
                                      const float* __restrict__ marg, const float* __restrict__ sq, float nrm_scale,
                                      float* __restrict__ pot, float* diffs, float* __restrict__ lse_out,
                                      float* __restrict__ bias2, float* dmax_x, float* dmax_y_clear) {
  __shared__ float red[32];
  if (blockIdx.x == 0 && threadIdx.x == 32) *dmax_y_clear = 0.f;       // the row pass has consumed it
  float acc = 0.f, dmx = 0.f;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    float m = pm[r], l = pl[r];
    for (int p = 1; p < parts; ++p) {
      const float m2 = pm[(int64_t)p * n + r], l2 = pl[(int64_t)p * n + r];
      const float mm = fmaxf(m, m2);
      l = l * ex2(m - mm) + l2 * ex2(m2 - mm);
      m = mm;
    }
    const float lse2 = m + log2f(l);
    const float bnat = logf(marg[r] + 1e-8f) - lse2 * LN2;
    const float pn = bnat + sq[r] * nrm_scale;
    acc += fabsf(pn - pot[r]);
    pot[r] = pn;
    lse_out[r] = lse2;
    const float nb = bnat * LOG2E;                                      // operand bias of the next column pass
    dmx = fmaxf(dmx, fabsf(nb - bias2[r]));
    bias2[r] = nb;
  }
  dmx = warp_max(dmx);
  if (threadIdx.x % 32 == 0) atomicMax(reinterpret_cast<unsigned*>(dmax_x), __float_as_uint(dmx));
  acc = warp_sum(acc);
  if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = acc;
  __syncthreads();
  if (threadIdx.x == 0 && diffs) { float t = 0; for (int w = 0; w < (int)blockDim.x / 32; ++w) t += red[w]; atomicAdd(&diffs[0], t); }
}

// stage 0: first iteration of a solve (prepares the operands, bias of the rows from u_local); stage >= 1: steady state
int sk_umma_sharded_step(const float* x_local, const float* y, int64_t n_local, int64_t M, int64_t dim, const float* a_local,
                         const float* b, float* u_local, float* v, double scale, double reg, int stage, void* const* peers_dev,
                         int world, int rank, void* xchg_local, int* ctrl, float* diffs, void* workspace, size_t workspace_bytes,
                         cudaStream_t st) {
  FsWork w;
  OTK_TRY(fs_carve(w, x_local, y, n_local, M, dim, workspace, workspace_bytes, st, stage == 0));
  const float nrm_scale = (float)(scale / reg), g2 = (float)(2.0 * scale / reg) * LOG2E;
  float *dmax_x = w.diff + 16, *dmax_y = w.diff + 18;
  if (stage == 0) {
    OTK_CUDA(cudaMemsetAsync(w.diff, 0, 64 * sizeof(float), st));
    OTK_CUDA(cudaMemsetAsync(w.biasY2, 0, (size_t)M * 4, st));
    fs_bias_kernel<<<fs_grid(n_local), 256, 0, st>>>(u_local, w.X.sq, nrm_scale, n_local, w.biasX2);
    OTK_LAUNCH_CHECK();
  }
  const bool bounded = stage >= 1;
  int parts = 0;
  OTK_TRY(fs_pass<false>(w.Y, w.X, w.biasX2, g2, w.sig + 1, dim, w.pm, w.pl, w.pc, nullptr, &parts, st,
                         bounded ? w.lseY : nullptr, dmax_x));
  fs_push_step_kernel<<<fs_grid(M), 256, 0, st>>>(w.pm, w.pl, parts, M, w.Y.sq, nrm_scale, w.lseY, peers_dev, world, rank, ctrl, diffs);
  fs_combine_step_kernel<<<fs_grid(M), 256, 0, st>>>(xchg_local, world, M, b, v, diffs, ctrl, w.Y.sq, nrm_scale, w.biasY2, dmax_y,
                                                    dmax_x);
  count_launch(1);
  OTK_LAUNCH_CHECK();
  OTK_TRY(fs_pass<false>(w.X, w.Y, w.biasY2, g2, w.sig + 1, dim, w.pm, w.pl, w.pc, nullptr, &parts, st,
                         bounded ? w.lseX : nullptr, dmax_y));
  fs_finish_step_kernel<<<fs_grid(n_local), 256, 0, st>>>(w.pm, w.pl, parts, n_local, a_local, w.X.sq, nrm_scale, u_local, diffs,
                                                         w.lseX, w.biasX2, dmax_x, dmax_y);
  OTK_LAUNCH_CHECK();
  return OTK_OK;
}

// Plan statistics of a row shard (x_local, u_local) against the replicated (y, v), on the prepared operands:
//   part[0] += <C, pi> over the local rows, part[1] += their mass, part[2] = max_i |sum_j pi_ij - a_i|  (local rows are
//   complete: y is replicated), row_marginal[n_local], and col_partial[M] = sum over the LOCAL rows of pi_ij - the caller
//   sums it over the ranks (it is a column marginal only then; part[3] is the local max |col_partial - b| and is
//   meaningful on one rank only).
int sk_umma_summary(const float* x_local, const float* y, int64_t n_local, int64_t M, int64_t dim, const float* a_local,
                    const float* b, const float* u_local, const float* v, double scale, double reg, int reuse_prepared,
                    double* part, float* row_marginal, float* col_partial, void* workspace, size_t workspace_bytes,
                    cudaStream_t st) {
  FsWork w;
  OTK_TRY(fs_carve(w, x_local, y, n_local, M, dim, workspace, workspace_bytes, st, !reuse_prepared));
  const float nrm_scale = (float)(scale / reg), g2 = (float)(2.0 * scale / reg) * LOG2E;
  fs_bias_kernel<<<fs_grid(n_local), 256, 0, st>>>(u_local, w.X.sq, nrm_scale, n_local, w.biasX2);
  fs_bias_kernel<<<fs_grid(M), 256, 0, st>>>(v, w.Y.sq, nrm_scale, M, w.biasY2);
  count_launch(1);
  OTK_LAUNCH_CHECK();
  OTK_CUDA(cudaMemsetAsync(part, 0, 4 * sizeof(double), st));
  int parts = 0;
  OTK_TRY(fs_pass<true>(w.X, w.Y, w.biasY2, g2, w.sig + 1, dim, w.pm, w.pl, w.pc, nullptr, &parts, st));
  fs_summary_kernel<<<fs_grid(n_local), 256, 0, st>>>(w.pm, w.pl, w.pc, parts, n_local, u_local, w.X.sq, nrm_scale, (float)scale,
                                                     a_local, part, 2, row_marginal);
  OTK_TRY(fs_pass<false>(w.Y, w.X, w.biasX2, g2, w.sig + 1, dim, w.pm, w.pl, w.pc, nullptr, &parts, st));
  fs_summary_kernel<<<fs_grid(M), 256, 0, st>>>(w.pm, w.pl, nullptr, parts, M, v, w.Y.sq, nrm_scale, (float)scale, b, part, 3,
                                               col_partial);
  count_launch(1);
  OTK_LAUNCH_CHECK();
  return OTK_OK;
}

}  // namespace otk
