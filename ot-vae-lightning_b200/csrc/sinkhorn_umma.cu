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
#include <cfloat>

#include "otk_ptx.cuh"
#include "sinkhorn_umma.cuh"
#include "tensormap.cuh"

namespace otk {

constexpr int FS_BM = 128, FS_BN = 128, FS_SLAB_BYTES = 128 * 128, FS_MAX_SLABS = 4, FS_QSTAGES = 2, FS_SBUF = 4;
constexpr int FS_THREADS = 320;
constexpr int FS_P_BYTES = FS_MAX_SLABS * FS_SLAB_BYTES;            // 64 KiB
constexpr int FS_Q_BYTES = FS_MAX_SLABS * FS_SLAB_BYTES;            // 64 KiB per stage
constexpr int FS_SMEM = FS_P_BYTES + FS_QSTAGES * FS_Q_BYTES + 1024 /*align*/ + 4096 /*bias staging, merge, barriers*/;
constexpr float FS_NEG = -1.0e30f;
constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;

struct FsState { int done; int iters; };

__device__ __forceinline__ float ex2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void named_bar(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// out: part_m / part_l [split][Np] (base-2 max and sum of 2^(t - max)); COST also part_c = sum 2^(t-max) (nq_q - 2 p.q)
template <bool COST>
__global__ void __launch_bounds__(FS_THREADS, 1)
fused_lse_kernel(const __grid_constant__ CUtensorMap mapP, const __grid_constant__ CUtensorMap mapQ,
                 const float* __restrict__ bias2, const float* __restrict__ nq, float g2, int Np, int Nq, int dim,
                 int tiles_per_split, float* __restrict__ part_m, float* __restrict__ part_l, float* __restrict__ part_c,
                 const FsState* __restrict__ state) {
  using namespace ptx;
  if (state && state->done) return;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sP = smem;
  uint8_t* sQ = smem + FS_P_BYTES;
  float* s_bias = reinterpret_cast<float*>(smem + FS_P_BYTES + FS_QSTAGES * FS_Q_BYTES);   // [2 wg][128]
  float* s_nq = s_bias + 256;                                                                // [2 wg][128]
  float* s_merge = s_nq + 256;                                                               // [3][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_merge + 384);
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
  const int nks = (dim + 31) / 32;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapP); tma_prefetch_desc(&mapQ);
    mbar_init(p_full, 1);
    for (int s = 0; s < FS_QSTAGES; ++s) { mbar_init(&q_full[s], 1); mbar_init(&q_empty[s], 1); }
    for (int b = 0; b < FS_SBUF; ++b) { mbar_init(&s_full[b], 1); mbar_init(&s_empty[b], 128); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(p_full, (uint32_t)nks * FS_SLAB_BYTES);
      for (int sl = 0; sl < nks; ++sl) tma_load_3d(sP + sl * FS_SLAB_BYTES, &mapP, sl * 32, pb * FS_BM, 0, p_full);
      for (int idx = 0; idx < n_tiles; ++idx) {
        const int s = idx % FS_QSTAGES, it = idx / FS_QSTAGES;
        mbar_wait(&q_empty[s], (it & 1) ^ 1);
        mbar_arrive_expect_tx(&q_full[s], (uint32_t)nks * FS_SLAB_BYTES);
        for (int sl = 0; sl < nks; ++sl)
          tma_load_3d(sQ + s * FS_Q_BYTES + sl * FS_SLAB_BYTES, &mapQ, sl * 32, (jt0 + idx) * FS_BN, 0, &q_full[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = idesc_tf32(FS_BM, FS_BN, 0, 0);
      mbar_wait(p_full, 0);
      for (int idx = 0; idx < n_tiles; ++idx) {
        const int s = idx % FS_QSTAGES, b = idx % FS_SBUF;
        mbar_wait(&q_full[s], (idx / FS_QSTAGES) & 1);
        mbar_wait(&s_empty[b], ((idx / FS_SBUF) & 1) ^ 1);
        tc_fence_after();
        const uint32_t pbase = smem_u32(sP), qbase = smem_u32(sQ + s * FS_Q_BYTES);
        for (int sl = 0; sl < nks; ++sl) {
          const int ksteps = min(4, (dim - sl * 32 + 7) / 8);
#pragma unroll 4
          for (int kk = 0; kk < ksteps; ++kk) {
            const uint64_t pd = smem_desc_sw128(pbase + sl * FS_SLAB_BYTES + kk * 32, 16, 1024);
            const uint64_t qd = smem_desc_sw128(qbase + sl * FS_SLAB_BYTES + kk * 32, 16, 1024);
            umma_tf32(tmem_base + b * FS_BN, pd, qd, idesc, (sl | kk) != 0);
          }
        }
        umma_commit(&q_empty[s]);
        umma_commit(&s_full[b]);
      }
    }
  } else {
    // ===== online log-sum-exp: thread <-> row of S, two warpgroups alternate over the tiles =====
    const int wg = (warp - 2) / 4;
    const int tid = (warp - 2) % 4 * 32 + lane;          // 0..127 inside the warpgroup
    const int q4 = warp % 4;                             // TMEM lane quarter of this warp
    float* sb = s_bias + wg * 128;
    float* sn = s_nq + wg * 128;
    float m = -3.0e38f, l = 0.f, lc = 0.f;
    for (int idx = wg; idx < n_tiles; idx += 2) {
      const int b = idx % FS_SBUF;
      const int qcol = (jt0 + idx) * FS_BN + tid;
      const float my_bias = qcol < Nq ? bias2[qcol] : FS_NEG;
      float my_nq = 0.f;
      if (COST) my_nq = qcol < Nq ? nq[qcol] : 0.f;
      mbar_wait(&s_full[b], (idx / FS_SBUF) & 1);
      tc_fence_after();
      named_bar(1 + wg, 128);                            // everyone is done reading the previous tile's staging
      sb[tid] = my_bias;
      if (COST) sn[tid] = my_nq;
      named_bar(1 + wg, 128);
      // software pipeline over the four 32-column chunks: the TMEM load of chunk c+1 is in flight while chunk c is reduced
      const uint32_t trow = tmem_base + ((uint32_t)(q4 * 32) << 16) + b * FS_BN;
      float va[32], vb[32];
      tmem_ld32(trow, va);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float* v = (c & 1) ? vb : va;
        float* vnext = (c & 1) ? va : vb;
        if (c < 3) tmem_ld32(trow + (c + 1) * 32, vnext);
        const int c0 = c * 32;
        float cmax = -3.0e38f;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 bb = *reinterpret_cast<const float4*>(sb + c0 + j);
          v[j] = fmaf(v[j], g2, bb.x); v[j + 1] = fmaf(v[j + 1], g2, bb.y);
          v[j + 2] = fmaf(v[j + 2], g2, bb.z); v[j + 3] = fmaf(v[j + 3], g2, bb.w);
          cmax = fmaxf(cmax, fmaxf(fmaxf(v[j], v[j + 1]), fmaxf(v[j + 2], v[j + 3])));
        }
        const float m_new = fmaxf(m, cmax);
        const float rescale = ex2(m - m_new);
        l *= rescale;
        if (COST) lc *= rescale;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float e0 = ex2(v[j] - m_new), e1 = ex2(v[j + 1] - m_new), e2 = ex2(v[j + 2] - m_new), e3 = ex2(v[j + 3] - m_new);
          a0 += e0; a1 += e1; a2 += e2; a3 += e3;
          if (COST) {
            // v holds t = g2 * s + bias2: recover the raw dot product s for the cost term
            const float4 nn = *reinterpret_cast<const float4*>(sn + c0 + j);
            const float4 bb = *reinterpret_cast<const float4*>(sb + c0 + j);
            const float ig = 1.f / g2;
            lc += e0 * fmaf((v[j] - bb.x) * ig, -2.f, nn.x) + e1 * fmaf((v[j + 1] - bb.y) * ig, -2.f, nn.y) +
                  e2 * fmaf((v[j + 2] - bb.z) * ig, -2.f, nn.z) + e3 * fmaf((v[j + 3] - bb.w) * ig, -2.f, nn.w);
          }
        }
        l += (a0 + a1) + (a2 + a3);
        m = m_new;
        if (c < 3) tmem_ld_wait();
      }
      tc_fence_before();
      mbar_arrive(&s_empty[b]);
    }
    if (wg == 1) { s_merge[tid] = m; s_merge[128 + tid] = l; if (COST) s_merge[256 + tid] = lc; }
    named_bar(3, 256);
    if (wg == 0) {
      const float m1 = s_merge[tid], l1 = s_merge[128 + tid];
      const float mm = fmaxf(m, m1);
      const float w0 = ex2(m - mm), w1 = ex2(m1 - mm);
      const int row = pb * FS_BM + q4 * 32 + lane;
      if (row < Np) {
        part_m[(int64_t)split * Np + row] = mm;
        part_l[(int64_t)split * Np + row] = l * w0 + l1 * w1;
        if (COST) part_c[(int64_t)split * Np + row] = lc * w0 + s_merge[256 + tid] * w1;
      }
    }
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ---- small kernels ----------------------------------------------------------------------------------------------
// hi plane (TF32-rounded points, the tensor-core operands) and squared norms of the ORIGINAL points (exact norms keep
// the rounding of the cross term zero-mean along both axes, so it averages out of the marginals); one warp per point
__global__ void fs_prep_kernel(const float* __restrict__ x, int64_t n, int64_t d, float* __restrict__ hi, float* __restrict__ sq) {
  const int lane = threadIdx.x % 32;
  const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / 32;
  if (row >= n) return;
  float acc = 0.f;
  for (int64_t k = lane; k < d; k += 32) {
    const float xv = x[row * d + k];
    hi[row * d + k] = ptx::tf32_rna(xv);
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
                                   float* __restrict__ out_l, float* __restrict__ diff, const FsState* state) {
  if (state && state->done) return;
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
    if (mode == 0) {
      const float L = (m + log2f(l)) * LN2;
      const float bnat = logf(marg[r] + 1e-8f) - L;
      const float pn = bnat + sq[r] * nrm_scale;
      acc += fabsf(pn - pot[r]);
      pot[r] = pn;
      bias2_out[r] = bnat * LOG2E;
    } else if (mode == 1) {
      out_m[r] = m * LN2;
      out_l[r] = l;
    } else {
      out_m[r] = sq[r] + m;
    }
  }
  if (mode == 0 && diff) {
    acc = warp_sum(acc);
    if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = acc;
    __syncthreads();
    if (threadIdx.x == 0) { float t = 0; for (int w = 0; w < blockDim.x / 32; ++w) t += red[w]; atomicAdd(diff, t); }
  }
}

// bias2[r] = (pot[r] - sq[r] * nrm_scale) * log2e   (bias of a side from its potentials)
__global__ void fs_bias_kernel(const float* __restrict__ pot, const float* __restrict__ sq, float nrm_scale, int64_t n,
                               float* __restrict__ bias2) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x)
    bias2[r] = (pot ? pot[r] - sq[r] * nrm_scale : -sq[r] * nrm_scale) * LOG2E;
}

__global__ void fs_check_kernel(float* diff, double threshold, FsState* state) {
  if (state->done) return;
  const double d = (double)diff[0] + (double)diff[1];
  diff[0] = 0.f; diff[1] = 0.f;
  state->iters += 1;
  if (d < threshold) state->done = 1;
}
__global__ void fs_init_state_kernel(FsState* st, float* diff) { st->done = 0; st->iters = 0; diff[0] = 0.f; diff[1] = 0.f; }

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
  return cost_kind == OTK_COST_SQEUCLIDEAN && dim >= 8 && dim <= 128 && dim % 4 == 0 && N >= 1 && M >= 1 &&
         N < (1ll << 30) && M < (1ll << 30) && tensormap_encoder() != nullptr;
}

static int fs_splits(int64_t p_rows, int64_t q_rows) {
  const int64_t pblocks = ceil_div(p_rows, FS_BM), qtiles = ceil_div(q_rows, FS_BN);
  int64_t want = ceil_div((int64_t)sm_count() * 6, pblocks);       // ~6 waves of CTAs
  if (want > qtiles) want = qtiles;
  if (want > 64) want = 64;
  if (want < 1) want = 1;
  return (int)want;
}

struct FsSide {           // one point cloud, prepared
  const float* raw; float* hi; float* sq; int64_t n;
  CUtensorMap map;
};

struct FsWork {
  FsSide X, Y;
  float *biasX2, *biasY2, *pm, *pl, *pc, *diff, *scratch_n;
  FsState* state;
  int max_parts;
};

size_t sk_umma_workspace_bytes(int64_t N, int64_t M, int64_t dim) {
  const int64_t mx = N > M ? N : M;
  return align_up((size_t)N * dim * 4, 256) + align_up((size_t)M * dim * 4, 256) + 6 * align_up((size_t)mx * 4, 256) +
         3 * align_up((size_t)64 * mx * 4, 256) + 4096;
}

static int fs_carve(FsWork& w, const float* x, const float* y, int64_t N, int64_t M, int64_t dim, void* workspace,
                    size_t workspace_bytes, cudaStream_t st) {
  Arena ar(workspace, workspace_bytes);
  const int64_t mx = N > M ? N : M;
  w.X = FsSide{x, ar.take<float>((size_t)N * dim), ar.take<float>((size_t)N), N, {}};
  w.Y = FsSide{y, ar.take<float>((size_t)M * dim), ar.take<float>((size_t)M), M, {}};
  w.biasX2 = ar.take<float>((size_t)N);
  w.biasY2 = ar.take<float>((size_t)M);
  w.scratch_n = ar.take<float>((size_t)mx);
  w.max_parts = 64;
  w.pm = ar.take<float>((size_t)64 * mx);
  w.pl = ar.take<float>((size_t)64 * mx);
  w.pc = ar.take<float>((size_t)64 * mx);
  w.diff = ar.take<float>(64);
  w.state = ar.take<FsState>(1);
  if (!ar.ok()) return OTK_ERR_WORKSPACE;
  fs_prep_kernel<<<(unsigned)ceil_div(N * 32, 256), 256, 0, st>>>(x, N, dim, w.X.hi, w.X.sq);
  fs_prep_kernel<<<(unsigned)ceil_div(M * 32, 256), 256, 0, st>>>(y, M, dim, w.Y.hi, w.Y.sq);
  count_launch(1);
  OTK_LAUNCH_CHECK();
  // 2-D maps [rows, dim], box 32 x 128, 128B swizzle (K-major operands)
  if (!encode_map_f32_3d(&w.X.map, w.X.hi, dim, N, 1, dim, N * dim, 32, FS_BM)) return OTK_ERR_CUDA;
  if (!encode_map_f32_3d(&w.Y.map, w.Y.hi, dim, M, 1, dim, M * dim, 32, FS_BM)) return OTK_ERR_CUDA;
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    OTK_CUDA(cudaFuncSetAttribute(fused_lse_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FS_SMEM));
    OTK_CUDA(cudaFuncSetAttribute(fused_lse_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FS_SMEM));
    attr_set[dev] = true;
  }
  return OTK_OK;
}

// one pass: partials of LSE_q(bias2_q + g2 * p.q) for every row of P; returns the number of parts written
template <bool COST>
static int fs_pass(const FsSide& P, const FsSide& Q, const float* biasQ2, float g2, int64_t dim, float* pm, float* pl,
                   float* pc, const FsState* state, int* parts_out, cudaStream_t st) {
  const int splits = fs_splits(P.n, Q.n);
  const int64_t qtiles = ceil_div(Q.n, FS_BN);
  const int tps = (int)ceil_div(qtiles, splits);
  const int parts = (int)ceil_div(qtiles, tps);
  dim3 grid((unsigned)ceil_div(P.n, FS_BM), (unsigned)parts);
  fused_lse_kernel<COST><<<grid, FS_THREADS, FS_SMEM, st>>>(P.map, Q.map, biasQ2, Q.sq, g2, (int)P.n, (int)Q.n, (int)dim,
                                                          tps, pm, pl, pc, state);
  OTK_LAUNCH_CHECK();
  *parts_out = parts;
  return OTK_OK;
}

static unsigned fs_grid(int64_t n) {
  int64_t b = ceil_div(n, 256), cap = (int64_t)sm_count() * 4;
  return (unsigned)(b < cap ? (b ? b : 1) : cap);
}

// max_ij |x_i - y_j|^2 on the device -> *out_dev (one fused pass with g2 = -2, bias = |y|^2: the running max is the answer)
static int fs_cost_max(FsWork& w, int64_t dim, float* out_dev, cudaStream_t st) {
  int parts = 0;
  fs_bias_kernel<<<fs_grid(w.Y.n), 256, 0, st>>>(nullptr, w.Y.sq, -1.f / LOG2E, w.Y.n, w.biasY2);  // bias2 = +|y|^2
  // the LSE pass maximises bias2 + g2 * s ; we need max(|y|^2 - 2 x.y): g2 = -2
  OTK_TRY(fs_pass<false>(w.X, w.Y, w.biasY2, -2.f, dim, w.pm, w.pl, w.pc, nullptr, &parts, st));
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
    OTK_TRY(fs_pass<false>(w.Y, w.X, w.biasX2, g2, dim, w.pm, w.pl, w.pc, w.state, &parts, st));
    fs_finalize_kernel<<<fs_grid(M), 256, 0, st>>>(w.pm, w.pl, parts, M, 0, b, w.Y.sq, nrm_scale, v, w.biasY2, nullptr, nullptr,
                                                  w.diff + 1, w.state);
    OTK_TRY(fs_pass<false>(w.X, w.Y, w.biasY2, g2, dim, w.pm, w.pl, w.pc, w.state, &parts, st));
    fs_finalize_kernel<<<fs_grid(N), 256, 0, st>>>(w.pm, w.pl, parts, N, 0, a, w.X.sq, nrm_scale, u, w.biasX2, nullptr, nullptr,
                                                  w.diff, w.state);
    fs_check_kernel<<<1, 1, 0, st>>>(w.diff, threshold, w.state);
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
    OTK_TRY(fs_pass<true>(w.X, w.Y, w.biasY2, g2, dim, w.pm, w.pl, w.pc, nullptr, &parts, st));
    fs_summary_kernel<<<fs_grid(N), 256, 0, st>>>(w.pm, w.pl, w.pc, parts, N, u, w.X.sq, nrm_scale, (float)scale, a, summary, 2, row_marginal);
    OTK_TRY(fs_pass<false>(w.Y, w.X, w.biasX2, g2, dim, w.pm, w.pl, w.pc, nullptr, &parts, st));
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

// row-sharded half-steps (stateless: the operands are re-prepared per call, ~2 passes over the points)
int sk_umma_colstep(const float* x_local, const float* y, int64_t n_local, int64_t M, int64_t dim, const float* u_local,
                    double scale, double reg, int precision, float* col_max, float* col_sum, void* workspace,
                    size_t workspace_bytes, cudaStream_t st) {
  (void)precision;
  FsWork w;
  OTK_TRY(fs_carve(w, x_local, y, n_local, M, dim, workspace, workspace_bytes, st));
  const float nrm_scale = (float)(scale / reg), g2 = (float)(2.0 * scale / reg) * LOG2E;
  fs_bias_kernel<<<fs_grid(n_local), 256, 0, st>>>(u_local, w.X.sq, nrm_scale, n_local, w.biasX2);
  int parts = 0;
  OTK_TRY(fs_pass<false>(w.Y, w.X, w.biasX2, g2, dim, w.pm, w.pl, w.pc, nullptr, &parts, st));
  // partial over the LOCAL rows of LSE_i(u_i + Cr_ij) = -nrm_j + LSE_i(bias_i + gamma x_i.y_j), natural log
  fs_finalize_kernel<<<fs_grid(M), 256, 0, st>>>(w.pm, w.pl, parts, M, 1, nullptr, nullptr, 0.f, nullptr, nullptr, col_max, col_sum,
                                                nullptr, nullptr);
  fs_fold_kernel<<<fs_grid(M), 256, 0, st>>>(col_max, w.Y.sq, nrm_scale, M);
  count_launch(2);
  OTK_LAUNCH_CHECK();
  return OTK_OK;
}

__global__ void fs_rowstep_finish_kernel(const float* __restrict__ pm, const float* __restrict__ pl, int parts, int64_t n,
                                         const float* __restrict__ marg, const float* __restrict__ sq, float nrm_scale,
                                         float* __restrict__ pot, float* diff) {
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
    const float pn = logf(marg[r] + 1e-8f) - (m + log2f(l)) * LN2 + sq[r] * nrm_scale;
    acc += fabsf(pn - pot[r]);
    pot[r] = pn;
  }
  acc = warp_sum(acc);
  if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = acc;
  __syncthreads();
  if (threadIdx.x == 0 && diff) { float t = 0; for (int w = 0; w < blockDim.x / 32; ++w) t += red[w]; atomicAdd(diff, t); }
}

int sk_umma_rowstep(const float* x_local, const float* y, int64_t n_local, int64_t M, int64_t dim, const float* a_local,
                    const float* v, double scale, double reg, int precision, float* u_local, float* diff, void* workspace,
                    size_t workspace_bytes, cudaStream_t st) {
  (void)precision;
  FsWork w;
  OTK_TRY(fs_carve(w, x_local, y, n_local, M, dim, workspace, workspace_bytes, st));
  const float nrm_scale = (float)(scale / reg), g2 = (float)(2.0 * scale / reg) * LOG2E;
  fs_bias_kernel<<<fs_grid(M), 256, 0, st>>>(v, w.Y.sq, nrm_scale, M, w.biasY2);
  int parts = 0;
  OTK_TRY(fs_pass<false>(w.X, w.Y, w.biasY2, g2, dim, w.pm, w.pl, w.pc, nullptr, &parts, st));
  fs_rowstep_finish_kernel<<<fs_grid(n_local), 256, 0, st>>>(w.pm, w.pl, parts, n_local, a_local, w.X.sq, nrm_scale, u_local, diff);
  count_launch(1);
  OTK_LAUNCH_CHECK();
  return OTK_OK;
}

}  // namespace otk
