// tcgen05 GEMM with fp32-accurate 3xTF32 arithmetic (sm_100a).
//
//   C[m,n] = alpha * sum_k A(m,k) B(n,k) + beta*C + bias[n] + diag_add*delta_mn      (optionally also emitted as a
//   TF32 hi/lo pair so that the next GEMM of a chain - the Newton-Schulz iteration - needs no conversion pass).
//
// Operands arrive as two TF32-exact planes (hi = rna_tf32(x), lo = rna_tf32(x - hi)); the product is
//   hi*hi' + hi*lo' + lo*hi'   accumulated in fp32 in TMEM (the dropped lo*lo' term is ~2^-22 relative).
// Data path: TMA (cp.async.bulk.tensor, 128B swizzle) -> shared memory -> tcgen05.mma kind::tf32 (one elected lane of a
// warp-uniform loop) -> 128 x BN fp32 accumulator in TMEM -> tcgen05.ld -> registers -> global.
// A is K-major ([M,K], K contiguous).  B is either K-major ([N,K]: "NT") or N-major ([K,N] row-major: "NN"); the
// N-major case uses the MN-major canonical UMMA layout so that a true A*B needs no transpose pass.
//
// The d x d products of the Newton-Schulz chain are latency-bound (a 512^3 product is 16 tiles of 128 x 128), so
//   * BN = 64 tiles are used when 128-wide tiles would leave most SMs idle,
//   * one launch can carry TWO independent products (Y <- Y T and Z <- T Z of an iteration) through blockIdx.z,
//   * a launch can be made conditional on a device-side iteration limit (`ctrl[0]`), so the host enqueues iterations
//     without synchronising and the ones past convergence retire immediately.
//
// Split-K over a thread-block cluster (KS = 2 or 4 CTAs along K): a d x d product streams (128 + BN) x K x 8 bytes of
// operand planes through ONE SM's L2 port per CTA - 768 KB at 512^3, i.e. ~12 of the 17 us such a launch took - while
// three quarters of the SMs idle.  With KS CTAs per output tile each streams 1/KS of K; the non-leaders park their fp32
// accumulator in their own (by then idle) pipeline shared memory and the leader adds it to its registers through
// distributed shared memory (ld.shared::cluster) before the usual epilogue.  No extra launch, no global workspace.
//
// CTA = 192 threads: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-5 = epilogue
// (warp w reads TMEM lanes 32*(w%4)..+31).  One 128 x BN output tile per CTA; UG_STAGES-deep smem ring over K.
#include <cuda.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#include "gemm.cuh"
#include "otk_ptx.cuh"
#include "tensormap.cuh"

namespace otk {

constexpr int UG_BM = 128, UG_BK = 32, UG_THREADS = 192;
template <int BN> __host__ __device__ constexpr int ug_stages() { return BN == 128 ? 3 : 4; }
constexpr int UG_TILE_BYTES = UG_BM * UG_BK * 4;                   // 16 KiB per A plane per stage
constexpr int UG_XPOSE = 32 * 33 * 4;                              // per-epilogue-warp transpose buffer
template <int BN> __host__ __device__ constexpr int ug_btile() { return BN * UG_BK * 4; }
template <int BN> __host__ __device__ constexpr int ug_stage() { return 2 * UG_TILE_BYTES + 2 * ug_btile<BN>(); }   // A_hi, A_lo, B_hi, B_lo
template <int BN> __host__ __device__ constexpr int ug_smem() { return ug_stages<BN>() * ug_stage<BN>() + 4 * UG_XPOSE + 1024 /*align*/ + 256 /*barriers*/; }

struct UmmaGemmParams {
  int M, N, K, passes, b_mn_major;
  float *C, *C_hi, *C_lo;
  int64_t ldc, strideC;
  float alpha, beta, diag_add;
  const float* bias;
  int64_t stride_bias;
  double* resid;
  const float *add_hi, *add_lo;   // optional addend planes (layout of C): C += add_scale * (add_hi + add_lo)
  float add_scale;
};
struct UmmaGemmMaps { CUtensorMap A_hi, A_lo, B_hi, B_lo; };

// Epilogue rows of the common case of the Newton-Schulz chain - full 32-row slice, result emitted as TF32 hi/lo planes
// only - without per-row pointer / bounds tests: the 32 rows are independent straight-line code the scheduler can
// interleave.  (With the tests inside the loop every row was a branchy dependent chain of ~190 cycles on the single
// epilogue warp of its scheduler: 3.1 us per 32-row chunk, 6.9 us of a 10 - 16 us launch - globaltimer stamps.)
template <bool RESID, bool ADD = false>
__device__ __forceinline__ void epilogue_rows_planes(const float (&v)[32], int mrow0, int n, int64_t ldc, float alpha,
                                                     float bn, float diag_add, float* __restrict__ Ch,
                                                     float* __restrict__ Cl, double& res,
                                                     const float* __restrict__ Ah = nullptr,
                                                     const float* __restrict__ Al = nullptr, float add_scale = 0.f) {
  float addend[32];
  if (ADD) {       // all 64 loads of the addend planes are issued before the first store of the result
#pragma unroll
    for (int r = 0; r < 32; ++r) {
      const int64_t idx = (int64_t)(mrow0 + r) * ldc + n;
      addend[r] = __ldg(Ah + idx) + __ldg(Al + idx);
    }
  }
#pragma unroll
  for (int r = 0; r < 32; ++r) {
    const int m = mrow0 + r;
    const float acc = v[r];
    if (RESID) { const float e = acc - (m == n ? 1.f : 0.f); res += (double)(e * e); }
    float t = alpha * acc + bn;
    if (ADD) t = fmaf(add_scale, addend[r], t);
    if (m == n) t += diag_add;
    float h, lo;
    ptx::split_tf32(t, h, lo);
    const int64_t idx = (int64_t)m * ldc + n;
    Ch[idx] = h;
    Cl[idx] = lo;
  }
}

template <int BN, int KS>
__global__ void __launch_bounds__(UG_THREADS, 1)
umma_gemm_kernel(const __grid_constant__ UmmaGemmMaps maps0, const __grid_constant__ UmmaGemmMaps maps1,
                 const UmmaGemmParams p0, const UmmaGemmParams p1, int batch_per_problem, const int* __restrict__ ctrl,
                 int ctrl_index, const NsCtrlEval ev) {
  using namespace ptx;
#ifdef OTK_GEMM_TIMING
  unsigned long long gt0, gts[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};       // per-thread stamps, printed once at the very end
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt0));
#define GT_STAMP(slot) { asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gts[slot])); }
#else
#define GT_STAMP(slot) {}
#endif
  constexpr int BTILE = ug_btile<BN>(), STAGE = ug_stage<BN>(), UG_STAGES = ug_stages<BN>();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* xpose = reinterpret_cast<float*>(smem + UG_STAGES * STAGE);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + UG_STAGES * STAGE + 4 * UG_XPOSE);
  uint64_t* empty = full + UG_STAGES;
  uint64_t* tmem_full = empty + UG_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const bool second = (int)blockIdx.z >= batch_per_problem;          // which of the two products of the launch
  const UmmaGemmMaps& maps = second ? maps1 : maps0;
  const UmmaGemmParams& p = second ? p1 : p0;
  const int ks = KS > 1 ? (int)cluster_ctarank() : 0;               // this CTA's share of K (cluster = KS CTAs along x)
  const int m0 = ((int)blockIdx.x / KS) * UG_BM, n0 = blockIdx.y * BN, batch = (int)blockIdx.z - (second ? batch_per_problem : 0);
  const int num_k_all = (p.K + UG_BK - 1) / UG_BK;
  const int kb0 = ks * num_k_all / KS, num_k = (ks + 1) * num_k_all / KS - kb0;   // k-blocks [kb0, kb0 + num_k), >= 1 each
  const bool three = p.passes == 3;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.A_hi); tma_prefetch_desc(&maps.B_hi);
    if (three) { tma_prefetch_desc(&maps.A_lo); tma_prefetch_desc(&maps.B_lo); }
    for (int s = 0; s < UG_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, BN); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above touched no global data, so it overlaps the tail of the previous
  // launch of the chain; the successor may be scheduled from here on, and nothing below runs before the predecessor
  // grid has completed (its planes, residuals and the iteration limit are then visible).
  GT_STAMP(0)
  pdl_launch_dependents();
  pdl_wait();
  GT_STAMP(1)
  if (ctrl && ctrl_index >= ctrl[0]) {             // past the device-side iteration limit: nothing to do
    if (warp == 1) tmem_dealloc(tmem_base, BN);
    return;
  }

  if (warp == 0) {
    // ===== TMA producer: warp-uniform loop, one elected lane issues (descriptors stay in uniform registers) =====
    const uint32_t stage_tx = (three ? 2u : 1u) * (UG_TILE_BYTES + BTILE);
    for (int kt = 0; kt < num_k; ++kt) {
      const int s = kt % UG_STAGES, it = kt / UG_STAGES;
      mbar_wait(&empty[s], (it & 1) ^ 1);
      uint8_t* st = smem + s * STAGE;
      const int k0 = (kb0 + kt) * UG_BK;
      if (elect_one()) {
        mbar_arrive_expect_tx(&full[s], stage_tx);
        tma_load_3d(st, &maps.A_hi, k0, m0, batch, &full[s]);
        if (three) tma_load_3d(st + UG_TILE_BYTES, &maps.A_lo, k0, m0, batch, &full[s]);
        if (!p.b_mn_major) {
          tma_load_3d(st + 2 * UG_TILE_BYTES, &maps.B_hi, k0, n0, batch, &full[s]);
          if (three) tma_load_3d(st + 2 * UG_TILE_BYTES + BTILE, &maps.B_lo, k0, n0, batch, &full[s]);
        } else {
          // B is [K, N] row-major: BN/32 boxes of 32(n) x 32(k) per plane, one per 128-byte MN slab
#pragma unroll
          for (int sl = 0; sl < BN / 32; ++sl) {
            tma_load_3d(st + 2 * UG_TILE_BYTES + sl * 4096, &maps.B_hi, n0 + 32 * sl, k0, batch, &full[s]);
            if (three) tma_load_3d(st + 2 * UG_TILE_BYTES + BTILE + sl * 4096, &maps.B_lo, n0 + 32 * sl, k0, batch, &full[s]);
          }
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===== MMA issuer: warp-uniform loop, one elected lane issues =====
    const uint32_t idesc = idesc_tf32(UG_BM, BN, 0, p.b_mn_major);
    const uint32_t b_kstep = p.b_mn_major ? 1024u : 32u;        // bytes per UMMA_K = 8 along K
    for (int kt = 0; kt < num_k; ++kt) {
      const int s = kt % UG_STAGES, it = kt / UG_STAGES;
      mbar_wait(&full[s], it & 1);
      tc_fence_after();
      if (kt == 0) GT_STAMP(2)
      if (kt == num_k - 1) GT_STAMP(3)
      const uint32_t base = smem_u32(smem + s * STAGE);
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < UG_BK / 8; ++kk) {
          const uint64_t a_hi = smem_desc_sw128(base + kk * 32, 16, 1024);
          const uint64_t a_lo = smem_desc_sw128(base + UG_TILE_BYTES + kk * 32, 16, 1024);
          const uint32_t b_hi_addr = base + 2 * UG_TILE_BYTES + kk * b_kstep, b_lo_addr = b_hi_addr + BTILE;
          const uint64_t b_hi = p.b_mn_major ? smem_desc_mn_tf32(b_hi_addr, 4096) : smem_desc_sw128(b_hi_addr, 16, 1024);
          const uint64_t b_lo = p.b_mn_major ? smem_desc_mn_tf32(b_lo_addr, 4096) : smem_desc_sw128(b_lo_addr, 16, 1024);
          if (three) {
            umma_tf32(tmem_base, a_lo, b_hi, idesc, (kt | kk) != 0);   // small terms first
            umma_tf32(tmem_base, a_hi, b_lo, idesc, 1);
            umma_tf32(tmem_base, a_hi, b_hi, idesc, 1);
          } else {
            umma_tf32(tmem_base, a_hi, b_hi, idesc, (kt | kk) != 0);
          }
        }
        umma_commit(&empty[s]);                    // frees the smem stage when these MMAs have read it
        if (kt == num_k - 1) umma_commit(tmem_full);   // accumulator complete
      }
      __syncwarp();
    }
  } else {
    // ===== epilogue: TMEM -> registers -> (per-warp 32x32 smem transpose) -> global =====
    // TMEM gives each lane one row; after the transpose lanes hold consecutive columns, so every store instruction
    // writes a full 128-byte line of C / C_hi / C_lo.
    // Split-K cluster (KS > 1): the tile is reduce-scattered - CTA `ks` finishes columns [ks, ks + 1) * BN / KS.  Once the
    // MMAs of every CTA of the cluster have retired (first cluster barrier: the pipeline smem is then free everywhere),
    // each CTA pushes the columns its peers own into the OWNER's smem with remote stores (slot = sender rank, column-major
    // [BN / KS][128] so that lanes write consecutive words); its own share stays in tensor memory.  Remote stores are fire
    // and forget, so the 21 B / clk distributed-shared-memory port is the only cost; the owner then reads local smem.
    if constexpr (KS > 1) {
      mbar_wait(tmem_full, 0);
      tc_fence_after();
    }
  }
  if constexpr (KS > 1) {
    cluster_sync_all();
    if (warp >= 2) {
      constexpr int SHARE = BN / KS;
      const int q = warp % 4;                      // TMEM lane quarter this warp may access
      const uint32_t land = smem_u32(smem) + (uint32_t)((ks * SHARE) * UG_BM + q * 32 + lane) * 4;   // my slot in any owner
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        const int owner = c0 / SHARE;
        if (owner == ks) continue;
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c0, v);
        tmem_ld_wait();
        const uint32_t dst = map_to_cta(land, (uint32_t)owner) + (uint32_t)((c0 - owner * SHARE) * UG_BM) * 4;
#pragma unroll
        for (int j = 0; j < 32; ++j) st_shared_cluster_f32(dst + (uint32_t)(j * UG_BM) * 4, v[j]);
      }
      tc_fence_before();
    }
    cluster_sync_all();                            // the peers' partial columns have landed in this CTA's smem
  }
  if (warp >= 2) {
    static_assert((BN / KS) % 32 == 0, "split-K share must be whole 32-column chunks");
    const int q = warp % 4;
    const uint32_t xp_addr = smem_u32(xpose + (warp - 2) * (32 * 33));
    const int mrow0 = m0 + q * 32;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    GT_STAMP(4)
    float* C = p.C ? p.C + (int64_t)batch * p.strideC : nullptr;
    float* Ch = p.C_hi ? p.C_hi + (int64_t)batch * p.strideC : nullptr;
    float* Cl = p.C_lo ? p.C_lo + (int64_t)batch * p.strideC : nullptr;
    const float* bias = p.bias ? p.bias + (int64_t)batch * p.stride_bias : nullptr;
    const float* Ah = p.add_hi ? p.add_hi + (int64_t)batch * p.strideC : nullptr;
    const float* Al = p.add_lo ? p.add_lo + (int64_t)batch * p.strideC : nullptr;
    double res = 0.0;
#pragma unroll 1
    for (int c0 = ks * (BN / KS); c0 < (ks + 1) * (BN / KS); c0 += 32) {
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c0, v);
      tmem_ld_wait();
      if (c0 == 0) GT_STAMP(6)
      if constexpr (KS > 1) {
        const uint32_t mine = smem_u32(smem) + (uint32_t)((c0 - ks * (BN / KS)) * UG_BM + q * 32 + lane) * 4;
#pragma unroll
        for (int r = 1; r < KS; ++r) {
          const uint32_t slot = mine + (uint32_t)((((ks + r) % KS) * (BN / KS)) * UG_BM) * 4;
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += lds32(slot + (uint32_t)(j * UG_BM) * 4);
        }
      }
      // 32 x 32 transpose through shared memory with explicit st.shared / ld.shared: through the generic `float*` the
      // compiler emitted generic LD.E / ST.E and, unable to tell them from the global stores of C_hi / C_lo, kept one
      // load -> split -> store chain per row in order (globaltimer stamps: 6.9 us of a 10 - 16 us launch were this
      // loop, for any d).  All 32 loads are issued before the first store.
#pragma unroll
      for (int j = 0; j < 32; ++j) sts32(xp_addr + (uint32_t)(lane * 33 + j) * 4, v[j]);
      __syncwarp();
#pragma unroll
      for (int r = 0; r < 32; ++r) v[r] = lds32(xp_addr + (uint32_t)(r * 33 + lane) * 4);
      if (c0 == 0) GT_STAMP(7)
      const int n = n0 + c0 + lane;
      if (n < p.N && !C && Ch && mrow0 + 32 <= p.M) {
        const float bn = bias ? bias[n] : 0.f;
        if (Ah && Al) epilogue_rows_planes<false, true>(v, mrow0, n, p.ldc, p.alpha, bn, p.diag_add, Ch, Cl, res, Ah, Al, p.add_scale);
        else if (p.resid) epilogue_rows_planes<true>(v, mrow0, n, p.ldc, p.alpha, bn, p.diag_add, Ch, Cl, res);
        else epilogue_rows_planes<false>(v, mrow0, n, p.ldc, p.alpha, bn, p.diag_add, Ch, Cl, res);
      } else if (n < p.N) {
        const float bn = bias ? bias[n] : 0.f;
#pragma unroll
        for (int r = 0; r < 32; ++r) {
          const int m = mrow0 + r;
          if (m >= p.M) continue;
          const float acc = v[r];
          const int64_t idx = (int64_t)m * p.ldc + n;
          if (p.resid) { const float e = acc - (m == n ? 1.f : 0.f); res += (double)(e * e); }
          float t = p.alpha * acc + bn;
          if (p.beta != 0.f && C) t += p.beta * C[idx];
          if (Ah) t += p.add_scale * (Ah[idx] + (Al ? Al[idx] : 0.f));
          if (m == n) t += p.diag_add;
          if (C) C[idx] = t;
          if (Ch) {
            float h, lo;
            split_tf32(t, h, lo);
            Ch[idx] = h;
            Cl[idx] = lo;
          }
        }
      }
      __syncwarp();
      if (c0 == 0) GT_STAMP(8)
    }
    GT_STAMP(9)
    if (p.resid) {
      res = warp_sum(res);
      if (lane == 0) atomicAdd(&p.resid[batch], res);
    }
    tc_fence_before();
    GT_STAMP(5)
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, BN); }
#ifdef OTK_GEMM_TIMING
  if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (threadIdx.x % 32) == 0 && ctrl_index == 7) {
    unsigned long long ge;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ge));
    printf("gemm k=7 grid %d warp %d: prologue %lld | pdl wait %lld | first stage %lld | last stage %lld | acc complete %lld | epilogue %lld | end %lld ns || tmem_ld0 %lld lds0 %lld rows0 %lld rows1 %lld\n",
           (int)(gridDim.x * gridDim.y * gridDim.z), (int)(threadIdx.x / 32), (long long)(gts[0] - gt0), (long long)(gts[1] - gt0), gts[2] ? (long long)(gts[2] - gt0) : -1LL,
           gts[3] ? (long long)(gts[3] - gt0) : -1LL, gts[4] ? (long long)(gts[4] - gt0) : -1LL, gts[5] ? (long long)(gts[5] - gt0) : -1LL, (long long)(ge - gt0),
           gts[6] ? (long long)(gts[6] - gt0) : -1LL, gts[7] ? (long long)(gts[7] - gt0) : -1LL, gts[8] ? (long long)(gts[8] - gt0) : -1LL, gts[9] ? (long long)(gts[9] - gt0) : -1LL);
  }
#endif
  // Newton-Schulz stopping rule (same logic as ns_ctrl_kernel), by one warp of one CTA: the residuals of iteration
  // ev.k were completed by the previous launch; the lowered limit is seen by the launches of iteration ev.k + 1 on.
  if (ev.ctrl && warp == 2 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && ev.k < ev.ctrl[0] &&
      (ev.k >= 5 || ev.k + 1 == ev.max_iters)) {
    double worst = 0;
    for (int64_t l = lane; l < ev.L; l += 32) {
      const double r = ev.resid_k[l];
      const double v = (r == r && r <= 1e30) ? r : 1e300;
      worst = v > worst ? v : worst;
    }
    worst = warp_max(worst);
    if (lane == 0) {
      if (worst >= 1e300) { ev.ctrl[0] = ev.k + 1; ev.ctrl[1] = 1; ev.ctrl[2] = 1; }
      else if (worst < ev.tol_done) { if (ev.k + 1 < ev.ctrl[0]) ev.ctrl[0] = ev.k + 1; ev.ctrl[1] = 0; }
      else if (worst < ev.tol_near) { if (ev.k + 2 < ev.ctrl[0]) ev.ctrl[0] = ev.k + 2; ev.ctrl[1] = 0; }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Throughput regime (batched operators, d >= 2048): persistent CTA-PAIR kernel on 256 x 256 tiles.
// The 128 x 128 kernel above streams 64 KB of operand planes per 3.1 Mflop k-block - 21 TB/s of L2 -> SM traffic at the
// TF32 peak, which the L2 cannot deliver (measured: 54 % of the tcgen05 TF32 ceiling at 10 x 1024^3) - and its epilogue is
// serial with its main loop.  Here a pair of CTAs (tcgen05 cta_group::2, UMMA 256 x 256 x 8) shares one tile: each CTA
// loads ITS 128 rows of A and ITS 128 of the 256 B columns (the same 64 KB per k-block, for twice the flops), both CTAs'
// TMA bytes are credited to the leader's `full` barrier, the leader issues the MMAs and multicasts the commits; two
// 256-column accumulators in tensor memory let the epilogue of tile i overlap the main loop of tile i + 1; the grid is
// persistent (one pair per two SMs, tiles dealt round-robin).  Plane operands and plane results only (what the
// Newton-Schulz chain uses); same residual / addend / device-side iteration limit / stop rule as the kernel above.
constexpr int G2_BN = 256, G2_STAGES = 3, G2_THREADS = 192;
constexpr int G2_PLANE = UG_BM * UG_BK * 4;                        // 16 KiB: 128 rows (A) or 128 columns (B half) x 32 k
constexpr int G2_STAGE = 4 * G2_PLANE;                             // A_hi, A_lo, B_hi, B_lo
constexpr int G2_SMEM = G2_STAGES * G2_STAGE + 4 * UG_XPOSE + 1024 + 256;

__global__ void __launch_bounds__(G2_THREADS, 1)
umma_gemm_pair_kernel(const __grid_constant__ UmmaGemmMaps maps0, const __grid_constant__ UmmaGemmMaps maps1,
                      const UmmaGemmParams p0, const UmmaGemmParams p1, int n_problems, int batch_per_problem,
                      const int* __restrict__ ctrl, int ctrl_index, const NsCtrlEval ev) {
  using namespace ptx;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* xpose = reinterpret_cast<float*>(smem + G2_STAGES * G2_STAGE);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + G2_STAGES * G2_STAGE + 4 * UG_XPOSE);
  uint64_t* empty = full + G2_STAGES;
  uint64_t* acc_full = empty + G2_STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x / 2, n_pairs = gridDim.x / 2;
  const int tiles_m = (p0.M + 2 * UG_BM - 1) / (2 * UG_BM), tiles_n = (p0.N + G2_BN - 1) / G2_BN;
  const int per_batch = tiles_m * tiles_n, per_problem = batch_per_problem * per_batch;
  const int total = n_problems * per_problem;
  const int num_k = (p0.K + UG_BK - 1) / UG_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps0.A_hi); tma_prefetch_desc(&maps0.A_lo); tma_prefetch_desc(&maps0.B_hi); tma_prefetch_desc(&maps0.B_lo);
    for (int s = 0; s < G2_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 8); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc_cg<2>(tmem_slot, 512); tmem_relinquish_cg<2>(); }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();
  if (ctrl && ctrl_index >= ctrl[0]) {             // past the device-side iteration limit: nothing to do (both CTAs agree)
    cluster_sync_all();
    if (warp == 1) tmem_dealloc_cg<2>(tmem_base, 512);
    return;
  }

  if (warp == 0) {
    // ===== TMA producer: this CTA's rows of A and its half of the B columns; bytes land on the LEADER's full barrier =====
    const uint32_t full_leader = map_to_cta(smem_u32(&full[0]), 0);
    int it = 0;
    for (int tile = pair; tile < total; tile += n_pairs) {
      const bool second = tile >= per_problem;
      const UmmaGemmMaps& maps = second ? maps1 : maps0;
      const UmmaGemmParams& p = second ? p1 : p0;
      const int rem = tile - (second ? per_problem : 0);
      const int batch = rem / per_batch, t2 = rem % per_batch;
      const int m0 = (t2 / tiles_n) * (2 * UG_BM) + (int)rank * UG_BM;
      const int n0 = (t2 % tiles_n) * G2_BN + (int)rank * 128;
      for (int kt = 0; kt < num_k; ++kt, ++it) {
        const int s = it % G2_STAGES;
        mbar_wait(&empty[s], ((it / G2_STAGES) & 1) ^ 1);
        uint8_t* st = smem + s * G2_STAGE;
        const int k0 = kt * UG_BK;
        if (elect_one()) {
          if (rank == 0) mbar_arrive_expect_tx(&full[s], 2 * G2_STAGE);
          const uint32_t bar = full_leader + s * 8;
          tma_load_3d_cg2(st, &maps.A_hi, k0, m0, batch, bar);
          tma_load_3d_cg2(st + G2_PLANE, &maps.A_lo, k0, m0, batch, bar);
          if (!p.b_mn_major) {
            tma_load_3d_cg2(st + 2 * G2_PLANE, &maps.B_hi, k0, n0, batch, bar);
            tma_load_3d_cg2(st + 3 * G2_PLANE, &maps.B_lo, k0, n0, batch, bar);
          } else {
#pragma unroll
            for (int sl = 0; sl < 4; ++sl) {       // B is [K, N] row-major: one 32(n) x 32(k) box per 128-byte MN slab
              tma_load_3d_cg2(st + 2 * G2_PLANE + sl * 4096, &maps.B_hi, n0 + 32 * sl, k0, batch, bar);
              tma_load_3d_cg2(st + 3 * G2_PLANE + sl * 4096, &maps.B_lo, n0 + 32 * sl, k0, batch, bar);
            }
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA only): warp-uniform loop, one elected lane issues =====
    if (rank == 0) {
      const uint32_t idesc = idesc_tf32(2 * UG_BM, G2_BN, 0, p0.b_mn_major);
      const uint32_t b_kstep = p0.b_mn_major ? 1024u : 32u;
      int it = 0, ti = 0;
      for (int tile = pair; tile < total; tile += n_pairs, ++ti) {
        const int a = ti & 1;
        mbar_wait(&acc_empty[a], ((ti >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t acc = tmem_base + a * G2_BN;
        for (int kt = 0; kt < num_k; ++kt, ++it) {
          const int s = it % G2_STAGES;
          mbar_wait(&full[s], (it / G2_STAGES) & 1);
          tc_fence_after();
          const uint32_t base = smem_u32(smem + s * G2_STAGE);
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < UG_BK / 8; ++kk) {
              const uint64_t a_hi = smem_desc_sw128(base + kk * 32, 16, 1024);
              const uint64_t a_lo = smem_desc_sw128(base + G2_PLANE + kk * 32, 16, 1024);
              const uint32_t b_hi_addr = base + 2 * G2_PLANE + kk * b_kstep, b_lo_addr = b_hi_addr + G2_PLANE;
              const uint64_t b_hi = p0.b_mn_major ? smem_desc_mn_tf32(b_hi_addr, 4096) : smem_desc_sw128(b_hi_addr, 16, 1024);
              const uint64_t b_lo = p0.b_mn_major ? smem_desc_mn_tf32(b_lo_addr, 4096) : smem_desc_sw128(b_lo_addr, 16, 1024);
              umma_tf32_ss<2>(acc, a_lo, b_hi, idesc, (kt | kk) != 0);   // small terms first
              umma_tf32_ss<2>(acc, a_hi, b_lo, idesc, 1);
              umma_tf32_ss<2>(acc, a_hi, b_hi, idesc, 1);
            }
            umma_commit_cg<2>(&empty[s]);
            if (kt == num_k - 1) umma_commit_cg<2>(&acc_full[a]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===== epilogue: this CTA's 128 x 256 half of the accumulator -> hi / lo planes (thread <-> row, 32-column chunks
    // transposed through shared memory so that every store instruction writes a full 128-byte line) =====
    const int q = warp % 4;
    const uint32_t xp_addr = smem_u32(xpose + (warp - 2) * (32 * 33));
    const uint32_t acc_empty_leader = map_to_cta(smem_u32(&acc_empty[0]), 0);
    int ti = 0;
    for (int tile = pair; tile < total; tile += n_pairs, ++ti) {
      const bool second = tile >= per_problem;
      const UmmaGemmParams& p = second ? p1 : p0;
      const int rem = tile - (second ? per_problem : 0);
      const int batch = rem / per_batch, t2 = rem % per_batch;
      const int mrow0 = (t2 / tiles_n) * (2 * UG_BM) + (int)rank * UG_BM + q * 32;
      const int n0 = (t2 % tiles_n) * G2_BN;
      const int a = ti & 1;
      float* Ch = p.C_hi + (int64_t)batch * p.strideC;
      float* Cl = p.C_lo + (int64_t)batch * p.strideC;
      const float* Ah = p.add_hi ? p.add_hi + (int64_t)batch * p.strideC : nullptr;
      const float* Al = p.add_lo ? p.add_lo + (int64_t)batch * p.strideC : nullptr;
      mbar_wait(&acc_full[a], (ti >> 1) & 1);
      tc_fence_after();
      double res = 0.0;
#pragma unroll 1
      for (int c0 = 0; c0 < G2_BN; c0 += 32) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + a * G2_BN + c0, v);
        tmem_ld_wait();
        if (c0 + 32 == G2_BN) {                    // accumulator drained into registers: the MMAs of tile i + 2 may overwrite it
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(acc_empty_leader + a * 8);
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) sts32(xp_addr + (uint32_t)(lane * 33 + j) * 4, v[j]);
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 32; ++r) v[r] = lds32(xp_addr + (uint32_t)(r * 33 + lane) * 4);
        const int n = n0 + c0 + lane;
        if (n < p.N && mrow0 + 32 <= p.M) {
          if (Ah && Al) epilogue_rows_planes<false, true>(v, mrow0, n, p.ldc, p.alpha, 0.f, p.diag_add, Ch, Cl, res, Ah, Al, p.add_scale);
          else if (p.resid) epilogue_rows_planes<true>(v, mrow0, n, p.ldc, p.alpha, 0.f, p.diag_add, Ch, Cl, res);
          else epilogue_rows_planes<false>(v, mrow0, n, p.ldc, p.alpha, 0.f, p.diag_add, Ch, Cl, res);
        } else if (n < p.N) {
#pragma unroll
          for (int r = 0; r < 32; ++r) {
            const int m = mrow0 + r;
            if (m >= p.M) continue;
            const float acc = v[r];
            const int64_t idx = (int64_t)m * p.ldc + n;
            if (p.resid) { const float e = acc - (m == n ? 1.f : 0.f); res += (double)(e * e); }
            float t = p.alpha * acc;
            if (Ah) t += p.add_scale * (Ah[idx] + (Al ? Al[idx] : 0.f));
            if (m == n) t += p.diag_add;
            float h, lo;
            split_tf32(t, h, lo);
            Ch[idx] = h;
            Cl[idx] = lo;
          }
        }
        __syncwarp();
      }
      if (p.resid) {
        res = warp_sum(res);
        if (lane == 0) atomicAdd(&p.resid[batch], res);
      }
    }
    tc_fence_before();
  }
  cluster_sync_all();
  if (warp == 1) { tc_fence_after(); tmem_dealloc_cg<2>(tmem_base, 512); }
  // Newton-Schulz stopping rule, as in umma_gemm_kernel
  if (ev.ctrl && warp == 2 && blockIdx.x == 0 && ev.k < ev.ctrl[0] && (ev.k >= 5 || ev.k + 1 == ev.max_iters)) {
    double worst = 0;
    for (int64_t l = lane; l < ev.L; l += 32) {
      const double r = ev.resid_k[l];
      const double v = (r == r && r <= 1e30) ? r : 1e300;
      worst = v > worst ? v : worst;
    }
    worst = warp_max(worst);
    if (lane == 0) {
      if (worst >= 1e300) { ev.ctrl[0] = ev.k + 1; ev.ctrl[1] = 1; ev.ctrl[2] = 1; }
      else if (worst < ev.tol_done) { if (ev.k + 1 < ev.ctrl[0]) ev.ctrl[0] = ev.k + 1; ev.ctrl[1] = 0; }
      else if (worst < ev.tol_near) { if (ev.k + 2 < ev.ctrl[0]) ev.ctrl[0] = ev.k + 2; ev.ctrl[1] = 0; }
    }
  }
}

// elementwise split of a strided [batch][rows][cols] operand into dense TF32 hi/lo planes [batch][rows][cols]
__global__ void split_planes_kernel(const float* __restrict__ x, int64_t rows, int64_t cols, int64_t ld, int64_t bstride,
                                    int64_t batch, float* __restrict__ hi, float* __restrict__ lo) {
  const int64_t total = batch * rows * cols;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = e / (rows * cols), r = (e / cols) % rows, c = e % cols;
    float h, l;
    ptx::split_tf32(x[b * bstride + r * ld + c], h, l);
    hi[e] = h;
    lo[e] = l;
  }
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

struct PreparedGemm {
  UmmaGemmMaps maps;
  UmmaGemmParams p;
};

// eligibility, operand planes (splitting A / B into scratch when no lo plane is given) and tensor maps of one product
static int prepare_gemm(const GemmArgs<float>& g, int64_t batch, int passes, int bn, cudaStream_t st, PreparedGemm* out) {
  if (g.a_off || g.sak != 1) return 0;                               // A must be K-major, no centring here
  const bool b_kmajor = (g.sbk == 1), b_nmajor = (g.sbn == 1);
  if (!b_kmajor && !b_nmajor) return 0;
  const int b_mn = b_kmajor ? 0 : 1;
  const int64_t lda = g.sam, ldb = b_kmajor ? g.sbn : g.sbk;
  if (g.M < 64 || g.N < 64 || g.K < 8) return 0;                     // tiny problems: FFMA engine
  if (g.M > INT32_MAX || g.N > INT32_MAX || g.K > INT32_MAX || batch > 32767) return 0;
  if ((lda % 4) || (ldb % 4) || (g.strideA % 4) || (g.strideB % 4)) return 0;   // TMA: 16-byte strides
  if (!aligned16(g.A) || !aligned16(g.B) || (g.A_lo && !aligned16(g.A_lo)) || (g.B_lo && !aligned16(g.B_lo))) return 0;
  if (!g.C && !(g.C_hi && g.C_lo)) return 0;
  const bool need_split_a = passes == 3 && !g.A_lo, need_split_b = passes == 3 && !g.B_lo;
  if ((need_split_a || need_split_b) && !g.scratch) return 0;
  if (!tensormap_encoder()) return 0;

  const float *A_hi = g.A, *A_lo = g.A_lo, *B_hi = g.B, *B_lo = g.B_lo;
  int64_t a_ld = lda, a_bs = g.strideA, b_ld = ldb, b_bs = g.strideB;
  const int64_t b_rows = b_mn ? g.K : g.N, b_cols = b_mn ? g.N : g.K;
  float* scratch = g.scratch;
  auto ew_grid = [](int64_t total) { int64_t b = ceil_div(total, 256), cap = (int64_t)sm_count() * 16; return (unsigned)(b < cap ? (b ? b : 1) : cap); };
  if (need_split_a) {
    const int64_t n = batch * g.M * g.K;
    float *h = scratch, *l = scratch + n;
    scratch += 2 * n;
    split_planes_kernel<<<ew_grid(n), 256, 0, st>>>(g.A, g.M, g.K, lda, g.strideA, batch, h, l);
    OTK_LAUNCH_CHECK();
    A_hi = h; A_lo = l; a_ld = g.K; a_bs = g.M * g.K;
  }
  if (need_split_b) {
    const int64_t n = batch * b_rows * b_cols;
    float *h = scratch, *l = scratch + n;
    split_planes_kernel<<<ew_grid(n), 256, 0, st>>>(g.B, b_rows, b_cols, ldb, g.strideB, batch, h, l);
    OTK_LAUNCH_CHECK();
    B_hi = h; B_lo = l; b_ld = b_cols; b_bs = b_rows * b_cols;
  }
  if ((a_ld % 4) || (b_ld % 4)) return 0;

  // tensor maps: [batch][rows][cols] fp32, 128B swizzle, box 32 cols x {128 | BN | 32} rows
  const int b_box_rows = b_mn ? 32 : bn;
  if (!encode_map_f32_3d(&out->maps.A_hi, A_hi, g.K, g.M, batch, a_ld, a_bs, 32, UG_BM)) return 0;
  if (!encode_map_f32_3d(&out->maps.B_hi, B_hi, b_cols, b_rows, batch, b_ld, b_bs, 32, b_box_rows, b_mn != 0)) return 0;
  out->maps.A_lo = out->maps.A_hi; out->maps.B_lo = out->maps.B_hi;
  if (passes == 3) {
    if (!encode_map_f32_3d(&out->maps.A_lo, A_lo, g.K, g.M, batch, a_ld, a_bs, 32, UG_BM)) return 0;
    if (!encode_map_f32_3d(&out->maps.B_lo, B_lo, b_cols, b_rows, batch, b_ld, b_bs, 32, b_box_rows, b_mn != 0)) return 0;
  }
  out->p = UmmaGemmParams{(int)g.M, (int)g.N, (int)g.K, passes == 3 ? 3 : 1, b_mn, g.C, g.C_hi, g.C_lo, g.ldc, g.strideC,
                          g.alpha, g.beta, g.diag_add, g.bias, g.stride_bias, g.resid, g.add, g.add_lo, g.add_scale};
  return 1;
}

template <int BN, int KS>
static int launch_gemm_ks(const PreparedGemm& a, const PreparedGemm& b, int n_problems, int64_t batch, const int* ctrl,
                          int ctrl_index, cudaStream_t st, const NsCtrlEval& ev) {
  auto kern = umma_gemm_kernel<BN, KS>;
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    OTK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ug_smem<BN>()));
    attr_set[dev] = true;
  }
  dim3 grid((unsigned)(ceil_div(a.p.M, UG_BM) * KS), (unsigned)ceil_div(a.p.N, BN), (unsigned)(batch * n_problems));
  if (grid.y > 65535 || grid.z > 65535) return 0;
  static const bool pdl_on = [] { const char* e = getenv("OTK_GEMM_PDL"); return !(e && e[0] == '0'); }();   // tuning aid
  if (KS == 1 && !pdl_on) {
    kern<<<grid, UG_THREADS, ug_smem<BN>(), st>>>(a.maps, b.maps, a.p, b.p, (int)batch, ctrl, ctrl_index, ev);
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(UG_THREADS);
    cfg.dynamicSmemBytes = ug_smem<BN>();
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    unsigned n_attr = 0;
    if (KS > 1) {
      attr[n_attr].id = cudaLaunchAttributeClusterDimension;
      attr[n_attr].val.clusterDim.x = KS;
      attr[n_attr].val.clusterDim.y = 1;
      attr[n_attr].val.clusterDim.z = 1;
      ++n_attr;
    }
    if (pdl_on) {
      attr[n_attr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[n_attr].val.programmaticStreamSerializationAllowed = 1;
      ++n_attr;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n_attr;
    OTK_CUDA(cudaLaunchKernelEx(&cfg, kern, a.maps, b.maps, a.p, b.p, (int)batch, ctrl, ctrl_index, ev));
  }
  OTK_LAUNCH_CHECK();
  return 1;
}

// the persistent CTA-pair kernel: plane operands and plane results, no bias / beta / fp32 result, K below the long-K rule
static bool pair_eligible(const GemmArgs<float>& g, int64_t batch, int n_problems) {
  static const bool on = [] { const char* e = getenv("OTK_GEMM_PAIR"); return !(e && e[0] == '0'); }();   // tuning aid
  if (!on || !g.A_lo || !g.B_lo || g.C || !g.C_hi || !g.C_lo || g.bias || g.beta != 0.f) return false;
  if (g.M < 2 * UG_BM || g.N < G2_BN || g.K >= 4096) return false;
  const int64_t tiles128 = ceil_div(g.M, UG_BM) * ceil_div(g.N, 128) * batch * n_problems;
  // threshold measured on whole operators (B200): > 200 tiles of 128 x 128 (d = 2048: 6.8 -> 4.9 ms, 3 x 768: 1.66 -> 1.54,
  // 10 x 1024: 9.1 -> 7.0, 40 x 512: 6.1 -> 4.6 ms); at > 120 the single 1024^3 and 4 x 512^3 chains lose 15 %: below the
  // threshold the 128-wide (split-K) tiles occupy the machine better.  OTK_GEMM_PAIR_MIN overrides (tuning aid).
  static const int64_t min_tiles = [] { const char* e = getenv("OTK_GEMM_PAIR_MIN"); return e ? (int64_t)atoi(e) : (int64_t)200; }();
  return tiles128 > min_tiles;
}
static int launch_gemm_pair(const PreparedGemm& a, const PreparedGemm& b, int n_problems, int64_t batch, const int* ctrl,
                            int ctrl_index, cudaStream_t st, const NsCtrlEval& ev) {
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    OTK_CUDA(cudaFuncSetAttribute(umma_gemm_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM));
    attr_set[dev] = true;
  }
  const int64_t total = ceil_div(a.p.M, 2 * UG_BM) * ceil_div(a.p.N, G2_BN) * batch * n_problems;
  if (total > INT32_MAX / 2) return 0;
  const int64_t pairs = total < sm_count() / 2 ? total : sm_count() / 2;
  OTK_CUDA(launch_pdl(umma_gemm_pair_kernel, dim3((unsigned)(2 * pairs)), dim3(G2_THREADS), (size_t)G2_SMEM, st, 2, a.maps, b.maps,
                      a.p, b.p, n_problems, (int)batch, ctrl, ctrl_index, ev));
  OTK_LAUNCH_CHECK();
  return 1;
}

template <int BN>
static int launch_gemm(const PreparedGemm& a, const PreparedGemm& b, int n_problems, int64_t batch, int ks, const int* ctrl,
                       int ctrl_index, cudaStream_t st, const NsCtrlEval& ev = NsCtrlEval{}) {
  if constexpr (BN / 4 >= 32) { if (ks == 4) return launch_gemm_ks<BN, 4>(a, b, n_problems, batch, ctrl, ctrl_index, st, ev); }
  if (ks == 4) ks = 2;                            // a share is at least one 32-column chunk
  if constexpr (BN / 2 >= 32) { if (ks == 2) return launch_gemm_ks<BN, 2>(a, b, n_problems, batch, ctrl, ctrl_index, st, ev); }
  return launch_gemm_ks<BN, 1>(a, b, n_problems, batch, ctrl, ctrl_index, st, ev);
}

// Tile width and split-K factor.  Small products are bound by the bytes ONE CTA streams through its SM's L2 port,
// (128 + BN) * K / KS * 8: among the (BN, KS) whose grid still fits one wave pick the one that minimises it (at least
// two k-blocks per CTA); products that fill the machine anyway use 128-wide tiles without a split.
static void pick_tile(int64_t M, int64_t N, int64_t K, int64_t batch, int n_problems, int* bn_out, int* ks_out) {
  // The split-K cluster reduce-scatters the tile (each CTA finishes BN / KS columns from its peers' parked partials), so the
  // K loop, the distributed-shared-memory reads and the epilogue all shrink by KS.  It also shortens the truncating
  // tensor-memory accumulation chains (relative error 9e-7 instead of 3.6e-6 at K = 512).  OTK_GEMM_SPLITK=0 turns it off.
  static const bool split_ok = [] { const char* e = getenv("OTK_GEMM_SPLITK"); return !(e && e[0] == '0'); }();
  const int64_t sms = sm_count(), num_k = ceil_div(K, UG_BK);
  // Long contractions: tensor memory accumulates with truncation (relative error ~6e-8 per K = 8 step), so at K >= 4096 the
  // K loop is cut over a cluster of 4 CTAs whose partial tiles are added in fp32 registers (round to nearest) through
  // distributed shared memory - chains of K / 4.  Measured on the d = 4096 map: error against the fp64 engine 1.6e-3
  // un-split (outside the 1e-3 budget, so the whole call was redone in fp64: 1.9 s) and 4.4e-4 split (89 ms).  At
  // K = 2048 the un-split product is accurate enough (4.1e-4) and twice as fast (7.5 vs 14 ms per map), so it stays whole.
  static const bool long_k_split = [] { const char* e = getenv("OTK_GEMM_LONGK_SPLIT"); return !(e && e[0] == '0'); }();
  if (long_k_split && K >= 4096 && M >= UG_BM && N >= 128) { *bn_out = 128; *ks_out = 4; return; }
  static const int forced = [] { const char* e = getenv("OTK_GEMM_TILE"); return e ? atoi(e) : 0; }();   // tuning aid: BN * 10 + KS
  if (forced > 0 && K < 4096) {
    const int bn = forced / 10, ks = forced % 10;
    if ((bn == 64 || bn == 128) && (ks == 1 || ks == 2 || ks == 4) && bn / ks >= 32 && num_k >= 2 * ks) { *bn_out = bn; *ks_out = ks; return; }
  }
  static const bool narrow_ok = [] { const char* e = getenv("OTK_GEMM_BN32"); return !(e && e[0] == '0'); }();   // tuning aid
  int best_bn = 128, best_ks = 1;
  int64_t best_cost = INT64_MAX;
  for (int bn : {128, 64, 32}) {
    if (bn == 32 && (N > 256 || !narrow_ok)) continue;   // 32-wide tiles: only for the smallest products (one epilogue chunk per CTA)
    const int64_t ctas = ceil_div(M, UG_BM) * ceil_div(N, bn) * batch * n_problems;
    for (int ks : {1, 2, 4}) {
      if (ks > 1 && (!split_ok || num_k < 4 * ks || bn / ks < 32)) continue;
      if (ctas * ks > sms && !(bn == 128 && ks == 1)) continue;      // (128, 1) is the fallback for large products
      const int64_t cost = (int64_t)(UG_BM + bn) * ceil_div(num_k, ks);
      if (cost < best_cost) { best_cost = cost; best_bn = bn; best_ks = ks; }
    }
  }
  *bn_out = best_bn;
  *ks_out = best_ks;
}

int gemm_umma_try(const GemmArgs<float>& g, int64_t batch, int passes, cudaStream_t st) {
  int bn, ks;
  pick_tile(g.M, g.N, g.K, batch, 1, &bn, &ks);
  PreparedGemm a;
  int r = prepare_gemm(g, batch, passes, bn, st, &a);
  if (r <= 0) return r;
  if (bn == 32) return launch_gemm<32>(a, a, 1, batch, ks, nullptr, 0, st);
  return bn == 64 ? launch_gemm<64>(a, a, 1, batch, ks, nullptr, 0, st) : launch_gemm<128>(a, a, 1, batch, ks, nullptr, 0, st);
}

// two independent products of identical shape in one launch, optionally conditional on ctrl[0] > ctrl_index
int gemm_umma_dual(const GemmArgs<float>& g0, const GemmArgs<float>* g1, int64_t batch, const int* ctrl, int ctrl_index,
                   cudaStream_t st, const NsCtrlEval* eval) {
  const int n_problems = g1 ? 2 : 1;
  if (g1 && (g0.M != g1->M || g0.N != g1->N || g0.K != g1->K)) return 0;
  if (pair_eligible(g0, batch, n_problems) && (!g1 || pair_eligible(*g1, batch, n_problems))) {
    PreparedGemm a, b;
    int r = prepare_gemm(g0, batch, 3, 128, st, &a);
    if (r <= 0) return r;
    if (g1) { r = prepare_gemm(*g1, batch, 3, 128, st, &b); if (r <= 0) return r; } else b = a;
    if (a.p.b_mn_major == b.p.b_mn_major) return launch_gemm_pair(a, b, n_problems, batch, ctrl, ctrl_index, st, eval ? *eval : NsCtrlEval{});
  }
  int bn, ks;
  pick_tile(g0.M, g0.N, g0.K, batch, n_problems, &bn, &ks);
  PreparedGemm a, b;
  int r = prepare_gemm(g0, batch, 3, bn, st, &a);
  if (r <= 0) return r;
  if (g1) { r = prepare_gemm(*g1, batch, 3, bn, st, &b); if (r <= 0) return r; } else b = a;
  const NsCtrlEval ev = eval ? *eval : NsCtrlEval{};
  if (bn == 32) return launch_gemm<32>(a, b, n_problems, batch, ks, ctrl, ctrl_index, st, ev);
  return bn == 64 ? launch_gemm<64>(a, b, n_problems, batch, ks, ctrl, ctrl_index, st, ev)
                  : launch_gemm<128>(a, b, n_problems, batch, ks, ctrl, ctrl_index, st, ev);
}

}  // namespace otk
