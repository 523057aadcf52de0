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
#include <cuda_fp16.h>

#include "otk_ptx.cuh"
#include "stats_umma.cuh"
#include "tensormap.cuh"

namespace otk {

constexpr int SH_T = 128, SH_BK = 64;
// Ring depths.  The raw ring and the A ring are as deep as there are converter sets, so a raw stage and an A slot belong
// to ONE set and their barriers advance by exactly one phase per tile of that set (a parity wait cannot tell "two phases
// behind" from "done").  The two-deep B ring is shared by the sets; see the wait order in the converter.
constexpr int SH_SETS = 3;
constexpr int SH_XS = SH_SETS, SH_BS = 2, SH_AS = SH_SETS, SH_ACC = 2;
constexpr int SH_THREADS = (2 + 4 * SH_SETS + 4) * 32;   // TMA, MMA | SH_SETS x 4 converter warps | 4 epilogue warps
constexpr int SH_SLAB = 32 * SH_BK * 4;              // 8 KiB: [64 rows x 32 features] fp32
constexpr int SH_RAW = 4 * SH_SLAB;                  // 32 KiB
constexpr int SH_BPLANE = SH_T * SH_BK * 2;          // 16 KiB: [128 features x 64 rows] fp16
constexpr int SH_BSTAGE = 2 * SH_BPLANE;             // hi + lo
constexpr int SH_SACC = SH_T * SH_T * 4;             // 64 KiB second-stage accumulators [column][row]
constexpr int SH_SUB = 1024;                         // rows accumulated in tensor memory per sub-chunk
constexpr int SH_ACOL0 = 256;                        // TMEM columns [0,256): two accumulators, [256,448): A ring (3 x 64)
constexpr int SH_SMEM = SH_XS * SH_RAW + SH_BS * SH_BSTAGE + SH_SACC + 1024 + 512;
constexpr float SH_LIMIT = 32768.f;                  // |x'| above this raises the overflow flag (FP16 max = 65504)

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

__global__ void __launch_bounds__(SH_THREADS, 1)
stats_h_kernel(const __grid_constant__ CUtensorMap mapX, const float* __restrict__ pivot, const float* __restrict__ scale,
               int rows, int dim, int parts, int range_len, double* __restrict__ ws_cov, double* __restrict__ ws_sum,
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
#pragma unroll
          for (int kk = 0; kk < SH_BK / 16; ++kk) {
            const uint64_t b_hi = smem_desc_sw128(bb + kk * 32, 16, 1024);
            const uint64_t b_lo = smem_desc_sw128(bb + SH_BPLANE + kk * 32, 16, 1024);
            umma_f16_ts(acc, ab + 32 + kk * 8, b_hi, idesc, !(first && kk == 0));   // lo * hi
            umma_f16_ts(acc, ab + kk * 8, b_lo, idesc, 1);                           // hi * lo
            umma_f16_ts(acc, ab + kk * 8, b_hi, idesc, 1);                           // hi * hi
          }
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
    double colsum = 0.0;
    float maxabs = 0.f;
    for (int it = cset; it < num_k; it += SH_SETS) {
      const int sx = it % SH_XS, sa = it % SH_AS, sb = it % SH_BS;
      const int valid = in ? min(SH_BK, r1 - (r0 + it * SH_BK)) : 0;     // rows past the range end contribute nothing
      // Order matters: a set's consecutive tiles are SH_SETS apart, more than one phase of the two-deep B ring.  Once the
      // MMAs of this set's previous tile (it - SH_SETS, same A slot) have retired, tile it-2 is the only user of the B
      // stage that can still be pending, i.e. empty_b is at most one phase behind.
      mbar_wait(&empty_a[sa], ((it / SH_AS) & 1) ^ 1);
      tc_fence_after();
      mbar_wait(&full_a[sx], (it / SH_XS) & 1);
      mbar_wait(&empty_b[sb], ((it / SH_BS) & 1) ^ 1);
      float part_sum = 0.f;
      const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + SH_ACOL0 + sa * 64;
      const uint32_t hb = brow + sb * SH_BSTAGE;
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {                                   // two halves of 32 rows
        uint32_t hw[16], lw[16];
#pragma unroll
        for (int p = 0; p < 16; ++p) {
          const int ra = h2 * 32 + 2 * p, rb = ra + 1;
          float xa0 = lds32(xbase + sx * SH_RAW + (uint32_t)ra * 128 + ((cc ^ (uint32_t)(ra & 3)) * 32) + within);
          float xa1 = lds32(xbase + sx * SH_RAW + (uint32_t)rb * 128 + ((cc ^ (uint32_t)(rb & 3)) * 32) + within);
          xa0 = ra < valid ? xa0 - c : 0.f;
          xa1 = rb < valid ? xa1 - c : 0.f;
          part_sum += xa0 + xa1;
          const float v0 = xa0 * s, v1 = xa1 * s;
          maxabs = fmaxf(maxabs, fmaxf(fabsf(v0), fabsf(v1)));
          const __half2 h = __floats2half2_rn(v0, v1);                  // .x (low half) = the even row
          const float2 hf = __half22float2(h);
          const __half2 lo = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
          hw[p] = *reinterpret_cast<const uint32_t*>(&h);
          lw[p] = *reinterpret_cast<const uint32_t*>(&lo);
        }
        tmem_st16u(ta + h2 * 16, hw);
        tmem_st16u(ta + 32 + h2 * 16, lw);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {                                 // 16-byte chunks (8 rows each) of this half
          const uint32_t off = (((uint32_t)(h2 * 4 + ch)) ^ sw) * 16;
          sts128u(hb + off, hw[4 * ch], hw[4 * ch + 1], hw[4 * ch + 2], hw[4 * ch + 3]);
          sts128u(hb + SH_BPLANE + off, lw[4 * ch], lw[4 * ch + 1], lw[4 * ch + 2], lw[4 * ch + 3]);
        }
      }
      colsum += (double)part_sum;
      fence_proxy_async_smem();   // generic-proxy writes of the B planes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) { mbar_arrive(&empty_ra[sx]); mbar_arrive(&ready_b[sb]); }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ready_a[sa]);
    }
    if (in && num_k > 0) atomicAdd(&ws_sum[(int64_t)l * dim + col], colsum);
    if (!(maxabs < SH_LIMIT)) atomicOr(overflow, 1);                    // also catches NaN / inf inputs
  } else {
    // ===== epilogue: 4 warps (warp <-> TMEM lane quarter); second-stage fp32 accumulation in shared memory =====
    const int q = warp % 4;
    const uint32_t srow = smem_u32(sacc) + (uint32_t)(q * 32 + lane) * 4;
    for (int sub = 0; sub < num_sub; ++sub) {
      const int a = sub % SH_ACC;
      mbar_wait(&acc_full[a], (sub / SH_ACC) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + a * SH_T;
#pragma unroll 1
      for (int c0 = 0; c0 < SH_T; c0 += 32) {
        float v[32];
        tmem_ld32(taddr + c0, v);
        tmem_ld_wait();
        if (c0 + 32 == SH_T) {                                           // accumulator fully read: hand it back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[a]);
        }
        if (sub == 0) {
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) sts32(srow + (uint32_t)(c0 + jj) * (SH_T * 4), v[jj]);
        } else {
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) {
            const uint32_t ad = srow + (uint32_t)(c0 + jj) * (SH_T * 4);
            sts32(ad, lds32(ad) + v[jj]);
          }
        }
      }
    }
    // flush: P'[gi][gj] for gi <= gj, stored TRANSPOSED (ws[gj][gi]) so that the 32 lanes of an atomic instruction hit
    // 32 consecutive doubles (the merge kernel reads the transposed position); the power-of-two scales divide out exactly
    const int gi = q * 32 + lane;
    if (num_k > 0 && gi < dim) {
      double* cov = ws_cov + (int64_t)l * dim * dim;
      const float* sc = scale + (int64_t)l * dim;
      const double inv_i = 1.0 / (double)sc[gi];
#pragma unroll 4
      for (int gj = 0; gj < dim; ++gj)
        if (gi <= gj) atomicAdd(&cov[(int64_t)gj * dim + gi], (double)lds32(srow + (uint32_t)gj * (SH_T * 4)) * (inv_i / (double)sc[gj]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// pivot[l, f] = mean of the first min(rows, 64) latents; scale[l, f] = power of two mapping the largest deviation from
// the pivot seen in those rows into [64, 128)  (1 if the feature is constant there).  Block = 32 features x 8 row groups.
__global__ void pivot_scale_kernel(const float* __restrict__ x, int64_t rows, int64_t dim, int64_t row_stride,
                                   int64_t batch_stride, float* __restrict__ pivot, float* __restrict__ scale) {
  __shared__ float part[8][33];
  __shared__ float piv[32];
  const int64_t l = blockIdx.y;
  const int64_t col = blockIdx.x * 32 + threadIdx.x;
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
  if (*flag == 0) return;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) p[e] = 0.0;
}

size_t stats_h_extra_workspace(int64_t L, int64_t dim) { return align_up((size_t)L * dim * 4, 256) + 256; }

bool stats_h_eligible(int64_t L, int64_t rows, int64_t dim) { return dim <= SH_T && rows >= 1 && L >= 1; }

// Launches pivot/scale + the FP16-split kernel.  *flag_out (device int) is non-zero afterwards iff a value left the FP16
// range, in which case the staging area is invalid and the caller must re-run with the TF32 kernel (stats_umma.cu).
int stats_h_launch(const float* x, int64_t L, int64_t rows, int64_t dim, int64_t row_stride, int64_t batch_stride,
                   float* pivot, double* ws_cov, double* ws_sum, Arena& ar, cudaStream_t st, int** flag_out) {
  float* scale = ar.take<float>((size_t)L * dim);
  int* flag = ar.take<int>(16);
  if (!ar.ok()) return OTK_ERR_WORKSPACE;
  OTK_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), st));
  pivot_scale_kernel<<<dim3((unsigned)ceil_div(dim, 32), (unsigned)L), dim3(32, 8), 0, st>>>(x, rows, dim, row_stride, batch_stride,
                                                                                             pivot, scale);
  OTK_LAUNCH_CHECK();
  CUtensorMap mX;
  if (!encode_map_f32_3d(&mX, x, dim, rows, L, row_stride, batch_stride, 32, SH_BK, /*atom32=*/true)) return 0;
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
  if (L * parts > INT32_MAX) return 0;
  stats_h_kernel<<<(unsigned)(L * parts), SH_THREADS, SH_SMEM, st>>>(mX, pivot, scale, (int)rows, (int)dim, (int)parts,
                                                                    (int)range_len, ws_cov, ws_sum, flag);
  OTK_LAUNCH_CHECK();
  *flag_out = flag;
  return 1;
}

int stats_zero_if(double* p, int64_t n, const int* flag, cudaStream_t st) {
  int64_t blocks = ceil_div(n, 256);
  if (blocks > 1024) blocks = 1024;
  zero_if_kernel<<<(unsigned)blocks, 256, 0, st>>>(p, n, flag);
  OTK_LAUNCH_CHECK();
  return OTK_OK;
}

}  // namespace otk
