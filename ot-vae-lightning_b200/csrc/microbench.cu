// Peak probes for the roofline denominators MEASURED_PEAKS.json does not hold ("builder must
// measure"): dense tcgen05 throughput of kind::tf32 and kind::f16 (operands resident in shared / tensor memory, no loads:
// the issue-bound ceiling of the MMA pipe) and the MUFU.EX2 rate (the binding unit of the fused Sinkhorn kernel).
// bench.py calls them once per run and reports the numbers next to the fractions they divide.
#include "otk_common.cuh"
#include "otk_ptx.cuh"

namespace otk {
using namespace ptx;

__device__ __forceinline__ void mb_umma_f16_ts_cg2(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d),
               "r"(a), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void mb_umma_tf32_ss_cg2(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
               "l"(a), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}

// CTA pairs (cta_group::2, M = 256, N = 256): one thread of the leader issues reps x 12 MMAs on whatever the shared /
// tensor memory holds.  kind 0: tf32, A and B from shared memory (K = 8 per MMA); kind 1: f16, A from tensor memory (K = 16).
__global__ void __launch_bounds__(128, 1) mb_mma_kernel(int kind, int reps) {
  extern __shared__ uint8_t mb_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)mb_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x / 32;
  const uint32_t rank = cluster_ctarank();
  for (int i = threadIdx.x; i < 49152; i += blockDim.x) ((float*)smem)[i] = 1.0f;   // 192 KB of ones
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc_cg<2>(&slot, 512); tmem_relinquish_cg<2>(); }
  fence_proxy_async_smem();
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tb = slot;
  if (threadIdx.x == 32 && rank == 0) {
    const uint32_t sb = smem_u32(smem);
    const uint32_t idesc = kind == 1 ? idesc_f16(256, 256) : idesc_tf32(256, 256, 0, 0);
    for (int r = 0; r < reps; ++r) {
      const uint32_t base = sb + (r % 4) * 32768;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint64_t b0 = smem_desc_sw128(base + kk * 32, 16, 1024), b1 = smem_desc_sw128(base + 16384 + kk * 32, 16, 1024);
        if (kind == 0) {
          const uint64_t a0 = smem_desc_sw128(sb + 131072 + kk * 32, 16, 1024), a1 = smem_desc_sw128(sb + 131072 + 16384 + kk * 32, 16, 1024);
          mb_umma_tf32_ss_cg2(tb, a1, b0, idesc, 1); mb_umma_tf32_ss_cg2(tb, a0, b1, idesc, 1); mb_umma_tf32_ss_cg2(tb, a0, b0, idesc, 1);
        } else {
          mb_umma_f16_ts_cg2(tb, tb + 256 + 32 + kk * 8, b0, idesc, 1); mb_umma_f16_ts_cg2(tb, tb + 256 + kk * 8, b1, idesc, 1);
          mb_umma_f16_ts_cg2(tb, tb + 256 + kk * 8, b0, idesc, 1);
        }
      }
    }
    umma_commit_cg<2>(&bar);
    mbar_wait(&bar, 0);
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) { tc_fence_after(); tmem_dealloc_cg<2>(tb, 512); }
}

// 8 independent ex2.approx chains per thread
__global__ void __launch_bounds__(256) mb_ex2_kernel(int reps, float seed, float* out) {
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = seed + 0.001f * (float)(threadIdx.x + i);
  for (int r = 0; r < reps; ++r) {
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = v[i] * 0.25f - 1.0f;      // keep the values in range (FMA pipe, not MUFU)
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
  if (s == 12345.678f) out[0] = s;                               // never true: keeps the chains alive
}

// 8 independent DFMA chains per thread (the fp64 engine of the matrix functions and the latency-mode statistics kernel)
__global__ void __launch_bounds__(256) mb_dfma_kernel(int reps, double seed, double* out) {
  double v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = seed + 1e-3 * (double)(threadIdx.x + i);
  const double a = 0.999999, b = 1e-7;
  for (int r = 0; r < reps; ++r) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = fma(v[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
  if (s == 12345.678) out[0] = s;
}

}  // namespace otk
using namespace otk;

// kind 0: tcgen05 kind::tf32 (TFLOP/s), 1: tcgen05 kind::f16 (TFLOP/s), 2: MUFU.EX2 (1e12 ex2/s), 3: DFMA (TFLOP/s).  Synchronous: times
// its own launches with CUDA events on `stream` (best of 3).
extern "C" int otk_microbench_peak(int kind, double* result_host, otk_stream_t stream) {
  OTK_TRY(require_device());
  OTK_REQUIRE(result_host && kind >= 0 && kind <= 3, "microbench_peak: bad arguments");
  cudaStream_t st = as_stream(stream);
  cudaEvent_t e0, e1;
  OTK_CUDA(cudaEventCreate(&e0));
  OTK_CUDA(cudaEventCreate(&e1));
  const int sms = sm_count();
  double best = 0.0;
  float* sink = nullptr;
  OTK_CUDA(cudaMalloc(&sink, 256));
  for (int rep = 0; rep < 4; ++rep) {
    float ms = 0.f;
    double work = 0.0;
    if (kind <= 1) {
      const int reps = 4000, smem = 200 * 1024, grid = sms / 2 * 2;
      OTK_CUDA(cudaFuncSetAttribute(mb_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.stream = st;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      OTK_CUDA(cudaEventRecord(e0, st));
      OTK_CUDA(cudaLaunchKernelEx(&cfg, mb_mma_kernel, kind, reps));
      OTK_CUDA(cudaEventRecord(e1, st));
      work = 2.0 * 256 * 256 * (kind == 1 ? 16 : 8) * 12.0 * reps * (grid / 2);        // flop
    } else if (kind == 3) {
      const int reps = 4000, blocks = sms * 8;
      OTK_CUDA(cudaEventRecord(e0, st));
      mb_dfma_kernel<<<blocks, 256, 0, st>>>(reps, 0.5, reinterpret_cast<double*>(sink));
      OTK_CUDA(cudaEventRecord(e1, st));
      work = 2.0 * 8.0 * reps * 256.0 * blocks;                                        // flop
    } else {
      const int reps = 20000, blocks = sms * 8;
      OTK_CUDA(cudaEventRecord(e0, st));
      mb_ex2_kernel<<<blocks, 256, 0, st>>>(reps, 0.5f, sink);
      OTK_CUDA(cudaEventRecord(e1, st));
      work = 8.0 * reps * 256.0 * blocks;                                              // ex2
    }
    OTK_CUDA(cudaEventSynchronize(e1));
    OTK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    count_launch(1);
    if (rep > 0 && ms > 0.f) { const double r = work / (ms * 1e-3) / 1e12; if (r > best) best = r; }
  }
  cudaFree(sink);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  OTK_CUDA(cudaGetLastError());
  *result_host = best;
  return OTK_OK;
}
