// tcgen05.mma issue/execute rate probe: one thread issues REPS batches of 12 MMAs back to back (garbage operands).
#include <cstdio>
#include <cstdlib>
#include "../ot-vae-lightning_b200/csrc/otk_ptx.cuh"
using namespace otk::ptx;

__device__ __forceinline__ void umma_f16_ts_cg(int CG, uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (CG == 1) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_tf32_ss_cg2(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// mode: 0 = tf32 TS, B K-major; 1 = tf32 TS, B MN-major; 2 = tf32 SS (A, B K-major); 3 = f16 TS, B K-major
template <int CG>
__global__ void __launch_bounds__(128, 1) rate_kernel(int mode, int N, int reps, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x / 32;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;
  for (int i = threadIdx.x; i < 49152; i += blockDim.x) ((float*)smem)[i] = 1.0f;   // 192 KB of ones
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc_cg<CG>(&slot, 512); tmem_relinquish_cg<CG>(); }
  fence_proxy_async_smem();
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  if (threadIdx.x == 32 && rank == 0) {
    const uint32_t sb = smem_u32(smem);
    uint32_t idesc;
    if (mode == 3) idesc = idesc_f16(128 * CG, N);
    else idesc = idesc_tf32(128 * CG, N, 0, mode == 1 ? 1 : 0);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const uint32_t base = sb + (r % 4) * 32768;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint64_t b0, b1;
        if (mode == 1) { b0 = smem_desc_mn_tf32(base + kk * 1024, 4096); b1 = smem_desc_mn_tf32(base + 16384 + kk * 1024, 4096); }
        else { b0 = smem_desc_sw128(base + kk * 32, 16, 1024); b1 = smem_desc_sw128(base + 16384 + kk * 32, 16, 1024); }
        if (mode == 2) {
          const uint64_t a0 = smem_desc_sw128(sb + 131072 + kk * 32, 16, 1024), a1 = smem_desc_sw128(sb + 131072 + 16384 + kk * 32, 16, 1024);
          if (CG == 1) { umma_tf32(tb, a1, b0, idesc, 1); umma_tf32(tb, a0, b1, idesc, 1); umma_tf32(tb, a0, b0, idesc, 1); }
          else { umma_tf32_ss_cg2(tb, a1, b0, idesc, 1); umma_tf32_ss_cg2(tb, a0, b1, idesc, 1); umma_tf32_ss_cg2(tb, a0, b0, idesc, 1); }
        } else if (mode == 3) {
          umma_f16_ts_cg(CG, tb, tb + 256 + 32 + kk * 8, b0, idesc, 1); umma_f16_ts_cg(CG, tb, tb + 256 + kk * 8, b1, idesc, 1); umma_f16_ts_cg(CG, tb, tb + 256 + kk * 8, b0, idesc, 1);
        } else {
          umma_tf32_ts<CG>(tb, tb + 256 + 32 + kk * 8, b0, idesc, 1); umma_tf32_ts<CG>(tb, tb + 256 + kk * 8, b1, idesc, 1); umma_tf32_ts<CG>(tb, tb + 256 + kk * 8, b0, idesc, 1);
        }
      }
    }
    long long t1 = clock64();
    umma_commit_cg<CG>(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc_cg<CG>(tb, 512); }
}

template <int CG>
void run(const char* name, int mode, int N, int grid) {
  long long* out; cudaMalloc(&out, 16); cudaMemset(out, 0, 16);
  const int reps = 2000, smem = 200 * 1024;
  cudaFuncSetAttribute(rate_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaLaunchKernelEx(&cfg, rate_kernel<CG>, mode, N, reps, out);
  cudaEventRecord(e0);
  cudaLaunchKernelEx(&cfg, rate_kernel<CG>, mode, N, reps, out);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaGetLastError();
  const double mmas = reps * 12.0;
  const double kdepth = mode == 3 ? 16 : 8;
  const double flops = 2.0 * 128 * CG * N * kdepth * mmas * (grid / CG);
  printf("%-40s CG=%d N=%3d grid=%3d: issue %.1f clk/MMA, complete %.1f clk/MMA, %.3f ms, %.0f TFLOP/s %s\n", name, CG, N, grid, h[0] / mmas, h[1] / mmas, ms,
         flops / ms / 1e9, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  run<1>("tf32 TS B K-major", 0, 128, 148);
  run<1>("tf32 TS B K-major", 0, 256, 148);
  run<1>("tf32 TS B MN-major", 1, 128, 148);
  run<1>("tf32 SS K-major", 2, 128, 148);
  run<2>("tf32 TS B K-major", 0, 128, 148);
  run<2>("tf32 TS B K-major", 0, 256, 148);
  run<2>("tf32 TS B MN-major", 1, 128, 148);
  run<2>("tf32 TS B MN-major", 1, 256, 148);
  run<2>("tf32 SS K-major", 2, 256, 148);
  run<1>("f16 TS B K-major", 3, 128, 148);
  run<2>("f16 TS B K-major", 3, 256, 148);
  run<2>("f16 TS B K-major", 3, 128, 148);
  run<2>("tf32 TS B K-major (1 pair)", 0, 256, 2);
  run<2>("tf32 TS B MN-major (1 pair)", 1, 128, 2);
  return 0;
}
