// is the ~56 clk tcgen05.mma issue cost per thread or per SM?  1 vs 2 issuing threads (different warps), small-N MMAs.
#include <cstdio>
#include "../ot-vae-lightning_b200/csrc/otk_ptx.cuh"
using namespace otk::ptx;
__global__ void __launch_bounds__(128, 1) k(int nthreads, int N, int reps, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar[2];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x / 32;
  for (int i = threadIdx.x; i < 32768; i += blockDim.x) ((float*)smem)[i] = 1.0f;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem(); tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = slot;
  if (nthreads == 3) {   // warp-uniform issue: the whole warp runs the loop, one elected lane issues
    if (warp == 1) {
      const uint32_t sb = smem_u32(smem);
      const uint32_t idesc = idesc_tf32(128, N, 0, 0);
      const uint64_t b0 = smem_desc_sw128(sb, 16, 1024);
      long long t0 = clock64();
      for (int r = 0; r < reps; ++r) { if (elect_one()) umma_tf32_ts<1>(tb, tb + 256, b0, idesc, 1); __syncwarp(); }
      long long t1 = clock64();
      if (elect_one()) umma_commit(&bar[0]);
      __syncwarp();
      mbar_wait(&bar[0], 0);
      long long t2 = clock64();
      if (threadIdx.x == 32) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
  } else if ((threadIdx.x == 32 || (threadIdx.x == 64 && nthreads == 2))) {
    const int me = threadIdx.x == 32 ? 0 : 1;
    const uint32_t sb = smem_u32(smem);
    const uint32_t idesc = idesc_tf32(128, N, 0, 0);
    const uint64_t b0 = smem_desc_sw128(sb, 16, 1024);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) umma_tf32_ts<1>(tb + me * 128, tb + 256, b0, idesc, 1);
    long long t1 = clock64();
    umma_commit(&bar[me]); mbar_wait(&bar[me], 0);
    long long t2 = clock64();
    out[me * 2] = t1 - t0; out[me * 2 + 1] = t2 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 512); }
}
int main() {
  long long* out; cudaMalloc(&out, 64);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024);
  long long h[4];
  for (int N : {16, 32, 64, 128}) for (int nt = 1; nt <= 3; ++nt) {
    cudaMemset(out, 0, 64);
    k<<<1, 128, 140 * 1024>>>(nt, N, 4000, out); cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost);
    printf("N=%3d threads=%d: thread0 issue %.1f / done %.1f clk per MMA; thread1 issue %.1f / done %.1f  (exec floor %d)  %s\n", N, nt, h[0] / 4000.0, h[1] / 4000.0,
           h[2] / 4000.0, h[3] / 4000.0, N / 2, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
