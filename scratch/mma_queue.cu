// how deep is the tcgen05.mma issue queue, and what do commit / try_wait between MMA batches cost the tensor pipe?
#include <cstdio>
#include "../ot-vae-lightning_b200/csrc/otk_ptx.cuh"
using namespace otk::ptx;
__global__ void __launch_bounds__(128, 1) q_kernel(int variant, int reps, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar, dummy[4], done;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x / 32;
  for (int i = threadIdx.x; i < 32768; i += blockDim.x) ((float*)smem)[i] = 1.0f;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&done, 1); for (int i = 0; i < 4; ++i) mbar_init(&dummy[i], 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem(); tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = slot;
  if (threadIdx.x == 32) {
    const uint32_t sb = smem_u32(smem);
    const uint32_t idesc = idesc_tf32(128, 128, 0, 0);
    const uint64_t b0 = smem_desc_sw128(sb, 16, 1024);
    if (variant < 0) {   // queue depth: timestamps after each issue
      long long t0 = clock64();
      for (int i = 0; i < 40; ++i) { umma_tf32_ts<1>(tb, tb + 256, b0, idesc, 1); out[i] = clock64() - t0; }
    } else {
      long long t0 = clock64();
      for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int k = 0; k < 12; ++k) umma_tf32_ts<1>(tb, tb + 256 + (k % 4) * 8, b0, idesc, 1);
        if (variant >= 1) umma_commit(&dummy[r & 1]);
        if (variant == 1 || variant == 2) umma_commit(&dummy[2 + (r & 1)]);
        if (variant >= 2) { mbar_wait(&done, 1); }          // parity 1 of a fresh barrier: already complete
        if (variant == 2) { mbar_wait(&done, 1); }
        if (variant >= 2) tc_fence_after();
      }
      out[0] = clock64() - t0;
    }
    umma_commit(&bar); mbar_wait(&bar, 0);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 512); }
}
int main() {
  long long* out; cudaMalloc(&out, 64 * 8);
  cudaFuncSetAttribute(q_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024);
  long long h[64];
  q_kernel<<<1, 128, 140 * 1024>>>(-1, 0, out); cudaMemcpy(h, out, 40 * 8, cudaMemcpyDeviceToHost);
  printf("issue timestamps: "); for (int i = 0; i < 40; ++i) printf("%lld ", h[i]); printf("\n");
  const char* names[] = {"12 MMA", "12 MMA + 2 commit", "12 MMA + 2 commit + 2 wait + fence", "12 MMA + 1 commit + 1 wait + fence"};
  for (int v = 0; v < 4; ++v) {
    q_kernel<<<1, 128, 140 * 1024>>>(v, 1000, out); cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost);
    printf("%-40s: %.1f clk per batch (ideal 768)  %s\n", names[v], h[0] / 1000.0, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
