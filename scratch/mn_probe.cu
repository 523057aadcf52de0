// Probe of UMMA MN-major descriptor semantics (tf32, SW128): one 128x128x32 product, B stored N-major.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../ot-vae-lightning_b200/csrc/otk_ptx.cuh"
using namespace otk::ptx;

struct Variant { uint32_t lbo, sbo, kstep; int a_mn; const char* name; };
__device__ __forceinline__ uint64_t desc_b32(uint32_t addr, uint32_t lbo, uint32_t sbo) { uint64_t d = smem_desc_sw128(addr, lbo, sbo); d &= ~((uint64_t)7 << 61); d |= (uint64_t)1 << 61; return d; }

__global__ void probe(const float* A, const float* B, float* C, uint32_t lbo, uint32_t sbo, uint32_t kstep, int a_mn) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem; uint8_t* sB = smem + 16384;
  uint64_t* bar = (uint64_t*)(smem + 32768);
  uint32_t* slot = (uint32_t*)(bar + 1);
  int tid = threadIdx.x;
  // A: element (m,k), m<128, k<32
  for (int e = tid; e < 128 * 32; e += blockDim.x) {
    int m = e / 32, k = e % 32;
    uint32_t off;
    if (!a_mn) off = m * 128 + (((k / 4) ^ (m % 8)) * 16) + (k % 4) * 4;           // K-major SW128
    else { int slab = m / 32, chunk = (m % 32) / 8; off = slab * 4096 + k * 128 + ((chunk ^ (k % 4)) * 32) + (m % 8) * 4; }
    *(float*)(sA + off) = A[m * 32 + k];
  }
  for (int e = tid; e < 128 * 32; e += blockDim.x) {
    int n = e / 32, k = e % 32;
    int slab = n / 32, chunk = (n % 32) / 8;
    uint32_t off = slab * 4096 + k * 128 + ((chunk ^ (k % 4)) * 32) + (n % 8) * 4;   // N-major, 32-wide slabs of 32 k rows
    *(float*)(sB + off) = B[n * 32 + k];
  }
  if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (tid < 32) { tmem_alloc(slot, 128); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  uint32_t tm = *slot;
  if (tid == 0) {
    uint32_t idesc = idesc_tf32(128, 128, a_mn, 1);
    for (int kk = 0; kk < 4; ++kk) {
      uint64_t ad = a_mn ? desc_b32(smem_u32(sA) + kk * kstep, lbo, sbo) : smem_desc_sw128(smem_u32(sA) + kk * 32, 16, 1024);
      uint64_t bd = desc_b32(smem_u32(sB) + kk * kstep, lbo, sbo);
      umma_tf32(tm, ad, bd, idesc, kk != 0);
    }
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  int warp = tid / 32, lane = tid % 32;
  for (int c0 = 0; c0 < 128; c0 += 32) {
    float v[32];
    tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) C[(warp * 32 + lane) * 128 + c0 + j] = v[j];
  }
  tc_fence_before(); __syncthreads();
  if (tid < 32) tmem_dealloc(tm, 128);
}

int main() {
  std::vector<float> A(128 * 32), B(128 * 32), C(128 * 128), R(128 * 128);
  srand(1);
  for (auto& x : A) x = (float)(rand() % 17 - 8);      // small integers: exact in tf32
  for (auto& x : B) x = (float)(rand() % 13 - 6);
  for (int m = 0; m < 128; ++m) for (int n = 0; n < 128; ++n) { float s = 0; for (int k = 0; k < 32; ++k) s += A[m * 32 + k] * B[n * 32 + k]; R[m * 128 + n] = s; }
  float *dA, *dB, *dC;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dC, C.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  Variant vs[] = {{4096, 512, 1024, 0, "B mn base32: lbo=4096 sbo=512 kstep=1024"},
                  {4096, 1024, 1024, 0, "B mn base32: lbo=4096 sbo=1024"},
                  {512, 4096, 1024, 0, "B mn base32: lbo=512 sbo=4096 (swapped)"},
                  {4096, 512, 1024, 1, "A+B mn base32: lbo=4096 sbo=512"},
                  {1024, 4096, 1024, 0, "B mn base32: lbo=1024 sbo=4096"}};
  for (auto& v : vs) {
    cudaMemset(dC, 0xff, C.size() * 4);
    probe<<<1, 128, 40000>>>(dA, dB, dC, v.lbo, v.sbo, v.kstep, v.a_mn);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s -> CUDA error %s\n", v.name, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, nz = 0; int bad = 0;
    for (size_t i = 0; i < C.size(); ++i) { double d = fabs((double)C[i] - R[i]); if (!(d < 1e-3)) ++bad; if (d > maxerr) maxerr = d; if (C[i] != 0) nz++; }
    printf("%-48s maxerr %.3g  wrong %d / 16384  nonzero %.0f  C[0,0..3]=%g %g %g %g  ref=%g %g %g %g\n", v.name, maxerr, bad, nz,
           C[0], C[1], C[2], C[3], R[0], R[1], R[2], R[3]);
  }
  return 0;
}
