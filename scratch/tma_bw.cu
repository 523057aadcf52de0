// TMA streaming micro-benchmark: each CTA streams [box_rows x 32 fp32] boxes through an mbarrier ring (no math).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../ot-vae-lightning_b200/csrc/otk_ptx.cuh"
#include "../ot-vae-lightning_b200/csrc/tensormap.cuh"
using namespace otk::ptx;

template <int STAGES, int LOADS>   // LOADS boxes of 16 KB per stage
__global__ void __launch_bounds__(64, 1) stream_kernel(const __grid_constant__ CUtensorMap map, int rows, int cols, int iters_total,
                                                       unsigned long long* sink) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + STAGES * LOADS * 16384);
  uint64_t* empty = full + STAGES;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (threadIdx.x == 0) { for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); } fence_barrier_init(); }
  __syncthreads();
  const int row_tiles = rows / 128, col_tiles = cols / 32;
  const int tiles = row_tiles * col_tiles / LOADS;
  const int total = tiles * (iters_total > 0 ? iters_total : 1);
  if (warp == 0 && lane == 0) {
    int it = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
      const int s = it % STAGES;
      mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
      mbar_arrive_expect_tx(&full[s], LOADS * 16384);
      for (int j = 0; j < LOADS; ++j) {
        const int box = (t % tiles) * LOADS + j;
        const int rt = box / col_tiles, ct = box % col_tiles;     // consecutive boxes walk along K (columns) of a row tile
        tma_load_3d(smem + (s * LOADS + j) * 16384, &map, ct * 32, rt * 128, 0, &full[s]);
      }
    }
  } else if (warp == 1 && lane == 0) {
    int it = 0;
    unsigned long long acc = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
      const int s = it % STAGES;
      mbar_wait(&full[s], (it / STAGES) & 1);
      acc += *(volatile unsigned*)(smem + s * LOADS * 16384);
      mbar_arrive(&empty[s]);
    }
    if (acc == 0x1234567) *sink = acc;
  }
}

template <int STAGES, int LOADS>
void run(const char* name, float* d, int rows, int cols, int grid, int reps = 1) {
  CUtensorMap m;
  if (!otk::encode_map_f32_3d(&m, d, cols, rows, 1, cols, (int64_t)rows * cols, 32, 128)) { printf("encode failed\n"); return; }
  unsigned long long* sink; cudaMalloc(&sink, 8);
  int smem = STAGES * LOADS * 16384 + 2048;
  cudaFuncSetAttribute(stream_kernel<STAGES, LOADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  stream_kernel<STAGES, LOADS><<<grid, 64, smem>>>(m, rows, cols, reps, sink);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < 5; ++r) stream_kernel<STAGES, LOADS><<<grid, 64, smem>>>(m, rows, cols, reps, sink);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  cudaError_t e = cudaGetLastError();
  double bytes = (double)rows * cols * 4 * reps;
  printf("%-34s rows=%7d cols=%4d stages=%d loads=%d grid=%4d : %8.3f ms  %7.1f GB/s  %s\n", name, rows, cols, STAGES, LOADS, grid, ms,
         bytes / ms / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(sink);
}

int main() {
  float* d; size_t n = (size_t)1 << 28;  // 1 GiB
  cudaMalloc(&d, n * 4); cudaMemset(d, 0, n * 4);
  run<3, 1>("DRAM, d=512, 1 box/stage", d, 1 << 19, 512, 148);
  run<3, 2>("DRAM, d=512, 2 box/stage", d, 1 << 19, 512, 148);
  run<6, 2>("DRAM, d=512, 2 box/stage x6", d, 1 << 19, 512, 148);
  run<3, 4>("DRAM, d=512, 4 box/stage", d, 1 << 19, 512, 148);
  run<3, 4>("DRAM, d=128, 4 box/stage", d, 1 << 21, 128, 148);
  run<3, 4>("L2 (32 MB) x50, d=512, 4 box", d, 1 << 14, 512, 148, 50);
  run<3, 2>("L2 (32 MB) x50, d=512, 2 box", d, 1 << 14, 512, 148, 50);
  run<6, 2>("L2 (32 MB) x50, d=512, 2 box x6", d, 1 << 14, 512, 148, 50);
  run<3, 4>("L2 (64 MB) x50, d=512, 4 box", d, 1 << 15, 512, 148, 50);
  run<3, 4>("L2 (8 MB) x200, d=512, 4 box", d, 1 << 12, 512, 148, 200);
  run<3, 4>("L2 (1 MB) x1600, d=512, 4 box", d, 1 << 9, 512, 148, 1600);
  run<3, 4>("L2 (32 MB), d=128, 4 box/stage", d, 1 << 16, 128, 148);
  run<6, 2>("L2 (32 MB), d=512, 2 box x6", d, 1 << 14, 512, 148);
  run<3, 4>("DRAM d=512 grid 296 (2 CTA/SM)", d, 1 << 19, 512, 296);
  run<3, 4>("DRAM d=512 grid 444 (3 CTA/SM)", d, 1 << 19, 512, 444);
  return 0;
}
