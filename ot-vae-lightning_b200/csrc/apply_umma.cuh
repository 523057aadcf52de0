// tcgen05 engine hook for apply_transport (filled in by apply_umma.cu).
#pragma once
#include "otk_common.cuh"
namespace otk {
// 1 = handled, 0 = not eligible, <0 = error
int apply_umma_try(const float* x, int64_t L, int64_t rows, int64_t dim, const float* ms32, const float* mt32,
                   const float* T32, float* Thi_scratch, float* Tlo_scratch, float* y, cudaStream_t st);
}
