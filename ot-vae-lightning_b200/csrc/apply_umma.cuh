// tcgen05 engine hook for apply_transport (filled in by apply_umma.cu).
#pragma once
#include "otk_common.cuh"
namespace otk {
// 1 = handled, 0 = not eligible, <0 = error
// `ar`: remaining workspace (apply_h_workspace_bytes) for the FP16-split kernel, which runs first when eligible; the TF32
// kernel is then only a device-gated fallback
int apply_umma_try(const float* x, int64_t L, int64_t rows, int64_t dim, const float* ms32, const float* mt32,
                   const float* T32, float* Thi_scratch, float* Tlo_scratch, float* y, Arena& ar, cudaStream_t st);
// FP16-split engine (apply_h.cu)
size_t apply_h_workspace_bytes(int64_t L, int64_t dim);
bool apply_h_eligible(int64_t L, int64_t rows, int64_t dim);
int apply_h_try(const float* x, int64_t L, int64_t rows, int64_t dim, const float* ms32, const float* mt32, const float* T32,
                float* y, Arena& ar, bool pair, cudaStream_t st, int** flag_out);
}
