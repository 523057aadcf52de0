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
int apply_h_prepare(int64_t L, int64_t dim, const float* ms32, const float* T32, const float* var32, Arena& ar, cudaStream_t st);
// x_row_stride / x_batch_stride (elements; 0 = contiguous [L, rows, dim]): the latents may be a strided view
int apply_h_run_prepared(const float* x, int64_t L, int64_t rows, int64_t dim, const float* mt32, float* y, Arena& ar, bool pair,
                         cudaStream_t st, int** flag_out, int64_t x_row_stride = 0, int64_t x_batch_stride = 0);
// TF32 kernel on planes that are already split (prepared form); run_flag as in apply_umma_try
int apply_umma_run_planes(const float* x, int64_t L, int64_t rows, int64_t dim, const float* ms32, const float* mt32,
                          const float* Thi, const float* Tlo, float* y, bool pair, const int* run_flag, cudaStream_t st,
                          int64_t x_row_stride = 0, int64_t x_batch_stride = 0);
void apply_umma_split(const float* T32, int64_t n, float* Thi, float* Tlo, cudaStream_t st);
bool apply_umma_eligible(const float* x, const float* y, int64_t L, int64_t rows, int64_t dim);
bool apply_umma_pair(int64_t rows, int64_t dim);
int apply_h_try(const float* x, int64_t L, int64_t rows, int64_t dim, const float* ms32, const float* mt32, const float* T32,
                float* y, Arena& ar, bool pair, cudaStream_t st, int** flag_out);
}
