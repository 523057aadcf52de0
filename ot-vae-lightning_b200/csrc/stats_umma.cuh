// tcgen05 engine hook for the streaming statistics (filled in by stats_umma.cu).
#pragma once
#include "otk_common.cuh"
namespace otk {
size_t stats_umma_extra_workspace(int64_t L, int64_t dim);
// returns 1 if the tcgen05 kernel handled the update (and sets *tile to its output tile size and *pivot_out to the
// per-feature pivot: the staging area then holds P' = sum (x-c)(x-c)^T TRANSPOSED (element (i <= j) at [j][i]) and
// S' = sum (x-c)), 0 if the shape is not eligible, <0 on error.
int stats_umma_try(const float* x, int64_t L, int64_t rows, int64_t dim, int64_t row_stride, int64_t batch_stride,
                   double* ws_cov, double* ws_sum, Arena& ar, cudaStream_t st, int* tile, const float** pivot_out);
// FP16-split engine for dim <= 128 (stats_h.cu)
size_t stats_h_extra_workspace(int64_t L, int64_t dim);
bool stats_h_eligible(int64_t L, int64_t rows, int64_t dim);
int stats_h_launch(const float* x, int64_t L, int64_t rows, int64_t dim, int64_t row_stride, int64_t batch_stride,
                   float* pivot, double* ws_cov, double* ws_sum, Arena& ar, cudaStream_t st, int** flag_out);
int stats_zero_if(double* p, int64_t n, const int* flag, cudaStream_t st);
// same scheme on the 128 x 128 blocks of the upper block triangle for dim > 128 (stats_h.cu)
size_t stats_h2_extra_workspace(int64_t L, int64_t dim);
bool stats_h2_eligible(int64_t L, int64_t rows, int64_t dim);
int stats_h2_launch(const float* x, int64_t L, int64_t rows, int64_t dim, int64_t row_stride, int64_t batch_stride,
                    float* pivot, double* ws_cov, double* ws_sum, Arena& ar, cudaStream_t st, int** flag_out);
}  // namespace otk
