// tcgen05 engine hook for the streaming statistics (filled in by stats_umma.cu).
#pragma once
#include "otk_common.cuh"
namespace otk {
// the caller's running buffers and the accumulate / EMA rule (decay < 0: plain accumulation)
struct StatsRunning { void *n_obs, *sum, *sum_cov; int n_dtype, buf_dtype; double decay; };
size_t stats_umma_extra_workspace(int64_t L, int64_t dim);
// returns 1 if the tcgen05 kernel handled the update (and sets *tile to its output tile size and *pivot_out to the
// per-feature pivot: the staging area then holds P' = sum (x-c)(x-c)^T TRANSPOSED (element (i <= j) at [j][i]) and
// S' = sum (x-c)), 2 if it also merged the result into the running buffers `run` (FP16-split kernels: nothing is left to
// do), 0 if the shape is not eligible, <0 on error.  The staging area need NOT be zeroed by the caller for 1 and 2; for 0
// the caller zeroes it before running its own engine.
int stats_umma_try(const float* x, int64_t L, int64_t rows, int64_t dim, int64_t row_stride, int64_t batch_stride,
                   double* ws_cov, double* ws_sum, Arena& ar, cudaStream_t st, int* tile, const float** pivot_out,
                   const StatsRunning& run);
// FP16-split engine for dim <= 128 (stats_h.cu)
size_t stats_h_extra_workspace(int64_t L, int64_t dim);
bool stats_h_eligible(int64_t L, int64_t rows, int64_t dim);
// What the FP16-split launch left behind, for the merge that follows the (device-gated) TF32 fallback:
//   mode 1: per-CTA records of the narrow kernel (packed upper triangles, fp32), S' in ws_sum;
//   mode 2: per-item partial tiles of the wide kernel (one super-chunk), S' in ws_sum;
//   mode 0: the wide kernel needed several super-chunks and reduced them into the staging area itself (old merge applies).
struct StatsHPlan { int mode; const float* parts; int n_parts, n_units, upl, nB; const float* scale; int* flag; };
int stats_h_launch(const float* x, int64_t L, int64_t rows, int64_t dim, int64_t row_stride, int64_t batch_stride,
                   float* pivot, double* ws_cov, double* ws_sum, Arena& ar, cudaStream_t st, StatsHPlan* plan);
int stats_zero_if(double* p, int64_t n, const int* flag, cudaStream_t st);
// merges (records | partial tiles | the staging area refilled by the fallback, if the flag is up) into the running buffers
int stats_h_merge(const StatsHPlan& plan, const float* pivot, const double* ws_cov, const double* ws_sum, int64_t L,
                  int64_t rows, int64_t dim, const StatsRunning& run, cudaStream_t st);
// same scheme on the 128 x 128 blocks of the upper block triangle for dim > 128 (stats_h.cu)
size_t stats_h2_extra_workspace(int64_t L, int64_t dim);
bool stats_h2_eligible(int64_t L, int64_t rows, int64_t dim);
int stats_h2_launch(const float* x, int64_t L, int64_t rows, int64_t dim, int64_t row_stride, int64_t batch_stride,
                    float* pivot, double* ws_cov, double* ws_sum, Arena& ar, cudaStream_t st, StatsHPlan* plan);
}  // namespace otk
