// TEMPORARY (phase A): tcgen05 engines not yet linked
#include "gemm.cuh"
#include "stats_umma.cuh"
#include "apply_umma.cuh"
#include "sinkhorn_umma.cuh"
namespace otk {
bool sk_umma_eligible(int64_t, int64_t, int64_t, int) { return false; }
size_t sk_umma_workspace_bytes(int64_t, int64_t, int64_t) { return 0; }
int sk_umma_solve(const float*, const float*, int64_t, int64_t, int64_t, const float*, const float*, double, int, double, int, double, int, int, float*, float*, double*, int*, void*, size_t, cudaStream_t) { return OTK_ERR_INVALID_ARGUMENT; }
int sk_umma_colstep(const float*, const float*, int64_t, int64_t, int64_t, const float*, double, double, int, float*, float*, void*, size_t, cudaStream_t) { return OTK_ERR_INVALID_ARGUMENT; }
int sk_umma_rowstep(const float*, const float*, int64_t, int64_t, int64_t, const float*, const float*, double, double, int, float*, float*, void*, size_t, cudaStream_t) { return OTK_ERR_INVALID_ARGUMENT; }
}
