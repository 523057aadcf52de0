// Shared host/device helpers for libotk (B200 / sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/otk.h"

namespace otk {

// ---- error plumbing -----------------------------------------------------------------------------
char* last_error_buffer();  // thread-local, defined in api.cu
inline void set_last_error(const char* what, cudaError_t e) {
  snprintf(last_error_buffer(), 512, "%s: %s", what, cudaGetErrorString(e));
}
inline void set_last_error_msg(const char* what) { snprintf(last_error_buffer(), 512, "%s", what); }

#define OTK_CUDA(call)                                  \
  do {                                                  \
    cudaError_t e__ = (call);                           \
    if (e__ != cudaSuccess) {                           \
      ::otk::set_last_error(#call, e__);                \
      return OTK_ERR_CUDA;                              \
    }                                                   \
  } while (0)

void count_launch(int n);  // process-wide kernel-launch counter (api.cu), read by otk_launch_count()

#define OTK_LAUNCH_CHECK()                              \
  do {                                                  \
    ::otk::count_launch(1);                             \
    cudaError_t e__ = cudaGetLastError();               \
    if (e__ != cudaSuccess) {                           \
      ::otk::set_last_error("kernel launch", e__);      \
      return OTK_ERR_CUDA;                              \
    }                                                   \
  } while (0)

#define OTK_REQUIRE(cond, msg)                          \
  do {                                                  \
    if (!(cond)) {                                      \
      ::otk::set_last_error_msg(msg);                   \
      return OTK_ERR_INVALID_ARGUMENT;                  \
    }                                                   \
  } while (0)

#define OTK_TRY(expr)                                   \
  do {                                                  \
    int s__ = (expr);                                   \
    if (s__ != OTK_OK) return s__;                      \
  } while (0)

int require_device();  // OTK_OK on sm_100, defined in api.cu
int sm_count();

inline cudaStream_t as_stream(otk_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
inline size_t dtype_size(int dt) { return dt == OTK_F64 ? 8 : 4; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// bump allocator over a caller-provided workspace
struct Arena {
  char* base;
  size_t cap, off;
  Arena(void* p, size_t n) : base(static_cast<char*>(p)), cap(n), off(0) {}
  template <typename T>
  T* take(size_t count) {
    off = align_up(off, 256);
    T* r = reinterpret_cast<T*>(base + off);
    off += count * sizeof(T);
    return r;
  }
  bool ok() const { return off <= cap && (base != nullptr || off == 0); }
};

// ---- programmatic dependent launch ----------------------------------------------------------------
// launch `kern` so that its launch processing (and whatever it does before `griddepcontrol.wait`) overlaps the tail of the
// previous kernel of the stream; the kernel MUST execute ptx::pdl_wait() before it touches anything the predecessor
// wrote.  `cluster` > 1 adds a cluster dimension.  OTK_PDL=0 launches plainly (tuning aid).
bool pdl_enabled(int site = 0);   // api.cu; OTK_PDL is a bit mask over launch sites (tuning aid), default all
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_site(int site, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                   int cluster, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = (unsigned)cluster;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled(site)) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = (unsigned)n;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster,
                              Args&&... args) {
  return launch_pdl_site(0, kern, grid, block, smem, st, cluster, static_cast<Args&&>(args)...);
}

// ---- device helpers -----------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <typename T>
__device__ __forceinline__ T warp_max(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    T w = __shfl_xor_sync(0xffffffffu, v, o);
    v = w > v ? w : v;
  }
  return v;
}

// dtype-erased load/store for fp32/fp64 buffers (statistics and small matrices)
__device__ __forceinline__ double load_real(const void* p, int64_t i, int dt) {
  return dt == OTK_F64 ? static_cast<const double*>(p)[i] : static_cast<double>(static_cast<const float*>(p)[i]);
}
__device__ __forceinline__ void store_real(void* p, int64_t i, int dt, double v) {
  if (dt == OTK_F64) static_cast<double*>(p)[i] = v;
  else static_cast<float*>(p)[i] = static_cast<float>(v);
}

}  // namespace otk
