// Host-side TMA descriptor (CUtensorMap) encoding.  The driver entry point is resolved through the runtime
// (cudaGetDriverEntryPoint), so libotk does not link against libcuda.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace otk {

typedef CUresult (*TensorMapEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                           const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                           CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline TensorMapEncodeTiledFn tensormap_encoder() {
  static TensorMapEncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<TensorMapEncodeTiledFn>(p);
  }();
  return fn;
}

// fp32 tensor [batch][rows][cols] (cols contiguous, row stride `ld`, batch stride `bstride`, in elements);
// box = box_cols x box_rows x 1, 128-byte swizzle (box_cols * 4 must be <= 128), out-of-bounds reads give zeros.
// atom32 selects SWIZZLE_128B_ATOM_32B, the TMA pattern matching the UMMA layout of MN-major TF32 operands.
inline bool encode_map_f32_3d(CUtensorMap* map, const float* base, int64_t cols, int64_t rows, int64_t batch, int64_t ld,
                              int64_t bstride, int box_cols, int box_rows, bool atom32 = false) {
  TensorMapEncodeTiledFn enc = tensormap_encoder();
  if (!enc) return false;
  if (batch <= 1 || bstride <= 0) { batch = batch < 1 ? 1 : batch; if (bstride <= 0) bstride = rows * ld; }
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)bstride * 4};
  cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// fp16 tensor [rows][cols] (cols contiguous, row stride `ld` elements); box = box_cols x box_rows, 128-byte swizzle
inline bool encode_map_f16_2d(CUtensorMap* map, const void* base, int64_t cols, int64_t rows, int64_t ld, int box_cols,
                              int box_rows) {
  TensorMapEncodeTiledFn enc = tensormap_encoder();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// fp16 tensor [batch][rows][cols] (cols contiguous; ld, bstride in elements, both multiples of 8); box = box_cols x box_rows x 1,
// 128-byte swizzle (box_cols * 2 must be <= 128), out-of-bounds reads give zeros
inline bool encode_map_f16_3d(CUtensorMap* map, const void* base, int64_t cols, int64_t rows, int64_t batch, int64_t ld,
                              int64_t bstride, int box_cols, int box_rows) {
  TensorMapEncodeTiledFn enc = tensormap_encoder();
  if (!enc) return false;
  if (batch < 1) batch = 1;
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)bstride * 2};
  cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

}  // namespace otk
