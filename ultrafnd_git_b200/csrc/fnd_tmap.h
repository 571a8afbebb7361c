// fnd_tmap.h — host-side TMA tensor-map encoding without linking libcuda.
// cuTensorMapEncodeTiled is fetched from the driver at run time through cudaGetDriverEntryPoint, so the
// library links against cudart only and still loads on a box without a GPU (symbols are checked there).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

namespace fnd {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// 2-D bf16 tensor, row-major: `outer` rows of `inner` contiguous elements, row pitch `pitch_elems`.
// Box = box_inner x box_outer elements, 128-byte swizzle (box_inner must be 64), zero fill out of bounds.
inline int encode_bf16_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_elems,
                          uint32_t box_inner, uint32_t box_outer) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) return -1;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (pitch_elems * 2) % 16 != 0) return -2;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -100 - static_cast<int>(r);
}

// 3-D bf16 tensor [batch][rows][cols] over a row-major [batch * rows, pitch] matrix: dim0 = `cols` contiguous elements,
// dim1 = `rows` (stride pitch), dim2 = `batch` (stride rows * pitch). Box = 64 x box_rows x 1, 128-byte swizzle. Rows
// past `rows` are zero-filled PER SAMPLE, so a tile that overhangs a sequence never sees the next sample's tokens.
inline int encode_bf16_3d(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint64_t batch,
                          uint64_t pitch_elems, uint32_t box_rows) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) return -1;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (pitch_elems * 2) % 16 != 0) return -2;
  cuuint64_t dims[3] = {cols, rows, batch};
  cuuint64_t strides[2] = {pitch_elems * 2, rows * pitch_elems * 2};
  cuuint32_t box[3] = {64, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -100 - static_cast<int>(r);
}

}  // namespace fnd
