// Host-side CUtensorMap construction (TMA descriptors) without linking libcuda: the driver entry point
// is resolved at run time through the CUDA runtime.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdexcept>
#include <string>

namespace fseend {

void set_last_error(const std::string& s);

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p)
      throw std::runtime_error("cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// fp16 tensor, `rank` dims (innermost first), strides in ELEMENTS for dims 1..rank-1, 128B swizzle.
// box[0] must be 64 (128 bytes).
inline CUtensorMap make_tmap_f16(const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                                 const uint32_t* box) {
  CUtensorMap m;
  cuuint64_t gdim[5], gstr[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gstr[i - 1] = strides_elems[i - 1] * 2;
  }
  CUresult r = get_encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(base), gdim, gstr, bx, es,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled failed, code " + std::to_string((int)r));
  return m;
}

// [rows][cols] row-major (row stride ld elements) viewed as (cols, rows_per_seq, n_seq): the 3-D form lets
// per-sequence tiles clip/zero-fill at the sequence end.  box = (64, box_rows, 1).
inline CUtensorMap make_tmap_rows3d(const void* base, uint64_t cols, uint64_t ld, uint64_t rows_per_seq,
                                    uint64_t n_seq, uint32_t box_rows) {
  uint64_t dims[3] = {cols, rows_per_seq, n_seq};
  uint64_t str[2] = {ld, ld * rows_per_seq};
  uint32_t box[3] = {64, box_rows, 1};
  return make_tmap_f16(base, 3, dims, str, box);
}

}  // namespace fseend
