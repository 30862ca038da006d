// Host side of the TMA tensor maps: cuTensorMapEncodeTiled is reached through the runtime's driver
// entry-point query, so the library links against cudart only.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "common.cuh"

typedef CUresult (*gs_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                       const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                       CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline gs_encode_tiled_fn gs_encode_tiled() {
  static gs_encode_tiled_fn fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (gs_encode_tiled_fn)f;
  }
  return fn;
}

// fp32 NHWC activation [n, h, w, c] seen as the 4-D tensor (c, w, n, h) -- innermost first -- so that a box
// (kc, bw, img, bh) lands in shared memory as [bh][img][bw][kc]: image rows of `img` images interleaved.
// `swizzle_bytes` = 128 / 64 / 0 must equal kc * 4 when non-zero.
// `sh`, `sw_` > 1 describe a strided (sub-pixel phase) view: pixel (y, x) of the view is pixel (y*sh, x*sw_) of
// the underlying [n, h*sh, w*sw_, c] tensor starting at `base`.
static inline int gs_make_act_tmap(CUtensorMap* tm, const float* base, int n, int h, int w, int c, int kc, int bw,
                                   int img, int bh, int swizzle_bytes, int sh = 1, int sw_ = 1) {
  gs_encode_tiled_fn enc = gs_encode_tiled();
  if (!enc) {
    gs_set_error("cuTensorMapEncodeTiled is not available from this driver");
    return GS_ERR_CUDA;
  }
  cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)n, (cuuint64_t)h};
  cuuint64_t strides[3] = {(cuuint64_t)c * 4 * sw_, (cuuint64_t)h * sh * w * sw_ * c * 4, (cuuint64_t)w * sw_ * c * 4 * sh};
  cuuint32_t box[4] = {(cuuint32_t)kc, (cuuint32_t)bw, (cuuint32_t)img, (cuuint32_t)bh};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    gs_set_error("cuTensorMapEncodeTiled failed (%d) for [%d,%d,%d,%d] box (%d,%d,%d,%d)", (int)r, n, h, w, c, kc, bw, img, bh);
    return GS_ERR_CUDA;
  }
  return GS_OK;
}
