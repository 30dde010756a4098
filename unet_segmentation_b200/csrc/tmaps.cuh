// Host-side TMA descriptor construction (tiled 2-D and im2col 4-D) through the driver entry
// points, resolved at run time so the library links without libcuda.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <string.h>

#include "common.cuh"

namespace ub {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*PFN_encodeIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                     const cuuint64_t*, const cuuint64_t*, const int*, const int*,
                                     cuuint32_t, cuuint32_t, const cuuint32_t*,
                                     CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TmapApi {
    PFN_encodeTiled tiled = nullptr;
    PFN_encodeIm2col im2col = nullptr;
    int driver_version = 0;
    bool ok = false;
};

inline TmapApi& tmap_api() {
    static TmapApi api;
    if (!api.ok) {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) ==
                cudaSuccess && f)
            api.tiled = (PFN_encodeTiled)f;
        f = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f, cudaEnableDefault, &q) ==
                cudaSuccess && f)
            api.im2col = (PFN_encodeIm2col)f;
        cudaDriverGetVersion(&api.driver_version);
        api.ok = api.tiled && api.im2col;
    }
    return api;
}

// 2-D bf16 matrix [rows][inner] with a row pitch, box = {64 inner elements (128 B), box_rows},
// 128-byte swizzle. Used for packed weights (K-major B operand) and for [pixels][channels]
// activations/gradients read MN-major by the weight-gradient kernel.
inline int make_tmap_2d(CUtensorMap* out, const void* ptr, unsigned long long inner,
                        unsigned long long rows, unsigned long long pitch_bytes,
                        unsigned box_rows) {
    TmapApi& api = tmap_api();
    if (!api.ok) return -1;
    cuuint64_t dims[2] = {inner, rows};
    cuuint64_t strides[1] = {pitch_bytes};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = api.tiled(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims,
                           strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -(int)r - 1000;
}

// Tap-major packed weights [taps][n][c] (bf16): box = (64 channels, box_n rows, box_taps taps) lands
// in shared memory as box_taps consecutive K-major [box_n][64] tiles.
inline int make_tmap_weights(CUtensorMap* out, const void* ptr, unsigned long long c,
                             unsigned long long n, unsigned long long taps, unsigned box_n,
                             unsigned box_taps) {
    TmapApi& api = tmap_api();
    if (!api.ok) return -1;
    cuuint64_t dims[3] = {c, n, taps};
    cuuint64_t strides[2] = {c * 2, n * c * 2};
    cuuint32_t box[3] = {64, box_n, box_taps};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = api.tiled(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims,
                           strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -(int)r - 1000;
}

// [rows][cols] bf16 matrix (row pitch in bytes) seen as (64 channels, rows, cols/64 chunks): one
// box of (64, box_rows, box_chunks) lands in shared memory as [chunk][row][64 ch] with the
// 128-byte swizzle, i.e. box_chunks MN-major operand chunks in a single TMA operation.
inline int make_tmap_chunked(CUtensorMap* out, const void* ptr, unsigned long long cols,
                             unsigned long long rows, unsigned long long pitch_bytes,
                             unsigned box_rows, unsigned box_chunks) {
    TmapApi& api = tmap_api();
    if (!api.ok) return -1;
    cuuint64_t dims[3] = {64, rows, cols / 64};
    cuuint64_t strides[2] = {pitch_bytes, 128};
    cuuint32_t box[3] = {64, box_rows, box_chunks};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = api.tiled(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims,
                           strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -(int)r - 1000;
}

// Tiled 4-D map over an NHWC bf16 view, box = (64 channels, box_w pixels, box_h rows). Out-of-range
// coordinates (negative or past the edge) are zero filled: that is the padding of the data gradient.
inline int make_tmap_rows(CUtensorMap* out, const View& v, unsigned box_w, unsigned box_h = 1) {
    TmapApi& api = tmap_api();
    if (!api.ok) return -1;
    cuuint64_t dims[4] = {(cuuint64_t)v.C, (cuuint64_t)v.W, (cuuint64_t)v.H, (cuuint64_t)v.N};
    cuuint64_t strides[3] = {(cuuint64_t)v.sW * 2, (cuuint64_t)v.sH * 2, (cuuint64_t)v.sN * 2};
    cuuint32_t box[4] = {64, box_w, box_h, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = api.tiled(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(v.ptr), dims,
                           strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -(int)r - 1000;
}

// im2col map over an NHWC bf16 view. Base pixels traverse, in (w,h,n) raster order with step
// `tstride`, the box [lower, W-1+upper] x [lower, H-1+upper]; a load fetches `pixels` base pixels
// x 64 channels, each pixel displaced by the per-instruction filter-tap offset.
//   valid 3x3 conv (fprop / wgrad activations): lower 0, upper -2
//   its data gradient (full correlation over dY):  lower -2, upper 0
//   2x2 stride-2 transposed conv backward:         lower 0, upper -1, tstride 2
//   per-axis corners (make_tmap_im2col_wh): the "shifted dY" weight gradient walks the activations
//   with (w: 0 / 0, h: 0 / -2) and dY with (w: -2 / 0, h: 0 / 0) — the same base-pixel grid for both.
inline int make_tmap_im2col_wh(CUtensorMap* out, const View& v, int lower_w, int upper_w, int lower_h,
                               int upper_h, int tstride, unsigned pixels);
inline int make_tmap_im2col(CUtensorMap* out, const View& v, int lower, int upper, int tstride,
                            unsigned pixels) {
    return make_tmap_im2col_wh(out, v, lower, upper, lower, upper, tstride, pixels);
}
inline int make_tmap_im2col_wh(CUtensorMap* out, const View& v, int lower_w, int upper_w, int lower_h,
                               int upper_h, int tstride, unsigned pixels) {
    TmapApi& api = tmap_api();
    if (!api.ok) return -1;
    cuuint64_t dims[4] = {(cuuint64_t)v.C, (cuuint64_t)v.W, (cuuint64_t)v.H, (cuuint64_t)v.N};
    cuuint64_t strides[3] = {(cuuint64_t)v.sW * 2, (cuuint64_t)v.sH * 2, (cuuint64_t)v.sN * 2};
    int lo[2] = {lower_w, lower_h};
    int up[2] = {upper_w, upper_h};
    cuuint32_t estr[4] = {1, (cuuint32_t)tstride, (cuuint32_t)tstride, 1};
    CUresult r = api.im2col(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(v.ptr),
                            dims, strides, lo, up, 64, pixels, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return -(int)r - 1000;
    // Drivers up to 13.1 mis-encode im2col maps of tensors smaller than 128 KiB (same
    // work-around CUTLASS applies in copy_traits_sm90_im2col.hpp).
    unsigned long long bytes = (unsigned long long)v.N * (unsigned long long)v.sN * 2ull;
    if (api.driver_version <= 13010 && bytes < 131072ull)
        reinterpret_cast<unsigned long long*>(out)[1] &= ~(1ull << 21);
    return 0;
}

}  // namespace ub
