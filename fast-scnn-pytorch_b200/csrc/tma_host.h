// tma_host.h -- host side of the TMA halo-tile loads: builds the 5-D tensor map that lets ONE cp.async.bulk.tensor
// stage an NHWC bf16 halo tile directly in the SWIZZLE_NONE core-matrix A-operand layout of tcgen05.mma
//     smem[(c / 8) * PIN + pixel][c % 8],  pixel = y * IW + x      (LBO = PIN * 16 bytes, SBO = 128 bytes)
// with out-of-image coordinates (the convolution padding, partial tiles) zero-filled by the copy engine.
// Tensor dims (innermost first): {8 channels, W, H, C/8 channel groups, N}; verified on hardware by tma_probe.cu.
// The driver entry point is fetched through the runtime, so the library carries no link-time libcuda dependency
// (it must still load on a machine without a GPU driver).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace fscnn {

typedef CUresult (*TmaEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline cudaError_t tma_encode_fn(TmaEncodeFn* out) {
    static TmaEncodeFn enc = nullptr;
    if (!enc) {
        cudaDriverEntryPointQueryResult q;
        void* fn = nullptr;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
        if (e != cudaSuccess) return e;
        if (!fn || q != cudaDriverEntryPointSuccess) return cudaErrorNotSupported;
        enc = reinterpret_cast<TmaEncodeFn>(fn);
    }
    *out = enc;
    return cudaSuccess;
}

// plain tiled map, no swizzle, zero fill outside the tensor; strides[i] = byte stride of dimension i + 1
inline cudaError_t make_tiled_map(CUtensorMap* map, CUtensorMapDataType dt, int rank, const void* base, const cuuint64_t* dims,
                                  const cuuint64_t* strides, const cuuint32_t* box) {
    TmaEncodeFn enc;
    cudaError_t e = tma_encode_fn(&enc);
    if (e != cudaSuccess) return e;
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const CUresult r = enc(map, dt, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

inline cudaError_t make_nhwc_halo_map(CUtensorMap* map, const void* base, int n, int h, int w, int c, int box_h, int box_w) {
    const cuuint64_t dims[5] = {8, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)c / 8, (cuuint64_t)n};
    const cuuint64_t strides[4] = {(cuuint64_t)c * 2, (cuuint64_t)w * c * 2, 16, (cuuint64_t)h * w * c * 2};
    const cuuint32_t box[5] = {8, (cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)c / 8, 1};
    return make_tiled_map(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, dims, strides, box);
}

}  // namespace fscnn
