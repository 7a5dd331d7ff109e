// stem.cu -- LearningToDownsample.conv: dense 3x3, stride 2, pad 0, 3 -> 32, folded BN, ReLU.
// Replaces reference models/fast_scnn.py:153 (_ConvBNReLU, :49-61).  Reads the NCHW fp32 image
// and writes the NHWC stage tensor, so the layout change costs no extra pass.
//
// Thread tile: 4 horizontally adjacent output pixels x 8 output channels.  A warp covers 32
// pixels of one output row x all 32 channels (lane = 4*pixel_group + channel_group), so each
// global store instruction writes full 32-byte sectors and the 27x32 weights are read from shared
// memory as broadcast float4s (2 LDS.128 per 32 FMA).  A CTA is 8 warps = 8 output rows x 32 cols.
#include "kernels.h"

#include "../../include/fscnn_b200.h"

namespace fscnn {

template <typename T, int FMT>
__global__ void __launch_bounds__(kThreads) stem_kernel(const void* __restrict__ xin, StemIn prm, const float* __restrict__ wpk,
                                                         const float* __restrict__ bias, T* __restrict__ out,
                                                         int H, int W, int Ho, int Wo) {
    const float* x = reinterpret_cast<const float*>(xin);
    const unsigned char* xb = reinterpret_cast<const unsigned char*>(xin);
    __shared__ __align__(16) float ws[27 * 32];
    __shared__ float bs[32];
    for (int i = threadIdx.x; i < 27 * 32; i += kThreads) ws[i] = __ldg(wpk + i);
    if (threadIdx.x < 32) bs[threadIdx.x] = __ldg(bias + threadIdx.x);
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cg = lane & 3, pg = lane >> 2;
    const int oy = blockIdx.y * 8 + warp;
    const int ox0 = blockIdx.x * 32 + pg * 4;
    const int n = blockIdx.z;
    if (oy >= Ho || ox0 >= Wo) return;

    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = bs[cg * 8 + j];

    const int ix0 = ox0 * 2;                       // first input column of this thread
    const bool vec_ok = ((W & 3) == 0) && (ix0 + 8 < W);
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const float* row = x + (((size_t)n * 3 + ci) * H + (oy * 2 + ky)) * W + ix0;
            float r[9];
            if (FMT == FSCNN_IN_U8_NHWC) {   // raw uint8 HWC: ToTensor + Normalize fused into the load
                const unsigned char* rb = xb + (((size_t)n * H + (oy * 2 + ky)) * W + ix0) * 3 + ci;
#pragma unroll
                for (int c = 0; c < 9; ++c)
                    r[c] = (ix0 + c < W) ? ((float)__ldg(rb + 3 * c) * (1.f / 255.f) - prm.mean[ci]) * prm.inv_std[ci] : 0.f;
            } else if (vec_ok) {
                const float4 v0 = __ldg(reinterpret_cast<const float4*>(row));
                const float4 v1 = __ldg(reinterpret_cast<const float4*>(row) + 1);
                r[0] = v0.x; r[1] = v0.y; r[2] = v0.z; r[3] = v0.w;
                r[4] = v1.x; r[5] = v1.y; r[6] = v1.z; r[7] = v1.w;
                r[8] = __ldg(row + 8);
            } else {
#pragma unroll
                for (int c = 0; c < 9; ++c) r[c] = (ix0 + c < W) ? __ldg(row + c) : 0.f;
            }
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const float* wk = ws + ((ci * 3 + ky) * 3 + kx) * 32 + cg * 8;
                const float4 w0 = *reinterpret_cast<const float4*>(wk);
                const float4 w1 = *reinterpret_cast<const float4*>(wk + 4);
                const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(r[2 * i + kx], wv[j], acc[i][j]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (ox0 + i >= Wo) break;
        T* o = out + (((size_t)n * Ho + oy) * Wo + ox0 + i) * 32 + cg * 8;
        Act<T>::st4(o, make_float4(relu(acc[i][0]), relu(acc[i][1]), relu(acc[i][2]), relu(acc[i][3])));
        Act<T>::st4(o + 4, make_float4(relu(acc[i][4]), relu(acc[i][5]), relu(acc[i][6]), relu(acc[i][7])));
    }
}

template <typename T>
cudaError_t launch_stem(const void* x, const StemIn& in, const StemW& w, T* out, int n, int h, int wd, int ho, int wo,
                        cudaStream_t s) {
    dim3 grid(ceil_div(wo, 32), ceil_div(ho, 8), n);
    if (in.format == FSCNN_IN_U8_NHWC)
        stem_kernel<T, FSCNN_IN_U8_NHWC><<<grid, kThreads, 0, s>>>(x, in, w.w, w.b, out, h, wd, ho, wo);
    else
        stem_kernel<T, FSCNN_IN_F32_NCHW><<<grid, kThreads, 0, s>>>(x, in, w.w, w.b, out, h, wd, ho, wo);
    return cudaGetLastError();
}

template cudaError_t launch_stem<float>(const void*, const StemIn&, const StemW&, float*, int, int, int, int, int, cudaStream_t);
template cudaError_t launch_stem<bf16>(const void*, const StemIn&, const StemW&, bf16*, int, int, int, int, int, cudaStream_t);

}  // namespace fscnn
