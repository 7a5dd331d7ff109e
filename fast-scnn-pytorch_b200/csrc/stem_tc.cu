// stem_tc.cu -- bf16 LearningToDownsample.conv (dense 3x3, stride 2, pad 0, 3 -> 32, BN, ReLU;
// reference models/fast_scnn.py:153, :49-61) as an implicit contraction on the tensor cores:
// M = 128 output pixels (4 rows x 32 columns), K = 27 taps padded to 32, N = 32 channels.
//
// The kernel also absorbs the step in front of the path: with the
// FSCNN_IN_U8_NHWC input format it reads the raw uint8 HWC image and applies
// transforms.ToTensor() + Normalize(mean, std) (reference eval.py:22-25, demo.py:37-40) while
// staging, so the fp32 CHW tensor (4x the bytes) never exists.
//
// Per CTA (256 threads, ~18 KB smem -> many CTAs per SM hide the load latency):
//   stage the 9 x 65 x 3 input patch as fp32 planes -> gather the im2col A tile (bf16) ->
//   2 MMAs into TMEM[128 x 32] -> bias + ReLU -> bf16 NHWC.
#include "kernels.h"
#include "umma.cuh"

#include "../../include/fscnn_b200.h"

namespace fscnn {

namespace {
constexpr int kRows = 9, kCols = 65, kLd = 68;   // staged patch: 9 rows x 65 columns per channel, row pitch 68 floats
}

template <int FMT>
__global__ void __launch_bounds__(kThreads)
stem_tc_kernel(const void* __restrict__ x, StemIn prm, const bf16* __restrict__ w_img, const float* __restrict__ bias,
               bf16* __restrict__ out, int H, int W, int Ho, int Wo) {
    __shared__ __align__(16) float In[3 * kRows * kLd];
    __shared__ __align__(128) uint8_t sAraw[128 * 32 * 2];
    __shared__ __align__(128) uint8_t sBraw[32 * 32 * 2];
    __shared__ float bs[32];
    __shared__ __align__(8) uint64_t bar_mma;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n = blockIdx.z, oy0 = blockIdx.y * 4, ox0 = blockIdx.x * 32;
    const int iy0 = oy0 * 2, ix0 = ox0 * 2;

    if (tid == 0) { mbar_init(&bar_mma, 1); fence_mbar_init(); }
    if (warp == 0) { tmem_alloc(&tmem_base_s, 32); tmem_relinquish(); }
    if (tid < 32) bs[tid] = __ldg(bias + tid);
    if (tid < 128) reinterpret_cast<uint4*>(sBraw)[tid] = __ldg(reinterpret_cast<const uint4*>(w_img) + tid);

    // stage the patch: one warp per (channel, row) line, lanes along columns -- no integer divisions
    if (FMT == FSCNN_IN_F32_NCHW) {
        const float* xf = reinterpret_cast<const float*>(x);
        for (int line = warp; line < 3 * kRows; line += kThreads / 32) {
            const int ci = line / kRows, r = line - ci * kRows;
            const int iy = iy0 + r;
            const float* src = xf + (((size_t)n * 3 + ci) * H + iy) * W + ix0;
            float* dst = In + line * kLd;
#pragma unroll
            for (int c = lane; c < kCols; c += 32) dst[c] = (iy < H && ix0 + c < W) ? __ldg(src + c) : 0.f;
        }
    } else {
        const unsigned char* xb = reinterpret_cast<const unsigned char*>(x);
        for (int r = warp; r < kRows; r += kThreads / 32) {
            const int iy = iy0 + r;
            const unsigned char* src = xb + (((size_t)n * H + iy) * W + ix0) * 3;
#pragma unroll
            for (int c = lane; c < kCols; c += 32) {
                float v0 = 0.f, v1 = 0.f, v2 = 0.f;
                if (iy < H && ix0 + c < W) {
                    v0 = ((float)__ldg(src + 3 * c) * (1.f / 255.f) - prm.mean[0]) * prm.inv_std[0];
                    v1 = ((float)__ldg(src + 3 * c + 1) * (1.f / 255.f) - prm.mean[1]) * prm.inv_std[1];
                    v2 = ((float)__ldg(src + 3 * c + 2) * (1.f / 255.f) - prm.mean[2]) * prm.inv_std[2];
                }
                In[(0 * kRows + r) * kLd + c] = v0;
                In[(1 * kRows + r) * kLd + c] = v1;
                In[(2 * kRows + r) * kLd + c] = v2;
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const uint32_t sA = smem_u32(sAraw), sB = smem_u32(sBraw);

    // im2col gather: thread = (pixel p, pair of 8-tap chunks); k = ci*9 + ky*3 + kx as in the folded weight.
    // Both branches are fully unrolled, so every tap offset is a compile-time constant.
    {
        const int p = tid & 127, hi = tid >> 7;
        const int py = p >> 5, px = p & 31;
        const float* base = In + (2 * py) * kLd + 2 * px;
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
            float v[8];
            if (hi == 0) {
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const int k = kk * 8 + t;
                    v[t] = base[((k / 9) * kRows + (k % 9) / 3) * kLd + k % 3];
                }
            } else {
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const int k = 16 + kk * 8 + t;
                    v[t] = k < 27 ? base[((k / 9) * kRows + (k % 9) / 3) * kLd + k % 3] : 0.f;
                }
            }
            sts128(sA + a_tile_off(p, 2 * hi + kk), packbf(v[0], v[1]), packbf(v[2], v[3]), packbf(v[4], v[5]), packbf(v[6], v[7]));
        }
    }
    fence_async_proxy();
    __syncthreads();
    if (tid == 0) {
        tc_fence_after_sync();
        constexpr uint32_t idesc = make_idesc_bf16(128, 32);
#pragma unroll
        for (int k16 = 0; k16 < 2; ++k16)
            umma_bf16_ss(tmem, make_smem_desc(sA + k16 * 4096, 2048, 128), make_smem_desc(sB + k16 * 2 * 512, 512, 128), idesc, k16 > 0);
        umma_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, 0);
    tc_fence_after_sync();
    {
        const int q = warp & 3, half = warp >> 2;
        const int p = q * 32 + lane;
        const int oy = oy0 + (p >> 5), ox = ox0 + (p & 31);
        uint32_t r[16];
        tmem_ld_32x32b_x16(tmem + ((uint32_t)(q * 32) << 16) + half * 16, r);
        tmem_ld_wait();
        if (oy < Ho && ox < Wo) {
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = relu(__uint_as_float(r[i]) + bs[half * 16 + i]);
            uint4* o = reinterpret_cast<uint4*>(out + (((size_t)n * Ho + oy) * Wo + ox) * 32 + half * 16);
            o[0] = make_uint4(packbf(v[0], v[1]), packbf(v[2], v[3]), packbf(v[4], v[5]), packbf(v[6], v[7]));
            o[1] = make_uint4(packbf(v[8], v[9]), packbf(v[10], v[11]), packbf(v[12], v[13]), packbf(v[14], v[15]));
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 32);
}

cudaError_t launch_stem_tc(const void* x, const StemIn& in, const bf16* w_img, const float* bias, bf16* out, int n, int h,
                           int wd, int ho, int wo, cudaStream_t s) {
    dim3 grid(ceil_div(wo, 32), ceil_div(ho, 4), n);
    if (in.format == FSCNN_IN_U8_NHWC)
        stem_tc_kernel<FSCNN_IN_U8_NHWC><<<grid, kThreads, 0, s>>>(x, in, w_img, bias, out, h, wd, ho, wo);
    else
        stem_tc_kernel<FSCNN_IN_F32_NCHW><<<grid, kThreads, 0, s>>>(x, in, w_img, bias, out, h, wd, ho, wo);
    return cudaGetLastError();
}

}  // namespace fscnn
