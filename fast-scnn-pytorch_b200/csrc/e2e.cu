// e2e.cu -- the camera-frame wrapper around the forward path (reference export_onnx_fixed.py:34-98, EndToEndFastSCNN /
// EndToEndPreprocessing), the steps immediately before and after the network:
//   preprocess  : frames [N,3,h,w] uint8 or float32 (0..255) -> bilinear resize (align_corners=False) to base x base, / 255,
//                 optional (x - mean) / std -> fp32 NCHW, the network's input                      (:78-98)
//   postprocess : the network's LOW-resolution logits [N,hl,wl,ncp] (NHWC fp32) -> the x8 align_corners=True upsample to
//                 base x base (models/fast_scnn.py:40) composed with the resize back to the frame size (align_corners=False,
//                 :54-55) and the optional softmax over the classes (:57-58) -> fp32 NCHW [N,nc,h,w].  The full-resolution
//                 logits (nc x base x base floats per frame) are never materialised: each output pixel interpolates its four
//                 full-resolution taps straight from the low-resolution tile, in the reference's order of operations.
#include "kernels.h"

namespace fscnn {

namespace {
// ATen area_pixel_compute_source_index, align_corners = false
__device__ __forceinline__ void hp_coord(int dst, float scale, int n_in, int& i0, int& i1, float& lam) {
    float src = scale * ((float)dst + 0.5f) - 0.5f;
    src = src < 0.f ? 0.f : src;
    i0 = min((int)src, n_in - 1);
    i1 = min(i0 + 1, n_in - 1);
    lam = src - (float)i0;
}
// align_corners = true
__device__ __forceinline__ void ac_coord(int dst, float scale, int n_in, int& i0, int& i1, float& lam) {
    const float src = scale * (float)dst;
    i0 = min((int)src, n_in - 1);
    i1 = min(i0 + 1, n_in - 1);
    lam = src - (float)i0;
}
constexpr int kE2eMaxNc = 32;
}  // namespace

template <typename T>
__global__ void e2e_preprocess_kernel(const T* __restrict__ x, int n, int h, int w, int base, float3 mean, float3 stdv, int norm,
                                      float* __restrict__ out) {
    const int ox = blockIdx.x * blockDim.x + threadIdx.x, oy = blockIdx.y, img = blockIdx.z;
    if (ox >= base) return;
    int y0, y1, x0, x1;
    float ly, lx;
    hp_coord(oy, (float)h / (float)base, h, y0, y1, ly);
    hp_coord(ox, (float)w / (float)base, w, x0, x1, lx);
    const float hy = 1.f - ly, hx = 1.f - lx;
    const float m[3] = {mean.x, mean.y, mean.z}, sd[3] = {stdv.x, stdv.y, stdv.z};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const T* p = x + ((size_t)img * 3 + c) * h * w;
        float v;
        if (h == base && w == base) {
            v = (float)p[(size_t)oy * w + ox];
        } else {
            const float top = (float)p[(size_t)y0 * w + x0] * hx + (float)p[(size_t)y0 * w + x1] * lx;
            const float bot = (float)p[(size_t)y1 * w + x0] * hx + (float)p[(size_t)y1 * w + x1] * lx;
            v = hy * top + ly * bot;
        }
        v = v / 255.0f;
        if (norm) v = (v - m[c]) / sd[c];
        out[(((size_t)img * 3 + c) * base + oy) * base + ox] = v;
    }
}

__global__ void e2e_postprocess_kernel(const float* __restrict__ low, int nc, int ncp, int hl, int wl, int bh, int bw, int oh, int ow,
                                       int apply_softmax, float* __restrict__ out) {
    const int ox = blockIdx.x * blockDim.x + threadIdx.x, oy = blockIdx.y, img = blockIdx.z;
    if (ox >= ow) return;
    // the four full-resolution taps of this output pixel (resize base -> frame, align_corners = false) ...
    int Y[2], X[2];
    float LY, LX;
    if (bh == oh && bw == ow) { Y[0] = Y[1] = oy; X[0] = X[1] = ox; LY = 0.f; LX = 0.f; }
    else {
        hp_coord(oy, (float)bh / (float)oh, bh, Y[0], Y[1], LY);
        hp_coord(ox, (float)bw / (float)ow, bw, X[0], X[1], LX);
    }
    // ... each an align_corners = true interpolation of the low-resolution logits
    const float sy = bh > 1 ? (float)(hl - 1) / (float)(bh - 1) : 0.f, sx = bw > 1 ? (float)(wl - 1) / (float)(bw - 1) : 0.f;
    int y0[2], y1[2], x0[2], x1[2];
    float ly[2], lx[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        ac_coord(Y[k], sy, hl, y0[k], y1[k], ly[k]);
        ac_coord(X[k], sx, wl, x0[k], x1[k], lx[k]);
    }
    const float* base = low + (size_t)img * hl * wl * ncp;
    float v[kE2eMaxNc];
    float vmax = -INFINITY;
    for (int c = 0; c < nc; ++c) {
        float full[2][2];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const float p00 = __ldg(base + ((size_t)y0[a] * wl + x0[b]) * ncp + c), p01 = __ldg(base + ((size_t)y0[a] * wl + x1[b]) * ncp + c);
                const float p10 = __ldg(base + ((size_t)y1[a] * wl + x0[b]) * ncp + c), p11 = __ldg(base + ((size_t)y1[a] * wl + x1[b]) * ncp + c);
                const float top = p00 * (1.f - lx[b]) + p01 * lx[b], bot = p10 * (1.f - lx[b]) + p11 * lx[b];
                full[a][b] = (1.f - ly[a]) * top + ly[a] * bot;
            }
        const float top = full[0][0] * (1.f - LX) + full[0][1] * LX, bot = full[1][0] * (1.f - LX) + full[1][1] * LX;
        const float val = (bh == oh && bw == ow) ? full[0][0] : (1.f - LY) * top + LY * bot;
        v[c] = val;
        vmax = fmaxf(vmax, val);
    }
    float sum = 0.f;
    if (apply_softmax)
        for (int c = 0; c < nc; ++c) { v[c] = expf(v[c] - vmax); sum += v[c]; }
    const float inv = apply_softmax ? 1.f / sum : 1.f;
    for (int c = 0; c < nc; ++c) out[(((size_t)img * nc + c) * oh + oy) * ow + ox] = apply_softmax ? v[c] * inv : v[c];
}

cudaError_t launch_e2e_preprocess(const void* x, int is_u8, int n, int h, int w, int base, const float* mean3, const float* std3,
                                  float* out, cudaStream_t s) {
    if (n < 1 || h < 1 || w < 1 || base < 1 || n > 65535 || base > 65535) return cudaErrorInvalidValue;
    const int norm = mean3 && std3;
    const float3 mean = norm ? make_float3(mean3[0], mean3[1], mean3[2]) : make_float3(0.f, 0.f, 0.f);
    const float3 inv = norm ? make_float3(std3[0], std3[1], std3[2]) : make_float3(1.f, 1.f, 1.f);
    dim3 grid(ceil_div(base, 128), base, n);
    if (is_u8) e2e_preprocess_kernel<unsigned char><<<grid, 128, 0, s>>>(reinterpret_cast<const unsigned char*>(x), n, h, w, base, mean, inv, norm, out);
    else e2e_preprocess_kernel<float><<<grid, 128, 0, s>>>(reinterpret_cast<const float*>(x), n, h, w, base, mean, inv, norm, out);
    return cudaGetLastError();
}

cudaError_t launch_e2e_postprocess(const float* low, int nc, int ncp, int n, int hl, int wl, int bh, int bw, int oh, int ow,
                                   int apply_softmax, float* out, cudaStream_t s) {
    if (nc < 1 || nc > kE2eMaxNc || n < 1 || n > 65535 || oh < 1 || oh > 65535) return cudaErrorInvalidValue;
    dim3 grid(ceil_div(ow, 128), oh, n);
    e2e_postprocess_kernel<<<grid, 128, 0, s>>>(low, nc, ncp, hl, wl, bh, bw, oh, ow, apply_softmax, out);
    return cudaGetLastError();
}

}  // namespace fscnn
