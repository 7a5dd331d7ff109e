// fold.cu -- load-time weight preparation: fold eval-mode BatchNorm (eps 1e-5) into the preceding
// convolution and repack to the k-major layouts the kernels read.  Replaces, once per
// load_state_dict, the per-forward nn.BatchNorm2d calls of reference models/fast_scnn.py
// (:56, :71, :74, :87, :108, :200, :204, :27):
//     W'[k][co] = W[co][k] * gamma[co] / sqrt(var[co] + eps)
//     b'[co]    = beta[co] + (conv_bias[co] - mean[co]) * gamma[co] / sqrt(var[co] + eps)
// Layers without BN (classifier.conv.1, auxlayer.4) pass gamma == nullptr (scale 1, b' = conv bias).
#include "kernels.h"

namespace fscnn {

__global__ void fold_kernel(const float* __restrict__ w, const float* __restrict__ cbias, const float* __restrict__ gamma,
                            const float* __restrict__ beta, const float* __restrict__ mean, const float* __restrict__ var,
                            int cout, int kdim, int taps, int tap_major, float* __restrict__ out_w, int ld_out,
                            float* __restrict__ out_b, int accumulate_bias) {
    const int total = cout * kdim;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total + cout; i += gridDim.x * blockDim.x) {
        const int co = i < total ? i / kdim : i - total;
        float scale = 1.f;
        if (gamma) scale = gamma[co] / sqrtf(var[co] + 1e-5f);
        if (i < total) {
            int k = i % kdim;
            if (tap_major && taps > 1) {   // (ci, tap) -> (tap, ci)
                const int cin = kdim / taps;
                k = (k % taps) * cin + (k / taps);
            }
            out_w[(size_t)k * ld_out + co] = w[i] * scale;
        } else if (out_b) {
            const float cb = cbias ? cbias[co] : 0.f;
            const float b = gamma ? beta[co] + (cb - mean[co]) * scale : cb;
            out_b[co] = accumulate_bias ? out_b[co] + b : b;
        }
    }
}

cudaError_t launch_fold(const float* w, const float* cbias, const float* gamma, const float* beta, const float* mean,
                        const float* var, int cout, int kdim, int taps, int tap_major, float* out_w, int ld_out,
                        float* out_b, int accumulate_bias, cudaStream_t s) {
    const int total = cout * kdim + cout;
    fold_kernel<<<ceil_div(total, 256), 256, 0, s>>>(w, cbias, gamma, beta, mean, var, cout, kdim, taps, tap_major, out_w,
                                                     ld_out, out_b, accumulate_bias);
    return cudaGetLastError();
}


// bf16 B-operand images for tcgen05.mma (umma.cuh): W'[n][k] = W[n][k] * scale[n] cut into chunks of NC rows x KC
// columns (chunk index = (n/NC)*(K/KC) + k/KC); inside a chunk 8x8 core matrices are ordered [k/8][n/8], i.e. element
// (n, k) sits at bf16 index ((k/8)*(NC/8) + n/8)*64 + (n%8)*8 + k%8 -- exactly the shared-memory image the kernels
// bulk-copy (LBO = NC*16 bytes, SBO = 128 bytes).
__global__ void fold_umma_kernel(const float* __restrict__ w, const float* __restrict__ gamma, const float* __restrict__ var,
                                 int nrows, int kdim, int NC, int KC, __nv_bfloat16* __restrict__ out) {
    const int total = nrows * kdim;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int n = i / kdim, k = i % kdim;
        const float scale = gamma ? gamma[n] / sqrtf(var[n] + 1e-5f) : 1.f;
        const int chunk = (n / NC) * (kdim / KC) + k / KC;
        const int nl = n % NC, kl = k % KC;
        const size_t o = (size_t)chunk * NC * KC + ((size_t)(kl >> 3) * (NC >> 3) + (nl >> 3)) * 64 + (nl & 7) * 8 + (kl & 7);
        out[o] = __float2bfloat16_rn(w[i] * scale);
    }
}

// uint8-input stem: fold ToTensor + Normalize, y_c = (x_c/255 - mean_c) * inv_std_c, into the (already BN-folded)
// stem weights  w[k][n] (k = ci*9 + tap)  ->  bf16 image W'[n][k] = w[k][n] * inv_std_ci / 255  and
// b'[n] = b[n] - sum_k w[k][n] * mean_ci * inv_std_ci.  Exact because the stem convolution has no padding.
__global__ void stem_refold_kernel(const float* __restrict__ w, const float* __restrict__ b, StemIn in,
                                   __nv_bfloat16* __restrict__ img, float* __restrict__ bias) {
    const int n = threadIdx.x;   // 32 output channels
    if (n >= 32) return;
    float acc = b[n];
    for (int k = 0; k < 32; ++k) {
        float v = 0.f;
        if (k < 27) {
            const int ci = k / 9;
            v = w[k * 32 + n] * (in.inv_std[ci] * (1.f / 255.f));
            acc -= w[k * 32 + n] * in.mean[ci] * in.inv_std[ci];
        }
        img[((k >> 3) * 4 + (n >> 3)) * 64 + (n & 7) * 8 + (k & 7)] = __float2bfloat16_rn(v);
    }
    bias[n] = acc;
}

// Stem weights for the gather-free fused kernel (l2d_front_tc.cu): one 32 x 16 B-operand image per kernel row ky with
// k = kx * 4 + c over 4 pixels x RGBX; the 4th pixel and the 4th channel get zero weights (except the two bias slots, below).  w is the BN-folded
// [27][32] table (k = ci*9 + ky*3 + kx).  With `norm` the uint8 ToTensor + Normalize is folded in as in
// stem_refold_kernel and the matching bias is written.
__global__ void stem_pack_rgbx_kernel(const float* __restrict__ w, const float* __restrict__ b, StemIn in, int norm,
                                      __nv_bfloat16* __restrict__ img, float* __restrict__ bias) {
    const int n = threadIdx.x;   // 32 output channels
    if (n >= 32) return;
    float acc = b[n];
    for (int ky = 0; ky < 3; ++ky)
        for (int k = 0; k < 16; ++k) {
            const int kx = k >> 2, c = k & 3;
            float v = 0.f;
            if (kx < 3 && c < 3) {
                const float wv = w[(c * 9 + ky * 3 + kx) * 32 + n];
                v = norm ? wv * (in.inv_std[c] * (1.f / 255.f)) : wv;
                if (norm) acc -= wv * in.mean[c] * in.inv_std[c];
            }
            img[ky * 512 + ((k >> 3) * 4 + (n >> 3)) * 64 + (n & 7) * 8 + (k & 7)] = __float2bfloat16_rn(v);
        }
    // The kernel stores 1.0 in the X channel of every pixel, so the bias rides along in the contraction: its bf16 head in the
    // X slot of pixel 0 of kernel row 0 (k = 3), the remainder in the X slot of pixel 1 (k = 7); relative error 2^-17.
    const __nv_bfloat16 bh = __float2bfloat16_rn(acc);
    const __nv_bfloat16 bl = __float2bfloat16_rn(acc - __bfloat162float(bh));
    img[((n >> 3)) * 64 + (n & 7) * 8 + 3] = bh;
    img[((n >> 3)) * 64 + (n & 7) * 8 + 7] = bl;
    if (bias) bias[n] = acc;
}

cudaError_t launch_stem_pack_rgbx(const float* w, const float* b, const StemIn& in, int norm, bf16* img, float* bias, cudaStream_t s) {
    stem_pack_rgbx_kernel<<<1, 32, 0, s>>>(w, b, in, norm, img, bias);
    return cudaGetLastError();
}

cudaError_t launch_stem_refold(const float* w, const float* b, const StemIn& in, bf16* img, float* bias, cudaStream_t s) {
    stem_refold_kernel<<<1, 32, 0, s>>>(w, b, in, img, bias);
    return cudaGetLastError();
}

cudaError_t launch_fold_umma(const float* w, const float* gamma, const float* var, int nrows, int kdim, int nc, int kc,
                             bf16* out, cudaStream_t s) {
    fold_umma_kernel<<<ceil_div(nrows * kdim, 256), 256, 0, s>>>(w, gamma, var, nrows, kdim, nc, kc, out);
    return cudaGetLastError();
}

// bf16 path: the depthwise kernels multiply with FHFMA.BF16, so the BN-folded depthwise weights [9][C] are rounded to
// bf16.  Every depthwise input of the network is non-negative (post-ReLU) and spatially smooth, so the harmful part of
// the rounding error is its sum over the 9 taps (a gain error on the local mean).  Error-diffused rounding removes it:
// taps are visited from the largest to the smallest magnitude, each is rounded after adding the error carried from the
// previous ones; the sum of the 9 rounded weights then differs from the exact sum by at most half an ulp of the
// SMALLEST tap.  The table is rewritten in place with the bf16-representable values (the kernels' own conversion is
// then exact).  One thread per channel.
__global__ void dw_round_bf16_kernel(float* __restrict__ wd, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float w[9];
    int order[9];
    for (int t = 0; t < 9; ++t) { w[t] = wd[t * C + c]; order[t] = t; }
    for (int i = 0; i < 9; ++i)            // selection sort by |w| descending (9 elements)
        for (int j = i + 1; j < 9; ++j)
            if (fabsf(w[order[j]]) > fabsf(w[order[i]])) { const int o = order[i]; order[i] = order[j]; order[j] = o; }
    float carry = 0.f;
    for (int i = 0; i < 9; ++i) {
        const int t = order[i];
        const float v = w[t] + carry;
        const float q = __bfloat162float(__float2bfloat16_rn(v));
        carry = v - q;
        wd[t * C + c] = q;
    }
}

cudaError_t launch_dw_round_bf16(float* wd, int c, cudaStream_t s) {
    dw_round_bf16_kernel<<<ceil_div(c, 128), 128, 0, s>>>(wd, c);
    return cudaGetLastError();
}

// Transposed stride-1 bottleneck kernel (bottleneck_s1t_tc.cu): the expand weights are the A operand, cut into chunks of 128
// expanded channels (rows) x (cin + 16) K columns, 8x8 core matrices ordered [k/8][row/8] (LBO = 2048 B, SBO = 128 B).  The 16
// extra K columns carry the BN-folded expand bias as bf16 head + remainder (relative error 2^-17), zeros elsewhere; rows past
// 6*cin (the half-empty last chunk of a 576-channel layer) are zero.  we is the folded fp32 table [cin][cexp].
__global__ void pack_s1t_we_kernel(const float* __restrict__ we, const float* __restrict__ be, int cin, int cexp,
                                   __nv_bfloat16* __restrict__ out) {
    const int KA = cin + 16, nch = (cexp + 127) / 128, total = nch * 128 * KA;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int chunk = i / (128 * KA), rem = i % (128 * KA), n = rem / KA, k = rem % KA;
        const int ch = chunk * 128 + n;
        float v = 0.f;
        if (ch < cexp) {
            if (k < cin) v = we[(size_t)k * cexp + ch];
            else if (k == cin) v = __bfloat162float(__float2bfloat16_rn(be[ch]));
            else if (k == cin + 1) v = be[ch] - __bfloat162float(__float2bfloat16_rn(be[ch]));
        }
        out[(size_t)chunk * 128 * KA + ((size_t)(k >> 3) * 16 + (n >> 3)) * 64 + (n & 7) * 8 + (k & 7)] = __float2bfloat16_rn(v);
    }
}

// per expanded channel a 32-byte record {bf16 w[9], 2 B pad, f32 depthwise bias, 8 B pad}, then f32 project bias [cout]
__global__ void pack_s1t_tab_kernel(const float* __restrict__ wd, const float* __restrict__ bd, const float* __restrict__ bp,
                                    int cexp, int cout, unsigned char* __restrict__ out) {
    const int nrec = (cexp + 127) / 128 * 128;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nrec; c += gridDim.x * blockDim.x) {
        __nv_bfloat16* w = reinterpret_cast<__nv_bfloat16*>(out + (size_t)c * 32);
        for (int t = 0; t < 16; ++t) w[t] = __float2bfloat16_rn(0.f);
        if (c < cexp) {
            for (int t = 0; t < 9; ++t) w[t] = __float2bfloat16_rn(wd[(size_t)t * cexp + c]);
            reinterpret_cast<float*>(out + (size_t)c * 32)[5] = bd[c];
        }
    }
    float* bpo = reinterpret_cast<float*>(out + (size_t)nrec * 32);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cout; i += gridDim.x * blockDim.x) bpo[i] = bp[i];
}

cudaError_t launch_pack_dw_tab(const float* wd, const float* bd, const float* bout, int c, int cout, unsigned char* tab, cudaStream_t s) {
    pack_s1t_tab_kernel<<<8, 128, 0, s>>>(wd, bd, bout, c, cout, tab);
    return cudaGetLastError();
}

cudaError_t launch_pack_s1t(const BneckW& w, int cin, int cout, bf16* we_img, unsigned char* tab, cudaStream_t s) {
    const int cexp = 6 * cin;
    pack_s1t_we_kernel<<<64, 256, 0, s>>>(w.we, w.be, cin, cexp, we_img);
    pack_s1t_tab_kernel<<<8, 128, 0, s>>>(w.wd, w.bd, w.bp, cexp, cout, tab);
    return cudaGetLastError();
}

}  // namespace fscnn
