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

}  // namespace fscnn
