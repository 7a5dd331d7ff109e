// aux.cu -- the optional auxiliary head (reference models/fast_scnn.py:24-31, 42-44):
// Conv2d(64, 32, 3, padding=1) + BN + ReLU, Dropout (identity in eval), Conv2d(32, nc, 1).
// The dense 3x3 runs as an implicit contraction with K = 9 taps x 64 channels in chunks of 32
// (one tap, half the channels): the operand tile is the input shifted by the tap, zero outside the
// image.  Output: low-resolution logits [n][h][w][ncp] fp32, upsampled later by up_logits.
#include "kernels.h"

namespace fscnn {

template <typename T>
__global__ void __launch_bounds__(kThreads, 2)
aux_kernel(const T* __restrict__ in, AuxW w, float* __restrict__ logits, int H, int W) {
    constexpr int KC = 32, CIN = 64, CMID = 32, TN = 2, LDC = CMID + 1;
    __shared__ __align__(16) float As[KC * 128];
    __shared__ __align__(16) float Bs[KC * CMID];
    __shared__ float Cs[128 * LDC];
    const int tid = threadIdx.x, n = blockIdx.z;
    const int oy0 = blockIdx.y * 8, ox0 = blockIdx.x * 16;
    const int tn = tid & 15, tp = tid >> 4;
    const int cv = tid & 7, pl = tid >> 3;

    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

#pragma unroll 1
    for (int kk = 0; kk < 18; ++kk) {
        const int tap = kk >> 1, c0 = (kk & 1) * 32;
        const int dy = tap / 3 - 1, dx = tap % 3 - 1;
        if (kk) __syncthreads();
        load_weight_tile<KC, CMID>(Bs, w.w + (size_t)(tap * CIN + c0) * CMID, CMID);
        for (int p = pl; p < 128; p += 32) {
            const int iy = oy0 + (p >> 4) + dy, ix = ox0 + (p & 15) + dx;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = Act<T>::ld4(in + (((size_t)n * H + iy) * W + ix) * CIN + c0 + 4 * cv);
            const int col = p ^ (cv << 2);
            As[(4 * cv + 0) * 128 + col] = v.x;
            As[(4 * cv + 1) * 128 + col] = v.y;
            As[(4 * cv + 2) * 128 + col] = v.z;
            As[(4 * cv + 3) * 128 + col] = v.w;
        }
        __syncthreads();
        contract_chunk<KC, TN, 128, CMID, true>(acc, As, Bs, tp, tn);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int co = ColMap<TN>::ch(tn, 0, j);
            Cs[(8 * tp + i) * LDC + co] = relu(acc[i][j] + __ldg(w.b + co));
        }
    __syncthreads();
    const int p = tid & 127, half = tid >> 7;
    const int oy = oy0 + (p >> 4), ox = ox0 + (p & 15);
    const int ncp = w.head.ncp;
    for (int g = half; g < ncp / 4; g += 2) {
        float4 l = __ldg(reinterpret_cast<const float4*>(w.head.b) + g);
#pragma unroll 8
        for (int k = 0; k < CMID; ++k) {
            const float a = Cs[p * LDC + k];
            const float4 b = __ldg(reinterpret_cast<const float4*>(w.head.w + k * ncp) + g);
            l.x = fmaf(a, b.x, l.x); l.y = fmaf(a, b.y, l.y); l.z = fmaf(a, b.z, l.z); l.w = fmaf(a, b.w, l.w);
        }
        if (oy < H && ox < W) *reinterpret_cast<float4*>(logits + (((size_t)n * H + oy) * W + ox) * ncp + 4 * g) = l;
    }
}

template <typename T>
cudaError_t launch_aux(const T* higher, const AuxW& w, float* logits, int n, int h, int wd, cudaStream_t s) {
    dim3 grid(ceil_div(wd, 16), ceil_div(h, 8), n);
    aux_kernel<T><<<grid, kThreads, 0, s>>>(higher, w, logits, h, wd);
    return cudaGetLastError();
}

template cudaError_t launch_aux<float>(const float*, const AuxW&, float*, int, int, int, cudaStream_t);
template cudaError_t launch_aux<bf16>(const bf16*, const AuxW&, float*, int, int, int, cudaStream_t);

}  // namespace fscnn
