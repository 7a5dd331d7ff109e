// ffm.cu -- FeatureFusionModule (reference models/fast_scnn.py:190-218) as ONE kernel:
//   out = relu( BN(conv1x1_64->128(higher)) + BN(conv1x1_128->128( relu(BN(DW3x3( up(lower) ))) )) )
// where up() is the size-driven bilinear align_corners=True resize to higher's HxW (:209-212).
// Both 1x1 convolutions are folded into one K = 192 contraction (k < 64: higher's channels,
// k >= 64: the depthwise output), their biases are pre-added at load time, and neither the
// upsampled tensor nor the depthwise output ever reaches HBM: per 32-channel chunk the CTA
// interpolates the 10x18 halo tile into shared memory, runs the depthwise 3x3 on it and feeds the
// result to the contraction as the operand tile.
#include "kernels.h"

namespace fscnn {

constexpr int kFfmDynSmem = 32 * 184 * 4;

template <typename T>
__global__ void __launch_bounds__(kThreads, 2)
ffm_kernel(const T* __restrict__ higher, const T* __restrict__ lower, FfmW w, T* __restrict__ out, int Hh, int Wh, int Hl,
           int Wl) {
    constexpr int KC = 32, COUT = 128, TN = 8, CH = 64, CL = 128;
    constexpr int IW = 18, PIN = 10 * 18, PINP = 184;
    using CM = ColMap<TN>;
    __shared__ __align__(16) float As[KC * 128];
    __shared__ __align__(16) float Bs[KC * COUT];
    extern __shared__ __align__(16) float Us[];   // [KC][PINP] resized halo tile (dynamic: static smem is capped at 48 KB)
    __shared__ float Wds[9 * KC];
    __shared__ float Bds[KC];
    // per halo pixel: the two source rows / columns and weights of the bilinear resize (or -1 outside the image)
    __shared__ int sy0[PINP], sx0[PINP], sy1[PINP], sx1[PINP];
    __shared__ float sly[PINP], slx[PINP];

    const int tid = threadIdx.x, n = blockIdx.z;
    const int oy0 = blockIdx.y * 8, ox0 = blockIdx.x * 16;
    const int tn = tid & 15, tp = tid >> 4;

    const float scy = Hh > 1 ? (float)(Hl - 1) / (float)(Hh - 1) : 0.f;
    const float scx = Wh > 1 ? (float)(Wl - 1) / (float)(Wh - 1) : 0.f;
    for (int pin = tid; pin < PINP; pin += kThreads) {
        const int y = oy0 - 1 + pin / IW, x = ox0 - 1 + pin % IW;
        if (pin < PIN && y >= 0 && y < Hh && x >= 0 && x < Wh) {
            const float fy = scy * (float)y, fx = scx * (float)x;
            const int y0 = min((int)fy, Hl - 1), x0 = min((int)fx, Wl - 1);
            sy0[pin] = y0; sy1[pin] = min(y0 + 1, Hl - 1); sly[pin] = fy - (float)y0;
            sx0[pin] = x0; sx1[pin] = min(x0 + 1, Wl - 1); slx[pin] = fx - (float)x0;
        } else {
            sy0[pin] = -1; sy1[pin] = 0; sx0[pin] = 0; sx1[pin] = 0; sly[pin] = 0.f; slx[pin] = 0.f;
        }
    }

    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

#pragma unroll 1
    for (int kc0 = 0; kc0 < CH + CL; kc0 += KC) {
        __syncthreads();   // coordinate tables ready / previous contraction done
        load_weight_tile<KC, COUT>(Bs, w.wcat + (size_t)kc0 * COUT, COUT);
        if (kc0 < CH) {
            // operand = higher's channels kc0..kc0+31 (lanes along channels, swizzled transposing store)
            const int cv = tid & 7, pl = tid >> 3;
            for (int p = pl; p < 128; p += 32) {
                const int oy = oy0 + (p >> 4), ox = ox0 + (p & 15);
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (oy < Hh && ox < Wh) v = Act<T>::ld4(higher + (((size_t)n * Hh + oy) * Wh + ox) * CH + kc0 + 4 * cv);
                const int col = p ^ (cv << 2);
                As[(4 * cv + 0) * 128 + col] = v.x;
                As[(4 * cv + 1) * 128 + col] = v.y;
                As[(4 * cv + 2) * 128 + col] = v.z;
                As[(4 * cv + 3) * 128 + col] = v.w;
            }
        } else {
            const int c0 = kc0 - CH;   // lower-branch channels c0..c0+31
            for (int i = tid; i < 9 * KC; i += kThreads) Wds[i] = __ldg(w.wd + (i / KC) * CL + c0 + (i % KC));
            if (tid < KC) Bds[tid] = __ldg(w.bd + c0 + tid);
            // bilinear resize of the halo tile; lanes run along pixels so the stores are conflict free
            for (int i = tid; i < 8 * PINP; i += kThreads) {
                const int pin = i % PINP, cv = i / PINP;
                float4 u = make_float4(0.f, 0.f, 0.f, 0.f);
                const int y0 = sy0[pin];
                if (y0 >= 0) {
                    const int y1 = sy1[pin], x0 = sx0[pin], x1 = sx1[pin];
                    const float ly = sly[pin], lx = slx[pin], hy = 1.f - ly, hx = 1.f - lx;
                    const T* base = lower + (size_t)n * Hl * Wl * CL + c0 + 4 * cv;
                    const float4 v00 = Act<T>::ld4(base + ((size_t)y0 * Wl + x0) * CL);
                    const float4 v01 = Act<T>::ld4(base + ((size_t)y0 * Wl + x1) * CL);
                    const float4 v10 = Act<T>::ld4(base + ((size_t)y1 * Wl + x0) * CL);
                    const float4 v11 = Act<T>::ld4(base + ((size_t)y1 * Wl + x1) * CL);
                    u.x = hy * (hx * v00.x + lx * v01.x) + ly * (hx * v10.x + lx * v11.x);
                    u.y = hy * (hx * v00.y + lx * v01.y) + ly * (hx * v10.y + lx * v11.y);
                    u.z = hy * (hx * v00.z + lx * v01.z) + ly * (hx * v10.z + lx * v11.z);
                    u.w = hy * (hx * v00.w + lx * v01.w) + ly * (hx * v10.w + lx * v11.w);
                }
                Us[(4 * cv + 0) * PINP + pin] = u.x;
                Us[(4 * cv + 1) * PINP + pin] = u.y;
                Us[(4 * cv + 2) * PINP + pin] = u.z;
                Us[(4 * cv + 3) * PINP + pin] = u.w;
            }
            __syncthreads();
            // depthwise 3x3 + ReLU -> operand tile (lanes along pixels; swizzle matches contract_chunk)
            for (int i = tid; i < KC * 128; i += kThreads) {
                const int p = i & 127, c = i >> 7;
                const float* up = Us + c * PINP + (p >> 4) * IW + (p & 15);
                float d = Bds[c];
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) d = fmaf(up[ky * IW + kx], Wds[(ky * 3 + kx) * KC + c], d);
                As[c * 128 + (p ^ swz(c))] = relu(d);
            }
        }
        __syncthreads();
        contract_chunk<KC, TN, 128, COUT, true>(acc, As, Bs, tp, tn);
    }

#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int p = 8 * tp + i;
        const int oy = oy0 + (p >> 4), ox = ox0 + (p & 15);
        if (oy >= Hh || ox >= Wh) continue;
        T* o = out + (((size_t)n * Hh + oy) * Wh + ox) * COUT;
#pragma unroll
        for (int q = 0; q < CM::NQ; ++q) {
            const int ch = CM::ch(tn, q, 0);
            const float4 b = __ldg(reinterpret_cast<const float4*>(w.bcat + ch));
            Act<T>::st4(o + ch, make_float4(relu(acc[i][q * 4 + 0] + b.x), relu(acc[i][q * 4 + 1] + b.y),
                                            relu(acc[i][q * 4 + 2] + b.z), relu(acc[i][q * 4 + 3] + b.w)));
        }
    }
}

template <typename T>
cudaError_t launch_ffm(const T* higher, const T* lower, const FfmW& w, T* out, int n, int hh, int wh, int hl, int wl,
                       cudaStream_t s) {
    static unsigned long long configured = 0;
    cudaError_t e = ensure_dyn_smem(ffm_kernel<T>, kFfmDynSmem, configured);
    if (e != cudaSuccess) return e;
    dim3 grid(ceil_div(wh, 16), ceil_div(hh, 8), n);
    ffm_kernel<T><<<grid, kThreads, kFfmDynSmem, s>>>(higher, lower, w, out, hh, wh, hl, wl);
    return cudaGetLastError();
}

template cudaError_t launch_ffm<float>(const float*, const float*, const FfmW&, float*, int, int, int, int, int, cudaStream_t);
template cudaError_t launch_ffm<bf16>(const bf16*, const bf16*, const FfmW&, bf16*, int, int, int, int, int, cudaStream_t);

}  // namespace fscnn
