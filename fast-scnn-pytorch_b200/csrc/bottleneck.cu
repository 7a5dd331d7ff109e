// bottleneck.cu -- fused LinearBottleneck: PW expand (c -> 6c) + BN + ReLU, DW 3x3 (stride s, pad 1)
// + BN + ReLU, PW project (6c -> c') + BN, optional residual -- one kernel, and the 6c-wide expanded
// tensor never leaves shared memory.  Replaces reference models/fast_scnn.py:95-115 (built nine
// times by GlobalFeatureExtractor._make_layer, :170-180).
//
// CTA = TH x 16 output pixels (8x16 for stride 1, 4x16 for stride 2).  The input tile with its 3x3
// halo is staged once as Xs[cin][pixel].  The expanded dimension is walked in chunks of CE = 32
// channels: expand contraction over the halo tile -> Es (zeroed outside the image, because the
// depthwise conv pads the EXPANDED tensor with zeros) -> depthwise -> Ds[ce][pixel] -> project
// contraction accumulated in registers across chunks (the project conv is linear).
#include "kernels.h"

namespace fscnn {

template <int CIN, int COUT, int STRIDE>
struct BneckCfg {
    static constexpr int TH = (STRIDE == 1) ? 8 : 4;
    static constexpr int TW = 16;
    static constexpr int P = TH * TW;                  // output pixels per CTA
    static constexpr int IH = (TH - 1) * STRIDE + 3;
    static constexpr int IW = (TW - 1) * STRIDE + 3;
    static constexpr int PIN = IH * IW;                // input (halo) pixels per CTA
    static constexpr int PINP = round_up(PIN, 8);
    static constexpr int CE = 32;
    static constexpr int CEXP = 6 * CIN;
    static constexpr int TM = P / 16;                  // project tile: TM pixels x TN channels
    static constexpr int TN = COUT / 16;
    static constexpr int EM = (STRIDE == 1) ? 8 : 4;   // expand tile: EM pixels x 4 channels
    // shared-memory carve-up (floats)
    static constexpr int oXs = 0;
    static constexpr int oEs = oXs + CIN * PINP;
    static constexpr int oDs = oEs + CE * PINP;
    static constexpr int oWe = oDs + CE * P;
    static constexpr int oWp = oWe + CIN * CE;
    static constexpr int oWd = oWp + CE * COUT;
    static constexpr int oBe = oWd + 9 * CE;
    static constexpr int oBd = oBe + CE;
    static constexpr int oEnd = oBd + CE;
    static constexpr size_t smem_bytes = (size_t)oEnd * 4 + PINP;   // + validity bytes
};

template <typename T, int CIN, int COUT, int STRIDE, bool RES>
__global__ void __launch_bounds__(kThreads, 1)
bottleneck_kernel(const T* __restrict__ in, BneckW w, T* __restrict__ out, int Hi, int Wi, int Ho, int Wo) {
    using C = BneckCfg<CIN, COUT, STRIDE>;
    using CM = ColMap<C::TN>;
    constexpr int P = C::P, PINP = C::PINP, CE = C::CE, IW = C::IW, TM = C::TM, TN = C::TN, EM = C::EM;
    extern __shared__ __align__(16) float sm[];
    float* Xs = sm + C::oXs;
    float* Es = sm + C::oEs;
    float* Ds = sm + C::oDs;
    float* Wes = sm + C::oWe;
    float* Wps = sm + C::oWp;
    float* Wds = sm + C::oWd;
    float* Bes = sm + C::oBe;
    float* Bds = sm + C::oBd;
    unsigned char* valid = reinterpret_cast<unsigned char*>(sm + C::oEnd);

    const int tid = threadIdx.x;
    const int n = blockIdx.z;
    const int oy0 = blockIdx.y * C::TH, ox0 = blockIdx.x * C::TW;
    const int iy0 = oy0 * STRIDE - 1, ix0 = ox0 * STRIDE - 1;
    const int tn = tid & 15, tp = tid >> 4;

    // ---- stage the input halo tile, transposed: Xs[c][pin]; lanes run along pixels ----
    for (int pin = tid; pin < PINP; pin += kThreads) {
        const int iy = iy0 + pin / IW, ix = ix0 + pin % IW;
        valid[pin] = (pin < C::PIN && iy >= 0 && iy < Hi && ix >= 0 && ix < Wi) ? 1 : 0;
    }
    {
        constexpr int NV = CIN / 4;
        for (int i = tid; i < NV * PINP; i += kThreads) {
            const int pin = i % PINP, cv = i / PINP;
            const int iy = iy0 + pin / IW, ix = ix0 + pin % IW;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pin < C::PIN && iy >= 0 && iy < Hi && ix >= 0 && ix < Wi)
                v = Act<T>::ld4(in + (((size_t)n * Hi + iy) * Wi + ix) * CIN + 4 * cv);
            Xs[(4 * cv + 0) * PINP + pin] = v.x;
            Xs[(4 * cv + 1) * PINP + pin] = v.y;
            Xs[(4 * cv + 2) * PINP + pin] = v.z;
            Xs[(4 * cv + 3) * PINP + pin] = v.w;
        }
    }

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

#pragma unroll 1
    for (int e0 = 0; e0 < C::CEXP; e0 += CE) {
        __syncthreads();   // Xs ready (first pass) / previous chunk's project contraction done
        // ---- weights of this chunk ----
        load_weight_tile<CIN, CE>(Wes, w.we + e0, C::CEXP);
        load_weight_tile<CE, COUT>(Wps, w.wp + (size_t)e0 * COUT, COUT);
        for (int i = tid; i < 9 * CE; i += kThreads) Wds[i] = __ldg(w.wd + (i / CE) * C::CEXP + e0 + (i % CE));
        if (tid < CE) { Bes[tid] = __ldg(w.be + e0 + tid); Bds[tid] = __ldg(w.bd + e0 + tid); }
        __syncthreads();

        // ---- expand: Es[ce][pin] = relu(be + sum_k Xs[k][pin] * We[k][ce]), 0 outside the image ----
        {
            constexpr int NT = (PINP / EM) * (CE / 4);
#pragma unroll 1
            for (int t = tid; t < NT; t += kThreads) {
                const int en = t & 7, em = t >> 3;
                float e[EM][4];
#pragma unroll
                for (int i = 0; i < EM; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) e[i][j] = Bes[4 * en + j];
#pragma unroll 8
                for (int k = 0; k < CIN; ++k) {
                    float a[EM];
                    const float4 a0 = *reinterpret_cast<const float4*>(Xs + k * PINP + EM * em);
                    a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
                    if (EM == 8) {
                        const float4 a1 = *reinterpret_cast<const float4*>(Xs + k * PINP + EM * em + 4);
                        a[EM - 4] = a1.x; a[EM - 3] = a1.y; a[EM - 2] = a1.z; a[EM - 1] = a1.w;
                    }
                    const float4 b = *reinterpret_cast<const float4*>(Wes + k * CE + 4 * en);
#pragma unroll
                    for (int i = 0; i < EM; ++i) {
                        e[i][0] = fmaf(a[i], b.x, e[i][0]); e[i][1] = fmaf(a[i], b.y, e[i][1]);
                        e[i][2] = fmaf(a[i], b.z, e[i][2]); e[i][3] = fmaf(a[i], b.w, e[i][3]);
                    }
                }
#pragma unroll
                for (int i = 0; i < EM; ++i) {
                    const int pin = EM * em + i;
                    const bool ok = valid[pin];
#pragma unroll
                    for (int j = 0; j < 4; ++j) Es[(4 * en + j) * PINP + pin] = ok ? relu(e[i][j]) : 0.f;
                }
            }
        }
        __syncthreads();

        // ---- depthwise: Ds[ce][p] = relu(bd + sum_taps Es[ce][nbr] * Wd[tap][ce]); lanes along pixels ----
        for (int i = tid; i < CE * P; i += kThreads) {
            const int p = i % P, ce = i / P;
            const int py = p / C::TW, px = p % C::TW;
            const float* ep = Es + ce * PINP + (py * STRIDE) * IW + px * STRIDE;
            float d = Bds[ce];
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) d = fmaf(ep[ky * IW + kx], Wds[(ky * 3 + kx) * CE + ce], d);
            Ds[ce * P + p] = relu(d);
        }
        __syncthreads();

        // ---- project: acc[p][co] += sum_ce Ds[ce][p] * Wp[ce][co] ----
#pragma unroll 8
        for (int k = 0; k < CE; ++k) {
            float a[TM];
            const float4 a0 = *reinterpret_cast<const float4*>(Ds + k * P + TM * tp);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
            if (TM == 8) {
                const float4 a1 = *reinterpret_cast<const float4*>(Ds + k * P + TM * tp + 4);
                a[TM - 4] = a1.x; a[TM - 3] = a1.y; a[TM - 2] = a1.z; a[TM - 1] = a1.w;
            }
            float b[TN];
#pragma unroll
            for (int q = 0; q < CM::NQ; ++q) {
                const float* bp = Wps + k * COUT + q * 16 * CM::VW + tn * CM::VW;
                if (CM::VW == 4) {
                    const float4 v = *reinterpret_cast<const float4*>(bp);
                    b[q * 4 + 0] = v.x; b[q * 4 + 1] = v.y; b[q * 4 + 2] = v.z; b[q * 4 + 3] = v.w;
                } else if (CM::VW == 2) {
                    const float2 v = *reinterpret_cast<const float2*>(bp);
                    b[q * 2 + 0] = v.x; b[q * 2 + 1] = v.y;
                } else {
                    b[q] = *bp;
                }
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
    }

    // ---- epilogue: + folded BN bias (+ residual from the staged input tile), linear output ----
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int p = TM * tp + i;
        const int py = p / C::TW, px = p % C::TW;
        const int oy = oy0 + py, ox = ox0 + px;
        if (oy >= Ho || ox >= Wo) continue;
        T* o = out + (((size_t)n * Ho + oy) * Wo + ox) * COUT;
#pragma unroll
        for (int q = 0; q < CM::NQ; ++q) {
            float v[CM::VW];
#pragma unroll
            for (int j = 0; j < CM::VW; ++j) {
                const int co = CM::ch(tn, q, j);
                v[j] = acc[i][q * CM::VW + j] + __ldg(w.bp + co);
                if (RES) v[j] += Xs[co * PINP + (py + 1) * IW + (px + 1)];   // stride 1: centre of the halo tile
            }
            store_vec<T, CM::VW>(o + CM::ch(tn, q, 0), v);
        }
    }
}

template <typename T, int CIN, int COUT, int STRIDE, bool RES>
static cudaError_t run(const T* in, const BneckW& w, T* out, int n, int hi, int wi, int ho, int wo, cudaStream_t s) {
    using C = BneckCfg<CIN, COUT, STRIDE>;
    static unsigned long long configured = 0;
    cudaError_t e = ensure_dyn_smem(bottleneck_kernel<T, CIN, COUT, STRIDE, RES>, C::smem_bytes, configured);
    if (e != cudaSuccess) return e;
    dim3 grid(ceil_div(wo, C::TW), ceil_div(ho, C::TH), n);
    bottleneck_kernel<T, CIN, COUT, STRIDE, RES><<<grid, kThreads, C::smem_bytes, s>>>(in, w, out, hi, wi, ho, wo);
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_bottleneck(int cin, int cout, int stride, const T* in, const BneckW& w, T* out, int n, int hi, int wi,
                              int ho, int wo, cudaStream_t s) {
    if (cin == 64 && cout == 64 && stride == 2) return run<T, 64, 64, 2, false>(in, w, out, n, hi, wi, ho, wo, s);
    if (cin == 64 && cout == 64 && stride == 1) return run<T, 64, 64, 1, true>(in, w, out, n, hi, wi, ho, wo, s);
    if (cin == 64 && cout == 96 && stride == 2) return run<T, 64, 96, 2, false>(in, w, out, n, hi, wi, ho, wo, s);
    if (cin == 96 && cout == 96 && stride == 1) return run<T, 96, 96, 1, true>(in, w, out, n, hi, wi, ho, wo, s);
    if (cin == 96 && cout == 128 && stride == 1) return run<T, 96, 128, 1, false>(in, w, out, n, hi, wi, ho, wo, s);
    if (cin == 128 && cout == 128 && stride == 1) return run<T, 128, 128, 1, true>(in, w, out, n, hi, wi, ho, wo, s);
    return cudaErrorInvalidValue;
}

template cudaError_t launch_bottleneck<float>(int, int, int, const float*, const BneckW&, float*, int, int, int, int, int,
                                              cudaStream_t);
template cudaError_t launch_bottleneck<bf16>(int, int, int, const bf16*, const BneckW&, bf16*, int, int, int, int, int,
                                             cudaStream_t);

}  // namespace fscnn
