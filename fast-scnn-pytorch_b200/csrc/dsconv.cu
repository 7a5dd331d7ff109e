// dsconv.cu -- fused depthwise-separable block (_DSConv): DW 3x3 (stride s, pad 1) + BN + ReLU,
// then PW 1x1 + BN + ReLU, in one kernel; the depthwise result lives only in shared memory.
// Replaces reference models/fast_scnn.py:64-79, used at :154-155 (LearningToDownsample) and
// :226-227 (Classifer); with HEAD the classifier's final Conv2d(128, nc, 1) (:228-231) is chained
// in the same kernel and only the low-resolution logits are written.
//
// CTA = 8x16 output pixels (128) x all COUT channels.  The input-channel dimension is processed in
// chunks of KC: the depthwise phase reads NHWC vectors straight from global/L1 (lanes run along
// channels, coalesced) and writes the transposed, swizzled operand tile As[k][pixel]; the pointwise
// phase is the 16x16 register-tile contraction of common.cuh.
#include "kernels.h"

namespace fscnn {

template <typename T, int CIN, int COUT, int STRIDE, bool HEAD>
__global__ void __launch_bounds__(kThreads, 2)
dsconv_kernel(const T* __restrict__ in, DsW w, T* __restrict__ out, HeadW head, float* __restrict__ logits,
              int Hi, int Wi, int Ho, int Wo) {
    constexpr int KC = (CIN % 32 == 0) ? 32 : 16;
    constexpr int NV = KC / 4;            // channel vectors per chunk
    constexpr int PP = kThreads / NV;     // pixels per depthwise pass
    constexpr int TN = COUT / 16;
    using CM = ColMap<TN>;
    static_assert(CIN % KC == 0 && COUT % 16 == 0, "unsupported channel counts");

    __shared__ __align__(16) float As[KC * 128];
    __shared__ __align__(16) float Bs[KC * COUT];
    __shared__ __align__(16) float Wds[9 * KC];
    __shared__ __align__(16) float Bds[KC];
    extern __shared__ __align__(16) float dyn[];   // HEAD: Cs[128][COUT+1] then Wh[COUT][ncp]

    const int tid = threadIdx.x;
    const int n = blockIdx.z;
    const int oy0 = blockIdx.y * 8, ox0 = blockIdx.x * 16;
    const int tn = tid & 15, tp = tid >> 4;
    const int cv = tid % NV, pl = tid / NV;

    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    for (int kc0 = 0; kc0 < CIN; kc0 += KC) {
        if (kc0) __syncthreads();   // previous chunk's contraction has finished reading As/Bs
        load_weight_tile<KC, COUT>(Bs, w.wp + (size_t)kc0 * COUT, COUT);
        for (int i = tid; i < 9 * KC; i += kThreads) Wds[i] = __ldg(w.wd + (i / KC) * CIN + kc0 + (i % KC));
        if (tid < KC) Bds[tid] = __ldg(w.bd + kc0 + tid);
        __syncthreads();
        // ---- depthwise phase ----
#pragma unroll 1
        for (int p = pl; p < 128; p += PP) {
            const int oy = oy0 + (p >> 4), ox = ox0 + (p & 15);
            float4 a = *reinterpret_cast<const float4*>(Bds + 4 * cv);
            const bool live = (oy < Ho) && (ox < Wo);
            if (live) {
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    const int iy = oy * STRIDE - 1 + ky;
                    if (iy < 0 || iy >= Hi) continue;
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const int ix = ox * STRIDE - 1 + kx;
                        if (ix < 0 || ix >= Wi) continue;
                        const float4 v = Act<T>::ld4(in + (((size_t)n * Hi + iy) * Wi + ix) * CIN + kc0 + 4 * cv);
                        const float4 k4 = *reinterpret_cast<const float4*>(Wds + (ky * 3 + kx) * KC + 4 * cv);
                        a.x = fmaf(v.x, k4.x, a.x); a.y = fmaf(v.y, k4.y, a.y);
                        a.z = fmaf(v.z, k4.z, a.z); a.w = fmaf(v.w, k4.w, a.w);
                    }
                }
                a.x = relu(a.x); a.y = relu(a.y); a.z = relu(a.z); a.w = relu(a.w);
            } else {
                a = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            const int col = p ^ ((cv & 7) << 2);   // == p ^ swz(4*cv + j) for j < 4
            As[(4 * cv + 0) * 128 + col] = a.x;
            As[(4 * cv + 1) * 128 + col] = a.y;
            As[(4 * cv + 2) * 128 + col] = a.z;
            As[(4 * cv + 3) * 128 + col] = a.w;
        }
        __syncthreads();
        // ---- pointwise phase ----
        contract_chunk<KC, TN, 128, COUT, true>(acc, As, Bs, tp, tn);
    }

    // ---- epilogue ----
    float bias[TN];
#pragma unroll
    for (int q = 0; q < CM::NQ; ++q)
#pragma unroll
        for (int j = 0; j < CM::VW; ++j) bias[q * CM::VW + j] = __ldg(w.bp + CM::ch(tn, q, j));

    if (!HEAD) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int p = 8 * tp + i;
            const int oy = oy0 + (p >> 4), ox = ox0 + (p & 15);
            if (oy >= Ho || ox >= Wo) continue;
            T* o = out + (((size_t)n * Ho + oy) * Wo + ox) * COUT;
#pragma unroll
            for (int q = 0; q < CM::NQ; ++q) {
                float v[CM::VW];
#pragma unroll
                for (int j = 0; j < CM::VW; ++j) v[j] = relu(acc[i][q * CM::VW + j] + bias[q * CM::VW + j]);
                store_vec<T, CM::VW>(o + CM::ch(tn, q, 0), v);
            }
        }
    } else {
        // chained 1x1 to the class logits: stage the activated tile as Cs[pixel][COUT+1]
        constexpr int LDC = COUT + 1;
        float* Cs = dyn;
        float* Whs = dyn + 128 * LDC;       // 128*LDC floats is a multiple of 16 bytes
        const int ncp = head.ncp;
        for (int i = tid; i < COUT * ncp; i += kThreads) Whs[i] = __ldg(head.w + i);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int q = 0; q < CM::NQ; ++q)
#pragma unroll
                for (int j = 0; j < CM::VW; ++j)
                    Cs[(8 * tp + i) * LDC + CM::ch(tn, q, j)] = relu(acc[i][q * CM::VW + j] + bias[q * CM::VW + j]);
        __syncthreads();
        const int p = tid & 127, half = tid >> 7;
        const int oy = oy0 + (p >> 4), ox = ox0 + (p & 15);
        const bool live = (oy < Ho) && (ox < Wo);
        for (int g = half; g < ncp / 4; g += 2) {
            float4 l = __ldg(reinterpret_cast<const float4*>(head.b) + g);
#pragma unroll 8
            for (int k = 0; k < COUT; ++k) {
                const float a = Cs[p * LDC + k];
                const float4 b = *reinterpret_cast<const float4*>(Whs + k * ncp + 4 * g);
                l.x = fmaf(a, b.x, l.x); l.y = fmaf(a, b.y, l.y); l.z = fmaf(a, b.z, l.z); l.w = fmaf(a, b.w, l.w);
            }
            if (live) *reinterpret_cast<float4*>(logits + (((size_t)n * Ho + oy) * Wo + ox) * ncp + 4 * g) = l;
        }
    }
}

template <typename T, int CIN, int COUT, int STRIDE, bool HEAD>
static cudaError_t run(const T* in, const DsW& w, T* out, const HeadW* head, float* logits, int n, int hi, int wi, int ho,
                       int wo, cudaStream_t s) {
    dim3 grid(ceil_div(wo, 16), ceil_div(ho, 8), n);
    size_t dyn = 0;
    HeadW h{};
    if (HEAD) {
        h = *head;
        dyn = (size_t)(128 * (COUT + 1) + COUT * h.ncp) * sizeof(float);
        static unsigned long long configured = 0;
        static size_t configured_bytes = 0;
        if (dyn > configured_bytes) { configured = 0; configured_bytes = dyn; }   // a model with more classes came along
        cudaError_t e = ensure_dyn_smem(dsconv_kernel<T, CIN, COUT, STRIDE, HEAD>, configured_bytes, configured);
        if (e != cudaSuccess) return e;
    }
    dsconv_kernel<T, CIN, COUT, STRIDE, HEAD><<<grid, kThreads, dyn, s>>>(in, w, out, h, logits, hi, wi, ho, wo);
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_dsconv(int cin, int cout, int stride, const T* in, const DsW& w, T* out, const HeadW* head,
                          float* logits, int n, int hi, int wi, int ho, int wo, cudaStream_t s) {
    if (cin == 32 && cout == 48 && stride == 2 && !head) return run<T, 32, 48, 2, false>(in, w, out, head, logits, n, hi, wi, ho, wo, s);
    if (cin == 48 && cout == 64 && stride == 2 && !head) return run<T, 48, 64, 2, false>(in, w, out, head, logits, n, hi, wi, ho, wo, s);
    if (cin == 128 && cout == 128 && stride == 1 && !head) return run<T, 128, 128, 1, false>(in, w, out, head, logits, n, hi, wi, ho, wo, s);
    if (cin == 128 && cout == 128 && stride == 1 && head) return run<T, 128, 128, 1, true>(in, w, out, head, logits, n, hi, wi, ho, wo, s);
    return cudaErrorInvalidValue;
}

template cudaError_t launch_dsconv<float>(int, int, int, const float*, const DsW&, float*, const HeadW*, float*, int, int,
                                          int, int, int, cudaStream_t);
template cudaError_t launch_dsconv<bf16>(int, int, int, const bf16*, const DsW&, bf16*, const HeadW*, float*, int, int, int,
                                         int, int, cudaStream_t);

}  // namespace fscnn
