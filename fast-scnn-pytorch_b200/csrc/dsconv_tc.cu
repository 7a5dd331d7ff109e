// dsconv_tc.cu -- bf16 _DSConv on the tensor cores: DW 3x3 (stride s, pad 1) + BN + ReLU on the CUDA
// cores (fp32), then the pointwise 1x1 + BN + ReLU as tcgen05.mma with the accumulator in TMEM.
// Replaces reference models/fast_scnn.py:64-79 for LearningToDownsample.dsconv1/2 (:154-155) and
// Classifer.dsconv1/2 (:226-227); with HEAD the classifier's Conv2d(128, nc, 1) (:228-231) runs as a
// second MMA on the activated tile and only fp32 low-resolution logits are written.
//
// CTA = 8x16 output pixels (M = 128), 256 threads:
//   halo tile  <- cp.async (coalesced 16-byte pieces, zero fill outside the image), rows swizzled
//   depthwise  : thread = (column, row strip, 8-channel chunk), fp32 accumulate -> bf16 A-operand tile
//   pointwise  : CIN/16 MMAs [128 x COUT]; weights arrive by one bulk copy of the pre-packed image
//   epilogue   : TMEM -> bias, ReLU -> bf16 NHWC store   (HEAD: -> A-operand tile -> MMA -> logits)
#include "kernels.h"
#include "umma.cuh"

namespace fscnn {

template <int CIN, int COUT, int STRIDE, bool HEAD>
struct DsTcCfg {
    static constexpr int TH = 8, TW = 16;
    static constexpr int IH = (TH - 1) * STRIDE + 3, IW = (TW - 1) * STRIDE + 3;
    static constexpr int PIN = IH * IW, PINP = round_up(PIN, 8);
    static constexpr int NCK = CIN / 8;                 // 16-byte channel chunks per pixel
    static constexpr int ROWB = CIN * 2;                // bytes per halo pixel
    static constexpr int RPS = (CIN >= 128) ? 4 : 2;    // output rows per depthwise strip
    static constexpr int NSTRIP = 16 * (TH / RPS) * NCK;
    static constexpr int H_BYTES = round_up(PINP * ROWB, 128);
    static constexpr int A2_BYTES = HEAD ? 128 * COUT * 2 : 0;
    // CIN = 128: the weight image is bulk-copied over the halo tile once the depthwise phase is done with it, and the
    // head's second A operand re-uses the first one's tile, so two CTAs fit in one SM's shared memory.
    static constexpr bool ALIAS_B = (CIN >= 128);
    static_assert(!HEAD || (ALIAS_B && COUT <= CIN), "HEAD re-uses the A tile");
    static constexpr int oH = 0;
    static constexpr int B_BYTES = COUT * CIN * 2;
    static constexpr int oA = oH + (ALIAS_B ? (H_BYTES > B_BYTES ? H_BYTES : B_BYTES) : H_BYTES);
    static constexpr int oB = ALIAS_B ? oH : oA + 128 * CIN * 2;
    static constexpr int oWd = ALIAS_B ? oA + 128 * CIN * 2 : oB + B_BYTES;
    static constexpr int oBd = oWd + 9 * CIN * 4;
    static constexpr int oBp = oBd + CIN * 4;
    static constexpr int oB2 = round_up(oBp + COUT * 4, 128);   // HEAD: head weight image (ncp16 x COUT), sized at run time
    static_assert(CIN % 16 == 0 && COUT % 16 == 0, "shape");
};

// bank-conflict avoiding permutation of a halo pixel's 16-byte chunks
template <int CIN>
__device__ __forceinline__ int chunk_swz(int pin, int k8) {
    if (CIN >= 64) return k8 ^ (pin & 7);
    if (CIN == 32) return k8 ^ ((pin >> 1) & 3);
    return k8;
}

template <int CIN, int COUT, int STRIDE, bool HEAD>
__global__ void __launch_bounds__(kThreads, 2)
dsconv_tc_kernel(const bf16* __restrict__ in, DsW w, const bf16* __restrict__ wp_img, bf16* __restrict__ out, HeadW head,
                 const bf16* __restrict__ wh_img, int ncp16, float* __restrict__ logits, int Hi, int Wi, int Ho, int Wo) {
    using C = DsTcCfg<CIN, COUT, STRIDE, HEAD>;
    constexpr int IW = C::IW, NCK = C::NCK, RPS = C::RPS;
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar_w, bar_mma, bar_mma2;
    __shared__ uint32_t tmem_base_s;
    float* Bds = reinterpret_cast<float*>(sm + C::oBd);
    float* Bps = reinterpret_cast<float*>(sm + C::oBp);
    const uint32_t sH = smem_u32(sm + C::oH), sA = smem_u32(sm + C::oA), sB = smem_u32(sm + C::oB), sB2 = smem_u32(sm + C::oB2);
    const uint32_t sWd = smem_u32(sm + C::oWd);   // depthwise weights [9][CIN] as bf16 (FHFMA operands)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n = blockIdx.z;
    const int oy0 = blockIdx.y * C::TH, ox0 = blockIdx.x * C::TW;
    const int iy0 = oy0 * STRIDE - 1, ix0 = ox0 * STRIDE - 1;
    const uint32_t tm_cols = HEAD ? ((COUT + ncp16 <= 256) ? 256u : 512u) : (COUT <= 64 ? 64u : 128u);

    if (tid == 0) {
        mbar_init(&bar_w, 1); mbar_init(&bar_mma, 1); mbar_init(&bar_mma2, 1);
        fence_mbar_init();
        if (!C::ALIAS_B) {
            mbar_arrive_expect_tx(&bar_w, COUT * CIN * 2);
            bulk_g2s(sm + C::oB, wp_img, COUT * CIN * 2, &bar_w);
        }
    }
    if (warp == 0) { tmem_alloc(&tmem_base_s, tm_cols); tmem_relinquish(); }

    // ---- halo tile: lanes along channel chunks -> coalesced global reads ----
    for (int i = tid; i < C::PINP * NCK; i += kThreads) {
        const int pin = i / NCK, k8 = i % NCK;
        const int iy = iy0 + pin / IW, ix = ix0 + pin % IW;
        const bool ok = (pin < C::PIN && iy >= 0 && iy < Hi && ix >= 0 && ix < Wi);
        const bf16* src = ok ? in + (((size_t)n * Hi + iy) * Wi + ix) * CIN + k8 * 8 : in;
        cp_async16z(sH + pin * C::ROWB + (chunk_swz<CIN>(pin, k8) << 4), src, ok);
    }
    for (int i = tid; i < 9 * CIN / 2; i += kThreads)
        reinterpret_cast<uint32_t*>(sm + C::oWd)[i] = packbf(__ldg(w.wd + 2 * i), __ldg(w.wd + 2 * i + 1));
    for (int i = tid; i < CIN; i += kThreads) Bds[i] = __ldg(w.bd + i);
    for (int i = tid; i < COUT; i += kThreads) Bps[i] = __ldg(w.bp + i);
    asm volatile("cp.async.wait_all;" ::: "memory");
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_base_s;

    // ---- depthwise 3x3 (fp32) -> A-operand tile ----
    if (CIN == 128 && STRIDE == 1) {
#pragma unroll 1
        for (int it = tid; it < 16 * (CIN / 4); it += kThreads)
            dw3x3_s1_col4<C::ROWB, IW>(sH, it & 15, it >> 4, sWd, CIN, Bds, sA, 0);
    } else {
    #pragma unroll 1
        for (int it = tid; it < C::NSTRIP; it += kThreads) {
            const int x = it & 15, rg = (it >> 4) % (C::TH / RPS), k8 = it / (16 * (C::TH / RPS));
            uint4 wk[9];
    #pragma unroll
            for (int t = 0; t < 9; ++t) wk[t] = lds128(sWd + (t * CIN + k8 * 8) * 2);
            float acc[RPS][8];
            {
                const float4 ba = *reinterpret_cast<const float4*>(Bds + k8 * 8);
                const float4 bb = *reinterpret_cast<const float4*>(Bds + k8 * 8 + 4);
    #pragma unroll
                for (int o = 0; o < RPS; ++o) {
                    acc[o][0] = ba.x; acc[o][1] = ba.y; acc[o][2] = ba.z; acc[o][3] = ba.w;
                    acc[o][4] = bb.x; acc[o][5] = bb.y; acc[o][6] = bb.z; acc[o][7] = bb.w;
                }
            }
            constexpr int NR = (RPS - 1) * STRIDE + 3;
    #pragma unroll
            for (int r = 0; r < NR; ++r) {
                const int iy = (RPS * rg) * STRIDE + r;
    #pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int pin = iy * IW + x * STRIDE + kx;
                    const uint4 v = lds128(sH + pin * C::ROWB + (chunk_swz<CIN>(pin, k8) << 4));
    #pragma unroll
                    for (int o = 0; o < RPS; ++o) {
                        const int ky = r - o * STRIDE;
                        if (ky >= 0 && ky < 3) fhfma8(acc[o], v, wk[ky * 3 + kx]);
                    }
                }
            }
    #pragma unroll
            for (int o = 0; o < RPS; ++o) {
                const int p = (RPS * rg + o) * 16 + x;
                sts128(sA + ((k8 * 16 + (p >> 3)) << 7) + ((p & 7) << 4), packbf_relu(acc[o][0], acc[o][1]),
                       packbf_relu(acc[o][2], acc[o][3]), packbf_relu(acc[o][4], acc[o][5]), packbf_relu(acc[o][6], acc[o][7]));
            }
        }
    }
    fence_async_proxy();
    __syncthreads();

    // ---- pointwise contraction on the tensor core ----
    if (tid == 0) {
        if (C::ALIAS_B) {   // the halo tile is dead: stream the weight image (and the head's) over it
            mbar_arrive_expect_tx(&bar_w, COUT * CIN * 2 + (HEAD ? ncp16 * COUT * 2 : 0));
            bulk_g2s(sm + C::oB, wp_img, COUT * CIN * 2, &bar_w);
            if (HEAD) bulk_g2s(sm + C::oB2, wh_img, ncp16 * COUT * 2, &bar_w);
        }
        mbar_wait(&bar_w, 0);
        tc_fence_after_sync();
        constexpr uint32_t idesc = make_idesc_bf16(128, COUT);
#pragma unroll
        for (int k16 = 0; k16 < CIN / 16; ++k16)
            umma_bf16_ss(tmem, make_smem_desc(sA + k16 * 4096, 2048, 128), make_smem_desc(sB + k16 * 2 * (COUT * 16), COUT * 16, 128),
                         idesc, k16 > 0);
        umma_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, 0);
    tc_fence_after_sync();

    // ---- epilogue: warp = (row quarter, column half) ----
    const int q = warp & 3, half = warp >> 2;
    const int p = q * 32 + lane;
    const int oy = oy0 + (p >> 4), ox = ox0 + (p & 15);
    const bool live = (oy < Ho) && (ox < Wo);
    constexpr int CH = COUT / 2;
    {
        uint32_t r[CH];
#pragma unroll
        for (int c0 = 0; c0 < CH; c0 += 8) tmem_ld_32x32b_x8(tmem + ((uint32_t)(q * 32) << 16) + half * CH + c0, r + c0);
        tmem_ld_wait();
#pragma unroll
        for (int c0 = 0; c0 < CH; c0 += 8) {
            const int co = half * CH + c0;
            const float4 ba = *reinterpret_cast<const float4*>(Bps + co);
            const float4 bb = *reinterpret_cast<const float4*>(Bps + co + 4);
            const uint32_t* q8 = r + c0;
            const uint32_t a = packbf_relu(__uint_as_float(q8[0]) + ba.x, __uint_as_float(q8[1]) + ba.y);
            const uint32_t b = packbf_relu(__uint_as_float(q8[2]) + ba.z, __uint_as_float(q8[3]) + ba.w);
            const uint32_t c = packbf_relu(__uint_as_float(q8[4]) + bb.x, __uint_as_float(q8[5]) + bb.y);
            const uint32_t d = packbf_relu(__uint_as_float(q8[6]) + bb.z, __uint_as_float(q8[7]) + bb.w);
            if (!HEAD) {
                if (live) *reinterpret_cast<uint4*>(out + (((size_t)n * Ho + oy) * Wo + ox) * COUT + co) = make_uint4(a, b, c, d);
            } else {
                sts128(sA + a_tile_off(p, co >> 3), a, b, c, d);   // second A operand over the first (its MMA has completed)
            }
        }
    }
    if (HEAD) {
        fence_async_proxy();
        tc_fence_before_sync();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after_sync();
            const uint32_t idesc2 = make_idesc_bf16(128, ncp16);
#pragma unroll
            for (int k16 = 0; k16 < COUT / 16; ++k16)
                umma_bf16_ss(tmem + COUT, make_smem_desc(sA + k16 * 4096, 2048, 128),
                             make_smem_desc(sB2 + k16 * 2 * (ncp16 * 16), ncp16 * 16, 128), idesc2, k16 > 0);
            umma_commit(&bar_mma2);
        }
        mbar_wait(&bar_mma2, 0);
        tc_fence_after_sync();
        // logits: quarter q rows; the two warp halves split the class groups of 8
        for (int c0 = half * 8; c0 < head.ncp; c0 += 16) {
            uint32_t r[8];
            tmem_ld_32x32b_x8(tmem + ((uint32_t)(q * 32) << 16) + COUT + c0, r);
            tmem_ld_wait();
            if (live) {
                float* lp = logits + (((size_t)n * Ho + oy) * Wo + ox) * head.ncp + c0;
#pragma unroll
                for (int g = 0; g < 2; ++g)
                    if (c0 + 4 * g < head.ncp)
                        *reinterpret_cast<float4*>(lp + 4 * g) =
                            make_float4(__uint_as_float(r[4 * g]) + __ldg(head.b + c0 + 4 * g), __uint_as_float(r[4 * g + 1]) + __ldg(head.b + c0 + 4 * g + 1),
                                        __uint_as_float(r[4 * g + 2]) + __ldg(head.b + c0 + 4 * g + 2), __uint_as_float(r[4 * g + 3]) + __ldg(head.b + c0 + 4 * g + 3));
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, tm_cols);
}

template <int CIN, int COUT, int STRIDE, bool HEAD>
static cudaError_t run_ds_tc(const bf16* in, const DsW& w, const bf16* wp_img, bf16* out, const HeadW* head, const bf16* wh_img,
                             float* logits, int n, int hi, int wi, int ho, int wo, cudaStream_t s) {
    using C = DsTcCfg<CIN, COUT, STRIDE, HEAD>;
    HeadW h{};
    int ncp16 = 0;
    if (HEAD) { h = *head; ncp16 = round_up(h.nc, 16); }
    const size_t smem = C::oB2 + (HEAD ? (size_t)ncp16 * COUT * 2 : 0);
    static unsigned long long configured = 0;
    static size_t configured_bytes = 0;
    if (smem > configured_bytes) { configured = 0; configured_bytes = smem; }
    cudaError_t e = ensure_dyn_smem(dsconv_tc_kernel<CIN, COUT, STRIDE, HEAD>, configured_bytes, configured);
    if (e != cudaSuccess) return e;
    dim3 grid(ceil_div(wo, C::TW), ceil_div(ho, C::TH), n);
    dsconv_tc_kernel<CIN, COUT, STRIDE, HEAD><<<grid, kThreads, smem, s>>>(in, w, wp_img, out, h, wh_img, ncp16, logits, hi, wi, ho, wo);
    return cudaGetLastError();
}

cudaError_t launch_dsconv_tc(int cin, int cout, int stride, const bf16* in, const DsW& w, const bf16* wp_img, bf16* out,
                             const HeadW* head, const bf16* wh_img, float* logits, int n, int hi, int wi, int ho, int wo,
                             cudaStream_t s) {
    if (cin == 32 && cout == 48 && stride == 2 && !head) return run_ds_tc<32, 48, 2, false>(in, w, wp_img, out, head, wh_img, logits, n, hi, wi, ho, wo, s);
    if (cin == 48 && cout == 64 && stride == 2 && !head) return run_ds_tc<48, 64, 2, false>(in, w, wp_img, out, head, wh_img, logits, n, hi, wi, ho, wo, s);
    if (cin == 128 && cout == 128 && stride == 1 && !head) return run_ds_tc<128, 128, 1, false>(in, w, wp_img, out, head, wh_img, logits, n, hi, wi, ho, wo, s);
    if (cin == 128 && cout == 128 && stride == 1 && head) return run_ds_tc<128, 128, 1, true>(in, w, wp_img, out, head, wh_img, logits, n, hi, wi, ho, wo, s);
    return cudaErrorInvalidValue;
}

}  // namespace fscnn
