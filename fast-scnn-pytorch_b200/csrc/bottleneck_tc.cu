// bottleneck_tc.cu -- bf16 LinearBottleneck on the 5th-gen tensor cores (tcgen05 + TMEM).
// Same fusion as bottleneck.cu (reference models/fast_scnn.py:95-115): expand 1x1 + ReLU -> DW 3x3
// (stride s) + ReLU -> project 1x1 (+ residual), the 6c-wide tensor never leaves the SM; but both
// 1x1 contractions run as tcgen05.mma (bf16 operands from shared memory, fp32 accumulators in
// TMEM), the depthwise stage runs on the CUDA cores in fp32 between them.
//
// CTA = 8x16 output pixels (M = 128 rows of the project MMA), 256 threads.
//   X  : input halo tile, bf16, as NMT MMA A-tiles of 128 rows (core-matrix layout [k/8][row/8])
//   per chunk of CE = 64 expanded channels:
//     We/Wp chunk  <- one bulk copy each (weights are pre-packed in the shared-memory image)
//     expand MMA   : TMEM[mt][128 x CE]  = X[mt] * We^T                    (one thread issues)
//     epilogue     : TMEM -> +bias, ReLU, zero outside the image -> bf16 -> E[halo px][CE] (swizzled rows)
//     depthwise    : E -> fp32 3x3 + bias + ReLU -> bf16 -> D[128][CE] (A-operand layout)
//     project MMA  : TMEM[128 x COUT] += D * Wp^T
//   final epilogue : TMEM -> +bias (+ residual from X) -> bf16 NHWC store
#include "kernels.h"
#include "umma.cuh"

namespace fscnn {

template <int CIN, int COUT, int STRIDE>
struct TcCfg {
    static constexpr int TH = 8, TW = 16, P = 128;
    static constexpr int IH = (TH - 1) * STRIDE + 3, IW = (TW - 1) * STRIDE + 3;
    static constexpr int PIN = IH * IW;
    static constexpr int NMT = (PIN + 127) / 128;          // expand M-tiles
    static constexpr int ROWS = NMT * 128;
    static constexpr int PINP = round_up(PIN, 8);
    static constexpr int CE = 64, CEXP = 6 * CIN, NCH = CEXP / CE;
    static constexpr int XT_BYTES = 128 * CIN * 2;         // one A-tile of X
    static constexpr int oX = 0;
    static constexpr int oE = oX + NMT * XT_BYTES;
    static constexpr int oD = oE + PINP * CE * 2;
    static constexpr int oWe = oD + P * CE * 2;
    static constexpr int oWp = oWe + CE * CIN * 2;
    static constexpr int oWd = oWp + COUT * CE * 2;
    static constexpr int oBe = oWd + 9 * CE * 4;
    static constexpr int oBd = oBe + CE * 4;
    static constexpr int oValid = oBd + CE * 4;
    static constexpr int smem_bytes = oValid + ROWS;
    static constexpr int TM_EXP = 0, TM_PROJ = NMT * CE;
    static constexpr int TM_COLS = (NMT * CE + COUT) <= 256 ? 256 : 512;
    static_assert(NMT * CE + COUT <= 512, "TMEM budget");
    static_assert(CEXP % CE == 0 && CIN % 16 == 0 && COUT % 16 == 0, "shape");
};

template <int CIN, int COUT, int STRIDE, bool RES>
__global__ void __launch_bounds__(kThreads, 1)
bottleneck_tc_kernel(const bf16* __restrict__ in, BneckW w, const bf16* __restrict__ we_img, const bf16* __restrict__ wp_img,
                     bf16* __restrict__ out, int Hi, int Wi, int Ho, int Wo) {
    using C = TcCfg<CIN, COUT, STRIDE>;
    constexpr int CE = C::CE, IW = C::IW, NMT = C::NMT;
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar_w, bar_exp, bar_proj;
    __shared__ uint32_t tmem_base_s;
    float* Wds = reinterpret_cast<float*>(sm + C::oWd);
    float* Bes = reinterpret_cast<float*>(sm + C::oBe);
    float* Bds = reinterpret_cast<float*>(sm + C::oBd);
    uint8_t* valid = sm + C::oValid;
    const uint32_t sX = smem_u32(sm + C::oX), sE = smem_u32(sm + C::oE), sD = smem_u32(sm + C::oD);
    const uint32_t sWe = smem_u32(sm + C::oWe), sWp = smem_u32(sm + C::oWp);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n = blockIdx.z;
    const int oy0 = blockIdx.y * C::TH, ox0 = blockIdx.x * C::TW;
    const int iy0 = oy0 * STRIDE - 1, ix0 = ox0 * STRIDE - 1;

    if (tid == 0) {
        mbar_init(&bar_w, 1); mbar_init(&bar_exp, 1); mbar_init(&bar_proj, 1);
        fence_mbar_init();
    }
    if (warp == 0) { tmem_alloc(&tmem_base_s, C::TM_COLS); tmem_relinquish(); }

    // ---- stage the input halo tile as MMA A-tiles (lanes along rows -> conflict-free 16-byte stores) ----
    for (int m = tid; m < C::ROWS; m += kThreads) {
        const int iy = iy0 + m / IW, ix = ix0 + m % IW;
        valid[m] = (m < C::PIN && iy >= 0 && iy < Hi && ix >= 0 && ix < Wi) ? 1 : 0;
    }
    for (int i = tid; i < C::ROWS * (CIN / 8); i += kThreads) {
        const int m = i % C::ROWS, k8 = i / C::ROWS;
        const int iy = iy0 + m / IW, ix = ix0 + m % IW;
        const bool ok = (m < C::PIN && iy >= 0 && iy < Hi && ix >= 0 && ix < Wi);
        const bf16* src = ok ? in + (((size_t)n * Hi + iy) * Wi + ix) * CIN + k8 * 8 : in;
        const uint32_t dst = sX + (m >> 7) * C::XT_BYTES + ((k8 * 16 + ((m & 127) >> 3)) << 7) + ((m & 7) << 4);
        cp_async16z(dst, src, ok);
    }
    cp_async_wait_all();
    fence_async_proxy();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_base_s;

    constexpr uint32_t idesc_exp = make_idesc_bf16(128, CE);
    constexpr uint32_t idesc_proj = make_idesc_bf16(128, COUT);

#pragma unroll 1
    for (int e = 0; e < C::NCH; ++e) {
        // ---- (1) weights of this chunk: two bulk copies, small fp32 tables by plain loads ----
        if (tid == 0) {
            if (e > 0) mbar_wait(&bar_proj, (e - 1) & 1);   // previous project MMA has finished reading Wp / D
            mbar_arrive_expect_tx(&bar_w, CE * CIN * 2 + COUT * CE * 2);
            bulk_g2s(sm + C::oWe, we_img + (size_t)e * CE * CIN, CE * CIN * 2, &bar_w);
            bulk_g2s(sm + C::oWp, wp_img + (size_t)e * COUT * CE, COUT * CE * 2, &bar_w);
        }
        for (int i = tid; i < 9 * CE; i += kThreads) Wds[i] = __ldg(w.wd + (i / CE) * C::CEXP + e * CE + (i % CE));
        if (tid < CE) { Bes[tid] = __ldg(w.be + e * CE + tid); Bds[tid] = __ldg(w.bd + e * CE + tid); }
        // ---- (2) expand MMAs ----
        if (tid == 0) {
            mbar_wait(&bar_w, e & 1);
            tc_fence_after_sync();
#pragma unroll
            for (int mt = 0; mt < NMT; ++mt)
#pragma unroll
                for (int k16 = 0; k16 < CIN / 16; ++k16) {
                    const uint64_t da = make_smem_desc(sX + mt * C::XT_BYTES + k16 * 2 * 2048, 2048, 128);
                    const uint64_t db = make_smem_desc(sWe + k16 * 2 * (CE * 16), CE * 16, 128);
                    umma_bf16_ss(tmem + C::TM_EXP + mt * CE, da, db, idesc_exp, k16 > 0);
                }
            umma_commit(&bar_exp);
        }
        __syncthreads();   // Wds / Bes / Bds visible
        // ---- (3) expand epilogue: TMEM -> bias, ReLU, image mask -> bf16 rows of E ----
        mbar_wait(&bar_exp, e & 1);
        tc_fence_after_sync();
        for (int task = warp; task < NMT * 4; task += kThreads / 32) {
            const int mt = task >> 2, q = task & 3;
            const int pin = mt * 128 + q * 32 + lane;
            const bool ok = valid[pin];
#pragma unroll
            for (int c0 = 0; c0 < CE; c0 += 32) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(tmem + ((uint32_t)(q * 32) << 16) + C::TM_EXP + mt * CE + c0, r);
                tmem_ld_wait();
                if (pin < C::PINP) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        uint32_t pk[4];
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            const int c = c0 + g * 8 + 2 * h;
                            const float a = ok ? relu(__uint_as_float(r[g * 8 + 2 * h]) + Bes[c]) : 0.f;
                            const float b = ok ? relu(__uint_as_float(r[g * 8 + 2 * h + 1]) + Bes[c + 1]) : 0.f;
                            pk[h] = packbf(a, b);
                        }
                        const uint32_t dst = sE + pin * (CE * 2) + ((((c0 >> 3) + g) ^ (pin & 7)) << 4);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
                    }
                }
            }
        }
        tc_fence_before_sync();
        __syncthreads();
        // ---- (4) depthwise 3x3 in fp32: thread = (column x, 4-row group rg, 8-channel chunk j) ----
        {
            const int x = tid & 15, rg = (tid >> 4) & 1, j = tid >> 5;
            float wk[9][8];
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const float4 a = *reinterpret_cast<const float4*>(Wds + t * CE + j * 8);
                const float4 b = *reinterpret_cast<const float4*>(Wds + t * CE + j * 8 + 4);
                wk[t][0] = a.x; wk[t][1] = a.y; wk[t][2] = a.z; wk[t][3] = a.w;
                wk[t][4] = b.x; wk[t][5] = b.y; wk[t][6] = b.z; wk[t][7] = b.w;
            }
            float acc[4][8];
#pragma unroll
            for (int o = 0; o < 4; ++o)
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[o][c] = Bds[j * 8 + c];
            constexpr int NR = 3 * STRIDE + 3;   // input rows feeding 4 output rows
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                const int iy = (4 * rg) * STRIDE + r;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int pin = iy * IW + x * STRIDE + kx;
                    uint4 v;
                    const uint32_t src = sE + pin * (CE * 2) + ((j ^ (pin & 7)) << 4);
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(src));
                    float f[8];
                    unpackbf8(v, f);
#pragma unroll
                    for (int o = 0; o < 4; ++o) {
                        const int ky = r - o * STRIDE;
                        if (ky >= 0 && ky < 3) {
#pragma unroll
                            for (int c = 0; c < 8; ++c) acc[o][c] = fmaf(f[c], wk[ky * 3 + kx][c], acc[o][c]);
                        }
                    }
                }
            }
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                const int p = (4 * rg + o) * 16 + x;
                const uint32_t dst = sD + ((j * 16 + (p >> 3)) << 7) + ((p & 7) << 4);
                const uint32_t p0 = packbf(relu(acc[o][0]), relu(acc[o][1])), p1 = packbf(relu(acc[o][2]), relu(acc[o][3]));
                const uint32_t p2 = packbf(relu(acc[o][4]), relu(acc[o][5])), p3 = packbf(relu(acc[o][6]), relu(acc[o][7]));
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(p0), "r"(p1), "r"(p2), "r"(p3) : "memory");
            }
        }
        fence_async_proxy();
        __syncthreads();
        // ---- (5) project MMAs, accumulated across chunks ----
        if (tid == 0) {
            tc_fence_after_sync();
#pragma unroll
            for (int k16 = 0; k16 < CE / 16; ++k16) {
                const uint64_t da = make_smem_desc(sD + k16 * 2 * 2048, 2048, 128);
                const uint64_t db = make_smem_desc(sWp + k16 * 2 * (COUT * 16), COUT * 16, 128);
                umma_bf16_ss(tmem + C::TM_PROJ, da, db, idesc_proj, (e | k16) != 0);
            }
            umma_commit(&bar_proj);
        }
    }

    // ---- final epilogue: + bias (+ residual), bf16 NHWC store; warp = (row quarter, column half) ----
    mbar_wait(&bar_proj, (C::NCH - 1) & 1);
    tc_fence_after_sync();
    {
        const int q = warp & 3, half = warp >> 2;
        const int p = q * 32 + lane;
        const int py = p >> 4, px = p & 15;
        const int oy = oy0 + py, ox = ox0 + px;
        const bool live = (oy < Ho) && (ox < Wo);
        constexpr int CH = COUT / 2;   // columns per warp-half (32, 48 or 64)
#pragma unroll
        for (int c0 = 0; c0 < CH; c0 += 16) {
            uint32_t r[16];
            tmem_ld_32x32b_x16(tmem + ((uint32_t)(q * 32) << 16) + C::TM_PROJ + half * CH + c0, r);
            tmem_ld_wait();
            if (live) {
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    const int co = half * CH + c0 + g * 8;
                    float v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[g * 8 + i]) + __ldg(w.bp + co + i);
                    if (RES) {   // stride 1: centre of the halo tile, still resident in X
                        const int m = (py + 1) * IW + (px + 1);
                        uint4 xv;
                        const uint32_t src = sX + (m >> 7) * C::XT_BYTES + (((co >> 3) * 16 + ((m & 127) >> 3)) << 7) + ((m & 7) << 4);
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(xv.x), "=r"(xv.y), "=r"(xv.z), "=r"(xv.w) : "r"(src));
                        float f[8];
                        unpackbf8(xv, f);
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] += f[i];
                    }
                    uint4 o;
                    o.x = packbf(v[0], v[1]); o.y = packbf(v[2], v[3]); o.z = packbf(v[4], v[5]); o.w = packbf(v[6], v[7]);
                    *reinterpret_cast<uint4*>(out + (((size_t)n * Ho + oy) * Wo + ox) * COUT + co) = o;
                }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, C::TM_COLS);
}

template <int CIN, int COUT, int STRIDE, bool RES>
static cudaError_t run_tc(const bf16* in, const BneckW& w, const bf16* we_img, const bf16* wp_img, bf16* out, int n, int hi,
                          int wi, int ho, int wo, cudaStream_t s) {
    using C = TcCfg<CIN, COUT, STRIDE>;
    static unsigned long long configured = 0;
    cudaError_t e = ensure_dyn_smem(bottleneck_tc_kernel<CIN, COUT, STRIDE, RES>, C::smem_bytes, configured);
    if (e != cudaSuccess) return e;
    dim3 grid(ceil_div(wo, C::TW), ceil_div(ho, C::TH), n);
    bottleneck_tc_kernel<CIN, COUT, STRIDE, RES><<<grid, kThreads, C::smem_bytes, s>>>(in, w, we_img, wp_img, out, hi, wi, ho, wo);
    return cudaGetLastError();
}

cudaError_t launch_bottleneck_tc(int cin, int cout, int stride, const bf16* in, const BneckW& w, const bf16* we_img,
                                 const bf16* wp_img, bf16* out, int n, int hi, int wi, int ho, int wo, cudaStream_t s) {
    if (cin == 64 && cout == 64 && stride == 2) return run_tc<64, 64, 2, false>(in, w, we_img, wp_img, out, n, hi, wi, ho, wo, s);
    if (cin == 64 && cout == 64 && stride == 1) return run_tc<64, 64, 1, true>(in, w, we_img, wp_img, out, n, hi, wi, ho, wo, s);
    if (cin == 64 && cout == 96 && stride == 2) return run_tc<64, 96, 2, false>(in, w, we_img, wp_img, out, n, hi, wi, ho, wo, s);
    if (cin == 96 && cout == 96 && stride == 1) return run_tc<96, 96, 1, true>(in, w, we_img, wp_img, out, n, hi, wi, ho, wo, s);
    if (cin == 96 && cout == 128 && stride == 1) return run_tc<96, 128, 1, false>(in, w, we_img, wp_img, out, n, hi, wi, ho, wo, s);
    if (cin == 128 && cout == 128 && stride == 1) return run_tc<128, 128, 1, true>(in, w, we_img, wp_img, out, n, hi, wi, ho, wo, s);
    return cudaErrorInvalidValue;
}

}  // namespace fscnn
