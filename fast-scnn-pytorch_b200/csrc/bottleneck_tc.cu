// bottleneck_tc.cu -- bf16 LinearBottleneck on the 5th-gen tensor cores (tcgen05 + TMEM).
// Same fusion as bottleneck.cu (reference models/fast_scnn.py:95-115): expand 1x1 + ReLU -> DW 3x3
// (stride s) + ReLU -> project 1x1 (+ residual), the 6c-wide tensor never leaves the SM; but both
// 1x1 contractions run as tcgen05.mma (bf16 operands from shared memory, fp32 accumulators in
// TMEM), the depthwise stage runs on the CUDA cores in fp32 between them.
//
// Measured alternative (round 1, rejected): running the depthwise 3x3 itself on the tensor core as 9 * CE/16
// block-diagonal MMAs (N = 16) over shifted views of a column-major expanded tile is numerically exact and needs 3x
// fewer CUDA-core instructions, but every SS-mode tcgen05.mma costs >= 64 cycles for its 128 x 16 A-operand read
// whatever N is (tc_probe: 64.1 cyc/MMA for N = 16..128), so 36 of them per chunk made the kernel tensor-pipe bound
// and 25-30 % slower than the CUDA-core depthwise below.
//
// CTA = 8x16 output pixels (M = 128 rows of the project MMA), 512 threads (16 warps keep the LDS / TMEM
// latencies of the CUDA-core phases covered while one elected thread feeds the tensor core).
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

constexpr int kNT = 512;   // threads per CTA of this kernel

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
    static constexpr int WE_BYTES = CE * CIN * 2, WP_BYTES = COUT * CE * 2;
    static constexpr int oWe = oD + P * CE * 2;            // 2 buffers
    static constexpr int oWp = oWe + 2 * WE_BYTES;         // 2 buffers
    static constexpr int oWd = oWp + 2 * WP_BYTES;         // fp32 tables of ALL chunks: Wd[9][CEXP], Be[CEXP], Bd[CEXP]
    static constexpr int oBe = oWd + 9 * CEXP * 4;
    static constexpr int oBd = oBe + CEXP * 4;
    static constexpr int oValid = oBd + CEXP * 4;
    static constexpr int smem_bytes = oValid + ROWS;
    static constexpr int TM_EXP = 0, TM_PROJ = NMT * CE;
    static constexpr int TM_COLS = (NMT * CE + COUT) <= 256 ? 256 : 512;
    static_assert(NMT * CE + COUT <= 512, "TMEM budget");
    static_assert(CEXP % CE == 0 && CIN % 16 == 0 && COUT % 16 == 0, "shape");
};

template <int CIN, int COUT, int STRIDE, bool RES>
__global__ void __launch_bounds__(kNT, 1)
bottleneck_tc_kernel(const bf16* __restrict__ in, BneckW w, const bf16* __restrict__ we_img, const bf16* __restrict__ wp_img,
                     bf16* __restrict__ out, int Hi, int Wi, int Ho, int Wo) {
    using C = TcCfg<CIN, COUT, STRIDE>;
    constexpr int CE = C::CE, IW = C::IW, NMT = C::NMT;
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar_we[2], bar_wp[2], bar_exp, bar_proj;
    __shared__ uint32_t tmem_base_s;
    const float* Wd_all = reinterpret_cast<const float*>(sm + C::oWd);
    const float* Be_all = reinterpret_cast<const float*>(sm + C::oBe);
    const float* Bd_all = reinterpret_cast<const float*>(sm + C::oBd);
    uint8_t* valid = sm + C::oValid;
    const uint32_t sX = smem_u32(sm + C::oX), sE = smem_u32(sm + C::oE), sD = smem_u32(sm + C::oD);
    const uint32_t sWe = smem_u32(sm + C::oWe), sWp = smem_u32(sm + C::oWp);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n = blockIdx.z;
    const int oy0 = blockIdx.y * C::TH, ox0 = blockIdx.x * C::TW;
    const int iy0 = oy0 * STRIDE - 1, ix0 = ox0 * STRIDE - 1;

    // weight chunk e lives in buffer e & 1; the bulk copies run one to two chunks ahead of their MMAs
    auto prefetch_we = [&](int e) {
        mbar_arrive_expect_tx(&bar_we[e & 1], C::WE_BYTES);
        bulk_g2s(sm + C::oWe + (e & 1) * C::WE_BYTES, we_img + (size_t)e * CE * CIN, C::WE_BYTES, &bar_we[e & 1]);
    };
    auto prefetch_wp = [&](int e) {
        mbar_arrive_expect_tx(&bar_wp[e & 1], C::WP_BYTES);
        bulk_g2s(sm + C::oWp + (e & 1) * C::WP_BYTES, wp_img + (size_t)e * COUT * CE, C::WP_BYTES, &bar_wp[e & 1]);
    };
    constexpr uint32_t idesc_exp = make_idesc_bf16(128, CE);
    constexpr uint32_t idesc_proj = make_idesc_bf16(128, COUT);
    auto issue_expand = [&](int e, uint32_t tmem) {
        mbar_wait(&bar_we[e & 1], (e >> 1) & 1);
        tc_fence_after_sync();
#pragma unroll
        for (int mt = 0; mt < NMT; ++mt)
#pragma unroll
            for (int k16 = 0; k16 < CIN / 16; ++k16) {
                const uint64_t da = make_smem_desc(sX + mt * C::XT_BYTES + k16 * 2 * 2048, 2048, 128);
                const uint64_t db = make_smem_desc(sWe + (e & 1) * C::WE_BYTES + k16 * 2 * (CE * 16), CE * 16, 128);
                umma_bf16_ss(tmem + C::TM_EXP + mt * CE, da, db, idesc_exp, k16 > 0);
            }
        umma_commit(&bar_exp);
    };

    if (tid == 0) {
        mbar_init(&bar_we[0], 1); mbar_init(&bar_we[1], 1); mbar_init(&bar_wp[0], 1); mbar_init(&bar_wp[1], 1);
        mbar_init(&bar_exp, 1); mbar_init(&bar_proj, 1);
        fence_mbar_init();
        prefetch_we(0);
        if (C::NCH > 1) prefetch_we(1);
        prefetch_wp(0);
    }
    if (warp == 0) { tmem_alloc(&tmem_base_s, C::TM_COLS); tmem_relinquish(); }

    // ---- stage the input halo tile as MMA A-tiles (lanes along rows -> conflict-free 16-byte stores) ----
    for (int m = tid; m < C::ROWS; m += kNT) {
        const int iy = iy0 + m / IW, ix = ix0 + m % IW;
        valid[m] = (m < C::PIN && iy >= 0 && iy < Hi && ix >= 0 && ix < Wi) ? 1 : 0;
    }
    for (int i = tid; i < C::ROWS * (CIN / 8); i += kNT) {
        const int m = i % C::ROWS, k8 = i / C::ROWS;
        const int iy = iy0 + m / IW, ix = ix0 + m % IW;
        const bool ok = (m < C::PIN && iy >= 0 && iy < Hi && ix >= 0 && ix < Wi);
        const bf16* src = ok ? in + (((size_t)n * Hi + iy) * Wi + ix) * CIN + k8 * 8 : in;
        const uint32_t dst = sX + (m >> 7) * C::XT_BYTES + ((k8 * 16 + ((m & 127) >> 3)) << 7) + ((m & 7) << 4);
        cp_async16z(dst, src, ok);
    }
    {   // fp32 depthwise weights and the two bias vectors of every chunk, once
        float* wd_s = reinterpret_cast<float*>(sm + C::oWd);
        float* be_s = reinterpret_cast<float*>(sm + C::oBe);
        float* bd_s = reinterpret_cast<float*>(sm + C::oBd);
        for (int i = tid; i < 9 * C::CEXP / 4; i += kNT) reinterpret_cast<float4*>(wd_s)[i] = __ldg(reinterpret_cast<const float4*>(w.wd) + i);
        for (int i = tid; i < C::CEXP / 4; i += kNT) {
            reinterpret_cast<float4*>(be_s)[i] = __ldg(reinterpret_cast<const float4*>(w.be) + i);
            reinterpret_cast<float4*>(bd_s)[i] = __ldg(reinterpret_cast<const float4*>(w.bd) + i);
        }
    }
    cp_async_wait_all();
    fence_async_proxy();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) issue_expand(0, tmem);

#pragma unroll 1
    for (int e = 0; e < C::NCH; ++e) {
        const float* Bes = Be_all + e * CE;
        const float* Bds = Bd_all + e * CE;
        const float* Wds = Wd_all + e * CE;      // tap t at Wds[t * CEXP + c]
        // ---- (3) expand epilogue: TMEM -> bias, ReLU, image mask -> bf16 rows of E ----
        mbar_wait(&bar_exp, e & 1);
        tc_fence_after_sync();
        for (int task = warp; task < NMT * 8; task += kNT / 32) {   // task = (M-tile, 32-column half, lane quarter)
            const int q = task & 3, ch = (task >> 2) & 1, mt = task >> 3;
            const int pin = mt * 128 + q * 32 + lane;
            const bool ok = valid[pin];
            const int c0 = ch * 32;
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
                    sts128(sE + pin * (CE * 2) + ((((c0 >> 3) + g) ^ (pin & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
                }
            }
        }
        tc_fence_before_sync();
        __syncthreads();
        // ---- the tensor core runs ahead: expand MMAs of the NEXT chunk overlap this chunk's depthwise phase ----
        if (tid == 0) {
            tc_fence_after_sync();
            if (e + 1 < C::NCH) issue_expand(e + 1, tmem);
            if (e + 2 < C::NCH) prefetch_we(e + 2);          // buffer e&1: expand(e) has completed (bar_exp waited above)
        }
        if (e > 0) mbar_wait(&bar_proj, (e - 1) & 1);        // project(e-1) done: D and Wp buffer (e+1)&1 are free
        if (tid == 0 && e + 1 < C::NCH) prefetch_wp(e + 1);
        // ---- (4) depthwise 3x3 in fp32: thread = (column x, 2-row group rg, 8-channel chunk j) ----
        {
            const int x = tid & 15, rg = (tid >> 4) & 3, j = tid >> 6;
            float acc[2][8];
#pragma unroll
            for (int o = 0; o < 2; ++o)
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[o][c] = Bds[j * 8 + c];
            constexpr int NR = STRIDE + 3;   // input rows feeding 2 output rows
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                const int iy = (2 * rg) * STRIDE + r;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int pin = iy * IW + x * STRIDE + kx;
                    float f[8];
                    unpackbf8(lds128(sE + pin * (CE * 2) + ((j ^ (pin & 7)) << 4)), f);
#pragma unroll
                    for (int o = 0; o < 2; ++o) {
                        const int ky = r - o * STRIDE;
                        if (ky >= 0 && ky < 3) {
                            const float4 wa = *reinterpret_cast<const float4*>(Wds + (ky * 3 + kx) * C::CEXP + j * 8);
                            const float4 wb = *reinterpret_cast<const float4*>(Wds + (ky * 3 + kx) * C::CEXP + j * 8 + 4);
                            acc[o][0] = fmaf(f[0], wa.x, acc[o][0]); acc[o][1] = fmaf(f[1], wa.y, acc[o][1]);
                            acc[o][2] = fmaf(f[2], wa.z, acc[o][2]); acc[o][3] = fmaf(f[3], wa.w, acc[o][3]);
                            acc[o][4] = fmaf(f[4], wb.x, acc[o][4]); acc[o][5] = fmaf(f[5], wb.y, acc[o][5]);
                            acc[o][6] = fmaf(f[6], wb.z, acc[o][6]); acc[o][7] = fmaf(f[7], wb.w, acc[o][7]);
                        }
                    }
                }
            }
#pragma unroll
            for (int o = 0; o < 2; ++o) {
                const int p = (2 * rg + o) * 16 + x;
                sts128(sD + a_tile_off(p, j), packbf(relu(acc[o][0]), relu(acc[o][1])), packbf(relu(acc[o][2]), relu(acc[o][3])),
                       packbf(relu(acc[o][4]), relu(acc[o][5])), packbf(relu(acc[o][6]), relu(acc[o][7])));
            }
        }
        fence_async_proxy();
        __syncthreads();
        // ---- (5) project MMAs, accumulated across chunks ----
        if (tid == 0) {
            mbar_wait(&bar_wp[e & 1], (e >> 1) & 1);
            tc_fence_after_sync();
#pragma unroll
            for (int k16 = 0; k16 < CE / 16; ++k16) {
                const uint64_t da = make_smem_desc(sD + k16 * 2 * 2048, 2048, 128);
                const uint64_t db = make_smem_desc(sWp + (e & 1) * C::WP_BYTES + k16 * 2 * (COUT * 16), COUT * 16, 128);
                umma_bf16_ss(tmem + C::TM_PROJ, da, db, idesc_proj, (e | k16) != 0);
            }
            umma_commit(&bar_proj);
        }
    }

    // ---- final epilogue: + bias (+ residual), bf16 NHWC store; warp = (row quarter, column half) ----
    mbar_wait(&bar_proj, (C::NCH - 1) & 1);
    tc_fence_after_sync();
    {
        const int q = warp & 3, part = warp >> 2;   // 16 warps = 4 lane quarters x 4 column parts
        const int p = q * 32 + lane;
        const int py = p >> 4, px = p & 15;
        const int oy = oy0 + py, ox = ox0 + px;
        const bool live = (oy < Ho) && (ox < Wo);
        constexpr int CP = COUT / 4;   // columns per part: 16, 24 or 32
        uint32_t r[CP];
#pragma unroll
        for (int c0 = 0; c0 < CP; c0 += 8) tmem_ld_32x32b_x8(tmem + ((uint32_t)(q * 32) << 16) + C::TM_PROJ + part * CP + c0, r + c0);
        tmem_ld_wait();
        if (live) {
#pragma unroll
            for (int c0 = 0; c0 < CP; c0 += 8) {
                const int co = part * CP + c0;
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[c0 + i]) + __ldg(w.bp + co + i);
                if (RES) {   // stride 1: centre of the halo tile, still resident in X
                    const int m = (py + 1) * IW + (px + 1);
                    float f[8];
                    unpackbf8(lds128(sX + (m >> 7) * C::XT_BYTES + a_tile_off(m & 127, co >> 3)), f);
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] += f[i];
                }
                *reinterpret_cast<uint4*>(out + (((size_t)n * Ho + oy) * Wo + ox) * COUT + co) =
                    make_uint4(packbf(v[0], v[1]), packbf(v[2], v[3]), packbf(v[4], v[5]), packbf(v[6], v[7]));
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
    bottleneck_tc_kernel<CIN, COUT, STRIDE, RES><<<grid, kNT, C::smem_bytes, s>>>(in, w, we_img, wp_img, out, hi, wi, ho, wo);
    return cudaGetLastError();
}

// expanded channels per chunk: the weight images handed to launch_bottleneck_tc must be cut accordingly
int bottleneck_tc_chunk(int) { return 64; }

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
