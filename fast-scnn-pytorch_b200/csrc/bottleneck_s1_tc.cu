// bottleneck_s1_tc.cu -- bf16 LinearBottleneck, stride 1 (reference models/fast_scnn.py:95-115), as a three-role
// pipeline on one persistent CTA per SM.  Same arithmetic as bottleneck_tc.cu (which keeps the stride-2 layers):
// expand 1x1 (tcgen05.mma) -> +bias, ReLU, zero padding -> bf16 E -> depthwise 3x3 (FHFMA.BF16) + bias + ReLU -> bf16 D
// -> project 1x1 (tcgen05.mma, accumulated over the chunks of 64 expanded channels) -> + bias (+ residual) -> bf16 NHWC.
//
//   control warp (lane 0) : TMA halo loads, weight-chunk bulk copies and every tcgen05.mma, driven by a non-blocking event
//                           loop over the mbarriers (an expand MMA is issued the moment its TMEM buffer, weights and halo
//                           tile are there, a project MMA the moment its D tile is)
//   8 expand warps (the higher warp ids: the arbiter prefers them, so the producers are never starved by the FMA-bound
//                           consumers): chunk c: TMEM[c&1] -> E[c&1]; after the first chunk of a tile: output epilogue of the tile before
//   8 depthwise warps     : chunk c: E[c&1] -> D[c&1]; 4 output rows x 8 channels per thread (18 + 9 LDS.128 per 4 outputs)
// E, D and the expand accumulator are double-buffered, so the expand warps run one chunk ahead of the depthwise warps and
// the tensor core two ahead; there is no CTA-wide barrier and no named barrier in the loop, every hand-off is an mbarrier.
#include "kernels.h"
#include "tma_host.h"
#include "umma.cuh"

namespace fscnn {

namespace {
constexpr int kRoleWarps = 8;                       // warps per compute role
constexpr int kS1Threads = (2 * kRoleWarps + 1) * 32;

__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"   // non-blocking (try_wait may suspend the thread)
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
}  // namespace

template <int CIN, int COUT>
struct S1Cfg {
    static constexpr int TH = 8, TW = 16, IH = 10, IW = 18, PIN = IH * IW, NMT = 2;
    static constexpr int CE = 64, CEXP = 6 * CIN, NCH = CEXP / CE;
    static constexpr int X_BYTES = PIN * CIN * 2;          // [CIN/8][PIN][8]: LBO = PIN*16, SBO = 128
    static constexpr int XS = round_up(X_BYTES, 128);
    static constexpr int E_BYTES = round_up(PIN, 8) * CE * 2;
    static constexpr int D_BYTES = 128 * CE * 2;
    static constexpr int WE_BYTES = CE * CIN * 2, WP_BYTES = COUT * CE * 2;
    static constexpr int TAB_BYTES = 9 * CEXP * 2 + 2 * CEXP * 4 + COUT * 4;   // bf16 Wd[9][CEXP] | f32 Be | f32 Bd | f32 Bp
    static constexpr int REST = 2 * E_BYTES + 2 * D_BYTES + 2 * WE_BYTES + 2 * WP_BYTES + TAB_BYTES;
    static constexpr int XB = (REST + 2 * XS <= 227 * 1024 - 512) ? 2 : 1;     // halo-tile buffers
    static constexpr int oX = 0;
    static constexpr int oE = XB * XS;
    static constexpr int oD = oE + 2 * E_BYTES;
    static constexpr int oWe = oD + 2 * D_BYTES;
    static constexpr int oWp = oWe + 2 * WE_BYTES;
    static constexpr int oTab = oWp + 2 * WP_BYTES;
    static constexpr int smem_bytes = oTab + TAB_BYTES;
    static constexpr int TM_PROJ = 2 * NMT * CE;           // expand accumulators: 2 buffers x 2 row tiles x 64 columns
    static constexpr int TM_COLS = 512;
    static_assert(TM_PROJ + COUT <= 512 && smem_bytes <= 227 * 1024, "budget");
    static_assert(CEXP % CE == 0 && CIN % 16 == 0 && COUT % 32 == 0, "shape");
    // the second expand row tile reads (256 - PIN) rows past the halo tile: they must stay inside the allocation
    static_assert((XB - 1) * XS + (CIN / 8 - 1) * PIN * 16 + NMT * 128 * 16 <= smem_bytes, "A-tile overrun");
};

#ifdef FSCNN_PHASE_TIMING   // debug build only: clock64 stamps of chunks 8..11 of CTA 5 of the <64,64> kernel, one thread per role
__device__ long long g_s1_phase[64];
#define S1_STAMP(role_tid, slot) do { if (CIN == 64 && blockIdx.x == 5 && tid == (role_tid) && g >= 8 && g < 12) g_s1_phase[(g - 8) * 8 + (slot)] = clock64(); } while (0)
extern "C" int fscnn_debug_s1_phases(long long* out64) {
    return cudaMemcpyFromSymbol(out64, g_s1_phase, sizeof(long long) * 64) == cudaSuccess ? 0 : -1;
}
#else
#define S1_STAMP(role_tid, slot) do { } while (0)
#endif

template <int CIN, int COUT, bool RES>
__global__ void __launch_bounds__(kS1Threads, 1)
bottleneck_s1_kernel(const __grid_constant__ CUtensorMap xmap, const bf16* __restrict__ in, const unsigned char* __restrict__ tab_img,
                     const bf16* __restrict__ we_img, const bf16* __restrict__ wp_img, bf16* __restrict__ out, int H, int W,
                     int tiles_x, int tiles_y, int ntiles) {
    using C = S1Cfg<CIN, COUT>;
    constexpr int CE = C::CE, IW = C::IW, NCH = C::NCH, PIN = C::PIN, XB = C::XB;
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar_we[2], bar_wp[2], bar_exp[2], bar_tmfree[2], bar_eready[2], bar_efree[2], bar_dready[2],
        bar_proj[2], bar_x[2], bar_tab, bar_projfree, bar_tiledone;
    __shared__ uint32_t tmem_base_s;
    const float* Be_all = reinterpret_cast<const float*>(sm + C::oTab + 9 * C::CEXP * 2);
    const float* Bd_all = Be_all + C::CEXP;
    const float* Bp_s = Bd_all + C::CEXP;
    const uint32_t sX = smem_u32(sm + C::oX), sE = smem_u32(sm + C::oE), sD = smem_u32(sm + C::oD);
    const uint32_t sWe = smem_u32(sm + C::oWe), sWp = smem_u32(sm + C::oWp), sWd = smem_u32(sm + C::oTab);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int gstep = gridDim.x;
    const int my_tiles = (ntiles - (int)blockIdx.x + gstep - 1) / gstep;
    const int total = my_tiles * NCH;                     // chunks this CTA processes, numbered g = lt * NCH + e
    auto tile_origin = [&](int lt, int& n, int& oy0, int& ox0) {
        const int tile = blockIdx.x + lt * gstep;
        const int tx = tile % tiles_x, r = tile / tiles_x;
        n = r / tiles_y; oy0 = (r % tiles_y) * C::TH; ox0 = tx * C::TW;
    };

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_we[i], 1); mbar_init(&bar_wp[i], 1); mbar_init(&bar_exp[i], 1); mbar_init(&bar_proj[i], 1); mbar_init(&bar_x[i], 1);
            mbar_init(&bar_tmfree[i], kRoleWarps); mbar_init(&bar_eready[i], kRoleWarps);
            mbar_init(&bar_efree[i], kRoleWarps); mbar_init(&bar_dready[i], kRoleWarps);
        }
        mbar_init(&bar_tab, 1); mbar_init(&bar_projfree, kRoleWarps); mbar_init(&bar_tiledone, 1);
        fence_mbar_init();
    }
    if (warp == 0) { tmem_alloc(&tmem_base_s, C::TM_COLS); tmem_relinquish(); }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_base_s;

    if (warp == 2 * kRoleWarps) {
        // =========================== control warp ===========================
        if (lane == 0) {
            auto prefetch_we = [&](int g) {
                mbar_arrive_expect_tx(&bar_we[g & 1], C::WE_BYTES);
                bulk_g2s(sm + C::oWe + (g & 1) * C::WE_BYTES, we_img + (size_t)(g % NCH) * CE * CIN, C::WE_BYTES, &bar_we[g & 1]);
            };
            auto prefetch_wp = [&](int g) {
                mbar_arrive_expect_tx(&bar_wp[g & 1], C::WP_BYTES);
                bulk_g2s(sm + C::oWp + (g & 1) * C::WP_BYTES, wp_img + (size_t)(g % NCH) * COUT * CE, C::WP_BYTES, &bar_wp[g & 1]);
            };
            auto load_x = [&](int lt) {
                int n, oy0, ox0;
                tile_origin(lt, n, oy0, ox0);
                const int xb = lt % XB;
                mbar_arrive_expect_tx(&bar_x[xb], C::X_BYTES);
                tma_load_halo(sX + xb * C::XS, &xmap, ox0 - 1, oy0 - 1, n, &bar_x[xb]);
            };
            constexpr uint32_t idesc_exp = make_idesc_bf16(128, CE);
            constexpr uint32_t idesc_proj = make_idesc_bf16(128, COUT);
            tma_prefetch_desc(&xmap);
            mbar_arrive_expect_tx(&bar_tab, C::TAB_BYTES);
            bulk_g2s(sm + C::oTab, tab_img, C::TAB_BYTES, &bar_tab);
            load_x(0);
            if (XB == 2 && my_tiles > 1) load_x(1);
            prefetch_we(0);
            if (total > 1) prefetch_we(1);
            prefetch_wp(0);
            if (total > 1) prefetch_wp(1);
            // event loop: ke / kp = next expand / project to issue, kw / kq = next expand / project weight chunk to fetch,
            // kx = next tile whose halo to load, exp_done / proj_done = MMAs known to have completed (tracked in order, so
            // a parity test never looks at a barrier that is more than one phase ahead)
            int ke = 0, kp = 0, kw = 2, kq = 2, kx = XB, exp_done = 0, proj_done = 0;
#pragma unroll 1
            while (kp < total) {
                // ---- expand(ke): needs its halo tile (first chunk of a tile), its weight chunk, and the TMEM buffer drained ----
                if (ke < total) {
                    const int lt = ke / NCH, e = ke - lt * NCH, xb = lt % XB;
                    bool ok = mbar_test(&bar_we[ke & 1], (ke >> 1) & 1);
                    if (ok && e == 0) ok = mbar_test(&bar_x[xb], (lt / XB) & 1);
                    if (ok && ke >= 2) ok = mbar_test(&bar_tmfree[ke & 1], ((ke - 2) >> 1) & 1);
                    if (ok) {
                        if (ke >= 2 && exp_done < ke - 1) exp_done = ke - 1;     // its epilogue ran: expand(ke-2) has completed
                        tc_fence_after_sync();
#pragma unroll
                        for (int mt = 0; mt < C::NMT; ++mt)
#pragma unroll
                            for (int k16 = 0; k16 < CIN / 16; ++k16) {
                                const uint64_t da = make_smem_desc(sX + xb * C::XS + mt * 2048 + k16 * 2 * (PIN * 16), PIN * 16, 128);
                                const uint64_t db = make_smem_desc(sWe + (ke & 1) * C::WE_BYTES + k16 * 2 * (CE * 16), CE * 16, 128);
                                umma_bf16_ss(tmem + (ke & 1) * (C::NMT * CE) + mt * CE, da, db, idesc_exp, k16 > 0);
                            }
                        umma_commit(&bar_exp[ke & 1]);
                        ++ke;
                    }
                }
                // ---- completion tracking of the expand MMAs, in order ----
                if (exp_done < ke && mbar_test(&bar_exp[exp_done & 1], (exp_done >> 1) & 1)) ++exp_done;
                // weight chunk kw goes where chunk kw-2 was: free once expand(kw-2) has completed
                if (kw < total && exp_done > kw - 2) { prefetch_we(kw); ++kw; }
                // halo tile kx goes where tile kx-XB was: free once that tile's last expand has completed
                if (kx < my_tiles && exp_done > (kx - XB) * NCH + NCH - 1) { load_x(kx); ++kx; }
                // ---- project(kp): needs D written, its weight chunk, and (first chunk of a tile) the previous tile's accumulator read ----
                {
                    const int lt = kp / NCH, e = kp - lt * NCH;
                    bool ok = mbar_test(&bar_dready[kp & 1], (kp >> 1) & 1);
                    if (ok) ok = mbar_test(&bar_wp[kp & 1], (kp >> 1) & 1);
                    if (ok && e == 0 && lt > 0) ok = mbar_test(&bar_projfree, (lt - 1) & 1);
                    if (ok) {
                        if (kp >= 2 && proj_done < kp - 1) proj_done = kp - 1;   // D[kp&1] was rewritten: project(kp-2) has completed
                        tc_fence_after_sync();
#pragma unroll
                        for (int k16 = 0; k16 < CE / 16; ++k16) {
                            const uint64_t da = make_smem_desc(sD + (kp & 1) * C::D_BYTES + k16 * 2 * 2048, 2048, 128);
                            const uint64_t db = make_smem_desc(sWp + (kp & 1) * C::WP_BYTES + k16 * 2 * (COUT * 16), COUT * 16, 128);
                            umma_bf16_ss(tmem + C::TM_PROJ, da, db, idesc_proj, (e | k16) != 0);
                        }
                        umma_commit(&bar_proj[kp & 1]);
                        // one phase per TILE for the expand warps' output epilogue: they do not follow bar_proj chunk by chunk,
                        // and a parity wait on a barrier that is several phases ahead would alias
                        if (e == NCH - 1) umma_commit(&bar_tiledone);
                        ++kp;
                    }
                }
                if (proj_done < kp && mbar_test(&bar_proj[proj_done & 1], (proj_done >> 1) & 1)) ++proj_done;
                if (kq < total && proj_done > kq - 2) { prefetch_wp(kq); ++kq; }
            }
        }
    } else if (warp >= kRoleWarps) {
        // =========================== expand warps (8..15): TMEM -> E, and the output epilogue ===========================
        // The SM's warp arbiter prefers the highest warp id among eligible warps, and a depthwise warp in its FMA phase is
        // always eligible: the producers must carry the higher ids or they are starved and the pipeline runs in lock-step.
        const int q = warp & 3, h = (warp >> 2) & 1;      // TMEM lane quarter, 32-column half
        mbar_wait(&bar_tab, 0);
        int n = 0, oy0 = 0, ox0 = 0, pn = 0, poy0 = 0, pox0 = 0;
        auto output_epilogue = [&](int lt, int tn, int toy0, int tox0) {   // + bias (+ residual from global, L2-resident), bf16 NHWC store
            mbar_wait(&bar_tiledone, lt & 1);                 // every project MMA of tile lt has completed
            tc_fence_after_sync();
            const int p = q * 32 + lane;
            const int oy = toy0 + (p >> 4), ox = tox0 + (p & 15);
            const bool live = (oy < H) && (ox < W);
            const size_t pix = ((size_t)tn * H + oy) * W + ox;
            constexpr int CP = COUT / 2;                  // columns per warp: 32, 48 or 64
#pragma unroll
            for (int c0 = 0; c0 < CP; c0 += 16) {
                uint32_t r[16];
                tmem_ld_32x32b_x16(tmem + ((uint32_t)(q * 32) << 16) + C::TM_PROJ + h * CP + c0, r);
                uint4 res[2];
                if (RES && live) {
                    res[0] = __ldg(reinterpret_cast<const uint4*>(in + pix * CIN + h * CP + c0));
                    res[1] = __ldg(reinterpret_cast<const uint4*>(in + pix * CIN + h * CP + c0 + 8));
                }
                tmem_ld_wait();
                if (live) {
#pragma unroll
                    for (int g8 = 0; g8 < 2; ++g8) {
                        const int co = h * CP + c0 + g8 * 8;
                        const float4 ba = *reinterpret_cast<const float4*>(Bp_s + co);
                        const float4 bb = *reinterpret_cast<const float4*>(Bp_s + co + 4);
                        const uint32_t* q8 = r + g8 * 8;
                        float v[8] = {__uint_as_float(q8[0]) + ba.x, __uint_as_float(q8[1]) + ba.y, __uint_as_float(q8[2]) + ba.z,
                                      __uint_as_float(q8[3]) + ba.w, __uint_as_float(q8[4]) + bb.x, __uint_as_float(q8[5]) + bb.y,
                                      __uint_as_float(q8[6]) + bb.z, __uint_as_float(q8[7]) + bb.w};
                        if (RES) {
                            float f[8];
                            unpackbf8(res[g8], f);
#pragma unroll
                            for (int i = 0; i < 8; ++i) v[i] += f[i];
                        }
                        *reinterpret_cast<uint4*>(out + pix * COUT + co) =
                            make_uint4(packbf(v[0], v[1]), packbf(v[2], v[3]), packbf(v[4], v[5]), packbf(v[6], v[7]));
                    }
                }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_projfree);       // the next tile's first project MMA may overwrite the accumulator
        };
#pragma unroll 1
        for (int g = 0; g < total; ++g) {
            const int lt = g / NCH, e = g - lt * NCH;
            if (e == 0) { pn = n; poy0 = oy0; pox0 = ox0; tile_origin(lt, n, oy0, ox0); }
            const int iy0 = oy0 - 1, ix0 = ox0 - 1;
            const float* Bes = Be_all + e * CE;
            const uint32_t sEb = sE + (g & 1) * C::E_BYTES;
            S1_STAMP(256, 0);
            mbar_wait(&bar_exp[g & 1], (g >> 1) & 1);            // expand(g) has completed
            tc_fence_after_sync();
            S1_STAMP(256, 1);
            if (g >= 2) mbar_wait(&bar_efree[g & 1], ((g - 2) >> 1) & 1);   // depthwise(g-2) has finished reading E[g&1]
            S1_STAMP(256, 2);
#pragma unroll
            for (int mt = 0; mt < C::NMT; ++mt) {
                const int pin = mt * 128 + q * 32 + lane;
                const int iy = iy0 + pin / IW, ix = ix0 + pin % IW;
                const bool ok = pin < PIN && iy >= 0 && iy < H && ix >= 0 && ix < W;
                const int c0 = h * 32;
                uint32_t r[32];
                tmem_ld_32x32b_x32(tmem + ((uint32_t)(q * 32) << 16) + (g & 1) * (C::NMT * CE) + mt * CE + c0, r);
                tmem_ld_wait();
                if (pin < PIN) {
                    if (ok) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float4 ba = *reinterpret_cast<const float4*>(Bes + c0 + k * 8);
                            const float4 bb = *reinterpret_cast<const float4*>(Bes + c0 + k * 8 + 4);
                            const uint32_t* q8 = r + k * 8;
                            sts128(sEb + pin * (CE * 2) + ((((c0 >> 3) + k) ^ (pin & 7)) << 4),
                                   packbf_relu(__uint_as_float(q8[0]) + ba.x, __uint_as_float(q8[1]) + ba.y),
                                   packbf_relu(__uint_as_float(q8[2]) + ba.z, __uint_as_float(q8[3]) + ba.w),
                                   packbf_relu(__uint_as_float(q8[4]) + bb.x, __uint_as_float(q8[5]) + bb.y),
                                   packbf_relu(__uint_as_float(q8[6]) + bb.z, __uint_as_float(q8[7]) + bb.w));
                        }
                    } else {     // outside the image: the depthwise zero padding
#pragma unroll
                        for (int k = 0; k < 4; ++k) sts128(sEb + pin * (CE * 2) + ((((c0 >> 3) + k) ^ (pin & 7)) << 4), 0u, 0u, 0u, 0u);
                    }
                }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) { mbar_arrive(&bar_tmfree[g & 1]); mbar_arrive(&bar_eready[g & 1]); }
            S1_STAMP(256, 3);
            if (e == 0 && lt >= 1) output_epilogue(lt - 1, pn, poy0, pox0);   // deferred by one chunk: keeps the next tile's pipeline fed
        }
        output_epilogue(my_tiles - 1, n, oy0, ox0);
    } else {
        // =========================== depthwise warps (0..7): E -> D ===========================
        const int dt = tid;
        const int x = dt & 15, half = (dt >> 4) & 1, j = dt >> 5;   // output column, rows 4*half .. 4*half+3, 8-channel chunk
        mbar_wait(&bar_tab, 0);
#pragma unroll 1
        for (int g = 0; g < total; ++g) {
            const int e = g % NCH;
            const float* Bds = Bd_all + e * CE;
            const uint32_t sWds = sWd + e * CE * 2;      // tap t, channel c at sWds + (t * CEXP + c) * 2
            const uint32_t sEb = sE + (g & 1) * C::E_BYTES, sDb = sD + (g & 1) * C::D_BYTES;
            uint4 wv[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) wv[t] = lds128(sWds + (t * C::CEXP + j * 8) * 2);
            float acc[4][8];
            {
                const float4 ba = *reinterpret_cast<const float4*>(Bds + j * 8);
                const float4 bb = *reinterpret_cast<const float4*>(Bds + j * 8 + 4);
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    acc[o][0] = ba.x; acc[o][1] = ba.y; acc[o][2] = ba.z; acc[o][3] = ba.w;
                    acc[o][4] = bb.x; acc[o][5] = bb.y; acc[o][6] = bb.z; acc[o][7] = bb.w;
                }
            }
            S1_STAMP(0, 4);
            mbar_wait(&bar_eready[g & 1], (g >> 1) & 1);         // E[g&1] written by the expand warps
            S1_STAMP(0, 5);
#pragma unroll
            for (int r = 0; r < 6; ++r) {
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int pin = (4 * half + r) * IW + x + kx;
                    const uint4 v = lds128(sEb + pin * (CE * 2) + ((j ^ (pin & 7)) << 4));
#pragma unroll
                    for (int o = 0; o < 4; ++o) {
                        const int ky = r - o;
                        if (ky >= 0 && ky < 3) fhfma8(acc[o], v, wv[ky * 3 + kx]);
                    }
                }
            }
            S1_STAMP(0, 6);
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_efree[g & 1]);       // every read of E[g&1] has been consumed: release it before the D hand-off
            if (g >= 2) mbar_wait(&bar_proj[g & 1], ((g - 2) >> 1) & 1);   // project(g-2) has completed: D[g&1] is free
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                const int p = (4 * half + o) * 16 + x;
                sts128(sDb + a_tile_off(p, j), packbf_relu(acc[o][0], acc[o][1]), packbf_relu(acc[o][2], acc[o][3]),
                       packbf_relu(acc[o][4], acc[o][5]), packbf_relu(acc[o][6], acc[o][7]));
            }
            fence_async_proxy();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_dready[g & 1]);
            S1_STAMP(0, 7);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, C::TM_COLS);
}

template <int CIN, int COUT, bool RES>
static cudaError_t run_s1(const bf16* in, const unsigned char* tab_img, const bf16* we_img, const bf16* wp_img, bf16* out, int n, int h,
                          int w, cudaStream_t s) {
    using C = S1Cfg<CIN, COUT>;
    static unsigned long long configured = 0;
    cudaError_t e = ensure_dyn_smem(bottleneck_s1_kernel<CIN, COUT, RES>, C::smem_bytes, configured);
    if (e != cudaSuccess) return e;
    CUtensorMap xmap;
    e = make_nhwc_halo_map(&xmap, in, n, h, w, CIN, C::IH, C::IW);
    if (e != cudaSuccess) return e;
    const int tiles_x = ceil_div(w, C::TW), tiles_y = ceil_div(h, C::TH), ntiles = tiles_x * tiles_y * n;
    const int grid = ntiles < num_sms() ? ntiles : num_sms();
    bottleneck_s1_kernel<CIN, COUT, RES><<<grid, kS1Threads, C::smem_bytes, s>>>(xmap, in, tab_img, we_img, wp_img, out, h, w, tiles_x,
                                                                                  tiles_y, ntiles);
    return cudaGetLastError();
}

cudaError_t launch_bottleneck_s1_tc(int cin, int cout, const bf16* in, const unsigned char* tab_img, const bf16* we_img,
                                    const bf16* wp_img, bf16* out, int n, int h, int w, cudaStream_t s) {
    if (cin == 64 && cout == 64) return run_s1<64, 64, true>(in, tab_img, we_img, wp_img, out, n, h, w, s);
    if (cin == 96 && cout == 96) return run_s1<96, 96, true>(in, tab_img, we_img, wp_img, out, n, h, w, s);
    if (cin == 96 && cout == 128) return run_s1<96, 128, false>(in, tab_img, we_img, wp_img, out, n, h, w, s);
    if (cin == 128 && cout == 128) return run_s1<128, 128, true>(in, tab_img, we_img, wp_img, out, n, h, w, s);
    return cudaErrorInvalidValue;
}

}  // namespace fscnn
