// bottleneck_s1t_tc.cu -- bf16 LinearBottleneck, stride 1 (reference models/fast_scnn.py:95-115), with the expand 1x1
// computed TRANSPOSED so that the expanded tensor never touches shared memory:
//
//   expand   E^T[128 ch x 192 px] = We^T[128 ch x CIN] * X^T   (tcgen05.mma, A = weight chunk, B = the TMA halo tile, which
//            is already K-major [pixel][cin]).  TMEM lane = expanded channel, TMEM column = halo pixel (10 x 18 = 180 used).
//            The BN-folded expand bias rides in 16 extra K columns (bf16 head + remainder) against a constant block of ones.
//   depthwise each thread owns ONE expanded channel (its TMEM lane) and two output rows: it pulls 4 halo rows x 18 columns
//            out of TMEM (tcgen05.ld), ReLU + bf16-packs them in registers (cvt.rn.relu.bf16x2) and runs the 3x3 taps as
//            FHFMA.BF16 with its 9 weights in registers.  No E tile, no LDS, no bank conflicts.
//   project  D[128 px x 128 ch] (bf16, MN-major A operand written by the depthwise threads) * Wp chunk -> TMEM accumulator
//            over the chunks of 128 expanded channels; + bias (+ residual) -> bf16 NHWC.
//
// One persistent CTA per SM: 16 compute warps (TMEM lane quarter = warp % 4, row strip = warp / 4) + two controller warps
// (converged, issuing through elect.sync): one issues the halo loads, expand weight copies and expand MMAs, the other the project weight copies and
// project MMAs, both in order with blocking mbarrier waits.  The expand accumulator and the
// weight chunks are double-buffered: the tensor core works two chunks ahead of the CUDA cores; every hand-off is an mbarrier.
#include "kernels.h"
#include "tma_host.h"
#include "umma.cuh"

namespace fscnn {

namespace {
constexpr int kTWarps = 16;
constexpr int kTThreads = (kTWarps + 2) * 32;   // 16 compute warps + expand controller + project controller

}  // namespace

template <int CIN, int COUT>
struct T1Cfg {
    static constexpr int TH = 8, TW = 16, IH = 10, IW = 18, PIN = IH * IW, NB = 192;
    static constexpr int CM = 128, CEXP = 6 * CIN, NCH = (CEXP + CM - 1) / CM;
    static constexpr int KA = CIN + 16;                        // expand K: input channels + the bias columns
    static constexpr int X_BYTES = PIN * CIN * 2;              // [CIN/8][PIN][8]: LBO = PIN*16, SBO = 128
    static constexpr int XS = X_BYTES;
    static constexpr int WE_BYTES = CM * KA * 2;               // A image [KA/8][128 rows][8]: LBO = 2048, SBO = 128
    static constexpr int WP_BYTES = COUT * CM * 2;             // B image [128/8][COUT rows][8]: LBO = COUT*16, SBO = 128
    // D (project A operand, 128 px x 128 ch) is stored MN-major: 16 bytes = 8 consecutive pixels of one channel, a core
    // matrix = 8 channels x 16 B, channel blocks 128 B apart (LBO), pixel blocks 2048 B apart (SBO).  A depthwise thread owns
    // one channel and a row of 16 pixels: two 16-byte stores per row, and a warp's store covers 512 contiguous bytes.
    static constexpr int D_LBO = 128, D_SBO = 2048;
    static constexpr int D_BYTES = 16 * D_SBO;
    static constexpr int ONES_BYTES = 256;                     // one core matrix of B rows {1, 1, 0 x 6} + one of zeros, re-read for every
                                                               // 8-row group (SBO = 0): the bias columns' partner
    static constexpr int LIMIT = 227 * 1024 - 256;
    // Every CTA needs every weight chunk for every tile: streamed, that is (We + Wp) x 148 SMs out of the same L2 lines per
    // chunk step, and the copies (not the MMAs or the FMAs) set the pace.  With 64 input channels all of it (108 KB) stays
    // resident in shared memory instead.
    static constexpr bool WRES = NCH * (WE_BYTES + WP_BYTES) + 2 * XS + 2 * D_BYTES + ONES_BYTES <= LIMIT;
    static constexpr int WEB = WRES ? NCH : 2;                 // expand weight buffers
    // streamed: weights first (a chunk's copy must overlap the chunk before), then D, then the halo tile
    static constexpr int FIXED = WEB * WE_BYTES + D_BYTES + ONES_BYTES + XS;
    static constexpr int WPB = WRES ? NCH : ((FIXED + 2 * WP_BYTES <= LIMIT) ? 2 : 1);
#ifdef S1T_PREFER_X
    static constexpr int XB = (FIXED + WPB * WP_BYTES + XS <= LIMIT) ? 2 : 1;
    static constexpr int DB = (FIXED + WPB * WP_BYTES + (XB - 1) * XS + D_BYTES <= LIMIT) ? 2 : 1;
#else
    static constexpr int DB = (FIXED + WPB * WP_BYTES + D_BYTES <= LIMIT) ? 2 : 1;
    static constexpr int XB = (FIXED + WPB * WP_BYTES + (DB - 1) * D_BYTES + XS <= LIMIT) ? 2 : 1;
#endif
    static constexpr int oX = 0;
    static constexpr int oWe = XB * XS;
    static constexpr int oWp = oWe + WEB * WE_BYTES;
    static constexpr int oD = oWp + WPB * WP_BYTES;
    static constexpr int oOnes = oD + DB * D_BYTES;
    static constexpr int smem_bytes = oOnes + ONES_BYTES;
    static constexpr int TM_PROJ = 2 * NB;                     // expand accumulators: 2 buffers x 192 columns
    static constexpr int TM_COLS = 512;
    static constexpr int TAB_BYTES = NCH * CM * 32 + COUT * 4; // per expanded channel {bf16 w[9], pad, f32 bd, pad} | f32 Bp[COUT]
    static_assert(TM_PROJ + COUT <= 512 && smem_bytes <= LIMIT, "budget");
    static_assert(CIN % 16 == 0 && COUT % 32 == 0 && XS % 128 == 0, "shape");
};

#ifdef FSCNN_PHASE_TIMING   // debug build only: clock64 stamps of chunks 12..19 of CTA 5 of the <DBG_CIN,*> kernel
#ifndef DBG_CIN
#define DBG_CIN 128
#endif
__device__ long long g_s1t_phase[8 * 16 * 3];   // [0,128): roles; [128,256): per-warp "D written"; [256,384): per-warp "top"
#define T_STAMP(cond, gg, slot) do { if (CIN == DBG_CIN && blockIdx.x == 5 && (cond) && (gg) >= 12 && (gg) < 20) g_s1t_phase[((slot) / 128) * 128 + ((gg) - 12) * 16 + ((slot) % 128)] = clock64(); } while (0)
extern "C" int fscnn_debug_s1t_phases(long long* out128) {
    return cudaMemcpyFromSymbol(out128, g_s1t_phase, sizeof(long long) * 384) == cudaSuccess ? 0 : -1;
}
#else
#define T_STAMP(cond, gg, slot) do { } while (0)
#endif

template <int CIN, int COUT, bool RES>
__global__ void __launch_bounds__(kTThreads, 1)   // 96 registers: the allocation granule is 1024 per warp (112 x 576 threads does not launch)
bottleneck_s1t_kernel(const __grid_constant__ CUtensorMap xmap, const bf16* __restrict__ in, const unsigned char* __restrict__ tab,
                      const bf16* __restrict__ we_img, const bf16* __restrict__ wp_img, bf16* __restrict__ out, int H, int W,
                      int tiles_x, int tiles_y, int ntiles) {
    using C = T1Cfg<CIN, COUT>;
    constexpr int IW = C::IW, NCH = C::NCH, PIN = C::PIN, XB = C::XB, WPB = C::WPB, DB = C::DB, CM = C::CM;
    constexpr bool WRES = C::WRES;
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar_we[2], bar_wp[2], bar_wres, bar_exp[2], bar_tmfree[2], bar_x[2], bar_dready[2], bar_proj[2],
        bar_projfree, bar_tiledone;
    __shared__ uint32_t tmem_base_s;
    const uint32_t sX = smem_u32(sm + C::oX), sWe = smem_u32(sm + C::oWe), sWp = smem_u32(sm + C::oWp), sD = smem_u32(sm + C::oD),
                   sOnes = smem_u32(sm + C::oOnes);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int gstep = gridDim.x;
    const int my_tiles = (ntiles - (int)blockIdx.x + gstep - 1) / gstep;
    const int total = my_tiles * NCH;                     // chunks this CTA processes, numbered g = lt * NCH + e
    auto tile_origin = [&](int lt, int& n, int& oy0, int& ox0) {
        const int tile = blockIdx.x + lt * gstep;
        const int tx = tile % tiles_x, r = tile / tiles_x;
        n = r / tiles_y; oy0 = (r % tiles_y) * C::TH; ox0 = tx * C::TW;
    };
    pdl_launch_dependents();

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_we[i], 1); mbar_init(&bar_wp[i], 1); mbar_init(&bar_exp[i], 1); mbar_init(&bar_x[i], 1);
            mbar_init(&bar_tmfree[i], kTWarps); mbar_init(&bar_dready[i], kTWarps); mbar_init(&bar_proj[i], 1);
        }
        mbar_init(&bar_projfree, kTWarps); mbar_init(&bar_tiledone, 1); mbar_init(&bar_wres, 1);
        fence_mbar_init();
    }
    if (tid < 16)                                         // the constant B block of the bias columns
        *reinterpret_cast<uint4*>(sm + C::oOnes + tid * 16) = make_uint4(tid < 8 ? 0x3F803F80u : 0u, 0u, 0u, 0u);
    fence_async_proxy();
    if (warp == 0) { tmem_alloc(&tmem_base_s, C::TM_COLS); tmem_relinquish(); }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_base_s;

    if (warp == kTWarps) {
        // =========================== expand controller ===========================
        // In-order, blocking: wait for what expand(ke) needs, issue it, wait for it to complete, recycle the buffers it read.
        // (An event loop polling every barrier from one thread reacted thousands of cycles late: each poll is a dependent
        // test + branch of a single warp sharing its scheduler with four FMA-bound warps.  mbarrier.try_wait parks the warp in
        // hardware instead, and the expand / project sides no longer queue behind each other's blocking MMA issue.)
        // (the whole warp walks the loop converged; asynchronous operations are issued by the lane elect.sync names: behind a
        // `lane == 0` branch every tcgen05.mma / bulk copy gets a vote loop over the active lanes, and this thread's instruction
        // latency sits on the weight-streaming path of the 96- / 128-channel layers -- see bottleneck_s2t_tc.cu)
        {
            auto prefetch_we = [&](int g) {
                if (elect_one()) {
                    mbar_arrive_expect_tx(&bar_we[g & 1], C::WE_BYTES);
                    bulk_g2s(sm + C::oWe + (g & 1) * C::WE_BYTES, we_img + (size_t)(g % NCH) * CM * C::KA, C::WE_BYTES, &bar_we[g & 1]);
                }
            };
            auto load_x = [&](int lt) {
                int n, oy0, ox0;
                tile_origin(lt, n, oy0, ox0);
                const int xb = lt % XB;
                if (elect_one()) {
                    mbar_arrive_expect_tx(&bar_x[xb], C::X_BYTES);
                    tma_load_halo(sX + xb * C::XS, &xmap, ox0 - 1, oy0 - 1, n, &bar_x[xb]);
                }
            };
            constexpr uint32_t idesc_exp = make_idesc_bf16(128, C::NB);
            if (elect_one()) tma_prefetch_desc(&xmap);
            if (WRES) {      // all weight chunks, once (the project controller waits on the same barrier)
                if (elect_one()) {
                    mbar_arrive_expect_tx(&bar_wres, NCH * (C::WE_BYTES + C::WP_BYTES));
                    for (int e = 0; e < NCH; ++e) {
                        bulk_g2s(sm + C::oWe + e * C::WE_BYTES, we_img + (size_t)e * CM * C::KA, C::WE_BYTES, &bar_wres);
                        bulk_g2s(sm + C::oWp + e * C::WP_BYTES, wp_img + (size_t)e * COUT * CM, C::WP_BYTES, &bar_wres);
                    }
                }
                __syncwarp();
            } else {
                prefetch_we(0);
                if (total > 1) prefetch_we(1);
            }
            pdl_wait();      // the weights are on their way; the halo tiles are the previous stage's output
            load_x(0);
            if (XB == 2 && my_tiles > 1) load_x(1);
            if (WRES) mbar_wait(&bar_wres, 0);
#pragma unroll 1
            for (int ke = 0; ke < total; ++ke) {
                const int lt = ke / NCH, e = ke - lt * NCH, xb = lt % XB;
                if (!WRES) mbar_wait(&bar_we[ke & 1], (ke >> 1) & 1);                    // its weight chunk
                if (e == 0) mbar_wait(&bar_x[xb], (lt / XB) & 1);                         // its halo tile
                if (ke >= 2) mbar_wait(&bar_tmfree[ke & 1], ((ke - 2) >> 1) & 1);         // its accumulator, drained by chunk ke-2
                T_STAMP(lane == 0, ke, 10);
                tc_fence_after_sync();
                // descriptors advance by adding to the 14-bit start-address field (bytes >> 4): a handful of instructions per MMA
                const uint64_t da0 = make_smem_desc(sWe + (WRES ? e : (ke & 1)) * C::WE_BYTES, 2048, 128);
                const uint64_t db0 = make_smem_desc(sX + xb * C::XS, PIN * 16, 128);
                const uint32_t dacc = tmem + (ke & 1) * C::NB;
                if (elect_one()) {
#pragma unroll
                    for (int k16 = 0; k16 < CIN / 16; ++k16)
                        umma_bf16_ss(dacc, da0 + (uint64_t)(k16 * ((2 * 2048) >> 4)), db0 + (uint64_t)(k16 * ((2 * PIN * 16) >> 4)), idesc_exp, k16 > 0);
                    umma_bf16_ss(dacc, da0 + (uint64_t)((CIN / 16) * ((2 * 2048) >> 4)), make_smem_desc(sOnes, 128, 0), idesc_exp, 1);
                    umma_commit(&bar_exp[ke & 1]);
                }
                __syncwarp();
                T_STAMP(lane == 0, ke, 11);
                const bool more_w = !WRES && ke + 2 < total, more_x = (e == NCH - 1) && (lt + XB < my_tiles);
                if (more_w || more_x) {
                    mbar_wait(&bar_exp[ke & 1], (ke >> 1) & 1);       // expand(ke) has completed: what it read may be overwritten
                    T_STAMP(lane == 0, ke, 12);
                    if (more_w) prefetch_we(ke + 2);
                    if (more_x) load_x(lt + XB);
                }
            }
        }
    } else if (warp == kTWarps + 1) {
        // =========================== project controller ===========================
        {
            auto prefetch_wp = [&](int g) {
                if (elect_one()) {
                    mbar_arrive_expect_tx(&bar_wp[g % WPB], C::WP_BYTES);
                    bulk_g2s(sm + C::oWp + (g % WPB) * C::WP_BYTES, wp_img + (size_t)(g % NCH) * COUT * CM, C::WP_BYTES, &bar_wp[g % WPB]);
                }
            };
            constexpr uint32_t idesc_proj = make_idesc_bf16(128, COUT) | (1u << 15);   // A (= D) is MN-major
            if (WRES) {
                mbar_wait(&bar_wres, 0);
            } else {
                prefetch_wp(0);
                if (WPB == 2 && total > 1) prefetch_wp(1);
            }
#pragma unroll 1
            for (int kp = 0; kp < total; ++kp) {
                const int lt = kp / NCH, e = kp - lt * NCH;
                if (!WRES) mbar_wait(&bar_wp[kp % WPB], (kp / WPB) & 1);                  // its weight chunk
                if (e == 0 && lt > 0) mbar_wait(&bar_projfree, (lt - 1) & 1);             // the previous tile's accumulator has been read
                mbar_wait(&bar_dready[kp % DB], (kp / DB) & 1);                           // D written by the depthwise threads
                T_STAMP(lane == 0, kp, 8);
                tc_fence_after_sync();
                const uint64_t da0 = make_smem_desc(sD + (kp % DB) * C::D_BYTES, C::D_LBO, C::D_SBO);
                const uint64_t db0 = make_smem_desc(sWp + (WRES ? e : kp % WPB) * C::WP_BYTES, COUT * 16, 128);
                const bool half = (C::CEXP - e * CM) < CM;       // the last chunk of a 576-channel layer holds 64 channels
                if (elect_one()) {
#pragma unroll
                    for (int k16 = 0; k16 < CM / 16; ++k16)
                        if (k16 < CM / 32 || !half)
                            umma_bf16_ss(tmem + C::TM_PROJ, da0 + (uint64_t)(k16 * ((2 * C::D_LBO) >> 4)),
                                         db0 + (uint64_t)(k16 * ((2 * COUT * 16) >> 4)), idesc_proj, (e | k16) != 0);
                    umma_commit(&bar_proj[kp % DB]);
                    if (e == NCH - 1) umma_commit(&bar_tiledone);   // one phase per TILE for the output epilogue
                }
                __syncwarp();
                T_STAMP(lane == 0, kp, 9);
                if (!WRES && kp + WPB < total) {
                    mbar_wait(&bar_proj[kp % DB], (kp / DB) & 1);     // project(kp) has completed: its weight buffer is free
                    T_STAMP(lane == 0, kp, 13);
                    prefetch_wp(kp + WPB);
                }
            }
        }
    } else {
        // =========================== compute warps ===========================
        pdl_wait();      // (the residual is read from global memory, the output written to it)
        const int q = warp & 3, s = warp >> 2;            // TMEM lane quarter = warp % 4 (hardware rule), row strip (output rows 2s, 2s+1)
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const float* Bp_g = reinterpret_cast<const float*>(tab + (size_t)NCH * CM * 32);
        int n = 0, oy0 = 0, ox0 = 0, pn = 0, poy0 = 0, pox0 = 0;
        auto output_epilogue = [&](int lt, int tn, int toy0, int tox0) {   // + bias (+ residual from global, L2-resident), bf16 NHWC store
            const int p = q * 32 + lane;
            const int oy = toy0 + (p >> 4), ox = tox0 + (p & 15);
            const bool live = (oy < H) && (ox < W);
            const size_t pix = ((size_t)tn * H + oy) * W + ox;
            constexpr int CP = COUT / 4;                  // columns per warp: 16, 24 or 32
            const int co0 = s * CP;
            // every global load is in flight before the first wait: the latency is paid once per tile, not once per piece
            uint4 res[CP / 8];
#pragma unroll
            for (int i = 0; i < CP / 8; ++i) {
                res[i] = make_uint4(0u, 0u, 0u, 0u);
                if (RES && live) res[i] = __ldg(reinterpret_cast<const uint4*>(in + pix * CIN + co0 + 8 * i));
            }
            mbar_wait(&bar_tiledone, lt & 1);                 // every project MMA of tile lt has completed
            tc_fence_after_sync();
            // accumulator columns in pieces of 16 (8 for the tail of a 24-column slice): keeps the register footprint small
#pragma unroll
            for (int c0 = 0; c0 < CP; c0 += 16) {
                constexpr int kDummy = 0; (void)kDummy;
                const int w16 = (CP - c0 >= 16) ? 16 : 8;
                uint32_t r[16];
                if (w16 == 16) tmem_ld_32x32b_x16(tmem + lane_base + C::TM_PROJ + co0 + c0, r);
                else tmem_ld_32x32b_x8(tmem + lane_base + C::TM_PROJ + co0 + c0, r);
                tmem_ld_wait();
                if (c0 + 16 >= CP) {                         // last piece read: the next tile's first project MMA may overwrite the accumulator
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar_projfree);
                }
                if (live) {
#pragma unroll
                    for (int i = 0; i < w16 / 8; ++i) {
                        const int co = co0 + c0 + 8 * i;
                        const float4 ba = __ldg(reinterpret_cast<const float4*>(Bp_g + co));
                        const float4 bb = __ldg(reinterpret_cast<const float4*>(Bp_g + co + 4));
                        const uint32_t* q8 = r + 8 * i;
                        float v[8] = {__uint_as_float(q8[0]) + ba.x, __uint_as_float(q8[1]) + ba.y, __uint_as_float(q8[2]) + ba.z,
                                      __uint_as_float(q8[3]) + ba.w, __uint_as_float(q8[4]) + bb.x, __uint_as_float(q8[5]) + bb.y,
                                      __uint_as_float(q8[6]) + bb.z, __uint_as_float(q8[7]) + bb.w};
                        if (RES) {
                            float f[8];
                            unpackbf8(res[c0 / 8 + i], f);
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[j] += f[j];
                        }
                        *reinterpret_cast<uint4*>(out + pix * COUT + co) =
                            make_uint4(packbf(v[0], v[1]), packbf(v[2], v[3]), packbf(v[4], v[5]), packbf(v[6], v[7]));
                    }
                }
            }
        };
#pragma unroll 1
        for (int g = 0; g < total; ++g) {
            const int lt = g / NCH, e = g - lt * NCH;
            if (e == 0) { pn = n; poy0 = oy0; pox0 = ox0; tile_origin(lt, n, oy0, ox0); }
            const bool active = (e * CM + q * 32) < C::CEXP;     // warp-uniform: the last chunk of a 576-channel layer is half empty
            // this thread's channel: 9 bf16 taps + fp32 bias from the (L1-resident) table, requested before the wait
            uint4 wa = make_uint4(0u, 0u, 0u, 0u), wb = wa;
            if (active) {
                const uint4* rec = reinterpret_cast<const uint4*>(tab + (size_t)(e * CM + q * 32 + lane) * 32);
                wa = __ldg(rec);
                wb = __ldg(rec + 1);
            }
            T_STAMP(tid == 0, g, 0);
            T_STAMP(lane == 0, g, 256 + warp);
            mbar_wait(&bar_exp[g & 1], (g >> 1) & 1);            // expand(g) has completed
            T_STAMP(tid == 0, g, 1);
            tc_fence_after_sync();
            uint32_t Ep[4][9];                                   // halo rows 2s .. 2s+3, column pairs (2i, 2i+1), ReLU'd bf16
            if (active) {
                uint32_t r[72];
                const uint32_t t0 = tmem + lane_base + (g & 1) * C::NB + (2 * s) * IW;
                tmem_ld_32x32b_x64(t0, r);
                tmem_ld_32x32b_x8(t0 + 64, r + 64);
                tmem_ld_wait();
#pragma unroll
                for (int rr = 0; rr < 4; ++rr)
#pragma unroll
                    for (int i = 0; i < 9; ++i)
                        Ep[rr][i] = packbf_relu(__uint_as_float(r[rr * IW + 2 * i]), __uint_as_float(r[rr * IW + 2 * i + 1]));
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_tmfree[g & 1]);      // expand(g+2) may overwrite this accumulator
            T_STAMP(tid == 0, g, 2);
            if (active) {
                // zero padding of the depthwise conv: halo columns / rows outside the image (border tiles only)
                const int ix0 = ox0 - 1, iy0 = oy0 - 1 + 2 * s;
                if (ix0 == -1 && ix0 + IW <= W) {               // left image border: only halo column 0 is outside
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr) Ep[rr][0] &= 0xFFFF0000u;
                } else if (ix0 >= 0 && ix0 + IW == W + 1 && (IW & 1) == 0) {   // right border, width a multiple of the tile: only the last column
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr) Ep[rr][IW / 2 - 1] &= 0x0000FFFFu;
                } else if (ix0 < 0 || ix0 + IW > W) {           // anything else (odd sizes): per-column masks
#pragma unroll
                    for (int i = 0; i < 9; ++i) {
                        const int xa = ix0 + 2 * i, xb2 = xa + 1;
                        const uint32_t m = ((xa >= 0 && xa < W) ? 0x0000FFFFu : 0u) | ((xb2 >= 0 && xb2 < W) ? 0xFFFF0000u : 0u);
#pragma unroll
                        for (int rr = 0; rr < 4; ++rr) Ep[rr][i] &= m;
                    }
                }
                if (iy0 < 0 || iy0 + 4 > H) {
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr)
                        if (iy0 + rr < 0 || iy0 + rr >= H) {
#pragma unroll
                            for (int i = 0; i < 9; ++i) Ep[rr][i] = 0u;
                        }
                }
                const uint32_t wq[5] = {wa.x, wa.y, wa.z, wa.w, wb.x};
                const float bd = __uint_as_float(wb.y);
                float acc[2][16];
                // straight-line: halo row rr feeds output rows o = rr - ky; the first tap of every accumulator adds the bias
#pragma unroll
                for (int rr = 0; rr < 4; ++rr)
#pragma unroll
                    for (int o = 0; o < 2; ++o) {
                        const int ky = rr - o;
                        if (ky >= 0 && ky < 3) {
#pragma unroll
                            for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                                for (int x = 0; x < 16; ++x)
                                    acc[o][x] = fhfma_sel((ky | kx) ? acc[o][x] : bd, Ep[rr][(x + kx) >> 1], (x + kx) & 1,
                                                          wq[(ky * 3 + kx) >> 1], (ky * 3 + kx) & 1);
                        }
                    }
                T_STAMP(tid == 0, g, 3);
                if (g >= DB) mbar_wait(&bar_proj[g % DB], (g / DB - 1) & 1);   // project(g-DB) has completed: this D buffer is free
                T_STAMP(tid == 0, g, 4);
                const int k = q * 32 + lane;
                const uint32_t d0 = sD + (g % DB) * C::D_BYTES + (k >> 3) * C::D_LBO + (k & 7) * 16 + (4 * s) * C::D_SBO;
#pragma unroll
                for (int o = 0; o < 2; ++o)
#pragma unroll
                    for (int hx = 0; hx < 2; ++hx)
                        sts128(d0 + (2 * o + hx) * C::D_SBO, packbf_relu(acc[o][8 * hx + 0], acc[o][8 * hx + 1]),
                               packbf_relu(acc[o][8 * hx + 2], acc[o][8 * hx + 3]), packbf_relu(acc[o][8 * hx + 4], acc[o][8 * hx + 5]),
                               packbf_relu(acc[o][8 * hx + 6], acc[o][8 * hx + 7]));
            } else if (g >= DB) {
                mbar_wait(&bar_proj[g % DB], (g / DB - 1) & 1);    // stay in step with the barrier's phases
            }
            fence_async_proxy();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_dready[g % DB]);
            T_STAMP(tid == 0, g, 5);
            T_STAMP(lane == 0, g, 128 + warp);
            T_STAMP(tid == 15 * 32, g, 6);
            if (e == 0 && lt >= 1) output_epilogue(lt - 1, pn, poy0, pox0);   // deferred by one chunk: keeps the pipeline fed
        }
        output_epilogue(my_tiles - 1, n, oy0, ox0);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, C::TM_COLS);
}

template <int CIN, int COUT, bool RES>
static cudaError_t run_s1t(const bf16* in, const unsigned char* tab, const bf16* we_img, const bf16* wp_img, bf16* out, int n, int h,
                           int w, cudaStream_t s) {
    using C = T1Cfg<CIN, COUT>;
    static unsigned long long configured = 0;
    cudaError_t e = ensure_dyn_smem(bottleneck_s1t_kernel<CIN, COUT, RES>, C::smem_bytes, configured);
    if (e != cudaSuccess) return e;
    CUtensorMap xmap;
    e = make_nhwc_halo_map(&xmap, in, n, h, w, CIN, C::IH, C::IW);
    if (e != cudaSuccess) return e;
    const int tiles_x = ceil_div(w, C::TW), tiles_y = ceil_div(h, C::TH), ntiles = tiles_x * tiles_y * n;
    const int grid = ntiles < num_sms() ? ntiles : num_sms();
    return launch_pdl(bottleneck_s1t_kernel<CIN, COUT, RES>, grid, kTThreads, C::smem_bytes, s, xmap, in, tab, we_img, wp_img, out, h, w, tiles_x,
                      tiles_y, ntiles);
}

cudaError_t launch_bottleneck_s1t_tc(int cin, int cout, const bf16* in, const unsigned char* tab, const bf16* we_img,
                                     const bf16* wp_img, bf16* out, int n, int h, int w, cudaStream_t s) {
    if (cin == 64 && cout == 64) return run_s1t<64, 64, true>(in, tab, we_img, wp_img, out, n, h, w, s);
    if (cin == 96 && cout == 96) return run_s1t<96, 96, true>(in, tab, we_img, wp_img, out, n, h, w, s);
    if (cin == 96 && cout == 128) return run_s1t<96, 128, false>(in, tab, we_img, wp_img, out, n, h, w, s);
    if (cin == 128 && cout == 128) return run_s1t<128, 128, true>(in, tab, we_img, wp_img, out, n, h, w, s);
    return cudaErrorInvalidValue;
}

// image sizes (bytes) of the transposed kernel's operands for a stride-1 layer
size_t bottleneck_s1t_we_bytes(int cin) { const int nch = (6 * cin + 127) / 128; return (size_t)nch * 128 * (cin + 16) * 2; }
size_t bottleneck_s1t_wp_bytes(int cin, int cout) { const int nch = (6 * cin + 127) / 128; return (size_t)nch * cout * 128 * 2; }
size_t bottleneck_s1t_tab_bytes(int cin, int cout) { const int nch = (6 * cin + 127) / 128; return (size_t)nch * 128 * 32 + (size_t)cout * 4; }

}  // namespace fscnn
