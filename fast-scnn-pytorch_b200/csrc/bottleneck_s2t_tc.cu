// bottleneck_s2t_tc.cu -- bf16 LinearBottleneck, stride 2 (reference models/fast_scnn.py:95-115; bottleneck1.0 and
// bottleneck2.0 of GlobalFeatureExtractor), with the transposed expand of bottleneck_s1t_tc.cu: expanded channels sit on
// the TMEM lanes, every depthwise thread owns one channel and convolves it in registers straight out of TMEM.
//
// An 8x16-pixel output tile needs a 17x33 input halo (561 pixels), more than an accumulator buffer can hold, so the tile is
// cut into 2x2 sub-tiles of 4x8 output pixels, each with its own 9x17 halo (153 pixels, one TMA tensor copy, N = 160):
//
//   for every chunk of 128 expanded channels:            (weights: We chunk = A operand, 16 extra K columns carry the bias)
//     for sub-tile 0..3:   ("unit" U = chunk * 4 + sub)
//       expand   E^T[128 ch x 160 px] = We^T * X_sub^T + bias * 1^T      tcgen05.mma -> TMEM buffer U & 1
//       depthwise warp (quarter q = warp % 4, output row s = warp / 4 of the sub-tile): 3 halo rows x 17 columns of its channel
//                tcgen05.ld -> ReLU + bf16 pack -> 72 FHFMA.BF16 (stride 2) -> 8 outputs -> ONE 16-byte store into the MN-major
//                D operand (pixel block = sub * 4 + s)
//     project  OUT[128 px x COUT] += D[128 px x 128 ch] * Wp chunk       after the fourth sub-tile
//   + bias -> bf16 NHWC (stride-2 layers have no residual)
//
// The four sub-tile halos of a tile stay resident for all chunks; the halo of the NEXT tile's sub-tile j is requested as soon
// as the last chunk's expand of sub-tile j has completed.  Controllers as in the stride-1 kernel: one warp issues halo
// loads, expand weight copies and expand MMAs, another the project weight copies and project MMAs, in order, with blocking
// mbarrier waits.
#include "kernels.h"
#include "tma_host.h"
#include "umma.cuh"

namespace fscnn {

namespace {
constexpr int kS2Warps = 16;
constexpr int kS2Threads = (kS2Warps + 2) * 32;
}  // namespace

template <int CIN, int COUT>
struct T2Cfg {
    static constexpr int TH = 8, TW = 16, SH = 4, SW = 8, NSUB = 4;     // output tile, sub-tile, sub-tiles per tile
    static constexpr int IH = 2 * SH + 1, IW = 2 * SW + 1, PIN = IH * IW, NB = 160;   // 9 x 17 halo, MMA N
    static constexpr int CM = 128, CEXP = 6 * CIN, NCH = (CEXP + CM - 1) / CM;
    static constexpr int KA = CIN + 16;
    static constexpr int X_BYTES = PIN * CIN * 2;              // [CIN/8][PIN][8]: LBO = PIN*16, SBO = 128
    static constexpr int XS = round_up(X_BYTES, 128);
    static constexpr int WE_BYTES = CM * KA * 2;
    static constexpr int WP_BYTES = COUT * CM * 2;
    static constexpr int D_LBO = 128, D_SBO = 2048;            // MN-major D: see bottleneck_s1t_tc.cu
    static constexpr int D_BYTES = 16 * D_SBO;
    static constexpr int ONES_BYTES = 256;
    static constexpr int LIMIT = 227 * 1024 - 256;
    static constexpr int FIXED = NSUB * XS + 2 * WE_BYTES + 2 * WP_BYTES + ONES_BYTES + D_BYTES;
    static constexpr int DB = (FIXED + D_BYTES <= LIMIT) ? 2 : 1;
    static constexpr int oX = 0;
    static constexpr int oWe = NSUB * XS;
    static constexpr int oWp = oWe + 2 * WE_BYTES;
    static constexpr int oD = oWp + 2 * WP_BYTES;
    static constexpr int oOnes = oD + DB * D_BYTES;
    static constexpr int smem_bytes = oOnes + ONES_BYTES;
    static constexpr int TM_PROJ = 2 * NB;
    static constexpr int TM_COLS = 512;
    static_assert(TM_PROJ + COUT <= 512 && smem_bytes <= LIMIT, "budget");
    static_assert(CIN % 16 == 0 && COUT % 32 == 0 && CEXP % CM == 0, "shape");
    // the B operand reads NB - PIN rows past each channel group of a halo tile: they must stay inside the allocation
    static_assert((NSUB - 1) * XS + (CIN / 8 - 1) * PIN * 16 + NB * 16 <= smem_bytes, "B-tile overrun");
};

#ifdef FSCNN_PHASE_TIMING   // debug build only: clock64 stamps of units 24..35 (the third tile of a 64-channel layer) of CTA 5
__device__ long long g_s2t_phase[12 * 16];
#define U_STAMP(cond, uu, slot) do { if (COUT == 64 && blockIdx.x == 5 && (cond) && (uu) >= 24 && (uu) < 36) g_s2t_phase[((uu) - 24) * 16 + (slot)] = clock64(); } while (0)
extern "C" int fscnn_debug_s2t_phases(long long* out192) {
    return cudaMemcpyFromSymbol(out192, g_s2t_phase, sizeof(long long) * 192) == cudaSuccess ? 0 : -1;
}
#else
#define U_STAMP(cond, uu, slot) do { } while (0)
#endif

template <int CIN, int COUT>
__global__ void __launch_bounds__(kS2Threads, 1)
bottleneck_s2t_kernel(const __grid_constant__ CUtensorMap xmap, const unsigned char* __restrict__ tab, const bf16* __restrict__ we_img,
                      const bf16* __restrict__ wp_img, bf16* __restrict__ out, int Hi, int Wi, int Ho, int Wo, int tiles_x,
                      int tiles_y, int ntiles) {
    using C = T2Cfg<CIN, COUT>;
    constexpr int IW = C::IW, NCH = C::NCH, PIN = C::PIN, DB = C::DB, CM = C::CM, NSUB = C::NSUB;
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar_we[2], bar_wp[2], bar_exp[2], bar_tmfree[2], bar_x[4], bar_dready[2], bar_proj[2],
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
            mbar_init(&bar_we[i], 1); mbar_init(&bar_wp[i], 1); mbar_init(&bar_exp[i], 1);
            mbar_init(&bar_tmfree[i], kS2Warps); mbar_init(&bar_dready[i], kS2Warps); mbar_init(&bar_proj[i], 1);
        }
        for (int i = 0; i < NSUB; ++i) mbar_init(&bar_x[i], 1);
        mbar_init(&bar_projfree, kS2Warps); mbar_init(&bar_tiledone, 1);
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

    if (warp == kS2Warps) {
        // =========================== expand controller ===========================
        // The whole warp walks the loop converged and every asynchronous operation is issued by the lane elect.sync names: behind a
        // `lane == 0` branch each tcgen05.mma / bulk copy is wrapped in a vote loop over the active lanes (6 extra instructions each),
        // and this single thread's instruction latency is what paces the stride-2 kernel (phase stamps: ~100 instructions = ~1 200
        // cycles per unit against 424 cycles of MMA work).
        {
            auto prefetch_we = [&](int g) {
                if (elect_one()) {
                    mbar_arrive_expect_tx(&bar_we[g & 1], C::WE_BYTES);
                    bulk_g2s(sm + C::oWe + (g & 1) * C::WE_BYTES, we_img + (size_t)(g % NCH) * CM * C::KA, C::WE_BYTES, &bar_we[g & 1]);
                }
            };
            auto load_x = [&](int lt, int sub) {      // input halo of sub-tile (sub >> 1, sub & 1): rows 2*oy - 1 .., columns 2*ox - 1 ..
                int n, oy0, ox0;
                tile_origin(lt, n, oy0, ox0);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&bar_x[sub], C::X_BYTES);
                    tma_load_halo(sX + sub * C::XS, &xmap, 2 * (ox0 + C::SW * (sub & 1)) - 1, 2 * (oy0 + C::SH * (sub >> 1)) - 1, n, &bar_x[sub]);
                }
            };
            constexpr uint32_t idesc_exp = make_idesc_bf16(128, C::NB);
            if (elect_one()) tma_prefetch_desc(&xmap);
            prefetch_we(0);
            if (total > 1) prefetch_we(1);
            pdl_wait();      // the weights are on their way; the halo tiles are the previous stage's output
            for (int j = 0; j < NSUB; ++j) load_x(0, j);
            // Reloads that must follow the completion of expand(u) (the weight buffer two chunks on, the sub-tile's halo of the next
            // tile) are issued one unit LATE: after expand(u+1) has been issued, when expand(u) has long completed.  Waiting for it right
            // after its own commit (the first version) parked this warp for a whole MMA sequence on every unit of a tile's last chunk
            // while the tensor core had nothing queued (phase stamps, tools/phase_timing.py: 770 of 1 700 cycles per unit there).
            int pend_u = -1, pend_g = 0, pend_lt = 0, pend_sub = 0;
            bool pend_w = false, pend_x = false;
            auto flush_pending = [&]() {
                if (pend_u < 0) return;
                mbar_wait(&bar_exp[pend_u & 1], (pend_u >> 1) & 1);      // expand(pend_u) has completed: what it read may be overwritten
                U_STAMP(lane == 0, pend_u, 11);
                if (pend_w) prefetch_we(pend_g + 2);
                if (pend_x) load_x(pend_lt + 1, pend_sub);
                pend_u = -1;
            };
#pragma unroll 1
            for (int g = 0; g < total; ++g) {
                const int lt = g / NCH, e = g - lt * NCH;
                mbar_wait(&bar_we[g & 1], (g >> 1) & 1);                                  // the chunk's expand weights
#pragma unroll 1
                for (int sub = 0; sub < NSUB; ++sub) {
                    const int u = g * NSUB + sub;                                          // unit: accumulator buffer u & 1
                    if (e == 0) mbar_wait(&bar_x[sub], lt & 1);                            // the sub-tile's halo
                    U_STAMP(lane == 0, u, 8);
                    if (u >= 2) mbar_wait(&bar_tmfree[u & 1], ((u - 2) >> 1) & 1);         // accumulator drained by unit u-2
                    U_STAMP(lane == 0, u, 9);
                    tc_fence_after_sync();
                    const uint64_t da0 = make_smem_desc(sWe + (g & 1) * C::WE_BYTES, 2048, 128);
                    const uint64_t db0 = make_smem_desc(sX + sub * C::XS, PIN * 16, 128);
                    const uint32_t dacc = tmem + (u & 1) * C::NB;
                    if (elect_one()) {
#pragma unroll
                        for (int k16 = 0; k16 < CIN / 16; ++k16)
                            umma_bf16_ss(dacc, da0 + (uint64_t)(k16 * ((2 * 2048) >> 4)), db0 + (uint64_t)(k16 * ((2 * PIN * 16) >> 4)), idesc_exp, k16 > 0);
                        umma_bf16_ss(dacc, da0 + (uint64_t)((CIN / 16) * ((2 * 2048) >> 4)), make_smem_desc(sOnes, 128, 0), idesc_exp, 1);
                        umma_commit(&bar_exp[u & 1]);
                    }
                    __syncwarp();
                    U_STAMP(lane == 0, u, 10);
                    flush_pending();                                  // unit u-1's reloads (bar_exp[(u-1) & 1]: no phase can alias, unit u+1 is not issued yet)
                    const bool more_w = (sub == NSUB - 1) && (g + 2 < total), more_x = (e == NCH - 1) && (lt + 1 < my_tiles);
                    if (more_w || more_x) { pend_u = u; pend_g = g; pend_lt = lt; pend_sub = sub; pend_w = more_w; pend_x = more_x; }
                }
            }
            flush_pending();
        }
    } else if (warp == kS2Warps + 1) {
        // =========================== project controller ===========================
        if (lane == 0) {
            auto prefetch_wp = [&](int g) {
                mbar_arrive_expect_tx(&bar_wp[g & 1], C::WP_BYTES);
                bulk_g2s(sm + C::oWp + (g & 1) * C::WP_BYTES, wp_img + (size_t)(g % NCH) * COUT * CM, C::WP_BYTES, &bar_wp[g & 1]);
            };
            constexpr uint32_t idesc_proj = make_idesc_bf16(128, COUT) | (1u << 15);   // A (= D) is MN-major
            prefetch_wp(0);
            if (total > 1) prefetch_wp(1);
#pragma unroll 1
            for (int kp = 0; kp < total; ++kp) {
                const int lt = kp / NCH, e = kp - lt * NCH;
                mbar_wait(&bar_wp[kp & 1], (kp >> 1) & 1);                                // its weight chunk
                if (e == 0 && lt > 0) mbar_wait(&bar_projfree, (lt - 1) & 1);             // the previous tile's accumulator has been read
                mbar_wait(&bar_dready[kp % DB], (kp / DB) & 1);                           // D written by the depthwise threads (4 sub-tiles)
                U_STAMP(true, kp * NSUB + 3, 12);
                tc_fence_after_sync();
                const uint64_t da0 = make_smem_desc(sD + (kp % DB) * C::D_BYTES, C::D_LBO, C::D_SBO);
                const uint64_t db0 = make_smem_desc(sWp + (kp & 1) * C::WP_BYTES, COUT * 16, 128);
#pragma unroll
                for (int k16 = 0; k16 < CM / 16; ++k16)
                    umma_bf16_ss(tmem + C::TM_PROJ, da0 + (uint64_t)(k16 * ((2 * C::D_LBO) >> 4)),
                                 db0 + (uint64_t)(k16 * ((2 * COUT * 16) >> 4)), idesc_proj, (e | k16) != 0);
                umma_commit(&bar_proj[kp % DB]);
                if (e == NCH - 1) umma_commit(&bar_tiledone);   // one phase per TILE for the output epilogue
                if (kp + 2 < total) {
                    mbar_wait(&bar_proj[kp % DB], (kp / DB) & 1);     // project(kp) has completed: its weight buffer is free
                    prefetch_wp(kp + 2);
                }
            }
        }
    } else {
        // =========================== compute warps ===========================
        pdl_wait();
        const int q = warp & 3, s = warp >> 2;            // TMEM lane quarter (32 expanded channels), output row of the sub-tile
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const float* Bp_g = reinterpret_cast<const float*>(tab + (size_t)NCH * CM * 32);
        int n = 0, oy0 = 0, ox0 = 0, pn = 0, poy0 = 0, pox0 = 0;
        auto output_epilogue = [&](int lt, int tn, int toy0, int tox0) {   // + bias, bf16 NHWC store
            const int p = q * 32 + lane;                  // accumulator row = pixel block (sub * 4 + row) * 8 + column
            const int sub = p >> 5, oy = toy0 + C::SH * (sub >> 1) + ((p >> 3) & 3), ox = tox0 + C::SW * (sub & 1) + (p & 7);
            const bool live = (oy < Ho) && (ox < Wo);
            const size_t pix = ((size_t)tn * Ho + oy) * Wo + ox;
            constexpr int CP = COUT / 4;                  // columns per warp: 16 or 24
            const int co0 = s * CP;
            mbar_wait(&bar_tiledone, lt & 1);             // every project MMA of tile lt has completed
            tc_fence_after_sync();
            uint32_t r[CP];
            tmem_ld_32x32b_x16(tmem + lane_base + C::TM_PROJ + co0, reinterpret_cast<uint32_t(&)[16]>(r[0]));
            if (CP == 24) tmem_ld_32x32b_x8(tmem + lane_base + C::TM_PROJ + co0 + 16, r + (CP == 24 ? 16 : 0));
            tmem_ld_wait();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_projfree);    // accumulator read: the next tile's first project MMA may overwrite it
            if (live) {
#pragma unroll
                for (int i = 0; i < CP / 8; ++i) {
                    const int co = co0 + 8 * i;
                    const float4 ba = __ldg(reinterpret_cast<const float4*>(Bp_g + co));
                    const float4 bb = __ldg(reinterpret_cast<const float4*>(Bp_g + co + 4));
                    const uint32_t* q8 = r + 8 * i;
                    *reinterpret_cast<uint4*>(out + pix * COUT + co) =
                        make_uint4(packbf(__uint_as_float(q8[0]) + ba.x, __uint_as_float(q8[1]) + ba.y),
                                   packbf(__uint_as_float(q8[2]) + ba.z, __uint_as_float(q8[3]) + ba.w),
                                   packbf(__uint_as_float(q8[4]) + bb.x, __uint_as_float(q8[5]) + bb.y),
                                   packbf(__uint_as_float(q8[6]) + bb.z, __uint_as_float(q8[7]) + bb.w));
                }
            }
        };
#pragma unroll 1
        for (int g = 0; g < total; ++g) {
            const int lt = g / NCH, e = g - lt * NCH;
            if (e == 0) { pn = n; poy0 = oy0; pox0 = ox0; tile_origin(lt, n, oy0, ox0); }
            // this thread's channel: 9 bf16 taps + fp32 bias from the (L1-resident) table
            const uint4* rec = reinterpret_cast<const uint4*>(tab + (size_t)(e * CM + q * 32 + lane) * 32);
            const uint4 wa = __ldg(rec), wb = __ldg(rec + 1);
            const uint32_t wq[5] = {wa.x, wa.y, wa.z, wa.w, wb.x};
            const float bd = __uint_as_float(wb.y);
            const int k = q * 32 + lane;
            const uint32_t d0 = sD + (g % DB) * C::D_BYTES + (k >> 3) * C::D_LBO + (k & 7) * 16 + s * C::D_SBO;
#pragma unroll 1
            for (int sub = 0; sub < NSUB; ++sub) {
                const int u = g * NSUB + sub;
                U_STAMP(tid == 0, u, 0);
                U_STAMP(tid == 15 * 32, u, 6);
                mbar_wait(&bar_exp[u & 1], (u >> 1) & 1);        // expand(u) has completed
                U_STAMP(tid == 0, u, 1);
                tc_fence_after_sync();
                uint32_t Ep[3][9];                               // halo rows 2s .. 2s+2, column pairs (2i, 2i+1), ReLU'd bf16
                {
                    uint32_t r[52];
                    const uint32_t t0 = tmem + lane_base + (u & 1) * C::NB + (2 * s) * IW;
                    tmem_ld_32x32b_x32(t0, reinterpret_cast<uint32_t(&)[32]>(r[0]));
                    tmem_ld_32x32b_x16(t0 + 32, reinterpret_cast<uint32_t(&)[16]>(r[32]));
                    tmem_ld_32x32b_x4(t0 + 48, r + 48);
                    tmem_ld_wait();
#pragma unroll
                    for (int rr = 0; rr < 3; ++rr)
#pragma unroll
                        for (int i = 0; i < 9; ++i)
                            Ep[rr][i] = packbf_relu(__uint_as_float(r[rr * IW + 2 * i]), __uint_as_float(r[rr * IW + (i < 8 ? 2 * i + 1 : 2 * i)]));
                }
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_tmfree[u & 1]);  // expand(u+2) may overwrite this accumulator
                U_STAMP(tid == 0, u, 2);
                U_STAMP(tid == 15 * 32, u, 7);
                // zero padding of the depthwise conv: halo columns / rows outside the image (border tiles only)
                const int ix0 = 2 * (ox0 + C::SW * (sub & 1)) - 1, iy0 = 2 * (oy0 + C::SH * (sub >> 1)) - 1 + 2 * s;
                if (ix0 == -1 && ix0 + IW <= Wi) {               // left image border: only halo column 0 is outside
#pragma unroll
                    for (int rr = 0; rr < 3; ++rr) Ep[rr][0] &= 0xFFFF0000u;
                } else if (ix0 >= 0 && ix0 + IW == Wi + 1 && (IW & 1) == 0) {   // right border, width a multiple of the tile: only the last column
#pragma unroll
                    for (int rr = 0; rr < 3; ++rr) Ep[rr][IW / 2 - 1] &= 0x0000FFFFu;
                } else if (ix0 < 0 || ix0 + IW > Wi) {           // anything else (odd sizes): per-column masks
#pragma unroll
                    for (int i = 0; i < 9; ++i) {
                        const int xa = ix0 + 2 * i, xb2 = xa + 1;
                        const uint32_t m = ((xa >= 0 && xa < Wi) ? 0x0000FFFFu : 0u) | ((xb2 >= 0 && xb2 < Wi) ? 0xFFFF0000u : 0u);
#pragma unroll
                        for (int rr = 0; rr < 3; ++rr) Ep[rr][i] &= m;
                    }
                }
                if (iy0 < 0 || iy0 + 3 > Hi) {
#pragma unroll
                    for (int rr = 0; rr < 3; ++rr)
                        if (iy0 + rr < 0 || iy0 + rr >= Hi) {
#pragma unroll
                            for (int i = 0; i < 9; ++i) Ep[rr][i] = 0u;
                        }
                }
                float acc[8];
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                        for (int x = 0; x < 8; ++x)
                            acc[x] = fhfma_sel((ky | kx) ? acc[x] : bd, Ep[ky][(2 * x + kx) >> 1], (2 * x + kx) & 1, wq[(ky * 3 + kx) >> 1],
                                               (ky * 3 + kx) & 1);
                U_STAMP(tid == 0, u, 3);
                if (sub == 0 && g >= DB) mbar_wait(&bar_proj[g % DB], (g / DB - 1) & 1);   // project(g-DB) has completed: this D buffer is free
                U_STAMP(tid == 0, u, 4);
                sts128(d0 + sub * (4 * C::D_SBO), packbf_relu(acc[0], acc[1]), packbf_relu(acc[2], acc[3]), packbf_relu(acc[4], acc[5]),
                       packbf_relu(acc[6], acc[7]));
            }
            fence_async_proxy();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_dready[g % DB]);
            U_STAMP(tid == 0, g * NSUB + 3, 5);
            if (e == 0 && lt >= 1) output_epilogue(lt - 1, pn, poy0, pox0);   // deferred by one chunk: keeps the pipeline fed
            U_STAMP(tid == 0, g * NSUB + 3, 13);
        }
        output_epilogue(my_tiles - 1, n, oy0, ox0);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, C::TM_COLS);
}

template <int CIN, int COUT>
static cudaError_t run_s2t(const bf16* in, const unsigned char* tab, const bf16* we_img, const bf16* wp_img, bf16* out, int n, int hi,
                           int wi, int ho, int wo, cudaStream_t s) {
    using C = T2Cfg<CIN, COUT>;
    static unsigned long long configured = 0;
    cudaError_t e = ensure_dyn_smem(bottleneck_s2t_kernel<CIN, COUT>, C::smem_bytes, configured);
    if (e != cudaSuccess) return e;
    CUtensorMap xmap;
    e = make_nhwc_halo_map(&xmap, in, n, hi, wi, CIN, C::IH, C::IW);
    if (e != cudaSuccess) return e;
    const int tiles_x = ceil_div(wo, C::TW), tiles_y = ceil_div(ho, C::TH), ntiles = tiles_x * tiles_y * n;
    const int grid = ntiles < num_sms() ? ntiles : num_sms();
    return launch_pdl(bottleneck_s2t_kernel<CIN, COUT>, grid, kS2Threads, C::smem_bytes, s, xmap, tab, we_img, wp_img, out, hi, wi, ho, wo, tiles_x,
                      tiles_y, ntiles);
}

cudaError_t launch_bottleneck_s2t_tc(int cin, int cout, const bf16* in, const unsigned char* tab, const bf16* we_img,
                                     const bf16* wp_img, bf16* out, int n, int hi, int wi, int ho, int wo, cudaStream_t s) {
    if (cin == 64 && cout == 64) return run_s2t<64, 64>(in, tab, we_img, wp_img, out, n, hi, wi, ho, wo, s);
    if (cin == 64 && cout == 96) return run_s2t<64, 96>(in, tab, we_img, wp_img, out, n, hi, wi, ho, wo, s);
    return cudaErrorInvalidValue;
}

}  // namespace fscnn
