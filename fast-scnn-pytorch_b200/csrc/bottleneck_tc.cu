// bottleneck_tc.cu -- bf16 LinearBottleneck on the 5th-gen tensor cores (tcgen05 + TMEM + TMA).
// Same fusion as bottleneck.cu (reference models/fast_scnn.py:95-115): expand 1x1 + ReLU -> DW 3x3
// (stride s) + ReLU -> project 1x1 (+ residual), the 6c-wide tensor never leaves the SM; both 1x1
// contractions run as tcgen05.mma (bf16 operands from shared memory, fp32 accumulators in TMEM), the
// depthwise stage runs on the CUDA cores between them (FHFMA.BF16: bf16 operands, fp32 accumulate).
//
// Persistent, warp-specialised: one CTA per SM walks 8x16-pixel output tiles.
//   control warp (lane 0)  : TMA halo-tile loads (one cp.async.bulk.tensor per tile, zero-filled padding, already in
//                            the MMA A-operand layout), bulk copies of the weight chunks, every tcgen05.mma + commit.
//                            It runs ahead of the compute warps: the next tile's halo load starts as soon as the last
//                            expand MMA of the current tile has completed, its first expand MMA overlaps the current
//                            tile's output epilogue.
//   16 compute warps       : per chunk of CE = 64 expanded channels
//       expand epilogue    : TMEM -> +bias, ReLU, zero outside the image -> bf16 -> E[halo px][CE] (swizzled rows)
//       depthwise          : E -> 3x3 + bias + ReLU -> bf16 -> D[128][CE] (A-operand layout)
//     after the last chunk : TMEM (project accumulators) -> +bias (+ residual) -> bf16 NHWC store
//   hand-offs are mbarriers (TMEM drained / D written -> control warp; MMA committed -> compute warps); the compute warps
//   synchronise among themselves with a named barrier, the control warp never joins it.
//
// Measured (round 1): a tcgen05.mma issued by a thread that also computes blocks that warp for
// 100-200 cycles per MMA behind the previous one; the depthwise phase is shared-memory-bandwidth bound (LDS.128 wavefronts
// + the MMA operand reads share the 128 B/clk port).  Rejected alternative: the depthwise 3x3 on the tensor core as
// 9 * CE/16 block-diagonal N = 16 MMAs is exact but every SS-mode MMA costs >= 64 cycles whatever N is.
#include "kernels.h"
#include "tma_host.h"
#include "umma.cuh"

namespace fscnn {

constexpr int kNT = 512;           // compute threads per CTA
constexpr int kNTall = kNT + 32;   // + the control warp

#ifdef FSCNN_PHASE_TIMING   // debug build only: clock64 stamps of the second tile of CTA 5 of the <64,64,1> kernel
__device__ long long g_bneck_phase[64];
#define BN_STAMP(i) do { if (CIN == 64 && COUT == 64 && tid == 0 && blockIdx.x == 5 && lt == 1) g_bneck_phase[i] = clock64(); } while (0)
#else
#define BN_STAMP(i) do { } while (0)
#endif

// 16-byte chunk permutation of an E row (128 B per halo pixel).  The expand epilogue writes 8 consecutive pixels per
// quarter-warp, the stride-2 depthwise reads every other pixel: XOR-ing with (pin ^ pin >> 3) keeps both conflict-free
// (pin & 7 alone makes the stride-2 reads 2-way conflicted: only 4 of the 8 chunk slots get used).
__device__ __forceinline__ int e_swz(int pin) { return (pin ^ (pin >> 3)) & 7; }

template <int CIN, int COUT, int STRIDE>
struct TcCfg {
    static constexpr int TH = 8, TW = 16, P = 128;
    static constexpr int IH = (TH - 1) * STRIDE + 3, IW = (TW - 1) * STRIDE + 3;
    static constexpr int PIN = IH * IW;
    static constexpr int NMT = (PIN + 127) / 128;          // expand M-tiles
    static constexpr int CE = 64, CEXP = 6 * CIN, NCH = CEXP / CE;
    static constexpr int X_BYTES = PIN * CIN * 2;          // [CIN/8][PIN][8]: LBO = PIN*16, SBO = 128
    static constexpr int oX = 0;
    static constexpr int oE = round_up(X_BYTES, 128);
    static constexpr int oD = oE + round_up(PIN, 8) * CE * 2;
    static constexpr int WE_BYTES = CE * CIN * 2, WP_BYTES = COUT * CE * 2;
    static constexpr int oWe = oD + P * CE * 2;            // 2 buffers
    static constexpr int oWp = oWe + 2 * WE_BYTES;         // 2 buffers
    static constexpr int oTab = oWp + 2 * WP_BYTES;        // bf16 Wd[9][CEXP] | f32 Be[CEXP] | f32 Bd[CEXP] | f32 Bp[COUT]
    static constexpr int TAB_BYTES = 9 * CEXP * 2 + 2 * CEXP * 4 + COUT * 4;
    static constexpr int smem_bytes = oTab + TAB_BYTES;
    static constexpr int TM_EXP = 0, TM_PROJ = NMT * CE;
    static constexpr int TM_COLS = (NMT * CE + COUT) <= 256 ? 256 : 512;
    static_assert(NMT * CE + COUT <= 512, "TMEM budget");
    static_assert(CEXP % CE == 0 && CIN % 16 == 0 && COUT % 16 == 0, "shape");
    // the last expand M-tile reads (NMT*128 - PIN) rows past the halo tile: they must stay inside the allocation
    static_assert((CIN / 8 - 1) * PIN * 16 + NMT * 128 * 16 <= smem_bytes, "A-tile overrun");
};

template <int CIN, int COUT, int STRIDE, bool RES>
__global__ void __launch_bounds__(kNTall, 1)
bottleneck_tc_kernel(const __grid_constant__ CUtensorMap xmap, const bf16* __restrict__ in, const unsigned char* __restrict__ tab_img,
                     const bf16* __restrict__ we_img, const bf16* __restrict__ wp_img, bf16* __restrict__ out, int Hi, int Wi,
                     int Ho, int Wo, int tiles_x, int tiles_y, int ntiles) {
    using C = TcCfg<CIN, COUT, STRIDE>;
    constexpr int CE = C::CE, IW = C::IW, NMT = C::NMT, NCH = C::NCH, PIN = C::PIN;
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar_we[2], bar_wp[2], bar_exp, bar_proj, bar_tm_free, bar_d_ready, bar_x, bar_tab;
    __shared__ uint32_t tmem_base_s;
    const float* Be_all = reinterpret_cast<const float*>(sm + C::oTab + 9 * C::CEXP * 2);
    const float* Bd_all = Be_all + C::CEXP;
    const float* Bp_s = Bd_all + C::CEXP;
    const uint32_t sX = smem_u32(sm + C::oX), sE = smem_u32(sm + C::oE), sD = smem_u32(sm + C::oD);
    const uint32_t sWe = smem_u32(sm + C::oWe), sWp = smem_u32(sm + C::oWp), sWd = smem_u32(sm + C::oTab);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total_chunks = my_tiles * NCH;
    auto tile_origin = [&](int tile, int& n, int& oy0, int& ox0) {
        const int tx = tile % tiles_x, r = tile / tiles_x;
        n = r / tiles_y; oy0 = (r % tiles_y) * C::TH; ox0 = tx * C::TW;
    };

    if (tid == 0) {
        mbar_init(&bar_we[0], 1); mbar_init(&bar_we[1], 1); mbar_init(&bar_wp[0], 1); mbar_init(&bar_wp[1], 1);
        mbar_init(&bar_exp, 1); mbar_init(&bar_proj, 1); mbar_init(&bar_x, 1); mbar_init(&bar_tab, 1);
        mbar_init(&bar_tm_free, kNT / 32); mbar_init(&bar_d_ready, kNT / 32);
        fence_mbar_init();
    }
    if (warp == 0) { tmem_alloc(&tmem_base_s, C::TM_COLS); tmem_relinquish(); }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_base_s;

    if (warp == kNT / 32) {
        // =========================== control warp ===========================
        if (lane == 0) {
            // weight chunk g (global chunk counter over all tiles of this CTA) lives in buffer g & 1
            auto prefetch_we = [&](int g) {
                mbar_arrive_expect_tx(&bar_we[g & 1], C::WE_BYTES);
                bulk_g2s(sm + C::oWe + (g & 1) * C::WE_BYTES, we_img + (size_t)(g % NCH) * CE * CIN, C::WE_BYTES, &bar_we[g & 1]);
            };
            auto prefetch_wp = [&](int g) {
                mbar_arrive_expect_tx(&bar_wp[g & 1], C::WP_BYTES);
                bulk_g2s(sm + C::oWp + (g & 1) * C::WP_BYTES, wp_img + (size_t)(g % NCH) * COUT * CE, C::WP_BYTES, &bar_wp[g & 1]);
            };
            auto load_x = [&](int tile) {
                int n, oy0, ox0;
                tile_origin(tile, n, oy0, ox0);
                mbar_arrive_expect_tx(&bar_x, C::X_BYTES);
                tma_load_halo(sX, &xmap, ox0 * STRIDE - 1, oy0 * STRIDE - 1, n, &bar_x);
            };
            constexpr uint32_t idesc_exp = make_idesc_bf16(128, CE);
            constexpr uint32_t idesc_proj = make_idesc_bf16(128, COUT);
            auto issue_expand = [&](int g) {
                mbar_wait(&bar_we[g & 1], (g >> 1) & 1);
                tc_fence_after_sync();
#pragma unroll
                for (int mt = 0; mt < NMT; ++mt)
#pragma unroll
                    for (int k16 = 0; k16 < CIN / 16; ++k16) {
                        const uint64_t da = make_smem_desc(sX + mt * 2048 + k16 * 2 * (PIN * 16), PIN * 16, 128);
                        const uint64_t db = make_smem_desc(sWe + (g & 1) * C::WE_BYTES + k16 * 2 * (CE * 16), CE * 16, 128);
                        umma_bf16_ss(tmem + C::TM_EXP + mt * CE, da, db, idesc_exp, k16 > 0);
                    }
                umma_commit(&bar_exp);
            };
            tma_prefetch_desc(&xmap);
            mbar_arrive_expect_tx(&bar_tab, C::TAB_BYTES);
            bulk_g2s(sm + C::oTab, tab_img, C::TAB_BYTES, &bar_tab);
            load_x(blockIdx.x);
            prefetch_we(0);
            if (total_chunks > 1) prefetch_we(1);
            prefetch_wp(0);
            int g = 0, lt = 0;
#pragma unroll 1
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++lt) {
                mbar_wait(&bar_x, lt & 1);                           // halo tile landed (TMEM exp is free: see last chunk)
                issue_expand(g);
#pragma unroll 1
                for (int e = 0; e < NCH; ++e, ++g) {
                    mbar_wait(&bar_tm_free, g & 1);                  // expand epilogue g has drained TMEM: expand(g) is complete
                    tc_fence_after_sync();
                    if (e + 1 < NCH) issue_expand(g + 1);
                    else if (tile + (int)gridDim.x < ntiles) load_x(tile + gridDim.x);   // X is dead: every expand MMA has read it
                    if (g + 2 < total_chunks) prefetch_we(g + 2);    // buffer g&1 is free
                    if (g > 0) mbar_wait(&bar_proj, (g - 1) & 1);    // project(g-1) done: Wp buffer (g+1)&1 is free
                    if (g + 1 < total_chunks) prefetch_wp(g + 1);
                    mbar_wait(&bar_d_ready, g & 1);                  // depthwise g has written D (writers fenced the async proxy)
                    mbar_wait(&bar_wp[g & 1], (g >> 1) & 1);
                    tc_fence_after_sync();
#pragma unroll
                    for (int k16 = 0; k16 < CE / 16; ++k16) {
                        const uint64_t da = make_smem_desc(sD + k16 * 2 * 2048, 2048, 128);
                        const uint64_t db = make_smem_desc(sWp + (g & 1) * C::WP_BYTES + k16 * 2 * (COUT * 16), COUT * 16, 128);
                        umma_bf16_ss(tmem + C::TM_PROJ, da, db, idesc_proj, (e | k16) != 0);
                    }
                    umma_commit(&bar_proj);
                }
            }
        }
    } else {
        // =========================== compute warps ===========================
        mbar_wait(&bar_tab, 0);
        int g = 0, lt = 0;
#pragma unroll 1
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++lt) {
            int n, oy0, ox0;
            tile_origin(tile, n, oy0, ox0);
            const int iy0 = oy0 * STRIDE - 1, ix0 = ox0 * STRIDE - 1;
            BN_STAMP(0);
#pragma unroll 1
            for (int e = 0; e < NCH; ++e, ++g) {
                const float* Bes = Be_all + e * CE;
                const float* Bds = Bd_all + e * CE;
                const uint32_t sWds = sWd + e * CE * 2;      // tap t, channel c at sWds + (t * CEXP + c) * 2
                // ---- expand epilogue: TMEM -> bias, ReLU, image mask -> bf16 rows of E ----
                mbar_wait(&bar_exp, g & 1);
                tc_fence_after_sync();
                BN_STAMP(1 + 3 * e);
                for (int task = warp; task < NMT * 8; task += kNT / 32) {   // task = (M-tile, 32-column half, lane quarter)
                    const int q = task & 3, ch = (task >> 2) & 1, mt = task >> 3;
                    const int pin = mt * 128 + q * 32 + lane;
                    const int iy = iy0 + pin / IW, ix = ix0 + pin % IW;
                    const bool ok = pin < PIN && iy >= 0 && iy < Hi && ix >= 0 && ix < Wi;
                    const int c0 = ch * 32;
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(tmem + ((uint32_t)(q * 32) << 16) + C::TM_EXP + mt * CE + c0, r);
                    tmem_ld_wait();
                    if (pin < PIN) {
                        if (ok) {
#pragma unroll
                            for (int h = 0; h < 4; ++h) {
                                const float4 ba = *reinterpret_cast<const float4*>(Bes + c0 + h * 8);
                                const float4 bb = *reinterpret_cast<const float4*>(Bes + c0 + h * 8 + 4);
                                const uint32_t* q8 = r + h * 8;
                                sts128(sE + pin * (CE * 2) + ((((c0 >> 3) + h) ^ e_swz(pin)) << 4),
                                       packbf_relu(__uint_as_float(q8[0]) + ba.x, __uint_as_float(q8[1]) + ba.y),
                                       packbf_relu(__uint_as_float(q8[2]) + ba.z, __uint_as_float(q8[3]) + ba.w),
                                       packbf_relu(__uint_as_float(q8[4]) + bb.x, __uint_as_float(q8[5]) + bb.y),
                                       packbf_relu(__uint_as_float(q8[6]) + bb.z, __uint_as_float(q8[7]) + bb.w));
                            }
                        } else {     // outside the image: the depthwise zero padding
#pragma unroll
                            for (int h = 0; h < 4; ++h) sts128(sE + pin * (CE * 2) + ((((c0 >> 3) + h) ^ e_swz(pin)) << 4), 0u, 0u, 0u, 0u);
                        }
                    }
                }
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_tm_free);            // the control warp may overwrite the expand accumulators
                named_bar_sync(1, kNT);                              // E complete
                BN_STAMP(2 + 3 * e);
                if (g > 0) mbar_wait(&bar_proj, (g - 1) & 1);        // project(g-1) done: D is free
                // ---- depthwise 3x3: thread = (column x, 2-row group rg, 8-channel chunk j) ----
                {
                    const int x = tid & 15, rg = (tid >> 4) & 3, j = tid >> 6;
                    float acc[2][8];
                    {
                        const float4 ba = *reinterpret_cast<const float4*>(Bds + j * 8);
                        const float4 bb = *reinterpret_cast<const float4*>(Bds + j * 8 + 4);
#pragma unroll
                        for (int o = 0; o < 2; ++o) {
                            acc[o][0] = ba.x; acc[o][1] = ba.y; acc[o][2] = ba.z; acc[o][3] = ba.w;
                            acc[o][4] = bb.x; acc[o][5] = bb.y; acc[o][6] = bb.z; acc[o][7] = bb.w;
                        }
                    }
                    uint4 wv[9];
#pragma unroll
                    for (int t = 0; t < 9; ++t) wv[t] = lds128(sWds + (t * C::CEXP + j * 8) * 2);
                    constexpr int NR = STRIDE + 3;   // input rows feeding 2 output rows
#pragma unroll
                    for (int r = 0; r < NR; ++r) {
                        const int iy = (2 * rg) * STRIDE + r;
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) {
                            const int pin = iy * IW + x * STRIDE + kx;
                            const uint4 v = lds128(sE + pin * (CE * 2) + ((j ^ e_swz(pin)) << 4));
#pragma unroll
                            for (int o = 0; o < 2; ++o) {
                                const int ky = r - o * STRIDE;
                                if (ky >= 0 && ky < 3) fhfma8(acc[o], v, wv[ky * 3 + kx]);
                            }
                        }
                    }
#pragma unroll
                    for (int o = 0; o < 2; ++o) {
                        const int p = (2 * rg + o) * 16 + x;
                        sts128(sD + a_tile_off(p, j), packbf_relu(acc[o][0], acc[o][1]), packbf_relu(acc[o][2], acc[o][3]),
                               packbf_relu(acc[o][4], acc[o][5]), packbf_relu(acc[o][6], acc[o][7]));
                    }
                }
                fence_async_proxy();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_d_ready);            // -> control warp issues the project MMAs of this chunk
                named_bar_sync(1, kNT);                              // every read of E is done before the next epilogue rewrites it
                BN_STAMP(3 + 3 * e);
            }
            // ---- output epilogue: + bias (+ residual from global, L2-resident), bf16 NHWC store ----
            mbar_wait(&bar_proj, (g - 1) & 1);
            tc_fence_after_sync();
            BN_STAMP(1 + 3 * NCH);
            {
                const int q = warp & 3, part = warp >> 2;   // 16 warps = 4 lane quarters x 4 column parts
                const int p = q * 32 + lane;
                const int oy = oy0 + (p >> 4), ox = ox0 + (p & 15);
                const bool live = (oy < Ho) && (ox < Wo);
                constexpr int CP = COUT / 4;   // columns per part: 16, 24 or 32
                uint32_t r[CP];
#pragma unroll
                for (int c0 = 0; c0 < CP; c0 += 8) tmem_ld_32x32b_x8(tmem + ((uint32_t)(q * 32) << 16) + C::TM_PROJ + part * CP + c0, r + c0);
                const size_t pix = ((size_t)n * Ho + oy) * Wo + ox;
                uint4 res[CP / 8];
                if (RES && live) {
#pragma unroll
                    for (int c0 = 0; c0 < CP; c0 += 8) res[c0 / 8] = __ldg(reinterpret_cast<const uint4*>(in + pix * CIN + part * CP + c0));
                }
                tmem_ld_wait();
                if (live) {
#pragma unroll
                    for (int c0 = 0; c0 < CP; c0 += 8) {
                        const int co = part * CP + c0;
                        const float4 ba = *reinterpret_cast<const float4*>(Bp_s + co);
                        const float4 bb = *reinterpret_cast<const float4*>(Bp_s + co + 4);
                        float v[8] = {__uint_as_float(r[c0]) + ba.x, __uint_as_float(r[c0 + 1]) + ba.y, __uint_as_float(r[c0 + 2]) + ba.z,
                                      __uint_as_float(r[c0 + 3]) + ba.w, __uint_as_float(r[c0 + 4]) + bb.x, __uint_as_float(r[c0 + 5]) + bb.y,
                                      __uint_as_float(r[c0 + 6]) + bb.z, __uint_as_float(r[c0 + 7]) + bb.w};
                        if (RES) {
                            float f[8];
                            unpackbf8(res[c0 / 8], f);
#pragma unroll
                            for (int i = 0; i < 8; ++i) v[i] += f[i];
                        }
                        *reinterpret_cast<uint4*>(out + pix * COUT + co) =
                            make_uint4(packbf(v[0], v[1]), packbf(v[2], v[3]), packbf(v[4], v[5]), packbf(v[6], v[7]));
                    }
                }
            }
            tc_fence_before_sync();   // orders these TMEM reads before the d_ready arrival that releases project(0) of the next tile
            BN_STAMP(2 + 3 * NCH);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, C::TM_COLS);
}

#ifdef FSCNN_PHASE_TIMING
extern "C" int fscnn_debug_bneck_phases(long long* out64) {
    return cudaMemcpyFromSymbol(out64, g_bneck_phase, sizeof(long long) * 64) == cudaSuccess ? 0 : -1;
}
#endif

template <int CIN, int COUT, int STRIDE, bool RES>
static cudaError_t run_tc(const bf16* in, const unsigned char* tab_img, const bf16* we_img, const bf16* wp_img, bf16* out, int n,
                          int hi, int wi, int ho, int wo, cudaStream_t s) {
    using C = TcCfg<CIN, COUT, STRIDE>;
    static unsigned long long configured = 0;
    cudaError_t e = ensure_dyn_smem(bottleneck_tc_kernel<CIN, COUT, STRIDE, RES>, C::smem_bytes, configured);
    if (e != cudaSuccess) return e;
    CUtensorMap xmap;
    e = make_nhwc_halo_map(&xmap, in, n, hi, wi, CIN, C::IH, C::IW);
    if (e != cudaSuccess) return e;
    const int tiles_x = ceil_div(wo, C::TW), tiles_y = ceil_div(ho, C::TH), ntiles = tiles_x * tiles_y * n;
    const int grid = ntiles < num_sms() ? ntiles : num_sms();
    bottleneck_tc_kernel<CIN, COUT, STRIDE, RES><<<grid, kNTall, C::smem_bytes, s>>>(xmap, in, tab_img, we_img, wp_img, out, hi, wi,
                                                                                       ho, wo, tiles_x, tiles_y, ntiles);
    return cudaGetLastError();
}

// expanded channels per chunk: the weight images handed to launch_bottleneck_tc must be cut accordingly
int bottleneck_tc_chunk(int) { return 64; }

size_t bottleneck_tc_tab_bytes(int cin, int cout) { return (size_t)9 * 6 * cin * 2 + (size_t)2 * 6 * cin * 4 + (size_t)cout * 4; }

cudaError_t launch_bottleneck_tc(int cin, int cout, int stride, const bf16* in, const unsigned char* tab_img, const bf16* we_img,
                                 const bf16* wp_img, bf16* out, int n, int hi, int wi, int ho, int wo, cudaStream_t s) {
    if (stride == 1) return launch_bottleneck_s1_tc(cin, cout, in, tab_img, we_img, wp_img, out, n, hi, wi, s);
    if (cin == 64 && cout == 64 && stride == 2) return run_tc<64, 64, 2, false>(in, tab_img, we_img, wp_img, out, n, hi, wi, ho, wo, s);
    if (cin == 64 && cout == 96 && stride == 2) return run_tc<64, 96, 2, false>(in, tab_img, we_img, wp_img, out, n, hi, wi, ho, wo, s);
    return cudaErrorInvalidValue;
}

}  // namespace fscnn
