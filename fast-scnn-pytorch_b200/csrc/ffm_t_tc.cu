// ffm_t_tc.cu -- bf16 FeatureFusionModule (reference models/fast_scnn.py:190-218) with the bilinear resize ON the tensor core
// and the depthwise 3x3 in registers (the transposed scheme of bottleneck_s1t_tc.cu):
//
//   resize    U^T[128 ch x 192 px] = L^T[128 ch x 48 src] * R^T[48 src x 192 px]      (tcgen05.mma)
//             L = the <= 5x8 patch of `lower` under the tile's 10x18 halo, dropped by ONE TMA tensor copy as [c/8][src][8 ch]:
//             that is an MN-major A operand (8 channels contiguous, LBO = 128 B between 8-pixel groups, SBO = 640 B between
//             8-channel groups).  R = the align_corners interpolation matrix of the halo pixels (reference :209-212), four
//             bf16 weights per row, all-zero rows for halo pixels outside the image (= the depthwise zero padding), rebuilt
//             per tile by the compute threads (one row each).
//   depthwise thread = one channel (its TMEM lane) x two output rows: tcgen05.ld -> bf16 pack -> FHFMA.BF16 -> MN-major D
//   fuse      OUT[128 px x 128] = higher[128 px x 64] * Wh + D[128 px x 128] * Wl   (4 + 8 tcgen05.mma into one accumulator)
//             + (pre-added) biases, ReLU -> bf16 NHWC
//
// Persistent CTA per SM, 16 compute warps + a resize controller (patch loads, resize MMAs) + a fuse controller (weights,
// `higher` tile loads, fuse MMAs); TMEM resize accumulators, patches, R and D are double-buffered, so the tensor core works
// two tiles ahead of the CUDA cores.  Interpolation weights are rounded to bf16 (2^-9 relative), the same rounding the
// resized tile itself gets in ffm_tc.cu.
#include "kernels.h"
#include "tma_host.h"
#include "umma.cuh"

namespace fscnn {

namespace {
constexpr int kFW = 16, kFThreads = (kFW + 2) * 32;
constexpr int kIW = 18, kPIN = 180, kNB = 192;
constexpr int kPH = 5, kPW = 8, kKR = 48;                     // lower patch 5 x 8 source pixels, resize K (padded to 48)
constexpr int kCL = 128, kCH = 64, kCO = 128;
constexpr int HI_BYTES = 128 * kCH * 2;                       // [64/8][128 px][8]: LBO 2048, SBO 128
constexpr int PATCH_BYTES = kPH * kPW * kCL * 2;              // [128/8][40 src][8]
constexpr int R_LBO = kNB * 16, R_BYTES = (kKR / 8) * R_LBO;  // [48/8][192 px][8]
constexpr int D_LBO = 128, D_SBO = 2048, D_BYTES = 16 * D_SBO;
constexpr int W_BYTES = kCO * (kCH + kCL) * 2;                // [192/8][128 cout][8]: LBO 2048, SBO 128
constexpr int oHi = 0, oPatch = oHi + 2 * HI_BYTES, oR = oPatch + 2 * PATCH_BYTES, oD = oR + 2 * R_BYTES, oW = oD + 2 * D_BYTES;
constexpr int oBias = oW + W_BYTES;                           // B block [2 k-blocks][128 cout][8]: {bias head, remainder, 0 x 6} | zeros
constexpr int BIAS_BYTES = 2 * kCO * 16;
constexpr int oOnes = oBias + BIAS_BYTES;                     // A block: one core matrix of rows {1, 1, 0 x 6} + one of zeros (SBO = 0)
constexpr int kSmemT = oOnes + 256;
static_assert(kSmemT <= 227 * 1024 - 256, "shared memory");
constexpr int TM_OUT = 2 * kNB;                               // resize accumulators 2 x 192 columns, fuse accumulator 128
}  // namespace

__global__ void __launch_bounds__(kFThreads, 1)
ffm_t_kernel(const __grid_constant__ CUtensorMap hmap, const __grid_constant__ CUtensorMap lmap, const unsigned char* __restrict__ tab,
             const bf16* __restrict__ wcat_img, bf16* __restrict__ out, int Hh, int Wh, int Hl, int Wl, int tiles_x, int tiles_y,
             int ntiles) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar_w, bar_hi[2], bar_patch[2], bar_r[2], bar_exp[2], bar_tmfree[2], bar_dready[2], bar_proj[2],
        bar_projfree;
    __shared__ uint32_t tmem_base_s;
    const uint32_t sHi = smem_u32(sm + oHi), sPatch = smem_u32(sm + oPatch), sR = smem_u32(sm + oR), sD = smem_u32(sm + oD),
                   sW = smem_u32(sm + oW);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int gstep = gridDim.x;
    const int my_tiles = (ntiles - (int)blockIdx.x + gstep - 1) / gstep;
    const float scy = Hh > 1 ? (float)(Hl - 1) / (float)(Hh - 1) : 0.f;
    const float scx = Wh > 1 ? (float)(Wl - 1) / (float)(Wh - 1) : 0.f;
    auto tile_origin = [&](int lt, int& n, int& oy0, int& ox0) {
        const int tile = blockIdx.x + lt * gstep;
        const int tx = tile % tiles_x, r = tile / tiles_x;
        n = r / tiles_y; oy0 = (r % tiles_y) * 8; ox0 = tx * 16;
    };
    // first source row / column under the tile's halo (the halo's first in-image pixel has the smallest source coordinate)
    auto patch_origin = [&](int oy0, int ox0, int& ry0, int& rx0) {
        ry0 = min((int)(scy * (float)max(oy0 - 1, 0)), Hl - 1);
        rx0 = min((int)(scx * (float)max(ox0 - 1, 0)), Wl - 1);
    };
    pdl_launch_dependents();

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_hi[i], 1); mbar_init(&bar_patch[i], 1); mbar_init(&bar_exp[i], 1); mbar_init(&bar_proj[i], 1);
            mbar_init(&bar_r[i], kFW); mbar_init(&bar_tmfree[i], kFW); mbar_init(&bar_dready[i], kFW);
        }
        mbar_init(&bar_w, 1); mbar_init(&bar_projfree, kFW);
        fence_mbar_init();
    }
    // R rows 180..191 stay zero for good; the patch buffers are cleared because the resize MMA's K padding (source pixels
    // 40..47) reads into the neighbouring channel group / buffer: those products meet zero weights, so they must be finite
    for (int i = tid; i < (2 * PATCH_BYTES + 2 * R_BYTES) / 16; i += kFThreads)
        *reinterpret_cast<uint4*>(sm + oPatch + i * 16) = make_uint4(0u, 0u, 0u, 0u);
    // the fused bias rides through the tensor core: OUT += ones[128 px x 16] * biasblock[128 cout x 16]^T (bf16 head + remainder)
    if (tid < kCO) {
        const float b = __ldg(reinterpret_cast<const float*>(tab + (size_t)kCL * 32) + tid);
        const __nv_bfloat16 bh = __float2bfloat16_rn(b), bl = __float2bfloat16_rn(b - __bfloat162float(bh));
        const uint32_t w0 = (uint32_t)(*reinterpret_cast<const uint16_t*>(&bh)) | ((uint32_t)(*reinterpret_cast<const uint16_t*>(&bl)) << 16);
        *reinterpret_cast<uint4*>(sm + oBias + tid * 16) = make_uint4(w0, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(sm + oBias + kCO * 16 + tid * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    if (tid < 16) *reinterpret_cast<uint4*>(sm + oOnes + tid * 16) = make_uint4(tid < 8 ? 0x3F803F80u : 0u, 0u, 0u, 0u);
    fence_async_proxy();
    if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_base_s;

    if (warp == kFW) {
        // =========================== resize controller ===========================
        if (lane == 0) {
            auto load_patch = [&](int lt) {
                int n, oy0, ox0, ry0, rx0;
                tile_origin(lt, n, oy0, ox0);
                patch_origin(oy0, ox0, ry0, rx0);
                mbar_arrive_expect_tx(&bar_patch[lt & 1], PATCH_BYTES);
                tma_load_halo(sPatch + (lt & 1) * PATCH_BYTES, &lmap, rx0, ry0, n, &bar_patch[lt & 1]);
            };
            constexpr uint32_t idesc_rs = make_idesc_bf16(128, kNB) | (1u << 15);   // A (= the patch) is MN-major
            tma_prefetch_desc(&lmap);
            pdl_wait();      // `lower` is the previous stage's output
            load_patch(0);
            if (my_tiles > 1) load_patch(1);
#pragma unroll 1
            for (int t = 0; t < my_tiles; ++t) {
                mbar_wait(&bar_patch[t & 1], (t >> 1) & 1);
                mbar_wait(&bar_r[t & 1], (t >> 1) & 1);                                   // interpolation matrix written
                if (t >= 2) mbar_wait(&bar_tmfree[t & 1], ((t - 2) >> 1) & 1);            // accumulator drained by tile t-2
                tc_fence_after_sync();
                const uint64_t da0 = make_smem_desc(sPatch + (t & 1) * PATCH_BYTES, 128, kPH * kPW * 16);
                const uint64_t db0 = make_smem_desc(sR + (t & 1) * R_BYTES, R_LBO, 128);
#pragma unroll
                for (int k16 = 0; k16 < kKR / 16; ++k16)
                    umma_bf16_ss(tmem + (t & 1) * kNB, da0 + (uint64_t)(k16 * ((2 * 128) >> 4)), db0 + (uint64_t)(k16 * ((2 * R_LBO) >> 4)),
                                 idesc_rs, k16 > 0);
                umma_commit(&bar_exp[t & 1]);
                if (t + 2 < my_tiles) {
                    mbar_wait(&bar_exp[t & 1], (t >> 1) & 1);          // resize(t) has completed: its patch buffer is free
                    load_patch(t + 2);
                }
            }
        }
    } else if (warp == kFW + 1) {
        // =========================== fuse controller ===========================
        if (lane == 0) {
            auto load_hi = [&](int lt) {
                int n, oy0, ox0;
                tile_origin(lt, n, oy0, ox0);
                mbar_arrive_expect_tx(&bar_hi[lt & 1], HI_BYTES);
                tma_load_halo(sHi + (lt & 1) * HI_BYTES, &hmap, ox0, oy0, n, &bar_hi[lt & 1]);
            };
            constexpr uint32_t idesc_hi = make_idesc_bf16(128, kCO);
            constexpr uint32_t idesc_lo = make_idesc_bf16(128, kCO) | (1u << 15);     // A (= D) is MN-major
            tma_prefetch_desc(&hmap);
            mbar_arrive_expect_tx(&bar_w, W_BYTES);
            bulk_g2s(sm + oW, wcat_img, W_BYTES, &bar_w);
            pdl_wait();      // the weights are on their way; `higher` is an earlier stage's output
            load_hi(0);
            if (my_tiles > 1) load_hi(1);
            mbar_wait(&bar_w, 0);
#pragma unroll 1
            for (int t = 0; t < my_tiles; ++t) {
                mbar_wait(&bar_hi[t & 1], (t >> 1) & 1);
                if (t > 0) mbar_wait(&bar_projfree, (t - 1) & 1);                          // the previous tile's accumulator has been read
                tc_fence_after_sync();
                const uint64_t dh0 = make_smem_desc(sHi + (t & 1) * HI_BYTES, 2048, 128);
                const uint64_t dw0 = make_smem_desc(sW, 2048, 128);
                umma_bf16_ss(tmem + TM_OUT, make_smem_desc(smem_u32(sm + oOnes), 128, 0), make_smem_desc(smem_u32(sm + oBias), kCO * 16, 128),
                             idesc_hi, 0);                                                 // bias
#pragma unroll
                for (int k16 = 0; k16 < kCH / 16; ++k16)                                   // higher's 64 channels: ready before D is
                    umma_bf16_ss(tmem + TM_OUT, dh0 + (uint64_t)(k16 * ((2 * 2048) >> 4)), dw0 + (uint64_t)(k16 * ((2 * 2048) >> 4)),
                                 idesc_hi, 1);
                mbar_wait(&bar_dready[t & 1], (t >> 1) & 1);                              // D written by the depthwise threads
                tc_fence_after_sync();
                const uint64_t dd0 = make_smem_desc(sD + (t & 1) * D_BYTES, D_LBO, D_SBO);
#pragma unroll
                for (int k16 = 0; k16 < kCL / 16; ++k16)
                    umma_bf16_ss(tmem + TM_OUT, dd0 + (uint64_t)(k16 * ((2 * D_LBO) >> 4)),
                                 dw0 + (uint64_t)((kCH / 16 + k16) * ((2 * 2048) >> 4)), idesc_lo, 1);
                umma_commit(&bar_proj[t & 1]);
                if (t + 2 < my_tiles) {
                    mbar_wait(&bar_proj[t & 1], (t >> 1) & 1);         // fuse(t) has completed: its `higher` buffer is free
                    load_hi(t + 2);
                }
            }
        }
    } else {
        // =========================== compute warps ===========================
        pdl_wait();
        const int q = warp & 3, s = warp >> 2;            // TMEM lane quarter, row strip (output rows 2s, 2s+1) / 32-column slice
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        // one row of the interpolation matrix per thread: lanes 0..22 of warps 0..7 (8 x 23 = 184 >= 180 rows): two builder warps
        // per scheduler, and half the warp-level instructions of spreading 12 lanes over all 16 warps
        auto build_r = [&](int lt) {
            const int pin = warp * 23 + lane;
            if (warp < 8 && lane < 23 && pin < kPIN) {
                int n, oy0, ox0, ry0, rx0;
                tile_origin(lt, n, oy0, ox0);
                patch_origin(oy0, ox0, ry0, rx0);
                const int y = oy0 - 1 + pin / kIW, x = ox0 - 1 + pin % kIW;
                const uint32_t row = sR + (lt & 1) * R_BYTES + pin * 16;
#pragma unroll
                for (int kb = 0; kb < kKR / 8; ++kb) sts128(row + kb * R_LBO, 0u, 0u, 0u, 0u);
                if (y >= 0 && y < Hh && x >= 0 && x < Wh) {
                    const float fy = scy * (float)y, fx = scx * (float)x;
                    const int y0 = min((int)fy, Hl - 1), x0 = min((int)fx, Wl - 1);
                    const int y1 = min(y0 + 1, Hl - 1), x1 = min(x0 + 1, Wl - 1);
                    const float ly = fy - (float)y0, lx = fx - (float)x0;
                    const float hy = 1.f - ly, hx = 1.f - lx;
                    float w00 = hy * hx, w01 = hy * lx, w10 = ly * hx, w11 = ly * lx;
                    if (x1 == x0) { w00 += w01; w10 += w11; w01 = 0.f; w11 = 0.f; }      // clamped at the right border
                    if (y1 == y0) { w00 += w10; w01 += w11; w10 = 0.f; w11 = 0.f; }      // clamped at the bottom border
                    const int jy0 = min(y0 - ry0, kPH - 1), jy1 = min(y1 - ry0, kPH - 1), jx0 = min(x0 - rx0, kPW - 1), jx1 = min(x1 - rx0, kPW - 1);
                    auto put = [&](int j, float wv) {
                        const __nv_bfloat16 b = __float2bfloat16_rn(wv);
                        asm volatile("st.shared.u16 [%0], %1;" ::"r"(row + (j >> 3) * R_LBO + (j & 7) * 2), "h"(*reinterpret_cast<const uint16_t*>(&b)) : "memory");
                    };
                    put(jy0 * kPW + jx0, w00);
                    if (x1 != x0) put(jy0 * kPW + jx1, w01);
                    if (y1 != y0) put(jy1 * kPW + jx0, w10);
                    if (x1 != x0 && y1 != y0) put(jy1 * kPW + jx1, w11);
                }
            }
        };
        auto publish_r = [&](int lt) {                    // after a fence.proxy.async of the writing threads
            if (lane == 0) mbar_arrive(&bar_r[lt & 1]);
        };
        auto epilogue = [&](int lt) {                     // ReLU -> bf16 NHWC (the bias is already in the accumulator)
            int n, oy0, ox0;
            tile_origin(lt, n, oy0, ox0);
            const int p = q * 32 + lane;
            const int oy = oy0 + (p >> 4), ox = ox0 + (p & 15);
            const bool live = (oy < Hh) && (ox < Wh);
            const size_t pix = ((size_t)n * Hh + oy) * Wh + ox;
            mbar_wait(&bar_proj[lt & 1], (lt >> 1) & 1);
            tc_fence_after_sync();
#pragma unroll
            for (int c0 = 0; c0 < 32; c0 += 16) {
                uint32_t r[16];
                tmem_ld_32x32b_x16(tmem + lane_base + TM_OUT + s * 32 + c0, r);
                tmem_ld_wait();
                if (c0 == 16) {
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar_projfree);   // accumulator read: the next tile's fuse MMAs may overwrite it
                }
                if (live) {
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const uint32_t* q8 = r + 8 * i;
                        *reinterpret_cast<uint4*>(out + pix * kCO + s * 32 + c0 + 8 * i) =
                            make_uint4(packbf_relu(__uint_as_float(q8[0]), __uint_as_float(q8[1])), packbf_relu(__uint_as_float(q8[2]), __uint_as_float(q8[3])),
                                       packbf_relu(__uint_as_float(q8[4]), __uint_as_float(q8[5])), packbf_relu(__uint_as_float(q8[6]), __uint_as_float(q8[7])));
                    }
                }
            }
        };
        // this thread's depthwise channel: 9 bf16 taps + fp32 bias
        const uint4* rec = reinterpret_cast<const uint4*>(tab + (size_t)(q * 32 + lane) * 32);
        const uint4 wa = __ldg(rec), wb = __ldg(rec + 1);
        const uint32_t wq[5] = {wa.x, wa.y, wa.z, wa.w, wb.x};
        const float bd = __uint_as_float(wb.y);
        build_r(0);
        if (my_tiles > 1) build_r(1);
        fence_async_proxy();
        __syncwarp();
        publish_r(0);
        if (my_tiles > 1) publish_r(1);
#pragma unroll 1
        for (int t = 0; t < my_tiles; ++t) {
            mbar_wait(&bar_exp[t & 1], (t >> 1) & 1);            // resize(t) has completed
            tc_fence_after_sync();
            uint32_t Ep[4][9];                                   // halo rows 2s .. 2s+3, column pairs, bf16
            {
                uint32_t r[72];
                const uint32_t t0 = tmem + lane_base + (t & 1) * kNB + (2 * s) * kIW;
                tmem_ld_32x32b_x64(t0, r);
                tmem_ld_32x32b_x8(t0 + 64, r + 64);
                tmem_ld_wait();
#pragma unroll
                for (int rr = 0; rr < 4; ++rr)
#pragma unroll
                    for (int i = 0; i < 9; ++i)
                        Ep[rr][i] = packbf(__uint_as_float(r[rr * kIW + 2 * i]), __uint_as_float(r[rr * kIW + 2 * i + 1]));
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_tmfree[t & 1]);      // resize(t+2) may overwrite this accumulator
            float acc[2][16];
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
            if (t >= 2) mbar_wait(&bar_proj[t & 1], ((t - 2) >> 1) & 1);   // fuse(t-2) has completed: D[t&1] is free
            const int k = q * 32 + lane;
            const uint32_t d0 = sD + (t & 1) * D_BYTES + (k >> 3) * D_LBO + (k & 7) * 16 + (4 * s) * D_SBO;
#pragma unroll
            for (int o = 0; o < 2; ++o)
#pragma unroll
                for (int hx = 0; hx < 2; ++hx)
                    sts128(d0 + (2 * o + hx) * D_SBO, packbf_relu(acc[o][8 * hx + 0], acc[o][8 * hx + 1]),
                           packbf_relu(acc[o][8 * hx + 2], acc[o][8 * hx + 3]), packbf_relu(acc[o][8 * hx + 4], acc[o][8 * hx + 5]),
                           packbf_relu(acc[o][8 * hx + 6], acc[o][8 * hx + 7]));
            if (t + 2 < my_tiles) build_r(t + 2);                // resize(t) has completed: R[t&1] may be rewritten; needed two tiles from now
            fence_async_proxy();                                 // one proxy fence (it costs a MEMBAR) for both the D and the R writes
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_dready[t & 1]);
            if (t + 2 < my_tiles) publish_r(t + 2);
            if (t >= 1) epilogue(t - 1);                         // deferred by one tile: fuse(t-1) ran during this tile's depthwise
        }
        epilogue(my_tiles - 1);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// The patch under a 10x18 halo must fit 5 x 8 source pixels: floor(9 s) + 2 <= 5 rows, floor(17 s) + 2 <= 8 columns.
bool ffm_t_supported(int hh, int wh, int hl, int wl) {
    const float scy = hh > 1 ? (float)(hl - 1) / (float)(hh - 1) : 0.f, scx = wh > 1 ? (float)(wl - 1) / (float)(wh - 1) : 0.f;
    return 9.f * scy < 2.99f && 17.f * scx < 5.99f && hl >= 1 && wl >= 1;
}

cudaError_t launch_ffm_t_tc(const bf16* higher, const bf16* lower, const unsigned char* tab, const bf16* wcat_img, bf16* out, int n,
                            int hh, int wh, int hl, int wl, cudaStream_t s) {
    static unsigned long long configured = 0;
    cudaError_t e = ensure_dyn_smem(ffm_t_kernel, (size_t)kSmemT, configured);
    if (e != cudaSuccess) return e;
    CUtensorMap hmap, lmap;
    e = make_nhwc_halo_map(&hmap, higher, n, hh, wh, kCH, 8, 16);
    if (e != cudaSuccess) return e;
    e = make_nhwc_halo_map(&lmap, lower, n, hl, wl, kCL, kPH, kPW);
    if (e != cudaSuccess) return e;
    const int tiles_x = ceil_div(wh, 16), tiles_y = ceil_div(hh, 8), ntiles = tiles_x * tiles_y * n;
    const int grid = ntiles < num_sms() ? ntiles : num_sms();
    return launch_pdl(ffm_t_kernel, grid, kFThreads, kSmemT, s, hmap, lmap, tab, wcat_img, out, hh, wh, hl, wl, tiles_x, tiles_y, ntiles);
}

}  // namespace fscnn
