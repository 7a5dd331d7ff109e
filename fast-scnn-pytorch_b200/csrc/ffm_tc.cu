// ffm_tc.cu -- bf16 FeatureFusionModule (reference models/fast_scnn.py:190-218) on the tensor cores.
// Same algebra as ffm.cu: one K = 192 contraction whose first 64 columns are `higher`'s channels and
// whose last 128 columns are relu(BN(DW3x3(bilinear_align_corners(lower)))), biases pre-added, ReLU.
//
// Persistent, warp-specialised, software-pipelined over 8x16-pixel output tiles (one CTA per SM):
//   control warp (lane 0) : the 128 x 192 weight image once; per tile one TMA tensor copy that drops `higher`'s 8x16x64
//                           tile straight into columns 0..63 of the A operand (its [c/8][pixel][8 ch] box layout IS the
//                           SWIZZLE_NONE core-matrix layout); the 12 MMAs + commit
//   16 compute warps, iteration t:
//       U(t)              : bilinear resize (align_corners, size-driven: reference :209-212; 4 corner weights per pixel
//                           from a per-tile tap table) of `lower` on the 10x18
//                           halo tile -> bf16, swizzled rows (lanes along channels: corner reads are 256-byte coalesced)
//       depthwise(t)      : 3x3 + bias + ReLU on U (FHFMA.BF16) -> A[t&1] columns 64..191; arrive -> MMA(t) is issued
//       epilogue(t-1)     : TMEM[(t-1)&1] -> + bias, ReLU -> bf16 NHWC store        (runs while MMA(t) multiplies)
//   A and the TMEM accumulator are double-buffered; U is single-buffered (two named barriers per tile among the
//   compute warps: U complete / U consumed).
#include "kernels.h"
#include "tma_host.h"
#include "umma.cuh"

namespace fscnn {

namespace {
constexpr int kIW = 18, kPIN = 180, kPINP = 184, kCL = 128, kCH = 64, kK = 192, kCO = 128;
constexpr int kFfNT = 512, kFfNTall = kFfNT + 32;
constexpr int kABytes = 128 * kK * 2;               // one A operand: 128 x 192 bf16
constexpr int oU = 0;                               // [184][256 B] resized halo tile
constexpr int oA = kPINP * kCL * 2;                 // 2 x A
constexpr int oB = oA + 2 * kABytes;                // 128 x 192 weight image (48 KB), resident
constexpr int oWd = oB + kCO * kK * 2;              // bf16 [9][128]
constexpr int oBd = oWd + 9 * kCL * 2;
constexpr int oBc = oBd + kCL * 4;
constexpr int oTab = oBc + kCO * 4;                 // 2 x [184] x {int4 source offsets, float4 corner weights}: bilinear taps of a halo tile
constexpr int kTabBytes = kPINP * 32;
constexpr int kSmem = oTab + 2 * kTabBytes;
static_assert(kSmem <= 227 * 1024, "shared memory");
}  // namespace

#ifdef FSCNN_PHASE_TIMING   // debug build only: clock64 stamps of the 3rd tile of CTA 9
__device__ long long g_ffm_phase[16];
#define FF_STAMP(i) do { if (tid == 0 && blockIdx.x == 9 && lt == 2) g_ffm_phase[i] = clock64(); } while (0)
extern "C" int fscnn_debug_ffm_phases(long long* out16) {
    return cudaMemcpyFromSymbol(out16, g_ffm_phase, sizeof(long long) * 16) == cudaSuccess ? 0 : -1;
}
#else
#define FF_STAMP(i) do { } while (0)
#endif

__global__ void __launch_bounds__(kFfNTall, 1)
ffm_tc_kernel(const __grid_constant__ CUtensorMap hmap, const bf16* __restrict__ lower, FfmW w, const bf16* __restrict__ wcat_img,
              bf16* __restrict__ out, int Hh, int Wh, int Hl, int Wl, int tiles_x, int tiles_y, int ntiles) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar_w, bar_hi[2], bar_a[2], bar_mma[2];
    __shared__ uint32_t tmem_base_s;
    const uint32_t sWd = smem_u32(sm + oWd);   // depthwise weights [9][128] as bf16 (FHFMA operands)
    float* Bds = reinterpret_cast<float*>(sm + oBd);
    float* Bcs = reinterpret_cast<float*>(sm + oBc);
    const uint32_t sU = smem_u32(sm + oU), sA = smem_u32(sm + oA), sB = smem_u32(sm + oB);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int gstep = gridDim.x;
    const int my_tiles = (ntiles - (int)blockIdx.x + gstep - 1) / gstep;
    auto tile_origin = [&](int lt, int& n, int& oy0, int& ox0) {
        const int tile = blockIdx.x + lt * gstep;
        const int tx = tile % tiles_x, r = tile / tiles_x;
        n = r / tiles_y; oy0 = (r % tiles_y) * 8; ox0 = tx * 16;
    };

    if (tid == 0) {
        mbar_init(&bar_w, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&bar_hi[i], 1); mbar_init(&bar_mma[i], 1); mbar_init(&bar_a[i], kFfNT / 32); }
        fence_mbar_init();
    }
    if (warp == 0) { tmem_alloc(&tmem_base_s, 256); tmem_relinquish(); }
    for (int i = tid; i < 9 * kCL / 2; i += kFfNTall)
        reinterpret_cast<uint32_t*>(sm + oWd)[i] = packbf(__ldg(w.wd + 2 * i), __ldg(w.wd + 2 * i + 1));
    if (tid < kCL) { Bds[tid] = __ldg(w.bd + tid); Bcs[tid] = __ldg(w.bcat + tid); }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_base_s;

    if (warp == kFfNT / 32) {
        // =========================== control warp ===========================
        if (lane == 0) {
            auto load_hi = [&](int lt) {   // `higher` tile -> A[lt&1] columns 0..63 (8 channel groups x 128 pixels x 16 B)
                int n, oy0, ox0;
                tile_origin(lt, n, oy0, ox0);
                mbar_arrive_expect_tx(&bar_hi[lt & 1], 128 * kCH * 2);
                tma_load_halo(sA + (lt & 1) * kABytes, &hmap, ox0, oy0, n, &bar_hi[lt & 1]);
            };
            tma_prefetch_desc(&hmap);
            mbar_arrive_expect_tx(&bar_w, kCO * kK * 2);
            bulk_g2s(sm + oB, wcat_img, kCO * kK * 2, &bar_w);
            load_hi(0);
            if (my_tiles > 1) load_hi(1);
            mbar_wait(&bar_w, 0);
            constexpr uint32_t idesc = make_idesc_bf16(128, kCO);
#pragma unroll 1
            for (int lt = 0; lt < my_tiles; ++lt) {
                mbar_wait(&bar_a[lt & 1], (lt >> 1) & 1);        // depthwise(lt) written (writers fenced the async proxy)
                mbar_wait(&bar_hi[lt & 1], (lt >> 1) & 1);       // `higher` tile landed
                tc_fence_after_sync();
#pragma unroll
                for (int k16 = 0; k16 < kK / 16; ++k16)
                    umma_bf16_ss(tmem + (lt & 1) * kCO, make_smem_desc(sA + (lt & 1) * kABytes + k16 * 4096, 2048, 128),
                                 make_smem_desc(sB + k16 * 2 * (kCO * 16), kCO * 16, 128), idesc, k16 > 0);
                umma_commit(&bar_mma[lt & 1]);
                if (lt >= 1 && lt + 1 < my_tiles) {              // A[(lt+1)&1] is free once MMA(lt-1) has completed
                    mbar_wait(&bar_mma[(lt - 1) & 1], ((lt - 1) >> 1) & 1);
                    load_hi(lt + 1);
                }
            }
        }
    } else {
        // =========================== compute warps ===========================
        const float scy = Hh > 1 ? (float)(Hl - 1) / (float)(Hh - 1) : 0.f;
        const float scx = Wh > 1 ? (float)(Wl - 1) / (float)(Wh - 1) : 0.f;
        const int q = warp & 3, part = warp >> 2;   // epilogue: 4 lane quarters x 4 parts of 32 columns
        // bilinear taps of tile lt's halo pixels (align_corners=True, size-driven: reference :209-212), one thread per pixel:
        // {element offset of the top-left / bottom-left source pixel, step to the right neighbour, valid} + the 4 corner weights
        auto build_tab = [&](int lt) {
            if (tid < kPINP) {
                int n, oy0, ox0;
                tile_origin(lt, n, oy0, ox0);
                const int pin = tid;
                const int y = oy0 - 1 + pin / kIW, x = ox0 - 1 + pin % kIW;
                int4 t = make_int4(0, 0, 0, 0);
                float4 wgt = make_float4(0.f, 0.f, 0.f, 0.f);
                if (pin < kPIN && y >= 0 && y < Hh && x >= 0 && x < Wh) {
                    const float fy = scy * (float)y, fx = scx * (float)x;
                    const int y0 = min((int)fy, Hl - 1), x0 = min((int)fx, Wl - 1);
                    const int y1 = min(y0 + 1, Hl - 1), dx = min(x0 + 1, Wl - 1) - x0;   // dx: 0 at the right border, else 1
                    const float ly = fy - (float)y0, lx = fx - (float)x0;
                    const float hy = 1.f - ly, hx = 1.f - lx;
                    t = make_int4((y0 * Wl + x0) * kCL, (y1 * Wl + x0) * kCL, dx * kCL, 1);
                    wgt = make_float4(hy * hx, hy * lx, ly * hx, ly * lx);
                }
                uint8_t* e = sm + oTab + (lt & 1) * kTabBytes + pin * 32;
                *reinterpret_cast<int4*>(e) = t;
                *reinterpret_cast<float4*>(e + 16) = wgt;
            }
        };
        build_tab(0);
        named_bar_sync(1, kFfNT);
        auto epilogue = [&](int lt) {
            int n, oy0, ox0;
            tile_origin(lt, n, oy0, ox0);
            mbar_wait(&bar_mma[lt & 1], (lt >> 1) & 1);
            tc_fence_after_sync();
            uint32_t r[32];
            tmem_ld_32x32b_x32(tmem + ((uint32_t)(q * 32) << 16) + (lt & 1) * kCO + part * 32, r);
            tmem_ld_wait();
            // A lane owns one pixel (TMEM lane) x 32 channels; storing that directly touches 32 different 128-byte lines per
            // instruction.  Stage the warp's 32 x 32 block in a private 2 KB slab (XOR-swizzled 16-byte chunks) inside the
            // depthwise half of A[lt&1] -- its MMA has completed and the next writer (depthwise lt+2) is behind the named
            // barrier every warp passes after this epilogue -- and write it back with consecutive lanes on consecutive chunks.
            const uint32_t slab = sA + (lt & 1) * kABytes + 128 * kCH * 2 + warp * 2048;
#pragma unroll
            for (int c0 = 0; c0 < 32; c0 += 8) {
                const float4 ba = *reinterpret_cast<const float4*>(Bcs + part * 32 + c0);
                const float4 bb = *reinterpret_cast<const float4*>(Bcs + part * 32 + c0 + 4);
                const uint32_t* q8 = r + c0;
                sts128(slab + lane * 64 + (((c0 >> 3) ^ ((lane >> 1) & 3)) << 4),
                       packbf_relu(__uint_as_float(q8[0]) + ba.x, __uint_as_float(q8[1]) + ba.y),
                       packbf_relu(__uint_as_float(q8[2]) + ba.z, __uint_as_float(q8[3]) + ba.w),
                       packbf_relu(__uint_as_float(q8[4]) + bb.x, __uint_as_float(q8[5]) + bb.y),
                       packbf_relu(__uint_as_float(q8[6]) + bb.z, __uint_as_float(q8[7]) + bb.w));
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int px = (lane >> 2) + 8 * i, ch = lane & 3;
                const uint4 v = lds128(slab + px * 64 + ((ch ^ ((px >> 1) & 3)) << 4));
                const int pp = q * 32 + px;
                const int yy = oy0 + (pp >> 4), xx = ox0 + (pp & 15);
                if (yy < Hh && xx < Wh) *reinterpret_cast<uint4*>(out + (((size_t)n * Hh + yy) * Wh + xx) * kCO + part * 32 + ch * 8) = v;
            }
            __syncwarp();
        };
#pragma unroll 1
        for (int lt = 0; lt < my_tiles; ++lt) {
            int n, oy0, ox0;
            tile_origin(lt, n, oy0, ox0);
            FF_STAMP(0);
            if (lt > 0) named_bar_sync(1, kFfNT);                // every depthwise read of U(lt-1) is done; tab[lt&1] is visible
            // ---- U = resize(lower) on the halo tile; item = (halo pixel, 8-channel chunk), lanes along chunks ----
            FF_STAMP(1);
            const bf16* lbase = lower + (size_t)n * Hl * Wl * kCL;
            const uint8_t* tabp = sm + oTab + (lt & 1) * kTabBytes;
#pragma unroll 2
            for (int i = tid; i < kPINP * 16; i += kFfNT) {
                const int pin = i >> 4, k8 = i & 15;
                const int4 t = *reinterpret_cast<const int4*>(tabp + pin * 32);
                uint4 o = make_uint4(0, 0, 0, 0);
                if (t.w) {
                    const float4 wg = *reinterpret_cast<const float4*>(tabp + pin * 32 + 16);
                    const bf16* b0 = lbase + t.x + k8 * 8;
                    const bf16* b1 = lbase + t.y + k8 * 8;
                    float v00[8], v01[8], v10[8], v11[8];
                    unpackbf8(__ldg(reinterpret_cast<const uint4*>(b0)), v00);
                    unpackbf8(__ldg(reinterpret_cast<const uint4*>(b0 + t.z)), v01);
                    unpackbf8(__ldg(reinterpret_cast<const uint4*>(b1)), v10);
                    unpackbf8(__ldg(reinterpret_cast<const uint4*>(b1 + t.z)), v11);
                    float u[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) u[c] = fmaf(wg.w, v11[c], fmaf(wg.z, v10[c], fmaf(wg.y, v01[c], wg.x * v00[c])));
                    o = make_uint4(packbf(u[0], u[1]), packbf(u[2], u[3]), packbf(u[4], u[5]), packbf(u[6], u[7]));
                }
                sts128(sU + pin * (kCL * 2) + ((k8 ^ (pin & 7)) << 4), o.x, o.y, o.z, o.w);
            }
            FF_STAMP(2);
            named_bar_sync(2, kFfNT);                            // U complete
            FF_STAMP(3);
            // ---- depthwise 3x3 + bias + ReLU on U -> A[lt&1] columns 64..191 (one output column x 4 channels per thread);
            //      A[lt&1] is free: this warp waited for MMA(lt-2) in its epilogue ----
            dw3x3_s1_col4<kCL * 2, kIW>(sU, tid & 15, tid >> 4, sWd, kCL, Bds, sA + (lt & 1) * kABytes, 8);
            FF_STAMP(4);
            fence_async_proxy();
            tc_fence_before_sync();      // this warp's TMEM reads of tile lt-2 precede the MMA the arrival releases
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_a[lt & 1]);
            if (lt + 1 < my_tiles) build_tab(lt + 1);            // published by the named barrier at the top of the next iteration
            FF_STAMP(5);
            if (lt >= 1) epilogue(lt - 1);
            FF_STAMP(6);
        }
        epilogue(my_tiles - 1);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

cudaError_t launch_ffm_tc(const bf16* higher, const bf16* lower, const FfmW& w, const bf16* wcat_img, bf16* out, int n, int hh,
                          int wh, int hl, int wl, cudaStream_t s) {
    static unsigned long long configured = 0;
    cudaError_t e = ensure_dyn_smem(ffm_tc_kernel, (size_t)kSmem, configured);
    if (e != cudaSuccess) return e;
    CUtensorMap hmap;
    e = make_nhwc_halo_map(&hmap, higher, n, hh, wh, kCH, 8, 16);
    if (e != cudaSuccess) return e;
    const int tiles_x = ceil_div(wh, 16), tiles_y = ceil_div(hh, 8), ntiles = tiles_x * tiles_y * n;
    const int grid = ntiles < num_sms() ? ntiles : num_sms();
    ffm_tc_kernel<<<grid, kFfNTall, kSmem, s>>>(hmap, lower, w, wcat_img, out, hh, wh, hl, wl, tiles_x, tiles_y, ntiles);
    return cudaGetLastError();
}

}  // namespace fscnn
