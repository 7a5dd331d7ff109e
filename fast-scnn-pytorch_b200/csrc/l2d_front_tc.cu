// l2d_front_tc.cu -- bf16 LearningToDownsample.conv + dsconv1 in ONE kernel (reference
// models/fast_scnn.py:153-154, :157-160): dense 3x3 s2 p0 (3 -> 32) + BN + ReLU, depthwise 3x3 s2 p1 +
// BN + ReLU, pointwise 32 -> 48 + BN + ReLU.  The stem's 32-channel output (the largest tensor of the
// network: 67 MB fp32 / 33 MB bf16 per 1024x2048 image, written once and read once in the unfused plan)
// never reaches HBM; both contractions run on the tensor cores.
//
// CTA = 8x16 dsconv1 output pixels; it needs the 17x33 stem pixels around them (halo of the stride-2
// depthwise conv), i.e. a 35x67 input patch:
//   stage patch (fp32 NCHW or raw uint8 HWC + ToTensor/Normalize) -> im2col gather into five A tiles
//   [128 x 32] -> 10 MMAs into TMEM[5][128 x 32] -> bias, ReLU, zero outside the stem image -> bf16
//   E[561][32] (over the dead patch) -> depthwise s2 (fp32) -> A2[128 x 32] (over the dead A tiles) ->
//   2 MMAs [128 x 48] -> bias, ReLU -> bf16 NHWC.
// 82 KB of shared memory and 256 TMEM columns: two CTAs per SM overlap each other's phases.
#include "kernels.h"
#include "umma.cuh"

#include "../../include/fscnn_b200.h"

namespace fscnn {

#ifdef FSCNN_PHASE_TIMING   // debug build only: per-phase clock64 stamps of one CTA (see tools/phase_timing.py)
__device__ long long g_front_phase[16];
#define PHASE_STAMP(i) do { if (tid == 0 && blockIdx.x == 7 && blockIdx.y == 9 && blockIdx.z == 0) g_front_phase[i] = clock64(); } while (0)
#else
#define PHASE_STAMP(i) do { } while (0)
#endif

namespace {
constexpr int SH = 17, SW = 33, SPIX = SH * SW;       // stem pixels per CTA (561)
constexpr int NMT = 5;                                 // A tiles of 128 stem pixels
constexpr int PR = 35, PC = 67, PLD = 68;              // input patch rows / cols / pitch
constexpr int kRW = 52, kRawWords = PR * kRW;          // uint8 input: 52 words (208 bytes) per patch row
constexpr int oIn = 0;                                 // fp32 [3][35][68] = 28560 B ... later E: 561 x 64 B = 35904 B
constexpr int R0 = 35968;                              // region 0 size (multiple of 128)
constexpr int oE = 0;
constexpr int oA = R0;                                 // 5 x 8 KB im2col tiles ... later A2 (8 KB)
constexpr int oWs = oA + NMT * 8192;                   // stem weight image 32 x 32 bf16
constexpr int oWp = oWs + 2048;                        // pointwise image 48 x 32 bf16
constexpr int oWd = oWp + 3072;                        // fp32 [9][32]
constexpr int oBs = oWd + 9 * 32 * 4;                  // stem bias [32]
constexpr int oBd = oBs + 128;                         // dw bias [32]
constexpr int oBp = oBd + 128;                         // pw bias [48]
constexpr int kSmem = oBp + 192;
constexpr int TM_STEM = 0, TM_PW = NMT * 32;           // 160 + 48 columns -> allocate 256
}  // namespace

template <int FMT>
__global__ void __launch_bounds__(kThreads, 2)
l2d_front_kernel(const void* __restrict__ x, StemIn prm, const bf16* __restrict__ ws_img, const float* __restrict__ bs,
                 DsW w, const bf16* __restrict__ wp_img, bf16* __restrict__ out, int H, int W, int H1, int W1, int H2, int W2) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar_stem, bar_pw;
    __shared__ uint32_t tmem_base_s;
    float* In = reinterpret_cast<float*>(sm + oIn);
    float* Wds = reinterpret_cast<float*>(sm + oWd);
    float* Bss = reinterpret_cast<float*>(sm + oBs);
    float* Bds = reinterpret_cast<float*>(sm + oBd);
    float* Bps = reinterpret_cast<float*>(sm + oBp);
    const uint32_t sE = smem_u32(sm + oE), sA = smem_u32(sm + oA), sWs = smem_u32(sm + oWs), sWp = smem_u32(sm + oWp);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n = blockIdx.z, oy0 = blockIdx.y * 8, ox0 = blockIdx.x * 16;
    const int sy0 = 2 * oy0 - 1, sx0 = 2 * ox0 - 1;      // first stem pixel of the halo tile (may be -1)
    const int iy0 = 2 * sy0, ix0 = 2 * sx0;              // first input row / column of the patch (may be -2)

    PHASE_STAMP(0);
    if (tid == 0) { mbar_init(&bar_stem, 1); mbar_init(&bar_pw, 1); fence_mbar_init(); }
    if (warp == 0) { tmem_alloc(&tmem_base_s, 256); tmem_relinquish(); }
    PHASE_STAMP(1);
    if (tid < 128) reinterpret_cast<uint4*>(sm + oWs)[tid] = __ldg(reinterpret_cast<const uint4*>(ws_img) + tid);
    if (tid < 192) reinterpret_cast<uint4*>(sm + oWp)[tid] = __ldg(reinterpret_cast<const uint4*>(wp_img) + tid);
    for (int i = tid; i < 9 * 32; i += kThreads) Wds[i] = __ldg(w.wd + i);
    if (tid < 32) { Bss[tid] = __ldg(bs + tid); Bds[tid] = __ldg(w.bd + tid); }
    if (tid < 48) Bps[tid] = __ldg(w.bp + tid);

    // ---- stage the input patch.  fp32 NCHW: thread = (channel, column), walks the 35 rows with a constant pointer
    //      step; all 35 loads are issued before the first store, and there is no per-element index arithmetic ----
    if (FMT == FSCNN_IN_F32_NCHW) {
        const float* xf = reinterpret_cast<const float*>(x);
        if (tid < 3 * PC) {
            const int ci = tid / PC, c = tid - ci * PC;
            const int ix = ix0 + c;
            const bool cok = ix >= 0 && ix < W;
            const float* src = xf + (((size_t)n * 3 + ci) * H) * W + (cok ? ix : 0);
            float v[PR];
#pragma unroll
            for (int r = 0; r < PR; ++r) {
                const int iy = iy0 + r;
                v[r] = (cok && iy >= 0 && iy < H) ? __ldg(src + (size_t)iy * W) : 0.f;
            }
            float* dst = In + ci * PR * PLD + c;
#pragma unroll
            for (int r = 0; r < PR; ++r) dst[r * PLD] = v[r];
        }
    } else {
        // raw uint8 HWC: park each 201-byte patch row as 52 32-bit words (row start rounded down to 4 bytes).  ToTensor +
        // Normalize are affine per input channel and the stem has no padding, so they are folded into the stem weights
        // and bias (stem_refold_kernel); pixel values 0..255 are exact in bf16.
        const unsigned char* xb = reinterpret_cast<const unsigned char*>(x);
        uint32_t* Raw = reinterpret_cast<uint32_t*>(sm + oIn);
        const int rowb = W * 3;
        const int b0 = ix0 * 3 - 2;                                  // ix0*3 == 2 (mod 4)
        if ((W & 3) == 0) {
            // thread = (word column wq < 52, row group rg < 4): 9 rows each, all loads in flight before the stores
            if (tid < 4 * kRW) {
                const int rg = tid / kRW, wq = tid - rg * kRW;
                const int b = b0 + 4 * wq;
                const bool cok = b >= 0 && b < rowb;
                const unsigned char* src = xb + (size_t)n * H * rowb + (cok ? b : 0);
                uint32_t v[9];
#pragma unroll
                for (int j = 0; j < 9; ++j) {
                    const int r = rg * 9 + j, iy = iy0 + r;
                    v[j] = (cok && r < PR && iy >= 0 && iy < H) ? __ldg(reinterpret_cast<const uint32_t*>(src + (size_t)iy * rowb)) : 0u;
                }
#pragma unroll
                for (int j = 0; j < 9; ++j) {
                    const int r = rg * 9 + j;
                    if (r < PR) Raw[r * kRW + wq] = v[j];
                }
            }
        } else {   // unaligned row pitch: byte loads into the same layout
            unsigned char* rawb = reinterpret_cast<unsigned char*>(Raw);
            for (int i = tid; i < PR * kRW * 4; i += kThreads) {
                const int r = i / (kRW * 4), bb = i - r * (kRW * 4);
                const int iy = iy0 + r, b = b0 + bb;
                rawb[i] = (iy >= 0 && iy < H && b >= 0 && b < rowb) ? __ldg(xb + ((size_t)n * H + iy) * rowb + b) : 0;
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    PHASE_STAMP(2);   // patch staged

    // ---- im2col gather of the 561 stem pixels into five A tiles; k = ci*9 + ky*3 + kx ----
    if (FMT == FSCNN_IN_U8_NHWC) {
        const unsigned char* rawb = reinterpret_cast<const unsigned char*>(sm + oIn);
        const int pl = tid & 127, hi = tid >> 7;
#pragma unroll
        for (int mt = 0; mt < NMT; ++mt) {
            const int m = mt * 128 + pl;
            const int sr = m / SW, sc = m - sr * SW;
            const unsigned char* base = rawb + (2 * sr) * (kRW * 4) + 2 + (2 * sc) * 3;
            const bool ok = m < SPIX;
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
                float v[8];
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    v[t] = 0.f;
                    if (hi == 0) {
                        const int k = kk * 8 + t;
                        if (ok) v[t] = (float)base[((k % 9) / 3) * (kRW * 4) + (k % 3) * 3 + k / 9];
                    } else {
                        const int k = 16 + kk * 8 + t;
                        if (ok && k < 27) v[t] = (float)base[((k % 9) / 3) * (kRW * 4) + (k % 3) * 3 + k / 9];
                    }
                }
                sts128(sA + mt * 8192 + a_tile_off(pl, 2 * hi + kk), packbf(v[0], v[1]), packbf(v[2], v[3]), packbf(v[4], v[5]),
                       packbf(v[6], v[7]));
            }
        }
    } else {
        {
            const int pl = tid & 127, hi = tid >> 7;
    #pragma unroll
            for (int mt = 0; mt < NMT; ++mt) {
                const int m = mt * 128 + pl;
                const int sr = m / SW, sc = m - sr * SW;
                const float* base = In + (2 * sr) * PLD + 2 * sc;   // rows beyond the patch (m >= 561) are never read below
                const bool ok = m < SPIX;
    #pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                    float v[8];
                    if (hi == 0) {
    #pragma unroll
                        for (int t = 0; t < 8; ++t) {
                            const int k = kk * 8 + t;
                            v[t] = ok ? base[((k / 9) * PR + (k % 9) / 3) * PLD + k % 3] : 0.f;
                        }
                    } else {
    #pragma unroll
                        for (int t = 0; t < 8; ++t) {
                            const int k = 16 + kk * 8 + t;
                            v[t] = (ok && k < 27) ? base[((k / 9) * PR + (k % 9) / 3) * PLD + k % 3] : 0.f;
                        }
                    }
                    sts128(sA + mt * 8192 + a_tile_off(pl, 2 * hi + kk), packbf(v[0], v[1]), packbf(v[2], v[3]), packbf(v[4], v[5]),
                           packbf(v[6], v[7]));
                }
            }
        }
    }
    fence_async_proxy();
    __syncthreads();
    PHASE_STAMP(3);   // im2col gathered
    if (tid == 0) {
        tc_fence_after_sync();
        constexpr uint32_t idesc = make_idesc_bf16(128, 32);
#pragma unroll
        for (int mt = 0; mt < NMT; ++mt)
#pragma unroll
            for (int k16 = 0; k16 < 2; ++k16)
                umma_bf16_ss(tmem + TM_STEM + mt * 32, make_smem_desc(sA + mt * 8192 + k16 * 4096, 2048, 128),
                             make_smem_desc(sWs + k16 * 1024, 512, 128), idesc, k16 > 0);
        umma_commit(&bar_stem);
    }
    mbar_wait(&bar_stem, 0);
    tc_fence_after_sync();
    PHASE_STAMP(4);   // stem MMAs done

    // ---- stem epilogue: bias, ReLU, zero outside the stem image (the depthwise conv pads with zeros) -> E ----
    for (int task = warp; task < NMT * 4; task += kThreads / 32) {
        const int mt = task >> 2, q = task & 3;
        const int m = mt * 128 + q * 32 + lane;
        const int sr = m / SW, sc = m - sr * SW;
        const int sy = sy0 + sr, sx = sx0 + sc;
        const bool ok = m < SPIX && sy >= 0 && sy < H1 && sx >= 0 && sx < W1;
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem + ((uint32_t)(q * 32) << 16) + TM_STEM + mt * 32, r);
        tmem_ld_wait();
        if (m < SPIX) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                uint32_t pk[4];
#pragma unroll
                for (int h2 = 0; h2 < 4; ++h2) {
                    const int c = g * 8 + 2 * h2;
                    const float a = ok ? relu(__uint_as_float(r[c]) + Bss[c]) : 0.f;
                    const float b = ok ? relu(__uint_as_float(r[c + 1]) + Bss[c + 1]) : 0.f;
                    pk[h2] = packbf(a, b);
                }
                sts128(sE + m * 64 + ((g ^ ((m >> 1) & 3)) << 4), pk[0], pk[1], pk[2], pk[3]);
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    PHASE_STAMP(5);   // stem epilogue done

    // ---- depthwise 3x3 stride 2 (fp32): strip = (column x, 2-row group, 8-channel chunk) -> A2 (over the dead A tiles) ----
    {
        const int xq = tid & 15, rg = (tid >> 4) & 3, k8 = tid >> 6;
        float acc[2][8];
#pragma unroll
        for (int o = 0; o < 2; ++o)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[o][c] = Bds[k8 * 8 + c];
#pragma unroll
        for (int r = 0; r < 5; ++r) {
            const int sr = 4 * rg + r;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int m = sr * SW + 2 * xq + kx;
                float f[8];
                unpackbf8(lds128(sE + m * 64 + ((k8 ^ ((m >> 1) & 3)) << 4)), f);
#pragma unroll
                for (int o = 0; o < 2; ++o) {
                    const int ky = r - 2 * o;
                    if (ky >= 0 && ky < 3) {
                        const float4 wa = *reinterpret_cast<const float4*>(Wds + (ky * 3 + kx) * 32 + k8 * 8);
                        const float4 wb = *reinterpret_cast<const float4*>(Wds + (ky * 3 + kx) * 32 + k8 * 8 + 4);
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
            const int p = (2 * rg + o) * 16 + xq;
            sts128(sA + a_tile_off(p, k8), packbf(relu(acc[o][0]), relu(acc[o][1])), packbf(relu(acc[o][2]), relu(acc[o][3])),
                   packbf(relu(acc[o][4]), relu(acc[o][5])), packbf(relu(acc[o][6]), relu(acc[o][7])));
        }
    }
    fence_async_proxy();
    __syncthreads();
    PHASE_STAMP(6);   // depthwise done
    if (tid == 0) {
        tc_fence_after_sync();
        constexpr uint32_t idesc = make_idesc_bf16(128, 48);
#pragma unroll
        for (int k16 = 0; k16 < 2; ++k16)
            umma_bf16_ss(tmem + TM_PW, make_smem_desc(sA + k16 * 4096, 2048, 128), make_smem_desc(sWp + k16 * 2 * 768, 768, 128),
                         idesc, k16 > 0);
        umma_commit(&bar_pw);
    }
    mbar_wait(&bar_pw, 0);
    tc_fence_after_sync();
    PHASE_STAMP(7);   // pointwise MMAs done
    {
        const int q = warp & 3, half = warp >> 2;           // 24 channels per warp half
        const int p = q * 32 + lane;
        const int oy = oy0 + (p >> 4), ox = ox0 + (p & 15);
        uint32_t r[24];
#pragma unroll
        for (int c0 = 0; c0 < 24; c0 += 8) tmem_ld_32x32b_x8(tmem + ((uint32_t)(q * 32) << 16) + TM_PW + half * 24 + c0, r + c0);
        tmem_ld_wait();
        if (oy < H2 && ox < W2) {
            bf16* op = out + (((size_t)n * H2 + oy) * W2 + ox) * 48 + half * 24;
#pragma unroll
            for (int c0 = 0; c0 < 24; c0 += 8) {
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = relu(__uint_as_float(r[c0 + i]) + Bps[half * 24 + c0 + i]);
                *reinterpret_cast<uint4*>(op + c0) = make_uint4(packbf(v[0], v[1]), packbf(v[2], v[3]), packbf(v[4], v[5]), packbf(v[6], v[7]));
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    PHASE_STAMP(8);   // output written
    if (warp == 0) tmem_dealloc(tmem, 256);
    PHASE_STAMP(9);
}

cudaError_t launch_l2d_front_tc(const void* x, const StemIn& in, const bf16* ws_img, const float* bs, const DsW& w,
                                const bf16* wp_img, bf16* out, int n, int h, int wd, int h1, int w1, int h2, int w2, cudaStream_t s) {
    static unsigned long long cfg_f32 = 0, cfg_u8 = 0;
    dim3 grid(ceil_div(w2, 16), ceil_div(h2, 8), n);
    if (in.format == FSCNN_IN_U8_NHWC) {
        cudaError_t e = ensure_dyn_smem(l2d_front_kernel<FSCNN_IN_U8_NHWC>, kSmem, cfg_u8);
        if (e != cudaSuccess) return e;
        l2d_front_kernel<FSCNN_IN_U8_NHWC><<<grid, kThreads, kSmem, s>>>(x, in, ws_img, bs, w, wp_img, out, h, wd, h1, w1, h2, w2);
    } else {
        cudaError_t e = ensure_dyn_smem(l2d_front_kernel<FSCNN_IN_F32_NCHW>, kSmem, cfg_f32);
        if (e != cudaSuccess) return e;
        l2d_front_kernel<FSCNN_IN_F32_NCHW><<<grid, kThreads, kSmem, s>>>(x, in, ws_img, bs, w, wp_img, out, h, wd, h1, w1, h2, w2);
    }
    return cudaGetLastError();
}

#ifdef FSCNN_PHASE_TIMING
extern "C" int fscnn_debug_front_phases(long long* out16) {
    return cudaMemcpyFromSymbol(out16, g_front_phase, sizeof(long long) * 16) == cudaSuccess ? 0 : -1;
}
#endif

}  // namespace fscnn
