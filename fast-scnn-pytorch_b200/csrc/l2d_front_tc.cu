// l2d_front_tc.cu -- bf16 LearningToDownsample.conv + dsconv1 in ONE persistent kernel (reference
// models/fast_scnn.py:153-154, :157-160): dense 3x3 s2 p0 (3 -> 32) + BN + ReLU, depthwise 3x3 s2 p1 +
// BN + ReLU, pointwise 32 -> 48 + BN + ReLU.  The stem's 32-channel output (the largest tensor of the
// network: 67 MB fp32 / 33 MB bf16 per 1024x2048 image, written once and read once in the unfused plan)
// never reaches HBM; both contractions run on the tensor cores.
//
// One tile = 8x16 dsconv1 output pixels; it needs the 17x33 stem pixels around them (halo of the
// stride-2 depthwise conv), i.e. a 35x67 input patch.  Per tile:
//   raw patch (fp32 NCHW planes, or raw uint8 HWC rows with ToTensor/Normalize folded into the stem weights)
//   -> repacked once as bf16 RGBX pixels (8 bytes each), even and odd patch rows in separate planes of pitch 544 B.
//   The stem's im2col matrix is then never built: with 4-channel pixels a stride-2 step is exactly 16 bytes, the
//   row pitch of a SWIZZLE_NONE core matrix, so row (sr, sc) of the A operand of kernel row ky IS the 32 bytes at
//   plane[(2 sr + ky)] + 16 sc (4 pixels x 4 channels = K 16; the 4th pixel meets zero weights, the 4th channel is the constant 1.0 and
//   carries the folded BN bias through the contraction as a bf16 head + remainder pair).
//   Consecutive rows overlap by half (LBO = 16 B, SBO = 128 B; verified by tc_probe.cu), stem rows are 34 A rows apart.
//   -> 15 MMAs (5 row tiles x 3 kernel rows, K = 16, N = 32) into TMEM -> bias, ReLU, zero outside the stem image
//   -> bf16 E[561][32] (over the dead raw patch) -> depthwise s2 (FHFMA.BF16, fp32 accumulate) -> A2[128 x 32]
//   -> 2 MMAs [128 x 48] -> bias, ReLU -> bf16 NHWC.
// The kernel is PERSISTENT (2 CTAs per SM, each walking tiles blockIdx.x, +gridDim.x, ...): barriers, TMEM
// and weights are set up once, and the next tile's patch is fetched with cp.async into the other half of a
// double buffer while the current tile computes, so the global-load latency is off the critical path.
#include "kernels.h"
#include "tma_host.h"
#include "umma.cuh"

#include "../../include/fscnn_b200.h"

namespace fscnn {

namespace {
constexpr int SH = 17, SW = 33;                        // stem pixels per tile: 17 x 33 = 561
constexpr int SWP = 34, MROWS = SH * SWP;              // A rows: stem row pitch 34 (column 33 is a dummy) -> 578
constexpr int NMT = 5;                                 // MMA row tiles of 128 A rows
constexpr int PR = 35, PC = 67;                        // input patch rows / columns the tile needs
// The raw patch starts 2 pixels left of the first needed column so that its first byte is 16-byte aligned in global
// memory (a TMA requirement): fp32 planes of 69 used columns at pitch 72, uint8 rows of 56 words (224 bytes).
constexpr int PCR = 69, PLD = 72;
constexpr int kRW = 56;
constexpr int kBuf = 35968;                            // one half of the double buffer: raw patch (30240 B) or E (561 x 64 B)
constexpr int kPP = SWP * 16;                          // RGBX plane row pitch: 68 pixels x 8 B = 544 B
constexpr int oBuf = 0;
constexpr int oP2 = 2 * kBuf;                          // RGBX planes: 18 even rows, then 17 odd rows
constexpr int oP2o = oP2 + 18 * kPP;
constexpr int kP2Bytes = round_up(35 * kPP, 128);
constexpr int oA2 = oP2 + kP2Bytes;                    // A2[128 rows][64 B], row-group-major core matrices (8 KB)
constexpr int oWs = oA2 + 8192;                        // stem weight image: 3 kernel rows x (32 x 16) bf16
constexpr int oWp = oWs + 3072;                        // pointwise image 48 x 32 bf16
constexpr int oWd = oWp + 3072;                        // depthwise weights [9][32] bf16 (FHFMA operands)
constexpr int oBs = oWd + 640;                         // stem bias [32]
constexpr int oBd = oBs + 128;                         // dw bias [32]
constexpr int oBp = oBd + 128;                         // pw bias [48]
constexpr int kSmem = oBp + 192;
constexpr int TM_STEM = 0, TM_PW = NMT * 32;           // 160 + 48 columns -> allocate 256
static_assert(kSmem <= 115600, "two CTAs per SM: 2 x (kSmem + 1 KB reserved + static) must fit 228 KB");
// the last row tile reads A rows up to 639: they must stay inside the allocation (their results are never used)
static_assert(oP2o + (NMT * 128 + 1) * 16 <= kSmem, "A over-read");

__device__ __forceinline__ void cp_async4z(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
// E tile: 64 bytes (32 channels) per stem pixel, two pixels per 128-byte line; the 16-byte chunk slot within the line is
// XOR-permuted so that both the epilogue (8 consecutive pixels per quarter-warp) and the stride-2 depthwise (every other
// pixel) touch 8 distinct slots
__device__ __forceinline__ uint32_t e_off(int m, int k8) { return (uint32_t)(m >> 1) * 128 + (((((m & 1) << 2) | k8) ^ ((m >> 1) & 7)) << 4); }
// A operand with row-group-major core matrices: element (row m, 16-byte chunk k8) ; LBO = 128, SBO = 512
__device__ __forceinline__ uint32_t a_rg_off(int m, int k8) { return (uint32_t)(m >> 3) * 512 + k8 * 128 + (m & 7) * 16; }
}  // namespace

#ifdef FSCNN_PHASE_TIMING   // debug build only: clock64 stamps of the 4th tile of CTA 7
__device__ long long g_front_phase[16];
#define FR_STAMP(i) do { if (tid == 0 && blockIdx.x == 7 && it == 3) g_front_phase[i] = clock64(); } while (0)
#else
#define FR_STAMP(i) do { } while (0)
#endif

constexpr int kFrontThreads = kThreads + 32;   // 8 compute warps + the control warp (TMA patch loads, every tcgen05.mma)

template <int FMT>
__global__ void __launch_bounds__(kFrontThreads, 2)
l2d_front_kernel(const __grid_constant__ CUtensorMap xmap, int use_tma, const void* __restrict__ x, StemIn prm, const bf16* __restrict__ ws_img, const float* __restrict__ bs,
                 DsW w, const bf16* __restrict__ wp_img, bf16* __restrict__ out, int H, int W, int H1, int W1, int H2, int W2,
                 int tiles_x, int tiles_y, int ntiles) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar_stem[NMT], bar_pw, bar_patch[2], bar_repack, bar_dw;
    __shared__ uint32_t tmem_base_s;
    float* Bds = reinterpret_cast<float*>(sm + oBd);
    float* Bps = reinterpret_cast<float*>(sm + oBp);
    const uint32_t sBuf = smem_u32(sm + oBuf), sP2 = smem_u32(sm + oP2), sA = smem_u32(sm + oA2), sWs = smem_u32(sm + oWs), sWp = smem_u32(sm + oWp),
                   sWd = smem_u32(sm + oWd);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int i = 0; i < NMT; ++i) mbar_init(&bar_stem[i], 1);
        mbar_init(&bar_pw, 1);
        mbar_init(&bar_patch[0], 1); mbar_init(&bar_patch[1], 1);
        mbar_init(&bar_repack, kThreads / 32); mbar_init(&bar_dw, kThreads / 32);
        fence_mbar_init();
    }
    if (warp == 0) { tmem_alloc(&tmem_base_s, 256); tmem_relinquish(); }
    if (tid < 192) reinterpret_cast<uint4*>(sm + oWs)[tid] = __ldg(reinterpret_cast<const uint4*>(ws_img) + tid);
    for (int i = tid; i < kP2Bytes / 16; i += kFrontThreads) reinterpret_cast<uint4*>(sm + oP2)[i] = make_uint4(0u, 0u, 0u, 0u);   // pad pixels stay finite
    if (tid < 192) reinterpret_cast<uint4*>(sm + oWp)[tid] = __ldg(reinterpret_cast<const uint4*>(wp_img) + tid);
    if (tid < 9 * 32 / 2) reinterpret_cast<uint32_t*>(sm + oWd)[tid] = packbf(__ldg(w.wd + 2 * tid), __ldg(w.wd + 2 * tid + 1));
    if (tid < 32) Bds[tid] = __ldg(w.bd + tid);
    if (tid < 48) Bps[tid] = __ldg(w.bp + tid);

    // fallback when the rows are not 16-byte aligned (no TMA): the compute threads fetch the patch of a tile with 4-byte
    // cp.async into buffer b (no registers, no waiting)
    auto prefetch = [&](int txi, int tyi, int n, int b) {
        const int iy0 = 4 * (tyi * 8) - 2, ix0 = 4 * (txi * 16) - 4;   // ix0: raw patch start (needed columns start at ix0 + 2)
        const uint32_t dst0 = sBuf + b * kBuf;
        if (FMT == FSCNN_IN_F32_NCHW) {
            const float* xf = reinterpret_cast<const float*>(x);
            if (tid < 3 * PCR) {   // thread = (channel, column): 35 rows, constant pointer step
                const int ci = tid / PCR, c = tid - ci * PCR;
                const int ix = ix0 + c;
                uint32_t dst = dst0 + (ci * PR * PLD + c) * 4;
                if (iy0 >= 0 && iy0 + PR <= H && ix0 >= 0 && ix0 + PCR <= W) {   // interior tile (CTA-uniform): no predicates
                    const float* src = xf + ((((size_t)n * 3 + ci) * H) + iy0) * W + ix;
#pragma unroll
                    for (int r = 0; r < PR; ++r) {
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
                        dst += PLD * 4;
                        src += W;
                    }
                } else {
                    const bool cok = ix >= 0 && ix < W;
                    const float* src = xf + (((size_t)n * 3 + ci) * H) * W + (cok ? ix : 0);
#pragma unroll
                    for (int r = 0; r < PR; ++r) {
                        const int iy = iy0 + r;
                        const bool ok = cok && iy >= 0 && iy < H;
                        cp_async4z(dst + r * PLD * 4, ok ? src + (size_t)iy * W : xf, ok);
                    }
                }
            }
        } else {
            // raw uint8 HWC rows as 56 aligned 32-bit words each, starting at byte ix0*3 - 4 (a multiple of 16)
            const unsigned char* xb = reinterpret_cast<const unsigned char*>(x);
            const int rowb = W * 3, b0 = ix0 * 3 - 4;
            if ((W & 3) == 0) {
                if (tid < 4 * kRW) {   // thread = (word column, group of 9 rows)
                    const int rg = tid / kRW, wq = tid - rg * kRW;
                    const int bb = b0 + 4 * wq;
                    if (iy0 >= 0 && iy0 + PR <= H && b0 >= 0 && b0 + 4 * kRW <= rowb) {   // interior tile: no predicates
                        const unsigned char* src = xb + ((size_t)n * H + iy0 + rg * 9) * rowb + bb;
                        uint32_t dst = dst0 + (rg * 9 * kRW + wq) * 4;
#pragma unroll
                        for (int j = 0; j < 9; ++j) {
                            if (rg * 9 + j < PR) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
                            dst += kRW * 4;
                            src += rowb;
                        }
                    } else {
                        const bool cok = bb >= 0 && bb < rowb;
                        const unsigned char* src = xb + (size_t)n * H * rowb + (cok ? bb : 0);
#pragma unroll
                        for (int j = 0; j < 9; ++j) {
                            const int r = rg * 9 + j, iy = iy0 + r;
                            const bool ok = cok && iy >= 0 && iy < H;
                            if (r < PR) cp_async4z(dst0 + (r * kRW + wq) * 4, ok ? src + (size_t)iy * rowb : xb, ok);
                        }
                    }
                }
            } else {   // unaligned row pitch: plain byte loads into the same layout
                unsigned char* rawb = sm + oBuf + b * kBuf;
                for (int i = tid; i < PR * kRW * 4; i += kThreads) {
                    const int r = i / (kRW * 4), bb = i - r * (kRW * 4);
                    const int iy = iy0 + r, gb = b0 + bb;
                    rawb[i] = (iy >= 0 && iy < H && gb >= 0 && gb < rowb) ? __ldg(xb + ((size_t)n * H + iy) * rowb + gb) : 0;
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // tile coordinates advance by gridDim.x tiles per iteration: decomposed once, then carried (no divisions in the loop)
    int t = blockIdx.x;
    const int gstep = gridDim.x;
    const int step_x = gstep % tiles_x, step_y = (gstep / tiles_x) % tiles_y, step_n = gstep / (tiles_x * tiles_y);
    int nx_x = t % tiles_x, nx_y = (t / tiles_x) % tiles_y, nx_n = t / (tiles_x * tiles_y);   // coordinates of the NEXT tile to fetch
    auto advance = [&]() {
        nx_x += step_x; if (nx_x >= tiles_x) { nx_x -= tiles_x; ++nx_y; }
        nx_y += step_y; if (nx_y >= tiles_y) { nx_y -= tiles_y; ++nx_n; }
        nx_n += step_n;
    };
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_base_s;

    if (warp == kThreads / 32) {
        // =========================== control warp ===========================
        if (lane == 0) {
            auto tma_patch = [&](int b) {   // patch of tile (nx_x, nx_y, nx_n) -> buffer b: fp32 planes {72, 35, 3} or uint8 rows {224 bytes, 35}
                const int iy0 = 4 * (nx_y * 8) - 2, ix0 = 4 * (nx_x * 16) - 4;
                if (FMT == FSCNN_IN_F32_NCHW) {
                    mbar_arrive_expect_tx(&bar_patch[b], 3 * PR * PLD * 4);
                    tma_load_3d(sBuf + b * kBuf, &xmap, ix0, iy0, nx_n * 3, &bar_patch[b]);
                } else {
                    mbar_arrive_expect_tx(&bar_patch[b], PR * kRW * 4);
                    tma_load_3d(sBuf + b * kBuf, &xmap, ix0 * 3 - 4, iy0, nx_n, &bar_patch[b]);
                }
                advance();
            };
            int tl = t;                                        // next tile to load
            if (use_tma) {
                tma_prefetch_desc(&xmap);
                if (tl < ntiles) { tma_patch(0); tl += gstep; }
                if (tl < ntiles) { tma_patch(1); tl += gstep; }
            }
            constexpr uint32_t idesc_s = make_idesc_bf16(128, 32), idesc_p = make_idesc_bf16(128, 48);
#pragma unroll 1
            for (int it = 0; t < ntiles; t += gstep, ++it) {
                mbar_wait(&bar_repack, it & 1);               // RGBX planes written (writers fenced the async proxy), TMEM stem region drained
                tc_fence_after_sync();
#pragma unroll
                for (int mt = 0; mt < NMT; ++mt) {
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky)   // kernel row ky reads plane (ky & 1) from its row (ky >> 1) on; K = 16 = 4 pixels x RGBX
                        umma_bf16_ss(tmem + TM_STEM + mt * 32,
                                     make_smem_desc(sP2 + (ky & 1) * (18 * kPP) + (ky >> 1) * kPP + mt * (128 * 16), 16, 128),
                                     make_smem_desc(sWs + ky * 1024, 512, 128), idesc_s, ky > 0);
                    umma_commit(&bar_stem[mt]);   // one barrier per row tile: its epilogue starts while the later tiles still multiply
                }
                mbar_wait(&bar_dw, it & 1);                   // A2 written, E of this tile dead, TMEM pointwise region drained
                tc_fence_after_sync();
#pragma unroll
                for (int k16 = 0; k16 < 2; ++k16)
                    umma_bf16_ss(tmem + TM_PW, make_smem_desc(sA + k16 * 256, 128, 512), make_smem_desc(sWp + k16 * 2 * 768, 768, 128),
                                 idesc_p, k16 > 0);
                umma_commit(&bar_pw);
                if (use_tma && tl < ntiles) { tma_patch(it & 1); tl += gstep; }   // this tile's buffer is free: fetch tile it + 2
            }
        }
    } else {
    // =========================== compute warps ===========================
    if (!use_tma && t < ntiles) prefetch(nx_x, nx_y, nx_n, 0);
#pragma unroll 1
    for (int it = 0; t < ntiles; t += gstep, ++it) {
        const int b = it & 1;
        const int txi = nx_x, tyi = nx_y, n = nx_n;
        advance();
        const int oy0 = tyi * 8, ox0 = txi * 16;
        const int sy0 = 2 * oy0 - 1, sx0 = 2 * ox0 - 1;      // first stem pixel of the halo tile (may be -1)
        const uint32_t sIn = sBuf + b * kBuf;                 // this tile's patch; later its E tile
        FR_STAMP(0);
        if (use_tma) {
            mbar_wait(&bar_patch[b], (it >> 1) & 1);
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            named_bar_sync(2, kThreads);                       // patch visible to every warp; all reads of E(it-1) are done
            if (t + gstep < ntiles) prefetch(nx_x, nx_y, nx_n, b ^ 1);
        }
        FR_STAMP(1);

        FR_STAMP(2);
        // ---- repack the raw patch as bf16 RGBX pixels, even / odd rows in separate planes (this IS the A operand) ----
        {
            const uint8_t* patch = sm + oBuf + b * kBuf;
#pragma unroll 1
            for (int i = tid; i < PR * PC; i += kThreads) {
                const int r = i / PC, px = i - r * PC;
                float v0, v1, v2;
                if (FMT == FSCNN_IN_U8_NHWC) {
                    const unsigned char* q = patch + r * (kRW * 4) + 10 + 3 * px;
                    v0 = (float)q[0]; v1 = (float)q[1]; v2 = (float)q[2];
                } else {
                    const float* q = reinterpret_cast<const float*>(patch) + r * PLD + px + 2;
                    v0 = q[0]; v1 = q[PR * PLD]; v2 = q[2 * PR * PLD];
                }
                sts64(sP2 + (r & 1) * (18 * kPP) + (r >> 1) * kPP + px * 8, packbf(v0, v1), packbf(v2, 1.f));   // X = 1: carries the bias through the MMA
            }
        }
        fence_async_proxy();
        tc_fence_before_sync();      // this warp's TMEM reads of the previous tile precede the MMAs the arrival releases
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_repack);
        FR_STAMP(3);
        FR_STAMP(4);

        // ---- stem epilogue: bias, ReLU, zero outside the stem image (the depthwise conv pads with zeros) -> E (over the patch) ----
        for (int task = warp; task < NMT * 4; task += kThreads / 32) {
            const int mt = task >> 2, q = task & 3;
            mbar_wait(&bar_stem[mt], it & 1);
            tc_fence_after_sync();
            const int ma = mt * 128 + q * 32 + lane;             // A row: stem row pitch 34
            const int sr = ma / SWP, sc = ma - sr * SWP;
            const int m = sr * SW + sc;                           // E row
            const bool row = ma < MROWS && sc < SW;
            const int sy = sy0 + sr, sx = sx0 + sc;
            const bool ok = row && sy >= 0 && sy < H1 && sx >= 0 && sx < W1;
            uint32_t r[32];
            tmem_ld_32x32b_x32(tmem + ((uint32_t)(q * 32) << 16) + TM_STEM + mt * 32, r);
            tmem_ld_wait();
            if (row) {
                if (ok) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const uint32_t* q8 = r + g * 8;      // the bias is already in the accumulator (X-channel trick)
                        sts128(sIn + e_off(m, g),
                               packbf_relu(__uint_as_float(q8[0]), __uint_as_float(q8[1])),
                               packbf_relu(__uint_as_float(q8[2]), __uint_as_float(q8[3])),
                               packbf_relu(__uint_as_float(q8[4]), __uint_as_float(q8[5])),
                               packbf_relu(__uint_as_float(q8[6]), __uint_as_float(q8[7])));
                    }
                } else {
#pragma unroll
                    for (int g = 0; g < 4; ++g) sts128(sIn + e_off(m, g), 0u, 0u, 0u, 0u);
                }
            }
        }
        named_bar_sync(1, kThreads);                          // E complete (and every stem MMA of this tile has been waited for)
        FR_STAMP(5);

        // ---- depthwise 3x3 stride 2 (fp32): strip = (column x, 2-row group, 8-channel chunk) -> A2 ----
        {
            const int xq = tid & 15, rg = (tid >> 4) & 3, k8 = tid >> 6;
            float acc[2][8];
            {
                const float4 ba = *reinterpret_cast<const float4*>(Bds + k8 * 8);
                const float4 bb = *reinterpret_cast<const float4*>(Bds + k8 * 8 + 4);
#pragma unroll
                for (int o = 0; o < 2; ++o) {
                    acc[o][0] = ba.x; acc[o][1] = ba.y; acc[o][2] = ba.z; acc[o][3] = ba.w;
                    acc[o][4] = bb.x; acc[o][5] = bb.y; acc[o][6] = bb.z; acc[o][7] = bb.w;
                }
            }
            uint4 wv[9];
#pragma unroll
            for (int tp = 0; tp < 9; ++tp) wv[tp] = lds128(sWd + (tp * 32 + k8 * 8) * 2);
#pragma unroll
            for (int r = 0; r < 5; ++r) {
                const int sr = 4 * rg + r;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int m = sr * SW + 2 * xq + kx;
                    const uint4 v = lds128(sIn + e_off(m, k8));
#pragma unroll
                    for (int o = 0; o < 2; ++o) {
                        const int ky = r - 2 * o;
                        if (ky >= 0 && ky < 3) fhfma8(acc[o], v, wv[ky * 3 + kx]);
                    }
                }
            }
#pragma unroll
            for (int o = 0; o < 2; ++o) {
                const int p = (2 * rg + o) * 16 + xq;
                sts128(sA + a_rg_off(p, k8), packbf_relu(acc[o][0], acc[o][1]), packbf_relu(acc[o][2], acc[o][3]),
                       packbf_relu(acc[o][4], acc[o][5]), packbf_relu(acc[o][6], acc[o][7]));
            }
        }
        fence_async_proxy();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_dw);
        FR_STAMP(6);
        mbar_wait(&bar_pw, it & 1);
        tc_fence_after_sync();
        FR_STAMP(7);
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
                    const float4 ba = *reinterpret_cast<const float4*>(Bps + half * 24 + c0);
                    const float4 bb = *reinterpret_cast<const float4*>(Bps + half * 24 + c0 + 4);
                    const uint32_t* q8 = r + c0;
                    *reinterpret_cast<uint4*>(op + c0) =
                        make_uint4(packbf_relu(__uint_as_float(q8[0]) + ba.x, __uint_as_float(q8[1]) + ba.y),
                                   packbf_relu(__uint_as_float(q8[2]) + ba.z, __uint_as_float(q8[3]) + ba.w),
                                   packbf_relu(__uint_as_float(q8[4]) + bb.x, __uint_as_float(q8[5]) + bb.y),
                                   packbf_relu(__uint_as_float(q8[6]) + bb.z, __uint_as_float(q8[7]) + bb.w));
                }
            }
        }
        FR_STAMP(8);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    }   // compute warps
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

#ifdef FSCNN_PHASE_TIMING
extern "C" int fscnn_debug_front_phases(long long* out16) {
    return cudaMemcpyFromSymbol(out16, g_front_phase, sizeof(long long) * 16) == cudaSuccess ? 0 : -1;
}
#endif

cudaError_t launch_l2d_front_tc(const void* x, const StemIn& in, const bf16* ws_img, const float* bs, const DsW& w,
                                const bf16* wp_img, bf16* out, int n, int h, int wd, int h1, int w1, int h2, int w2, cudaStream_t s) {
    static unsigned long long cfg_f32 = 0, cfg_u8 = 0;
    static int num_sms = 0;
    if (!num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || num_sms <= 0) num_sms = 148;
    }
    const int tiles_x = ceil_div(w2, 16), tiles_y = ceil_div(h2, 8);
    const long long ntiles_ll = (long long)tiles_x * tiles_y * n;
    if (ntiles_ll > 0x7fffffff) return cudaErrorInvalidValue;
    const int ntiles = (int)ntiles_ll;
    const int grid = ntiles < 2 * num_sms ? ntiles : 2 * num_sms;   // persistent: two CTAs per SM
    // TMA needs 16-byte aligned rows; other shapes keep the per-thread cp.async path
    CUtensorMap xmap{};
    int use_tma = 0;
    if (in.format == FSCNN_IN_U8_NHWC) {
        if ((wd * 3) % 16 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
            const cuuint64_t dims[3] = {(cuuint64_t)wd * 3, (cuuint64_t)h, (cuuint64_t)n};
            const cuuint64_t strides[2] = {(cuuint64_t)wd * 3, (cuuint64_t)h * wd * 3};
            const cuuint32_t box[3] = {kRW * 4, PR, 1};
            use_tma = make_tiled_map(&xmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, x, dims, strides, box) == cudaSuccess;
        }
        cudaError_t e = ensure_dyn_smem(l2d_front_kernel<FSCNN_IN_U8_NHWC>, kSmem, cfg_u8);
        if (e != cudaSuccess) return e;
        l2d_front_kernel<FSCNN_IN_U8_NHWC><<<grid, kFrontThreads, kSmem, s>>>(xmap, use_tma, x, in, ws_img, bs, w, wp_img, out, h, wd, h1, w1, h2, w2,
                                                                          tiles_x, tiles_y, ntiles);
    } else {
        if (wd % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
            const cuuint64_t dims[3] = {(cuuint64_t)wd, (cuuint64_t)h, (cuuint64_t)n * 3};
            const cuuint64_t strides[2] = {(cuuint64_t)wd * 4, (cuuint64_t)h * wd * 4};
            const cuuint32_t box[3] = {PLD, PR, 3};
            use_tma = make_tiled_map(&xmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, x, dims, strides, box) == cudaSuccess;
        }
        cudaError_t e = ensure_dyn_smem(l2d_front_kernel<FSCNN_IN_F32_NCHW>, kSmem, cfg_f32);
        if (e != cudaSuccess) return e;
        l2d_front_kernel<FSCNN_IN_F32_NCHW><<<grid, kFrontThreads, kSmem, s>>>(xmap, use_tma, x, in, ws_img, bs, w, wp_img, out, h, wd, h1, w1, h2, w2,
                                                                           tiles_x, tiles_y, ntiles);
    }
    return cudaGetLastError();
}

}  // namespace fscnn
