// l2d_front_t_tc.cu -- bf16 LearningToDownsample.conv + dsconv1 (reference models/fast_scnn.py:153-154, :157-160) with the
// TRANSPOSED stem: stem channels sit on the TMEM lanes, the stride-2 depthwise runs in registers straight out of TMEM (the
// scheme of bottleneck_s2t_tc.cu), and the stem's 32-channel output touches neither HBM nor shared memory.
//
// The stem has only 32 output channels, so the 128 TMEM lanes hold FOUR sub-tiles: lane = sub-tile * 32 + channel.
// One tile = 8x16 dsconv1 output pixels = 2x2 sub-tiles of 4x8 pixels; sub-tile j needs 9x17 stem pixels (its halo for the
// stride-2 depthwise), i.e. a 19x35 window of the 35x67 input patch.  Per tile:
//   raw patch (fp32 NCHW planes or raw uint8 HWC rows, ONE TMA tensor copy, two tiles ahead)
//   -> repacked as bf16 RGBX pixels into FOUR per-sub-tile plane sets (even / odd input rows, row pitch 288 B = 18 stem pixels)
//   -> 12 tcgen05.mma (sub-tile j x kernel row ky; M = 128, N = 176, K = 16 = 4 pixels x RGBX):
//        A = the stem weights of kernel row ky sitting in rows 32j..32j+31 of an otherwise zero 128-row operand (one 224-row
//            buffer per ky with the weights in the middle, addressed at a start offset of (96 - 32 j) rows),
//        B = the overlapping-window view of sub-tile j's planes: row n = sr * 18 + sc is the 32 bytes at plane[2 sr + ky] + 16 sc
//            (a stride-2 step is 16 B = the row pitch of a core matrix: LBO = 16 B, SBO = 128 B, no im2col),
//      the BN bias rides on the constant X channel, the four products accumulate into ONE accumulator: column = stem pixel of
//      the sub-tile, lane quarter = sub-tile
//   -> depthwise warp (quarter = sub-tile, row s of the sub-tile): tcgen05.ld of 3 stem rows x 18 columns of its channel -> ReLU
//      + bf16 pack -> zero outside the stem image -> 72 FHFMA.BF16 -> 8 outputs -> ONE 16-byte store into the MN-major operand D
//   -> 2 tcgen05.mma  OUT[128 px x 48] = D[128 px x 32] * Wp  -> + bias, ReLU -> bf16 NHWC.
// One persistent CTA per SM: 16 compute warps + a stem controller (patch loads, stem MMAs) + a pointwise controller; planes,
// D and the stem accumulator are double-buffered, the stem of tile t+1 multiplies while tile t is convolved.
// Inputs whose rows are not 16-byte aligned (no TMA) keep the previous kernel (l2d_front_tc.cu).
#include "kernels.h"
#include "tma_host.h"
#include "umma.cuh"

#include "../../include/fscnn_b200.h"

namespace fscnn {

namespace {
constexpr int kCW = 16, kFTThreads = (kCW + 2) * 32;
constexpr int PR = 35, PC = 67;                        // input patch rows / columns the tile needs
constexpr int PLD = 72, kRW = 56;                      // raw patch: fp32 planes of pitch 72 (needed columns start at +2), uint8 rows of 56 words
constexpr int kRaw = 30720;                            // one raw patch buffer (fp32: 3 x 35 x 72 x 4 = 30240 B)
constexpr int kPitch = 288;                            // plane row pitch: 36 input pixels x 8 B = 18 stem pixels x 16 B
constexpr int kSubP = 19 * kPitch;                     // per sub-tile: 10 even rows, then 9 odd rows
constexpr int kOdd = 10 * kPitch;
constexpr int kPlanes = 4 * kSubP + 512;               // + slack: the N padding (rows 162..175) and the second K block over-read
constexpr int NB = 176;                                // MMA N: 9 x 18 = 162 stem pixels, padded
constexpr int kZ = 2 * 224 * 16;                       // A buffer of one kernel row: 2 K blocks x 224 rows x 16 B
constexpr int oRaw = 0, oPl = 2 * kRaw, oZ = oPl + 2 * kPlanes, oD = oZ + 3 * kZ, oWp = oD + 2 * 8192, oBp = oWp + 3072;
constexpr int oPB = oBp + 256;                         // pointwise bias as a B block [2 k-blocks][48][8]: {head, remainder, 0 x 6} | zeros
constexpr int oOne = oPB + 2 * 48 * 16;                // MN-major A block of ones: k = 0, 1 rows of eight 1.0, rest zero; re-read by every pixel block
constexpr int kSmemFT = oOne + 256;
constexpr int TM_PW = 2 * NB;                          // stem accumulators 2 x 176 columns, pointwise accumulator 48
static_assert(kSmemFT <= 227 * 1024 - 256, "shared memory");
}  // namespace

#ifdef FSCNN_PHASE_TIMING   // debug build only: clock64 stamps of tiles 8..15 of CTA 5 (fp32 input kernel)
__device__ long long g_front_t_phase[8 * 16];
#define F_STAMP(cond, tt, slot) do { if (FMT == FSCNN_IN_F32_NCHW && blockIdx.x == 5 && (cond) && (tt) >= 8 && (tt) < 16) g_front_t_phase[((tt) - 8) * 16 + (slot)] = clock64(); } while (0)
extern "C" int fscnn_debug_front_t_phases(long long* out128) {
    return cudaMemcpyFromSymbol(out128, g_front_t_phase, sizeof(long long) * 128) == cudaSuccess ? 0 : -1;
}
#else
#define F_STAMP(cond, tt, slot) do { } while (0)
#endif

template <int FMT>
__global__ void __launch_bounds__(kFTThreads, 1)
l2d_front_t_kernel(const __grid_constant__ CUtensorMap xmap, const bf16* __restrict__ ws_img, DsW w, const bf16* __restrict__ wp_img,
                   bf16* __restrict__ out, int H1, int W1, int H2, int W2, int tiles_x, int tiles_y, int ntiles) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar_patch[2], bar_planes[2], bar_exp[2], bar_tmfree[2], bar_dready[2], bar_proj[2], bar_projfree;
    __shared__ uint32_t tmem_base_s;
    float* Bps = reinterpret_cast<float*>(sm + oBp);
    const uint32_t sRaw = smem_u32(sm + oRaw), sPl = smem_u32(sm + oPl), sZ = smem_u32(sm + oZ), sD = smem_u32(sm + oD), sWp = smem_u32(sm + oWp);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int gstep = gridDim.x;
    const int my_tiles = (ntiles - (int)blockIdx.x + gstep - 1) / gstep;
    auto tile_origin = [&](int lt, int& n, int& oy0, int& ox0) {
        const int tile = blockIdx.x + lt * gstep;
        const int tx = tile % tiles_x, r = tile / tiles_x;
        n = r / tiles_y; oy0 = (r % tiles_y) * 8; ox0 = tx * 16;
    };
    pdl_launch_dependents();

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_patch[i], 1); mbar_init(&bar_exp[i], 1); mbar_init(&bar_proj[i], 1);
            mbar_init(&bar_planes[i], kCW); mbar_init(&bar_tmfree[i], kCW); mbar_init(&bar_dready[i], kCW);
        }
        mbar_init(&bar_projfree, kCW);
        fence_mbar_init();
    }
    if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
    // planes (pad pixels must stay finite) and the three A buffers (zero rows around the weights) start out cleared
    for (int i = tid; i < (2 * kPlanes + 3 * kZ) / 16; i += kFTThreads) reinterpret_cast<uint4*>(sm + oPl)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    // stem weights of kernel row ky (ws_img: [ky][k/8][32 rows][8]) into rows 96..127 of A buffer ky
    if (tid < 3 * 2 * 32) {
        const int ky = tid / 64, kb = (tid >> 5) & 1, n = tid & 31;
        *reinterpret_cast<uint4*>(sm + oZ + ky * kZ + kb * (224 * 16) + (96 + n) * 16) =
            __ldg(reinterpret_cast<const uint4*>(ws_img + ky * 512 + (kb * 4 + (n >> 3)) * 64 + (n & 7) * 8));
    }
    if (tid < 192) reinterpret_cast<uint4*>(sm + oWp)[tid] = __ldg(reinterpret_cast<const uint4*>(wp_img) + tid);
    if (tid < 48) {   // the pointwise bias rides through the tensor core (bf16 head + remainder against a block of ones)
        const float b = __ldg(w.bp + tid);
        Bps[tid] = b;
        const __nv_bfloat16 bh = __float2bfloat16_rn(b), bl = __float2bfloat16_rn(b - __bfloat162float(bh));
        const uint32_t w0 = (uint32_t)(*reinterpret_cast<const uint16_t*>(&bh)) | ((uint32_t)(*reinterpret_cast<const uint16_t*>(&bl)) << 16);
        *reinterpret_cast<uint4*>(sm + oPB + tid * 16) = make_uint4(w0, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(sm + oPB + 48 * 16 + tid * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    if (tid >= 64 && tid < 80) *reinterpret_cast<uint4*>(sm + oOne + (tid - 64) * 16) = (tid - 64) < 2 ? make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u) : make_uint4(0u, 0u, 0u, 0u);
    fence_async_proxy();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    pdl_wait();      // set-up done (weights only); the input patches and the output belong to the stream order

    if (warp == kCW) {
        // =========================== stem controller ===========================
        if (lane == 0) {
            auto load_patch = [&](int lt) {   // patch of tile lt -> buffer lt & 1: fp32 planes {72, 35, 3} or uint8 rows {224 bytes, 35}
                int n, oy0, ox0;
                tile_origin(lt, n, oy0, ox0);
                const int iy0 = 4 * oy0 - 2, ix0 = 4 * ox0 - 4;
                const int b = lt & 1;
                if (FMT == FSCNN_IN_F32_NCHW) {
                    mbar_arrive_expect_tx(&bar_patch[b], 3 * PR * PLD * 4);
                    tma_load_3d(sRaw + b * kRaw, &xmap, ix0, iy0, n * 3, &bar_patch[b]);
                } else {
                    mbar_arrive_expect_tx(&bar_patch[b], PR * kRW * 4);
                    tma_load_3d(sRaw + b * kRaw, &xmap, ix0 * 3 - 4, iy0, n, &bar_patch[b]);
                }
            };
            constexpr uint32_t idesc_s = make_idesc_bf16(128, NB);
            tma_prefetch_desc(&xmap);
            load_patch(0);
            if (my_tiles > 1) load_patch(1);
#pragma unroll 1
            for (int t = 0; t < my_tiles; ++t) {
                F_STAMP(true, t, 0);
                mbar_wait(&bar_planes[t & 1], (t >> 1) & 1);      // planes of tile t written: its raw patch buffer is free
                F_STAMP(true, t, 1);
                if (t + 2 < my_tiles) load_patch(t + 2);
                F_STAMP(true, t, 2);
                if (t >= 2) mbar_wait(&bar_tmfree[t & 1], ((t - 2) >> 1) & 1);
                F_STAMP(true, t, 3);
                tc_fence_after_sync();
                const uint32_t pl = sPl + (t & 1) * kPlanes, dacc = tmem + (t & 1) * NB;
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky)   // kernel row ky reads the even / odd plane from its row (ky >> 1) on
                        umma_bf16_ss(dacc, make_smem_desc(sZ + ky * kZ + (96 - 32 * j) * 16, 224 * 16, 128),
                                     make_smem_desc(pl + j * kSubP + (ky & 1) * kOdd + (ky >> 1) * kPitch, 16, 128), idesc_s, (j | ky) != 0);
                umma_commit(&bar_exp[t & 1]);
                F_STAMP(true, t, 4);
            }
        }
    } else if (warp == kCW + 1) {
        // =========================== pointwise controller ===========================
        if (lane == 0) {
            constexpr uint32_t idesc_p = make_idesc_bf16(128, 48) | (1u << 15);    // A (= D) is MN-major
#pragma unroll 1
            for (int t = 0; t < my_tiles; ++t) {
                if (t > 0) mbar_wait(&bar_projfree, (t - 1) & 1);
                mbar_wait(&bar_dready[t & 1], (t >> 1) & 1);
                tc_fence_after_sync();
                umma_bf16_ss(tmem + TM_PW, make_smem_desc(smem_u32(sm + oOne), 128, 0), make_smem_desc(smem_u32(sm + oPB), 768, 128), idesc_p, 0);   // bias
#pragma unroll
                for (int k16 = 0; k16 < 2; ++k16)
                    umma_bf16_ss(tmem + TM_PW, make_smem_desc(sD + (t & 1) * 8192 + k16 * 256, 128, 512),
                                 make_smem_desc(sWp + k16 * 2 * 768, 768, 128), idesc_p, 1);
                umma_commit(&bar_proj[t & 1]);
            }
        }
    } else {
        // =========================== compute warps ===========================
        const int q = warp & 3, s = warp >> 2;            // lane quarter = sub-tile, output row of the sub-tile / 16-column slice
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        // this thread's depthwise channel (= lane): 9 bf16 taps + fp32 bias, loop invariant
        uint32_t wq[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) wq[i] = packbf(__ldg(w.wd + (2 * i) * 32 + lane), i < 4 ? __ldg(w.wd + (2 * i + 1) * 32 + lane) : 0.f);
        const float bd = __ldg(w.bd + lane);
        // raw patch of tile lt -> the four per-sub-tile plane sets (pixels of the overlaps are written to 2 or 4 of them).
        // thread = (patch column, row phase): 67 columns x 7 phases, rows phase, phase + 7, ... ; every offset that depends on
        // the column only is computed once per kernel
        const int rp_px = tid % PC, rp_ph = tid / PC;
        const bool rp_on = tid < PC * 7;
        const int rp_c0 = rp_px <= 35 ? rp_px * 8 : -1;                               // sub-tile column 0: lc = px
        const int rp_c1 = rp_px >= 32 ? kSubP + (rp_px - 32) * 8 : -1;                // sub-tile column 1: lc = px - 32
        const int rp_src = (FMT == FSCNN_IN_U8_NHWC ? 10 + 3 * rp_px : (rp_px + 2) * 4) + rp_ph * (FMT == FSCNN_IN_U8_NHWC ? kRW * 4 : PLD * 4);
        int rp_r0[5], rp_r1[5];                           // plane-row offsets of this thread's 5 patch rows in sub-tile rows 0 / 1 (-1: not part)
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const int r = rp_ph + 7 * k;
            rp_r0[k] = r <= 18 ? (r & 1) * kOdd + (r >> 1) * kPitch : -1;
            rp_r1[k] = r >= 16 ? 2 * kSubP + (r & 1) * kOdd + ((r - 16) >> 1) * kPitch : -1;
        }
        auto repack = [&](int lt) {
            if (rp_on) {
                const uint8_t* patch = sm + oRaw + (lt & 1) * kRaw + rp_src;
                const uint32_t pl = sPl + (lt & 1) * kPlanes;
                // all fifteen loads first: the stores below are asm volatile with a memory clobber, which the compiler will not
                // move a load across, so a load -> pack -> store loop pays the shared-memory latency once per row (ncu: 28 % of the
                // kernel's stall samples sat on the packs waiting for their loads)
                float v[5][3];
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    if (FMT == FSCNN_IN_U8_NHWC) {
                        const unsigned char* qp = patch + 7 * k * (kRW * 4);
                        v[k][0] = (float)qp[0]; v[k][1] = (float)qp[1]; v[k][2] = (float)qp[2];
                    } else {
                        const float* qp = reinterpret_cast<const float*>(patch + 7 * k * (PLD * 4));
                        v[k][0] = qp[0]; v[k][1] = qp[PR * PLD]; v[k][2] = qp[2 * PR * PLD];
                    }
                }
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    const uint32_t lo = packbf(v[k][0], v[k][1]), hi = packbf(v[k][2], 1.f);   // X = 1: carries the bias through the MMA
                    if (rp_r0[k] >= 0) {
                        if (rp_c0 >= 0) sts64(pl + rp_r0[k] + rp_c0, lo, hi);
                        if (rp_c1 >= 0) sts64(pl + rp_r0[k] + rp_c1, lo, hi);
                    }
                    if (rp_r1[k] >= 0) {
                        if (rp_c0 >= 0) sts64(pl + rp_r1[k] + rp_c0, lo, hi);
                        if (rp_c1 >= 0) sts64(pl + rp_r1[k] + rp_c1, lo, hi);
                    }
                }
            }
        };
        auto epilogue = [&](int lt, int n, int oy0, int ox0) {   // + bias, ReLU -> bf16 NHWC; slices s = 0..2 hold 16 of the 48 channels each
            const int oy = oy0 + 4 * (q >> 1) + (lane >> 3), ox = ox0 + 8 * (q & 1) + (lane & 7);   // accumulator row = sub-tile * 32 + row * 8 + column
            mbar_wait(&bar_proj[lt & 1], (lt >> 1) & 1);
            tc_fence_after_sync();
            uint32_t r[16];
            if (s < 3) {
                tmem_ld_32x32b_x16(tmem + lane_base + TM_PW + s * 16, r);
                tmem_ld_wait();
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_projfree);
            if (s < 3 && oy < H2 && ox < W2) {
                bf16* op = out + (((size_t)n * H2 + oy) * W2 + ox) * 48 + s * 16;
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const uint32_t* q8 = r + 8 * i;      // the bias is already in the accumulator
                    *reinterpret_cast<uint4*>(op + 8 * i) =
                        make_uint4(packbf_relu(__uint_as_float(q8[0]), __uint_as_float(q8[1])), packbf_relu(__uint_as_float(q8[2]), __uint_as_float(q8[3])),
                                   packbf_relu(__uint_as_float(q8[4]), __uint_as_float(q8[5])), packbf_relu(__uint_as_float(q8[6]), __uint_as_float(q8[7])));
                }
            }
        };
        // the planes are written TWO tiles ahead, next to the D stores, so that one proxy fence per tile covers both (the fence
        // costs a MEMBAR that waits for the epilogue's global stores: it sits as far behind them as the loop allows)
        mbar_wait(&bar_patch[0], 0);
        repack(0);
        if (my_tiles > 1) { mbar_wait(&bar_patch[1], 0); repack(1); }
        fence_async_proxy();
        __syncwarp();
        if (lane == 0) { mbar_arrive(&bar_planes[0]); if (my_tiles > 1) mbar_arrive(&bar_planes[1]); }
        // tile coordinates advance by gridDim.x tiles per iteration: decomposed once, then carried (no divisions in the loop)
        const int step_x = gstep % tiles_x, step_y = (gstep / tiles_x) % tiles_y, step_n = gstep / (tiles_x * tiles_y);
        int tx = (int)blockIdx.x % tiles_x, ty = ((int)blockIdx.x / tiles_x) % tiles_y, n = (int)blockIdx.x / (tiles_x * tiles_y);
        int oy0 = 0, ox0 = 0, pn = 0, poy0 = 0, pox0 = 0, cn = 0;
#pragma unroll 1
        for (int t = 0; t < my_tiles; ++t) {
            pn = cn; poy0 = oy0; pox0 = ox0;
            cn = n; oy0 = ty * 8; ox0 = tx * 16;
            tx += step_x; if (tx >= tiles_x) { tx -= tiles_x; ++ty; }
            ty += step_y; if (ty >= tiles_y) { ty -= tiles_y; ++n; }
            n += step_n;
            F_STAMP(tid == 0, t, 5);
            mbar_wait(&bar_exp[t & 1], (t >> 1) & 1);            // stem(t) has completed
            F_STAMP(tid == 0, t, 6);
            tc_fence_after_sync();
            uint32_t Ep[3][9];                                   // stem rows 2s .. 2s+2, column pairs (2i, 2i+1), ReLU'd bf16
            {
                uint32_t r[56];
                const uint32_t t0 = tmem + lane_base + (t & 1) * NB + (2 * s) * 18;
                tmem_ld_32x32b_x32(t0, reinterpret_cast<uint32_t(&)[32]>(r[0]));
                tmem_ld_32x32b_x16(t0 + 32, reinterpret_cast<uint32_t(&)[16]>(r[32]));
                tmem_ld_32x32b_x8(t0 + 48, r + 48);
                tmem_ld_wait();
#pragma unroll
                for (int rr = 0; rr < 3; ++rr)
#pragma unroll
                    for (int i = 0; i < 9; ++i)
                        Ep[rr][i] = packbf_relu(__uint_as_float(r[rr * 18 + 2 * i]), __uint_as_float(r[rr * 18 + 2 * i + 1]));
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_tmfree[t & 1]);      // stem(t+2) may overwrite this accumulator
            F_STAMP(tid == 0, t, 7);
            // zero padding of the depthwise conv: stem pixels outside the stem image (border tiles only)
            const int ix0 = 2 * (ox0 + 8 * (q & 1)) - 1, iy0 = 2 * (oy0 + 4 * (q >> 1)) - 1 + 2 * s;
            if (ix0 < 0 || ix0 + 18 > W1) {
#pragma unroll
                for (int i = 0; i < 9; ++i) {
                    const int xa = ix0 + 2 * i, xb2 = xa + 1;
                    const uint32_t m = ((xa >= 0 && xa < W1) ? 0x0000FFFFu : 0u) | ((xb2 >= 0 && xb2 < W1) ? 0xFFFF0000u : 0u);
#pragma unroll
                    for (int rr = 0; rr < 3; ++rr) Ep[rr][i] &= m;
                }
            }
            if (iy0 < 0 || iy0 + 3 > H1) {
#pragma unroll
                for (int rr = 0; rr < 3; ++rr)
                    if (iy0 + rr < 0 || iy0 + rr >= H1) {
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
            F_STAMP(tid == 0, t, 8);
            if (t + 2 < my_tiles) {       // planes[t&1] are free (stem(t) has completed); the stem of tile t+2 runs during tile t+1
                mbar_wait(&bar_patch[t & 1], ((t + 2) >> 1) & 1);
                F_STAMP(tid == 0, t, 9);
                repack(t + 2);
            }
            F_STAMP(tid == 0, t, 10);
            if (t >= 2) mbar_wait(&bar_proj[t & 1], ((t - 2) >> 1) & 1);   // pointwise(t-2) has completed: D[t&1] is free
            sts128(sD + (t & 1) * 8192 + (q * 4 + s) * 512 + (lane >> 3) * 128 + (lane & 7) * 16, packbf_relu(acc[0], acc[1]),
                   packbf_relu(acc[2], acc[3]), packbf_relu(acc[4], acc[5]), packbf_relu(acc[6], acc[7]));
            fence_async_proxy();
            __syncwarp();
            if (lane == 0) { mbar_arrive(&bar_dready[t & 1]); if (t + 2 < my_tiles) mbar_arrive(&bar_planes[t & 1]); }
            F_STAMP(tid == 0, t, 11);
            if (t >= 1) epilogue(t - 1, pn, poy0, pox0);
            F_STAMP(tid == 0, t, 12);
        }
        epilogue(my_tiles - 1, cn, oy0, ox0);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// returns cudaErrorNotSupported when the input rows are not 16-byte aligned (the caller then uses launch_l2d_front_tc)
cudaError_t launch_l2d_front_t_tc(const void* x, const StemIn& in, const bf16* ws_img, const DsW& w, const bf16* wp_img, bf16* out, int n,
                                  int h, int wd, int h1, int w1, int h2, int w2, cudaStream_t s) {
    static unsigned long long cfg_f32 = 0, cfg_u8 = 0;
    const int tiles_x = ceil_div(w2, 16), tiles_y = ceil_div(h2, 8);
    const long long ntiles_ll = (long long)tiles_x * tiles_y * n;
    if (ntiles_ll > 0x7fffffff) return cudaErrorInvalidValue;
    const int ntiles = (int)ntiles_ll;
    const int grid = ntiles < num_sms() ? ntiles : num_sms();
    CUtensorMap xmap{};
    if ((reinterpret_cast<uintptr_t>(x) & 15) != 0) return cudaErrorNotSupported;
    if (in.format == FSCNN_IN_U8_NHWC) {
        if ((wd * 3) % 16 != 0) return cudaErrorNotSupported;
        const cuuint64_t dims[3] = {(cuuint64_t)wd * 3, (cuuint64_t)h, (cuuint64_t)n};
        const cuuint64_t strides[2] = {(cuuint64_t)wd * 3, (cuuint64_t)h * wd * 3};
        const cuuint32_t box[3] = {kRW * 4, PR, 1};
        if (make_tiled_map(&xmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, x, dims, strides, box) != cudaSuccess) return cudaErrorNotSupported;
        cudaError_t e = ensure_dyn_smem(l2d_front_t_kernel<FSCNN_IN_U8_NHWC>, kSmemFT, cfg_u8);
        if (e != cudaSuccess) return e;
        e = launch_pdl(l2d_front_t_kernel<FSCNN_IN_U8_NHWC>, grid, kFTThreads, kSmemFT, s, xmap, ws_img, w, wp_img, out, h1, w1, h2, w2, tiles_x, tiles_y, ntiles);
        if (e != cudaSuccess) return e;
    } else {
        if (wd % 4 != 0) return cudaErrorNotSupported;
        const cuuint64_t dims[3] = {(cuuint64_t)wd, (cuuint64_t)h, (cuuint64_t)n * 3};
        const cuuint64_t strides[2] = {(cuuint64_t)wd * 4, (cuuint64_t)h * wd * 4};
        const cuuint32_t box[3] = {PLD, PR, 3};
        if (make_tiled_map(&xmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, x, dims, strides, box) != cudaSuccess) return cudaErrorNotSupported;
        cudaError_t e = ensure_dyn_smem(l2d_front_t_kernel<FSCNN_IN_F32_NCHW>, kSmemFT, cfg_f32);
        if (e != cudaSuccess) return e;
        e = launch_pdl(l2d_front_t_kernel<FSCNN_IN_F32_NCHW>, grid, kFTThreads, kSmemFT, s, xmap, ws_img, w, wp_img, out, h1, w1, h2, w2, tiles_x, tiles_y, ntiles);
        if (e != cudaSuccess) return e;
    }
    return cudaGetLastError();
}

}  // namespace fscnn
