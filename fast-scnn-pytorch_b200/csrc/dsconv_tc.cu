// dsconv_tc.cu -- bf16 _DSConv on the tensor cores: DW 3x3 (stride s, pad 1) + BN + ReLU on the CUDA
// cores (FHFMA.BF16, fp32 accumulate), then the pointwise 1x1 + BN + ReLU as tcgen05.mma with the
// accumulator in TMEM.  Replaces reference models/fast_scnn.py:64-79 for LearningToDownsample.dsconv1/2
// (:154-155) and Classifer.dsconv1/2 (:226-227); with HEAD the classifier's Conv2d(128, nc, 1) (:228-231)
// runs as a second MMA on the activated tile and only fp32 low-resolution logits are written.
//
// Persistent, warp-specialised, software-pipelined over 8x16-pixel output tiles (one CTA per SM):
//   control warp (elect)  : one TMA tensor copy per halo tile (zero-filled padding, [c/8][pixel][8 ch] layout, double
//                           buffered, two tiles ahead), the pointwise weights once, every tcgen05.mma + commit
//   16 compute warps, iteration t:
//       depthwise(t)      : halo[t&1] -> 3x3 + bias + ReLU -> bf16 A[t&1] (A-operand layout); arrive -> MMA(t) is issued
//       epilogue(t-1)     : TMEM[(t-1)&1] -> + bias, ReLU -> bf16 NHWC store        (runs while MMA(t) multiplies)
//     HEAD: epilogue 1 (t-1) writes the activated tile back over A[(t-1)&1] as the head's A operand, the control warp
//           issues the head MMA, epilogue 2 (t-2) stores the logits one iteration later.
//   There is no CTA-wide barrier in the loop: every hand-off is an mbarrier (halo landed / A written / MMA committed).
#include "kernels.h"
#include "tma_host.h"
#include "umma.cuh"

namespace fscnn {

namespace {
constexpr int kDsNT = 512;            // compute threads
constexpr int kDsNTall = kDsNT + 32;  // + control warp
}  // namespace

template <int CIN, int COUT, int STRIDE, bool HEAD>
struct DsTcCfg {
    static constexpr int TH = 8, TW = 16;
    static constexpr int IH = (TH - 1) * STRIDE + 3, IW = (TW - 1) * STRIDE + 3;
    static constexpr int PIN = IH * IW;
    static constexpr int NCK = CIN / 8;                 // 16-byte channel chunks per pixel
    // depthwise work item = (column, RPS-row strip, 8-channel chunk); RPS chosen so that ~512 items exist
    static constexpr int RPS = (NCK >= 16) ? 4 : (NCK >= 6 ? 2 : 1);
    static constexpr int NITEM = 16 * (TH / RPS) * NCK;
    static constexpr int H_BYTES = round_up(PIN * CIN * 2, 128);
    static constexpr int A_BYTES = 128 * CIN * 2;
    static constexpr int B_BYTES = COUT * CIN * 2;
    static constexpr int oH = 0;
    static constexpr int oA = 2 * H_BYTES;
    static constexpr int oB = oA + 2 * A_BYTES;
    static constexpr int oWd = oB + B_BYTES;             // bf16 [9][CIN]
    static constexpr int oBd = oWd + round_up(9 * CIN * 2, 16);
    static constexpr int oBp = oBd + CIN * 4;
    static constexpr int oBias = round_up(oBp + COUT * 4, 128);  // B block [2 k-blocks][COUT][8]: {bias head, remainder, 0 x 6} | zeros
    static constexpr int oOnes = oBias + 2 * COUT * 16;          // A block: rows {1, 1, 0 x 6} + zeros, re-read by every row group (SBO = 0)
    static constexpr int oB2 = oOnes + 256;                      // HEAD: head weight image (ncp16 x COUT), sized at run time;
                                                                 // else: per-warp output staging (16 x 32 pixels x CP channels)
    static constexpr int STAGE_BYTES = HEAD ? 0 : 16 * 32 * (COUT / ((COUT % 32 == 0) ? 4 : COUT / 16)) * 2;
    static constexpr int NPART = (COUT % 32 == 0) ? 4 : COUT / 16;   // epilogue column parts (16 warps = 4 quarters x 4 parts)
    static constexpr int CP = COUT / NPART;                           // columns per part: 32, 16 or 16
    static_assert(CIN % 16 == 0 && COUT % 16 == 0 && CP % 8 == 0 && NPART <= 4, "shape");
    static_assert(!HEAD || COUT <= CIN, "HEAD re-uses the A tile");
};

template <int CIN, int COUT, int STRIDE, bool HEAD>
__global__ void __launch_bounds__(kDsNTall, 1)
dsconv_tc_kernel(const __grid_constant__ CUtensorMap xmap, DsW w, const bf16* __restrict__ wp_img, bf16* __restrict__ out, HeadW head,
                 const bf16* __restrict__ wh_img, int ncp16, float* __restrict__ logits, int Ho, int Wo, int tiles_x, int tiles_y,
                 int ntiles) {
    using C = DsTcCfg<CIN, COUT, STRIDE, HEAD>;
    constexpr int IW = C::IW, RPS = C::RPS, PIN = C::PIN, CP = C::CP;
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar_w, bar_h[2], bar_a[2], bar_mma[2], bar_a2[2], bar_head[2];
    __shared__ uint32_t tmem_base_s;
    float* Bds = reinterpret_cast<float*>(sm + C::oBd);
    float* Bps = reinterpret_cast<float*>(sm + C::oBp);
    const uint32_t sH = smem_u32(sm + C::oH), sA = smem_u32(sm + C::oA), sB = smem_u32(sm + C::oB), sB2 = smem_u32(sm + C::oB2);
    const uint32_t sWd = smem_u32(sm + C::oWd);   // depthwise weights [9][CIN] as bf16 (FHFMA operands)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t tm_cols = HEAD ? 512u : (2 * COUT <= 128 ? 128u : 256u);   // 2 x COUT (+ 2 x ncp16 for the head)
    constexpr uint32_t TM_HEAD = 2 * COUT;
    const int gstep = gridDim.x;
    const int my_tiles = (ntiles - (int)blockIdx.x + gstep - 1) / gstep;
    auto tile_origin = [&](int tile, int& n, int& oy0, int& ox0) {
        const int tx = tile % tiles_x, r = tile / tiles_x;
        n = r / tiles_y; oy0 = (r % tiles_y) * C::TH; ox0 = tx * C::TW;
    };

    pdl_launch_dependents();
    if (tid == 0) {
        mbar_init(&bar_w, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_h[i], 1); mbar_init(&bar_mma[i], 1); mbar_init(&bar_head[i], 1);
            mbar_init(&bar_a[i], kDsNT / 32); mbar_init(&bar_a2[i], kDsNT / 32);
        }
        fence_mbar_init();
    }
    if (warp == 0) { tmem_alloc(&tmem_base_s, tm_cols); tmem_relinquish(); }
    for (int i = tid; i < 9 * CIN / 2; i += kDsNTall)
        reinterpret_cast<uint32_t*>(sm + C::oWd)[i] = packbf(__ldg(w.wd + 2 * i), __ldg(w.wd + 2 * i + 1));
    for (int i = tid; i < CIN; i += kDsNTall) Bds[i] = __ldg(w.bd + i);
    for (int i = tid; i < COUT; i += kDsNTall) Bps[i] = __ldg(w.bp + i);
    // the pointwise bias rides through the tensor core: OUT = ones[128 x 16] * biasblock^T + A * W^T (bf16 head + remainder)
    for (int i = tid; i < COUT; i += kDsNTall) {
        const float b = __ldg(w.bp + i);
        const __nv_bfloat16 bh = __float2bfloat16_rn(b), bl = __float2bfloat16_rn(b - __bfloat162float(bh));
        const uint32_t w0 = (uint32_t)(*reinterpret_cast<const uint16_t*>(&bh)) | ((uint32_t)(*reinterpret_cast<const uint16_t*>(&bl)) << 16);
        *reinterpret_cast<uint4*>(sm + C::oBias + i * 16) = make_uint4(w0, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(sm + C::oBias + COUT * 16 + i * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    if (tid < 16) *reinterpret_cast<uint4*>(sm + C::oOnes + tid * 16) = make_uint4(tid < 8 ? 0x3F803F80u : 0u, 0u, 0u, 0u);
    fence_async_proxy();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_base_s;

    if (warp == kDsNT / 32) {
        // =========================== control warp ===========================
        // (converged warp, asynchronous operations by the lane elect.sync names: no per-lane vote loop around every tcgen05.mma /
        // TMA instruction, see bottleneck_s2t_tc.cu)
        {
            auto load_halo = [&](int lt) {
                int n, oy0, ox0;
                tile_origin(blockIdx.x + lt * gstep, n, oy0, ox0);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&bar_h[lt & 1], PIN * CIN * 2);
                    tma_load_halo(sH + (lt & 1) * C::H_BYTES, &xmap, ox0 * STRIDE - 1, oy0 * STRIDE - 1, n, &bar_h[lt & 1]);
                }
            };
            if (elect_one()) {
                tma_prefetch_desc(&xmap);
                mbar_arrive_expect_tx(&bar_w, C::B_BYTES + (HEAD ? ncp16 * COUT * 2 : 0));
                bulk_g2s(sm + C::oB, wp_img, C::B_BYTES, &bar_w);
                if (HEAD) bulk_g2s(sm + C::oB2, wh_img, ncp16 * COUT * 2, &bar_w);
            }
            __syncwarp();
            pdl_wait();      // the weights are on their way; the halo tiles are the previous stage's output
            load_halo(0);
            if (my_tiles > 1) load_halo(1);
            mbar_wait(&bar_w, 0);
            constexpr uint32_t idesc = make_idesc_bf16(128, COUT);
            const uint32_t idesc2 = make_idesc_bf16(128, HEAD ? ncp16 : 16);
            auto issue_head = [&](int lt) {   // head MMA of tile lt: A2 (over A[lt&1]) x head weights -> TMEM head[lt&1]
                mbar_wait(&bar_a2[lt & 1], (lt >> 1) & 1);
                tc_fence_after_sync();
                if (elect_one()) {
#pragma unroll
                    for (int k16 = 0; k16 < COUT / 16; ++k16)
                        umma_bf16_ss(tmem + TM_HEAD + (lt & 1) * ncp16, make_smem_desc(sA + (lt & 1) * C::A_BYTES + k16 * 4096, 2048, 128),
                                     make_smem_desc(sB2 + k16 * 2 * (ncp16 * 16), ncp16 * 16, 128), idesc2, k16 > 0);
                    umma_commit(&bar_head[lt & 1]);
                }
                __syncwarp();
            };
#pragma unroll 1
            for (int lt = 0; lt < my_tiles; ++lt) {
                mbar_wait(&bar_a[lt & 1], (lt >> 1) & 1);        // depthwise(lt) written (writers fenced the async proxy); halo[lt&1] is dead
                tc_fence_after_sync();
                if (elect_one()) {
                    umma_bf16_ss(tmem + (lt & 1) * COUT, make_smem_desc(smem_u32(sm + C::oOnes), 128, 0),
                                 make_smem_desc(smem_u32(sm + C::oBias), COUT * 16, 128), idesc, 0);   // bias
#pragma unroll
                    for (int k16 = 0; k16 < CIN / 16; ++k16)
                        umma_bf16_ss(tmem + (lt & 1) * COUT, make_smem_desc(sA + (lt & 1) * C::A_BYTES + k16 * 4096, 2048, 128),
                                     make_smem_desc(sB + k16 * 2 * (COUT * 16), COUT * 16, 128), idesc, 1);
                    umma_commit(&bar_mma[lt & 1]);
                }
                __syncwarp();
                if (lt + 2 < my_tiles) load_halo(lt + 2);
                if (HEAD && lt >= 1) issue_head(lt - 1);
            }
            if (HEAD) issue_head(my_tiles - 1);
        }
    } else {
        // =========================== compute warps ===========================
        pdl_wait();
        // epilogue of tile lt: TMEM[lt&1] -> bias, ReLU -> global (or, HEAD, the head's A operand over A[lt&1])
        const int q = warp & 3, part = warp >> 2;
        auto epilogue = [&](int lt) {
            int n, oy0, ox0;
            tile_origin(blockIdx.x + lt * gstep, n, oy0, ox0);
            const int p = q * 32 + lane;
            mbar_wait(&bar_mma[lt & 1], (lt >> 1) & 1);
            tc_fence_after_sync();
            if (part < C::NPART) {
                uint32_t r[CP];
#pragma unroll
                for (int c0 = 0; c0 < CP; c0 += 8)
                    tmem_ld_32x32b_x8(tmem + ((uint32_t)(q * 32) << 16) + (lt & 1) * COUT + part * CP + c0, r + c0);
                tmem_ld_wait();
                // !HEAD: a lane owns one pixel (TMEM lane) and CP channels; storing that directly touches 32 different 128-byte
                // lines per instruction.  The warp stages its 32 x CP block in a private shared-memory slab (XOR-swizzled
                // 16-byte chunks) and writes it back with consecutive lanes on consecutive chunks: 4x fewer lines per store.
                constexpr int NCK8 = CP / 8;                              // 16-byte chunks per pixel in this warp's block: 4 or 2
                const uint32_t slab = sB2 + warp * (32 * CP * 2);
#pragma unroll
                for (int c0 = 0; c0 < CP; c0 += 8) {
                    const int co = part * CP + c0;
                    const uint32_t* q8 = r + c0;      // the bias is already in the accumulator
                    const uint32_t a = packbf_relu(__uint_as_float(q8[0]), __uint_as_float(q8[1]));
                    const uint32_t b = packbf_relu(__uint_as_float(q8[2]), __uint_as_float(q8[3]));
                    const uint32_t c = packbf_relu(__uint_as_float(q8[4]), __uint_as_float(q8[5]));
                    const uint32_t d = packbf_relu(__uint_as_float(q8[6]), __uint_as_float(q8[7]));
                    if (!HEAD) {
                        const int sw = NCK8 == 4 ? (lane >> 1) & 3 : (lane >> 2) & 1;
                        sts128(slab + lane * (CP * 2) + (((c0 >> 3) ^ sw) << 4), a, b, c, d);
                    } else {
                        sts128(sA + (lt & 1) * C::A_BYTES + a_tile_off(p, co >> 3), a, b, c, d);   // its MMA has completed
                    }
                }
                if (!HEAD) {
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < NCK8; ++i) {
                        const int px = lane / NCK8 + i * (32 / NCK8), ch = lane % NCK8;   // pixel within the warp's 32, chunk
                        const int sw = NCK8 == 4 ? (px >> 1) & 3 : (px >> 2) & 1;
                        const uint4 v = lds128(slab + px * (CP * 2) + ((ch ^ sw) << 4));
                        const int pp = q * 32 + px;
                        const int yy = oy0 + (pp >> 4), xx = ox0 + (pp & 15);
                        if (yy < Ho && xx < Wo)
                            *reinterpret_cast<uint4*>(out + (((size_t)n * Ho + yy) * Wo + xx) * COUT + part * CP + ch * 8) = v;
                    }
                    __syncwarp();
                }
            }
            if (HEAD) {
                fence_async_proxy();
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_a2[lt & 1]);
            }
        };
        auto head_out = [&](int lt) {   // logits of tile lt: quarter q rows; the four parts split the class groups of 8
            int n, oy0, ox0;
            tile_origin(blockIdx.x + lt * gstep, n, oy0, ox0);
            const int p = q * 32 + lane;
            const int oy = oy0 + (p >> 4), ox = ox0 + (p & 15);
            const bool live = (oy < Ho) && (ox < Wo);
            for (int c0 = part * 8; c0 < head.ncp; c0 += 32) {
                uint32_t r[8];
                tmem_ld_32x32b_x8(tmem + ((uint32_t)(q * 32) << 16) + TM_HEAD + (lt & 1) * ncp16 + c0, r);
                tmem_ld_wait();
                if (live) {
                    float* lp = logits + (((size_t)n * Ho + oy) * Wo + ox) * head.ncp + c0;
#pragma unroll
                    for (int g = 0; g < 2; ++g)
                        if (c0 + 4 * g < head.ncp)
                            *reinterpret_cast<float4*>(lp + 4 * g) =
                                make_float4(__uint_as_float(r[4 * g]) + __ldg(head.b + c0 + 4 * g), __uint_as_float(r[4 * g + 1]) + __ldg(head.b + c0 + 4 * g + 1),
                                            __uint_as_float(r[4 * g + 2]) + __ldg(head.b + c0 + 4 * g + 2), __uint_as_float(r[4 * g + 3]) + __ldg(head.b + c0 + 4 * g + 3));
                }
            }
        };

#pragma unroll 1
        for (int lt = 0; lt < my_tiles; ++lt) {
            // the head MMA of tile lt-2 read its A operand from A[lt&1]: it must have completed before the depthwise rewrites it
            if (HEAD && lt >= 2) { mbar_wait(&bar_head[lt & 1], ((lt - 2) >> 1) & 1); tc_fence_after_sync(); }
            mbar_wait(&bar_h[lt & 1], (lt >> 1) & 1);
            // ---- depthwise 3x3: item = (column x, RPS-row strip rg, 8-channel chunk k8); lanes run along x ----
            {
                const uint32_t sHb = sH + (lt & 1) * C::H_BYTES, sAb = sA + (lt & 1) * C::A_BYTES;
#pragma unroll 1
                for (int item = tid; item < C::NITEM; item += kDsNT) {
                    const int x = item & 15, rg = (item >> 4) % (C::TH / RPS), k8 = item / (16 * (C::TH / RPS));
                    uint4 wk[9];
#pragma unroll
                    for (int t = 0; t < 9; ++t) wk[t] = lds128(sWd + (t * CIN + k8 * 8) * 2);
                    float acc[RPS][8];
                    {
                        const float4 ba = *reinterpret_cast<const float4*>(Bds + k8 * 8);
                        const float4 bb = *reinterpret_cast<const float4*>(Bds + k8 * 8 + 4);
#pragma unroll
                        for (int o = 0; o < RPS; ++o) {
                            acc[o][0] = ba.x; acc[o][1] = ba.y; acc[o][2] = ba.z; acc[o][3] = ba.w;
                            acc[o][4] = bb.x; acc[o][5] = bb.y; acc[o][6] = bb.z; acc[o][7] = bb.w;
                        }
                    }
                    constexpr int NR = (RPS - 1) * STRIDE + 3;
#pragma unroll
                    for (int r = 0; r < NR; ++r) {
                        const int iy = (RPS * rg) * STRIDE + r;
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) {
                            const int pin = iy * IW + x * STRIDE + kx;
                            const uint4 v = lds128(sHb + (k8 * PIN + pin) * 16);
#pragma unroll
                            for (int o = 0; o < RPS; ++o) {
                                const int ky = r - o * STRIDE;
                                if (ky >= 0 && ky < 3) fhfma8(acc[o], v, wk[ky * 3 + kx]);
                            }
                        }
                    }
#pragma unroll
                    for (int o = 0; o < RPS; ++o) {
                        const int p = (RPS * rg + o) * 16 + x;
                        sts128(sAb + a_tile_off(p, k8), packbf_relu(acc[o][0], acc[o][1]), packbf_relu(acc[o][2], acc[o][3]),
                               packbf_relu(acc[o][4], acc[o][5]), packbf_relu(acc[o][6], acc[o][7]));
                    }
                }
            }
            fence_async_proxy();
            tc_fence_before_sync();      // this warp's TMEM reads of tile lt-2 precede the MMA the arrival releases
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_a[lt & 1]);
            if (lt >= 1) epilogue(lt - 1);
            if (HEAD && lt >= 2) head_out(lt - 2);               // bar_head[(lt-2)&1] was waited for at the top of this iteration
        }
        // drain the pipeline
        epilogue(my_tiles - 1);
        if (HEAD) {
            if (my_tiles >= 2) {
                mbar_wait(&bar_head[(my_tiles - 2) & 1], ((my_tiles - 2) >> 1) & 1);
                tc_fence_after_sync();
                head_out(my_tiles - 2);
            }
            mbar_wait(&bar_head[(my_tiles - 1) & 1], ((my_tiles - 1) >> 1) & 1);
            tc_fence_after_sync();
            head_out(my_tiles - 1);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, tm_cols);
}

template <int CIN, int COUT, int STRIDE, bool HEAD>
static cudaError_t run_ds_tc(const bf16* in, const DsW& w, const bf16* wp_img, bf16* out, const HeadW* head, const bf16* wh_img,
                             float* logits, int n, int hi, int wi, int ho, int wo, cudaStream_t s) {
    using C = DsTcCfg<CIN, COUT, STRIDE, HEAD>;
    HeadW h{};
    int ncp16 = 0;
    if (HEAD) { h = *head; ncp16 = round_up(h.nc, 16); }
    if (HEAD && 2 * COUT + 2 * ncp16 > 512) return cudaErrorInvalidValue;   // TMEM budget (nc <= 128)
    const size_t smem = C::oB2 + (HEAD ? (size_t)ncp16 * COUT * 2 : (size_t)C::STAGE_BYTES);
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    static unsigned long long configured = 0;
    static size_t configured_bytes = 0;
    if (smem > configured_bytes) { configured = 0; configured_bytes = smem; }
    cudaError_t e = ensure_dyn_smem(dsconv_tc_kernel<CIN, COUT, STRIDE, HEAD>, configured_bytes, configured);
    if (e != cudaSuccess) return e;
    CUtensorMap xmap;
    e = make_nhwc_halo_map(&xmap, in, n, hi, wi, CIN, C::IH, C::IW);
    if (e != cudaSuccess) return e;
    const int tiles_x = ceil_div(wo, C::TW), tiles_y = ceil_div(ho, C::TH), ntiles = tiles_x * tiles_y * n;
    const int grid = ntiles < num_sms() ? ntiles : num_sms();
    return launch_pdl(dsconv_tc_kernel<CIN, COUT, STRIDE, HEAD>, grid, kDsNTall, smem, s, xmap, w, wp_img, out, h, wh_img, ncp16, logits, ho, wo,
                      tiles_x, tiles_y, ntiles);
}

cudaError_t launch_dsconv_tc(int cin, int cout, int stride, const bf16* in, const DsW& w, const bf16* wp_img, bf16* out,
                             const HeadW* head, const bf16* wh_img, float* logits, int n, int hi, int wi, int ho, int wo,
                             cudaStream_t s) {
    if (cin == 32 && cout == 48 && stride == 2 && !head) return run_ds_tc<32, 48, 2, false>(in, w, wp_img, out, head, wh_img, logits, n, hi, wi, ho, wo, s);
    if (cin == 48 && cout == 64 && stride == 2 && !head) return run_ds_tc<48, 64, 2, false>(in, w, wp_img, out, head, wh_img, logits, n, hi, wi, ho, wo, s);
    if (cin == 128 && cout == 128 && stride == 1 && !head) return run_ds_tc<128, 128, 1, false>(in, w, wp_img, out, head, wh_img, logits, n, hi, wi, ho, wo, s);
    if (cin == 128 && cout == 128 && stride == 1 && head) return run_ds_tc<128, 128, 1, true>(in, w, wp_img, out, head, wh_img, logits, n, hi, wi, ho, wo, s);
    return cudaErrorInvalidValue;
}

}  // namespace fscnn
