// ffm_tc.cu -- bf16 FeatureFusionModule (reference models/fast_scnn.py:190-218) on the tensor cores.
// Same algebra as ffm.cu: one K = 192 contraction whose first 64 columns are `higher`'s channels and
// whose last 128 columns are relu(BN(DW3x3(bilinear_align_corners(lower)))), biases pre-added, ReLU.
//
// CTA = 8x16 output pixels (M = 128), 256 threads:
//   A[:, 0:64]   <- cp.async of `higher` straight into the A-operand tile
//   U            <- bilinear resize of `lower` on the 10x18 halo tile (lanes along channels: the four
//                   corner reads are 256-byte coalesced), bf16, swizzled rows
//   A[:, 64:192] <- depthwise 3x3 + bias + ReLU on U (fp32 accumulate)
//   TMEM[128x128] = A * Wcat^T  (12 MMAs), epilogue: + bias, ReLU -> bf16 NHWC
#include "kernels.h"
#include "umma.cuh"

namespace fscnn {

namespace {
constexpr int kIW = 18, kPIN = 180, kPINP = 184, kCL = 128, kCH = 64, kK = 192, kCO = 128;
constexpr int oU = 0;                               // [184][256 B] resized halo tile ...
constexpr int oB = 0;                               // ... later overwritten by the 128 x 192 weight image (48 KB)
constexpr int oA = kCO * kK * 2;                    // 128 x 192 bf16
constexpr int oWd = oA + 128 * kK * 2;              // fp32 [9][128]
constexpr int oBd = oWd + 9 * kCL * 4;
constexpr int oBc = oBd + kCL * 4;
constexpr int oTab = oBc + kCO * 4;                 // per halo pixel: int4 {top-left idx, bottom-left idx, dx, -} or x = -1 outside
constexpr int kSmem = oTab + kPINP * 16;
}  // namespace

__global__ void __launch_bounds__(kThreads, 2)
ffm_tc_kernel(const bf16* __restrict__ higher, const bf16* __restrict__ lower, FfmW w, const bf16* __restrict__ wcat_img,
              bf16* __restrict__ out, int Hh, int Wh, int Hl, int Wl) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar_w, bar_mma;
    __shared__ uint32_t tmem_base_s;
    const uint32_t sWd = smem_u32(sm + oWd);   // depthwise weights [9][128] as bf16 (FHFMA operands)
    float* Bds = reinterpret_cast<float*>(sm + oBd);
    float* Bcs = reinterpret_cast<float*>(sm + oBc);
    int4* tab = reinterpret_cast<int4*>(sm + oTab);
    const uint32_t sU = smem_u32(sm + oU), sA = smem_u32(sm + oA), sB = smem_u32(sm + oB);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n = blockIdx.z, oy0 = blockIdx.y * 8, ox0 = blockIdx.x * 16;

    if (tid == 0) {
        mbar_init(&bar_w, 1); mbar_init(&bar_mma, 1);
        fence_mbar_init();
    }
    if (warp == 0) { tmem_alloc(&tmem_base_s, 128); tmem_relinquish(); }

    // `higher` -> A columns 0..63
    for (int i = tid; i < 128 * (kCH / 8); i += kThreads) {
        const int p = i >> 3, k8 = i & 7;
        const int oy = oy0 + (p >> 4), ox = ox0 + (p & 15);
        const bool ok = oy < Hh && ox < Wh;
        const bf16* src = ok ? higher + (((size_t)n * Hh + oy) * Wh + ox) * kCH + k8 * 8 : higher;
        cp_async16z(sA + a_tile_off(p, k8), src, ok);
    }
    // bilinear source table of the halo tile (align_corners=True, size-driven: reference :209-212)
    const float scy = Hh > 1 ? (float)(Hl - 1) / (float)(Hh - 1) : 0.f;
    const float scx = Wh > 1 ? (float)(Wl - 1) / (float)(Wh - 1) : 0.f;
    for (int pin = tid; pin < kPINP; pin += kThreads) {
        const int y = oy0 - 1 + pin / kIW, x = ox0 - 1 + pin % kIW;
        int4 t = make_int4(-1, 0, 0, 0);
        if (pin < kPIN && y >= 0 && y < Hh && x >= 0 && x < Wh) {
            const float fy = scy * (float)y, fx = scx * (float)x;
            const int y0 = min((int)fy, Hl - 1), x0 = min((int)fx, Wl - 1);
            const int y1 = min(y0 + 1, Hl - 1), x1 = min(x0 + 1, Wl - 1);
            t.x = y0 * Wl + x0;                       // top-left source pixel
            t.y = y1 * Wl + x0;                       // bottom-left source pixel
            t.z = x1 - x0;                            // 0 at the right border, else 1
        }
        tab[pin] = t;
    }
    for (int i = tid; i < 9 * kCL / 2; i += kThreads)
        reinterpret_cast<uint32_t*>(sm + oWd)[i] = packbf(__ldg(w.wd + 2 * i), __ldg(w.wd + 2 * i + 1));
    if (tid < kCL) { Bds[tid] = __ldg(w.bd + tid); Bcs[tid] = __ldg(w.bcat + tid); }
    __syncthreads();

    // U = resize(lower) on the halo tile; item = (halo pixel, 8-channel chunk), lanes along chunks
    for (int i = tid; i < kPINP * 16; i += kThreads) {
        const int pin = i >> 4, k8 = i & 15;
        const int4 t = tab[pin];
        uint4 o = make_uint4(0, 0, 0, 0);
        if (t.x >= 0) {
            const int y = oy0 - 1 + pin / kIW, x = ox0 - 1 + pin % kIW;
            const float fy = scy * (float)y, fx = scx * (float)x;
            const float ly = fy - (float)min((int)fy, Hl - 1), lx = fx - (float)min((int)fx, Wl - 1);
            const float hy = 1.f - ly, hx = 1.f - lx;
            const bf16* base = lower + (size_t)n * Hl * Wl * kCL + k8 * 8;
            float v00[8], v01[8], v10[8], v11[8];
            unpackbf8(__ldg(reinterpret_cast<const uint4*>(base + (size_t)t.x * kCL)), v00);
            unpackbf8(__ldg(reinterpret_cast<const uint4*>(base + (size_t)(t.x + t.z) * kCL)), v01);
            unpackbf8(__ldg(reinterpret_cast<const uint4*>(base + (size_t)t.y * kCL)), v10);
            unpackbf8(__ldg(reinterpret_cast<const uint4*>(base + (size_t)(t.y + t.z) * kCL)), v11);
            float u[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) u[c] = hy * (hx * v00[c] + lx * v01[c]) + ly * (hx * v10[c] + lx * v11[c]);
            o = make_uint4(packbf(u[0], u[1]), packbf(u[2], u[3]), packbf(u[4], u[5]), packbf(u[6], u[7]));
        }
        sts128(sU + pin * (kCL * 2) + ((k8 ^ (pin & 7)) << 4), o.x, o.y, o.z, o.w);
    }
    cp_async_wait_all();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_base_s;

    // depthwise 3x3 + bias + ReLU on U -> A columns 64..191 (one output column x 4 channels per item)
#pragma unroll 1
    for (int it = tid; it < 16 * (kCL / 4); it += kThreads)
        dw3x3_s1_col4<kCL * 2, kIW>(sU, it & 15, it >> 4, sWd, kCL, Bds, sA, 8);
    fence_async_proxy();
    __syncthreads();

    if (tid == 0) {
        mbar_arrive_expect_tx(&bar_w, kCO * kK * 2);     // U is dead: the weight image streams over it
        bulk_g2s(sm + oB, wcat_img, kCO * kK * 2, &bar_w);
        mbar_wait(&bar_w, 0);
        tc_fence_after_sync();
        constexpr uint32_t idesc = make_idesc_bf16(128, kCO);
#pragma unroll
        for (int k16 = 0; k16 < kK / 16; ++k16)
            umma_bf16_ss(tmem, make_smem_desc(sA + k16 * 4096, 2048, 128), make_smem_desc(sB + k16 * 2 * (kCO * 16), kCO * 16, 128),
                         idesc, k16 > 0);
        umma_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, 0);
    tc_fence_after_sync();

    const int q = warp & 3, half = warp >> 2;
    const int p = q * 32 + lane;
    const int oy = oy0 + (p >> 4), ox = ox0 + (p & 15);
    const bool live = oy < Hh && ox < Wh;
    {
        uint32_t r[64];
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 8) tmem_ld_32x32b_x8(tmem + ((uint32_t)(q * 32) << 16) + half * 64 + c0, r + c0);
        tmem_ld_wait();
        if (live) {
            bf16* op = out + (((size_t)n * Hh + oy) * Wh + ox) * kCO + half * 64;
#pragma unroll
            for (int c0 = 0; c0 < 64; c0 += 8) {
                const float4 ba = *reinterpret_cast<const float4*>(Bcs + half * 64 + c0);
                const float4 bb = *reinterpret_cast<const float4*>(Bcs + half * 64 + c0 + 4);
                const uint32_t* q8 = r + c0;
                *reinterpret_cast<uint4*>(op + c0) =
                    make_uint4(packbf_relu(__uint_as_float(q8[0]) + ba.x, __uint_as_float(q8[1]) + ba.y),
                               packbf_relu(__uint_as_float(q8[2]) + ba.z, __uint_as_float(q8[3]) + ba.w),
                               packbf_relu(__uint_as_float(q8[4]) + bb.x, __uint_as_float(q8[5]) + bb.y),
                               packbf_relu(__uint_as_float(q8[6]) + bb.z, __uint_as_float(q8[7]) + bb.w));
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}

cudaError_t launch_ffm_tc(const bf16* higher, const bf16* lower, const FfmW& w, const bf16* wcat_img, bf16* out, int n, int hh,
                          int wh, int hl, int wl, cudaStream_t s) {
    static unsigned long long configured = 0;
    cudaError_t e = ensure_dyn_smem(ffm_tc_kernel, (size_t)kSmem, configured);
    if (e != cudaSuccess) return e;
    dim3 grid(ceil_div(wh, 16), ceil_div(hh, 8), n);
    ffm_tc_kernel<<<grid, kThreads, kSmem, s>>>(higher, lower, w, wcat_img, out, hh, wh, hl, wl);
    return cudaGetLastError();
}

}  // namespace fscnn
