// ppm_tc.cu -- bf16 PyramidPooling output stage (reference models/fast_scnn.py:134-135, :143-144) on the tensor core.
// ppm.cu's third launch computes  out = relu(b + x * Wx + sum_S bilinear_align_corners(z_S))  with an fp32 register-tile
// contraction and 13 interpolation taps per pixel and channel group.  Both terms are matrix products:
//     out[px][co] = relu(b[co] + X[px][0:128] * Wx[0:128][co] + R[px][0:64] * Z[0:64][co])
// R = the interpolation matrix of the 50 pooled bins (S = 1, 2, 3, 6; align_corners=True, zero-padded to 64 columns): it
// depends only on the pixel position, is built once per forward by ppm_fill_r_kernel (bf16, NHWC [h][w][64]) and arrives
// like X through one TMA tensor copy per 8x16-pixel tile, already in the K-major core-matrix A layout.  Z = the per-image
// bin vectors projected through the out conv (ppm_branch_kernel), written as a bf16 B-operand image [64/8][128 co][8].
// One CTA per tile: 12 tcgen05.mma (K = 128 + 64) into a 128-column TMEM accumulator, 4 epilogue warps (+ bias, ReLU,
// bf16, each lane stores its pixel's 256 contiguous bytes).  Several CTAs share an SM, so the loads of one overlap the
// epilogue of another.
#include "kernels.h"
#include "tma_host.h"
#include "umma.cuh"

namespace fscnn {

namespace {
constexpr int kC = 128, kBinsP = 64;
constexpr int X_BYTES = 128 * kC * 2, R_BYTES = 128 * kBinsP * 2, W_BYTES = kC * kC * 2, Z_BYTES = kC * kBinsP * 2;
constexpr int oXs = 0, oRs = oXs + X_BYTES, oWs = oRs + R_BYTES, oZs = oWs + W_BYTES, kSmemP = oZs + Z_BYTES;
constexpr int kPThreads = 160;
__device__ __forceinline__ int pbin_start(int i, int n, int s) { return (i * n) / s; }
}  // namespace

// R[y][x][0:64]: bilinear align_corners weights of the 1 + 4 + 9 + 36 pooled bins at pixel (y, x); columns 50..63 zero
__global__ void ppm_fill_r_kernel(bf16* __restrict__ r, int h, int wd) {
    pdl_launch_dependents();
    pdl_wait();      // r may still be read by the previous forward's output stage
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= h * wd) return;
    const int y = p / wd, x = p % wd;
    float w[kBinsP];
#pragma unroll
    for (int i = 0; i < kBinsP; ++i) w[i] = 0.f;
    w[0] = 1.f;                                                   // S = 1: a broadcast
    const int scales[3] = {2, 3, 6}, base[3] = {1, 5, 14};
#pragma unroll
    for (int si = 0; si < 3; ++si) {
        const int s = scales[si];
        const float sy = h > 1 ? (float)(s - 1) / (float)(h - 1) : 0.f, sx = wd > 1 ? (float)(s - 1) / (float)(wd - 1) : 0.f;
        const float fy = sy * (float)y, fx = sx * (float)x;
        const int y0 = min((int)fy, s - 1), x0 = min((int)fx, s - 1);
        const int y1 = min(y0 + 1, s - 1), x1 = min(x0 + 1, s - 1);
        const float ly = fy - (float)y0, lx = fx - (float)x0, hy = 1.f - ly, hx = 1.f - lx;
        w[base[si] + y0 * s + x0] += hy * hx;
        w[base[si] + y0 * s + x1] += hy * lx;
        w[base[si] + y1 * s + x0] += ly * hx;
        w[base[si] + y1 * s + x1] += ly * lx;
    }
    bf16* o = r + (size_t)p * kBinsP;
#pragma unroll
    for (int i = 0; i < kBinsP; i += 2) *reinterpret_cast<uint32_t*>(o + i) = packbf(w[i], w[i + 1]);
}

__global__ void __launch_bounds__(kPThreads)
ppm_out_tc_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap rmap, const bf16* __restrict__ wx_img,
                  const bf16* __restrict__ z_img, const float* __restrict__ bias, bf16* __restrict__ out, int h, int wd, int tiles_x,
                  int tiles_y) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar_ld, bar_mma;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.x, n = blockIdx.y;
    const int oy0 = (tile / tiles_x) * 8, ox0 = (tile % tiles_x) * 16;
    pdl_launch_dependents();
    if (tid == 0) { mbar_init(&bar_ld, 1); mbar_init(&bar_mma, 1); fence_mbar_init(); }
    if (warp == 0) { tmem_alloc(&tmem_base_s, 128); tmem_relinquish(); }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    pdl_wait();
    if (warp == 4) {
        if (lane == 0) {
            mbar_arrive_expect_tx(&bar_ld, X_BYTES + R_BYTES + W_BYTES + Z_BYTES);
            tma_load_halo(smem_u32(sm + oXs), &xmap, ox0, oy0, n, &bar_ld);
            tma_load_halo(smem_u32(sm + oRs), &rmap, ox0, oy0, 0, &bar_ld);
            bulk_g2s(sm + oWs, wx_img, W_BYTES, &bar_ld);
            bulk_g2s(sm + oZs, z_img + (size_t)n * kC * kBinsP, Z_BYTES, &bar_ld);
            mbar_wait(&bar_ld, 0);
            tc_fence_after_sync();
            constexpr uint32_t idesc = make_idesc_bf16(128, kC);
            const uint64_t dx = make_smem_desc(smem_u32(sm + oXs), 2048, 128), dw = make_smem_desc(smem_u32(sm + oWs), 2048, 128);
            const uint64_t dr = make_smem_desc(smem_u32(sm + oRs), 2048, 128), dz = make_smem_desc(smem_u32(sm + oZs), 2048, 128);
#pragma unroll
            for (int k16 = 0; k16 < kC / 16; ++k16)
                umma_bf16_ss(tmem, dx + (uint64_t)(k16 * ((2 * 2048) >> 4)), dw + (uint64_t)(k16 * ((2 * 2048) >> 4)), idesc, k16 > 0);
#pragma unroll
            for (int k16 = 0; k16 < kBinsP / 16; ++k16)
                umma_bf16_ss(tmem, dr + (uint64_t)(k16 * ((2 * 2048) >> 4)), dz + (uint64_t)(k16 * ((2 * 2048) >> 4)), idesc, 1);
            umma_commit(&bar_mma);
        }
    } else {
        const int p = warp * 32 + lane;
        const int oy = oy0 + (p >> 4), ox = ox0 + (p & 15);
        const bool live = oy < h && ox < wd;
        bf16* o = out + (((size_t)n * h + oy) * wd + ox) * kC;
        mbar_wait(&bar_mma, 0);
        tc_fence_after_sync();
#pragma unroll
        for (int c0 = 0; c0 < kC; c0 += 32) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
            tmem_ld_wait();
            if (live) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 ba = __ldg(reinterpret_cast<const float4*>(bias + c0 + 8 * i));
                    const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + c0 + 8 * i + 4));
                    const uint32_t* q8 = r + 8 * i;
                    *reinterpret_cast<uint4*>(o + c0 + 8 * i) =
                        make_uint4(packbf_relu(__uint_as_float(q8[0]) + ba.x, __uint_as_float(q8[1]) + ba.y),
                                   packbf_relu(__uint_as_float(q8[2]) + ba.z, __uint_as_float(q8[3]) + ba.w),
                                   packbf_relu(__uint_as_float(q8[4]) + bb.x, __uint_as_float(q8[5]) + bb.y),
                                   packbf_relu(__uint_as_float(q8[6]) + bb.z, __uint_as_float(q8[7]) + bb.w));
                }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}

cudaError_t launch_ppm_out_tc(const bf16* in, const bf16* wx_img, const bf16* z_img, const float* bias, bf16* r_img, bf16* out, int n,
                              int h, int wd, cudaStream_t s) {
    static unsigned long long configured = 0;
    cudaError_t e = ensure_dyn_smem(ppm_out_tc_kernel, (size_t)kSmemP, configured);
    if (e != cudaSuccess) return e;
    e = launch_pdl(ppm_fill_r_kernel, ceil_div(h * wd, 128), 128, 0, s, r_img, h, wd);
    if (e != cudaSuccess) return e;
    CUtensorMap xmap, rmap;
    e = make_nhwc_halo_map(&xmap, in, n, h, wd, kC, 8, 16);
    if (e != cudaSuccess) return e;
    e = make_nhwc_halo_map(&rmap, r_img, 1, h, wd, kBinsP, 8, 16);
    if (e != cudaSuccess) return e;
    const int tiles_x = ceil_div(wd, 16), tiles_y = ceil_div(h, 8);
    return launch_pdl(ppm_out_tc_kernel, dim3(tiles_x * tiles_y, n), kPThreads, kSmemP, s, xmap, rmap, wx_img, z_img, bias, out, h, wd, tiles_x, tiles_y);
}

}  // namespace fscnn
