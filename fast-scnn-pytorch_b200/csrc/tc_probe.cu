// tc_probe.cu -- standalone bring-up probe for the tcgen05 building blocks used by the bf16 kernels
// (not part of libfscnn_b200.so):  nvcc -gencode arch=compute_100a,code=sm_100a -O3 tc_probe.cu -o tc_probe
//   1. correctness of tcgen05.mma (kind::f16, bf16 in, fp32 accumulate in TMEM) with SWIZZLE_NONE K-major
//      shared-memory descriptors for two core-matrix orders, and of the 32x32b TMEM load mapping;
//   2. rough per-SM throughput of MMA issue and of tcgen05.ld, to size the fused kernels' phases.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "umma.cuh"

using namespace fscnn;

constexpr int M = 128, K = 64, N = 64;

__global__ void probe_gemm(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D, int variant, int swap_fields) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    uint8_t* sa = smem;
    uint8_t* sb = smem + M * K * 2;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // core-matrix (8 rows x 16 bytes) orders: variant 0 = [row/8][k/8], variant 1 = [k/8][row/8]
    const uint32_t lbo_a = variant == 0 ? 128 : (M / 8) * 128, sbo_a = variant == 0 ? (K / 8) * 128 : 128;
    const uint32_t lbo_b = variant == 0 ? 128 : (N / 8) * 128, sbo_b = variant == 0 ? (K / 8) * 128 : 128;
    for (int i = tid; i < M * (K / 8); i += blockDim.x) {
        const int m = i / (K / 8), k8 = i % (K / 8);
        *reinterpret_cast<uint4*>(sa + (m / 8) * sbo_a + k8 * lbo_a + (m % 8) * 16) =
            *reinterpret_cast<const uint4*>(A + m * K + k8 * 8);
    }
    for (int i = tid; i < N * (K / 8); i += blockDim.x) {
        const int n = i / (K / 8), k8 = i % (K / 8);
        *reinterpret_cast<uint4*>(sb + (n / 8) * sbo_b + k8 * lbo_b + (n % 8) * 16) =
            *reinterpret_cast<const uint4*>(B + n * K + k8 * 8);
    }
    if (tid == 0) { mbar_init(&mbar, 1); fence_mbar_init(); }
    if (warp == 0) { tmem_alloc(&tmem_base_s, 64); tmem_relinquish(); }
    fence_async_proxy();            // generic-proxy smem writes -> visible to the tensor core (async proxy)
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t taddr = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = make_idesc_bf16(M, N);
        for (int k16 = 0; k16 < K / 16; ++k16) {
            uint64_t da = make_smem_desc(smem_u32(sa) + k16 * 2 * lbo_a, swap_fields ? sbo_a : lbo_a, swap_fields ? lbo_a : sbo_a);
            uint64_t db = make_smem_desc(smem_u32(sb) + k16 * 2 * lbo_b, swap_fields ? sbo_b : lbo_b, swap_fields ? lbo_b : sbo_b);
            umma_bf16_ss(taddr, da, db, idesc, k16 > 0);
        }
        umma_commit(&mbar);
    }
    mbar_wait(&mbar, 0);
    tc_fence_after_sync();
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr + ((uint32_t)(warp * 32) << 16) + c0, r);
        tmem_ld_wait();
        for (int i = 0; i < 32; ++i) D[(warp * 32 + lane) * N + c0 + i] = __uint_as_float(r[i]);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(taddr, 64);
}

// Layout probe for the depthwise-on-tensor-core trick: A rows are 16-byte pieces of a [k8][pin][16B] array; the
// 8-row groups of the MMA tile start every `grp` pins (SBO = grp*16 bytes, NOT a multiple of 128), the tile starts at an
// arbitrary 16-byte offset `pin0`, chunks along K are LBO = npin*16 bytes apart; N = 16.
__global__ void probe_strided(const __nv_bfloat16* A /*[128][K]*/, const __nv_bfloat16* B /*[16][K]*/, float* D, int grp, int pin0, int npin) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t* sa = smem;
    uint8_t* sb = smem + (K / 8) * npin * 16;
    for (int i = tid; i < (K / 8) * npin; i += blockDim.x) reinterpret_cast<uint4*>(sa)[i] = make_uint4(0x7fc07fc0u, 0x7fc07fc0u, 0x7fc07fc0u, 0x7fc07fc0u);  // NaN filler
    __syncthreads();
    for (int i = tid; i < M * (K / 8); i += blockDim.x) {
        const int m = i / (K / 8), k8 = i % (K / 8);
        const int pin = pin0 + (m / 8) * grp + (m % 8);
        *reinterpret_cast<uint4*>(sa + (k8 * npin + pin) * 16) = *reinterpret_cast<const uint4*>(A + m * K + k8 * 8);
    }
    for (int i = tid; i < 16 * (K / 8); i += blockDim.x) {
        const int n = i / (K / 8), k8 = i % (K / 8);
        *reinterpret_cast<uint4*>(sb + (k8 * 2 + n / 8) * 128 + (n % 8) * 16) = *reinterpret_cast<const uint4*>(B + n * K + k8 * 8);
    }
    if (tid == 0) { mbar_init(&mbar, 1); fence_mbar_init(); }
    if (warp == 0) { tmem_alloc(&tmem_base_s, 32); tmem_relinquish(); }
    fence_async_proxy();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t taddr = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = make_idesc_bf16(M, 16);
        for (int k16 = 0; k16 < K / 16; ++k16)
            umma_bf16_ss(taddr, make_smem_desc(smem_u32(sa) + pin0 * 16 + k16 * 2 * npin * 16, npin * 16, grp * 16),
                         make_smem_desc(smem_u32(sb) + k16 * 2 * 256, 256, 128), idesc, k16 > 0);
        umma_commit(&mbar);
    }
    mbar_wait(&mbar, 0);
    tc_fence_after_sync();
    uint32_t r[16];
    tmem_ld_32x32b_x16(taddr + ((uint32_t)(warp * 32) << 16), r);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) D[(warp * 32 + lane) * 16 + i] = __uint_as_float(r[i]);
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(taddr, 32);
}

// throughput: `reps` x (M=128, N=256, K=64) MMAs per CTA on garbage data, then `reps` x32 TMEM loads per warp
// Overlapping-window A operand: row m of the 128 x 16 tile is the 32 bytes at S + 16*m of a linear bf16 array, i.e.
// the second core matrix along K starts 16 bytes after the first (LBO = 16 B, SBO = 128 B) and consecutive rows
// overlap by half.  This is what lets a stride-2 3x3 convolution over 4-channel bf16 pixels read its im2col rows
// straight out of the input patch (l2d_front_tc.cu).
__global__ void probe_overlap(const __nv_bfloat16* S /*[8*130]*/, const __nv_bfloat16* B /*[16][16]*/, float* D) {
    __shared__ __align__(128) uint8_t sa[130 * 16];
    __shared__ __align__(128) uint8_t sb[16 * 16 * 2];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 130; i += blockDim.x) reinterpret_cast<uint4*>(sa)[i] = reinterpret_cast<const uint4*>(S)[i];
    for (int i = tid; i < 16 * 2; i += blockDim.x) {
        const int n = i / 2, k8 = i % 2;
        *reinterpret_cast<uint4*>(sb + (k8 * 2 + n / 8) * 128 + (n % 8) * 16) = *reinterpret_cast<const uint4*>(B + n * 16 + k8 * 8);
    }
    if (tid == 0) { mbar_init(&mbar, 1); fence_mbar_init(); }
    if (warp == 0) { tmem_alloc(&tmem_base_s, 32); tmem_relinquish(); }
    fence_async_proxy();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t taddr = tmem_base_s;
    if (tid == 0) {
        umma_bf16_ss(taddr, make_smem_desc(smem_u32(sa), 16, 128), make_smem_desc(smem_u32(sb), 256, 128), make_idesc_bf16(128, 16), 0);
        umma_commit(&mbar);
    }
    mbar_wait(&mbar, 0);
    tc_fence_after_sync();
    uint32_t r[16];
    tmem_ld_32x32b_x16(taddr + ((uint32_t)(warp * 32) << 16), r);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) D[(warp * 32 + lane) * 16 + i] = __uint_as_float(r[i]);
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(taddr, 32);
}

__global__ void probe_rate(long long* cycles, int reps) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (128 * 64 * 2 + 256 * 64 * 2) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) { mbar_init(&mbar, 1); fence_mbar_init(); }
    if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
    fence_async_proxy();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t taddr = tmem_base_s;
    long long t0 = clock64();
    if (tid == 0) {
        const uint32_t idesc = make_idesc_bf16(128, 256);
        for (int r = 0; r < reps; ++r)
            for (int k16 = 0; k16 < 4; ++k16) {
                uint64_t da = make_smem_desc(smem_u32(smem) + k16 * 256, 128, 1024);
                uint64_t db = make_smem_desc(smem_u32(smem + 128 * 64 * 2) + k16 * 256, 128, 1024);
                umma_bf16_ss(taddr + (r & 1) * 256, da, db, idesc, 1);
            }
        umma_commit(&mbar);
    }
    mbar_wait(&mbar, 0);
    tc_fence_after_sync();
    long long t1 = clock64();
    uint32_t acc = 0;
    for (int r = 0; r < reps; ++r) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + ((uint32_t)((warp & 3) * 32) << 16) + ((r * 32) & 511), v);
        tmem_ld_wait();
        for (int i = 0; i < 32; ++i) acc ^= v[i];
    }
    long long t2 = clock64();
    if (tid == 0 && blockIdx.x == 0) { cycles[0] = t1 - t0; cycles[1] = t2 - t1; cycles[2] = acc; }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(taddr, 512);
}

// cycles for `reps` MMAs of shape 128 x N x 16 with the given A descriptor strides (start offset / SBO in bytes)
__global__ void probe_small(long long* cycles, int reps, int n, int a_off, int a_sbo, int a_lbo) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 48 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) { mbar_init(&mbar, 1); fence_mbar_init(); }
    if (warp == 0) { tmem_alloc(&tmem_base_s, 256); tmem_relinquish(); }
    fence_async_proxy();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t taddr = tmem_base_s;
    long long t0 = clock64();
    if (tid == 0) {
        const uint32_t idesc = make_idesc_bf16(128, n);
        for (int r = 0; r < reps; ++r) {
            uint64_t da = make_smem_desc(smem_u32(smem) + a_off + (r % 9) * 16, a_lbo, a_sbo);
            uint64_t db = make_smem_desc(smem_u32(smem + 40 * 1024) + (r & 3) * 512, n * 16, 128);
            umma_bf16_ss(taddr + (r & 3) * 16, da, db, idesc, 1);
        }
        umma_commit(&mbar);
    }
    mbar_wait(&mbar, 0);
    tc_fence_after_sync();
    long long t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(taddr, 256);
}

int main() {
    std::vector<__nv_bfloat16> hA(M * K), hB(N * K);
    std::vector<float> fA(M * K), fB(N * K), ref(M * N), got(M * N);
    srand(1);
    for (int i = 0; i < M * K; ++i) { hA[i] = __float2bfloat16((rand() % 200 - 100) / 64.f); fA[i] = __bfloat162float(hA[i]); }
    for (int i = 0; i < N * K; ++i) { hB[i] = __float2bfloat16((rand() % 200 - 100) / 64.f); fB[i] = __bfloat162float(hB[i]); }
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            float s = 0;
            for (int k = 0; k < K; ++k) s += fA[m * K + k] * fB[n * K + k];
            ref[m * N + n] = s;
        }
    __nv_bfloat16 *dA, *dB;
    float* dD;
    long long* dC;
    cudaMalloc(&dA, M * K * 2); cudaMalloc(&dB, N * K * 2); cudaMalloc(&dD, M * N * 4); cudaMalloc(&dC, 64);
    cudaMemcpy(dA, hA.data(), M * K * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), N * K * 2, cudaMemcpyHostToDevice);
    int bad = 0;
    for (int variant = 0; variant < 2; ++variant)
        for (int swap = 0; swap < 1; ++swap) {   // swapped fields were tried once during bring-up: wrong results / faults
            cudaMemset(dD, 0, M * N * 4);
            probe_gemm<<<1, 128, (M + N) * K * 2>>>(dA, dB, dD, variant, swap);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("variant %d swap %d: CUDA error %s\n", variant, swap, cudaGetErrorString(e)); return 2; }
            cudaMemcpy(got.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
            double err = 0;
            for (int i = 0; i < M * N; ++i) err = fmax(err, fabs(got[i] - ref[i]));
            printf("gemm M=%d N=%d K=%d  core-matrix order %s  desc fields %s : max abs err %.4g  (ref absmax ~%.1f)\n", M, N, K,
                   variant ? "[k/8][row/8]" : "[row/8][k/8]", swap ? "SWAPPED" : "lbo=K-step,sbo=row-step", err, 40.0);
            if (!swap && err > 1e-3) bad = 1;
        }
    {   // strided-group layout: groups every 10 pins (SBO = 160 B), tile start at pin 13 (208 B), 180 pins per chunk (LBO = 2880 B)
        const int grp = 10, pin0 = 13, npin = 16 * grp + pin0 + 8;
        std::vector<float> ref16(M * 16), got16(M * 16);
        for (int m = 0; m < M; ++m)
            for (int n2 = 0; n2 < 16; ++n2) {
                float s2 = 0;
                for (int k = 0; k < K; ++k) s2 += fA[m * K + k] * fB[n2 * K + k];
                ref16[m * 16 + n2] = s2;
            }
        cudaMemset(dD, 0, M * N * 4);
        probe_strided<<<1, 128, (K / 8) * npin * 16 + 16 * K * 2>>>(dA, dB, dD, grp, pin0, npin);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("strided probe: CUDA error %s\n", cudaGetErrorString(e)); return 2; }
        cudaMemcpy(got16.data(), dD, M * 16 * 4, cudaMemcpyDeviceToHost);
        double err = 0;
        for (int i = 0; i < M * 16; ++i) err = fmax(err, fabs(got16[i] - ref16[i]));
        printf("strided A (SBO=160B, start +208B, LBO=%dB), N=16: max abs err %.4g\n", npin * 16, err);
        if (!(err < 1e-3)) bad = 1;
    }
    {   // overlapping windows: A[m][k] = S[8*m + k]
        std::vector<float> ref16(M * 16), got16(M * 16);
        for (int m = 0; m < 128; ++m)
            for (int n2 = 0; n2 < 16; ++n2) {
                float s2 = 0;
                for (int k = 0; k < 16; ++k) s2 += fA[8 * m + k] * fB[n2 * 16 + k];
                ref16[m * 16 + n2] = s2;
            }
        cudaMemset(dD, 0, M * N * 4);
        probe_overlap<<<1, 128>>>(dA, dB, dD);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("overlap probe: CUDA error %s\n", cudaGetErrorString(e)); return 2; }
        cudaMemcpy(got16.data(), dD, M * 16 * 4, cudaMemcpyDeviceToHost);
        double err = 0;
        for (int i = 0; i < 128 * 16; ++i) err = fmax(err, fabs(got16[i] - ref16[i]));
        printf("overlapping-window A (LBO=16B, SBO=128B), N=16, K=16: max abs err %.4g\n", err);
        if (!(err < 1e-3)) bad = 1;
    }
    for (int threads : {128, 256}) {
        cudaFuncSetAttribute(probe_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        const int reps = 64;
        probe_rate<<<148, threads, 64 * 1024>>>(dC, reps);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("rate probe: CUDA error %s\n", cudaGetErrorString(e)); return 2; }
        long long c[3];
        cudaMemcpy(c, dC, 24, cudaMemcpyDeviceToHost);
        printf("rate (%d thr/CTA, 148 CTAs): %d x [128x256x64] MMA = %lld cyc (%.0f MAC/clk/SM); %d x ld.32x32b.x32 per warp = %lld cyc (%.1f B/clk/SM)\n",
               threads, reps, c[0], 128.0 * 256 * 64 * reps / c[0], reps, c[1], (double)reps * 4096 * (threads / 32) / c[1]);
    }
    {
        cudaFuncSetAttribute(probe_small, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
        struct { int n, off, sbo, lbo; const char* what; } cfgs[] = {
            {16, 0, 128, 2048, "N=16 aligned A (SBO 128)"}, {16, 0, 160, 2944, "N=16 strided A (SBO 160, 16B-step starts)"},
            {64, 0, 128, 2048, "N=64 aligned A"}, {64, 0, 160, 2944, "N=64 strided A (SBO 160)"}, {128, 0, 128, 2048, "N=128 aligned A"}};
        for (auto& c : cfgs) {
            const int reps = 288;
            probe_small<<<148, 128, 48 * 1024>>>(dC, reps, c.n, c.off, c.sbo, c.lbo);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("small-MMA probe: CUDA error %s\n", cudaGetErrorString(e)); return 2; }
            long long cyc;
            cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost);
            printf("%-44s: %d MMAs (128 x N x 16) = %lld cyc -> %.1f cyc/MMA\n", c.what, reps, cyc, (double)cyc / reps);
        }
    }
    printf(bad ? "PROBE FAILED\n" : "PROBE OK\n");
    return bad;
}
