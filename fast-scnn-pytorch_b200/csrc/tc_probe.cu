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

// throughput: `reps` x (M=128, N=256, K=64) MMAs per CTA on garbage data, then `reps` x32 TMEM loads per warp
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
    printf(bad ? "PROBE FAILED\n" : "PROBE OK\n");
    return bad;
}
