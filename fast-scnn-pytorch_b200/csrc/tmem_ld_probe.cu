// tmem_ld_probe.cu -- tcgen05.ld (32x32b) throughput per SM: how many bytes per clock the compute warps of the transposed kernels
// (bottleneck_s1t / s2t, ffm_t, l2d_front_t: every depthwise thread pulls its halo rows out of TMEM) can read, by number of warps
// and load width, alone and interleaved with FHFMA-like FMA work.  Build: make probes; run on a B200.
#include <cstdio>
#include "umma.cuh"
using namespace fscnn;

template <int X, bool FMA>
__global__ void __launch_bounds__(512, 1) probe(long long* cycles, float* sink, int reps) {
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t taddr = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 32;   // lane quarter = warp % 4 (hardware rule)
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < reps; ++r) {
        uint32_t v[64];
        if (X == 64) tmem_ld_32x32b_x64(taddr, v);
        else if (X == 32) tmem_ld_32x32b_x32(taddr, reinterpret_cast<uint32_t(&)[32]>(v[0]));
        else tmem_ld_32x32b_x16(taddr, reinterpret_cast<uint32_t(&)[16]>(v[0]));
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < X; ++i) acc[i & 7] += __uint_as_float(v[i]);                 // consume every register
        if (FMA) {                                                                        // ~4 FMAs per loaded value, like the 3x3 taps
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int i = 0; i < X; ++i) acc[i & 7] = fmaf(__uint_as_float(v[i]), 1.0001f + k, acc[(i + k) & 7]);
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (tid == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += acc[i];
    if (s == 123.456f) sink[tid] = s;
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base_s, 512);
}

template <int X, bool FMA>
static void run(long long* dC, float* dS, int warps, const char* what) {
    const int reps = 2000;
    probe<X, FMA><<<148, warps * 32>>>(dC, dS, reps);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("probe: CUDA error %s\n", cudaGetErrorString(e)); exit(2); }
    long long c;
    cudaMemcpy(&c, dC, 8, cudaMemcpyDeviceToHost);
    const double bytes = (double)warps * X * 128.0 * reps;
    printf("%-34s warps %2d  x%-2d  %8.1f cycles / round  %7.1f B/clk/SM\n", what, warps, X, (double)c / reps, bytes / (double)c);
}

int main() {
    long long* dC;
    float* dS;
    cudaMalloc(&dC, 64);
    cudaMalloc(&dS, 4096);
    for (int warps : {1, 4, 8, 16}) {
        run<16, false>(dC, dS, warps, "tcgen05.ld.32x32b + wait");
        run<32, false>(dC, dS, warps, "tcgen05.ld.32x32b + wait");
        run<64, false>(dC, dS, warps, "tcgen05.ld.32x32b + wait");
    }
    for (int warps : {4, 8, 16}) {
        run<32, true>(dC, dS, warps, "ld + wait + 4 FMA per value");
        run<64, true>(dC, dS, warps, "ld + wait + 4 FMA per value");
    }
    return 0;
}
