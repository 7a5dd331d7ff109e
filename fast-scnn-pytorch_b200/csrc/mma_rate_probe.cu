// mma_rate_probe.cu -- cycles per tcgen05.mma (kind::f16, M = 128, K = 16) for the operand layouts of the transposed
// bottleneck kernel: N, the B operand's LBO (the TMA halo tile has LBO = 180 * 16 = 2880 B, not a multiple of 128), and an
// MN-major A operand.  Build: make probes; run on a B200.
#include <cstdio>
#include "umma.cuh"
using namespace fscnn;

__global__ void probe(long long* cycles, int reps, int n, int a_lbo, int a_sbo, int a_mn, int b_lbo, int b_sbo) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) { mbar_init(&mbar, 1); fence_mbar_init(); }
    if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
    fence_async_proxy();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t taddr = tmem_base_s;
    long long t0 = clock64();
    if (tid == 0) {
        const uint32_t idesc = make_idesc_bf16(128, n) | (a_mn ? (1u << 15) : 0u);
        for (int r = 0; r < reps; ++r) {
            uint64_t da = make_smem_desc(smem_u32(smem) + (r & 3) * 4096, a_lbo, a_sbo);
            uint64_t db = make_smem_desc(smem_u32(smem + 64 * 1024) + (r & 3) * 2 * b_lbo, b_lbo, b_sbo);
            umma_bf16_ss(taddr + (r & 1) * 256, da, db, idesc, 1);
        }
        umma_commit(&mbar);
    }
    mbar_wait(&mbar, 0);
    tc_fence_after_sync();
    long long t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(taddr, 512);
}

int main() {
    long long* dC;
    cudaMalloc(&dC, 64);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    struct { int n, a_lbo, a_sbo, a_mn, b_lbo, b_sbo; const char* what; } cfgs[] = {
        {64, 2048, 128, 0, 1024, 128, "N=64  A K-major, B aligned"},
        {128, 2048, 128, 0, 2048, 128, "N=128 A K-major, B aligned"},
        {192, 2048, 128, 0, 3072, 128, "N=192 A K-major, B LBO 3072"},
        {192, 2048, 128, 0, 2880, 128, "N=192 A K-major, B LBO 2880 (halo tile)"},
        {192, 2048, 128, 0, 128, 0, "N=192 A K-major, B ones block (SBO 0)"},
        {256, 2048, 128, 0, 4096, 128, "N=256 A K-major, B aligned"},
        {64, 128, 2048, 1, 1024, 128, "N=64  A MN-major (LBO 128, SBO 2048)"},
        {128, 128, 2048, 1, 2048, 128, "N=128 A MN-major (LBO 128, SBO 2048)"},
        {96, 128, 2048, 1, 1536, 128, "N=96  A MN-major"},
        {64, 2064, 128, 0, 1024, 128, "N=64  A K-major LBO 2064"},
    };
    for (auto& c : cfgs) {
        const int reps = 256;
        probe<<<148, 128, 160 * 1024>>>(dC, reps, c.n, c.a_lbo, c.a_sbo, c.a_mn, c.b_lbo, c.b_sbo);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("probe: CUDA error %s\n", cudaGetErrorString(e)); return 2; }
        long long cyc;
        cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost);
        printf("%-44s: %d MMAs = %lld cyc -> %.1f cyc/MMA\n", c.what, reps, cyc, (double)cyc / reps);
    }
    return 0;
}
