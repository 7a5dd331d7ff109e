// tma_probe.cu -- standalone check of the 5-D tensor map the bf16 kernels use to stage an NHWC halo tile straight
// into the SWIZZLE_NONE core-matrix A-operand layout ([c/8][pixel][8 ch]) with one cp.async.bulk.tensor, including
// negative / out-of-range coordinates (zero fill = the convolution padding).  Not part of the library (`make probes`).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap tmap, uint16_t* out, int bytes, int ix0, int iy0, int n, long long* cyc) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t sbar = (uint32_t)__cvta_generic_to_shared(&bar), sdst = (uint32_t)__cvta_generic_to_shared(sm);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sbar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    long long t0 = 0;
    if (threadIdx.x == 0) {
        t0 = clock64();
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sbar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                     ::"r"(sdst), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(0), "r"(ix0), "r"(iy0), "r"(0), "r"(n), "r"(sbar) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}" ::"r"(sbar) : "memory");
    if (threadIdx.x == 0) *cyc = clock64() - t0;
    for (int i = threadIdx.x; i < bytes / 2; i += blockDim.x) out[i] = reinterpret_cast<uint16_t*>(sm)[i];
}

static int run(EncodeFn enc, int N, int H, int W, int C, int IH, int IW, int ix0, int iy0, int n) {
    std::vector<uint16_t> h((size_t)N * H * W * C);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint16_t)(1 + (i * 2654435761u >> 7) % 60000);   // never zero
    uint16_t* d; cudaMalloc(&d, h.size() * 2);
    cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap tm;
    const cuuint64_t dims[5] = {8, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C / 8, (cuuint64_t)N};
    const cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, 16, (cuuint64_t)H * W * C * 2};
    const cuuint32_t box[5] = {8, (cuuint32_t)IW, (cuuint32_t)IH, (cuuint32_t)C / 8, 1};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed: %d (C=%d)\n", (int)r, C); return 1; }
    const int pin = IH * IW, bytes = pin * C * 2;
    uint16_t* dout; cudaMalloc(&dout, bytes); long long* dc; cudaMalloc(&dc, 8);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    probe<<<1, 128, bytes>>>(tm, dout, bytes, ix0, iy0, n, dc);
    probe<<<1, 128, bytes>>>(tm, dout, bytes, ix0, iy0, n, dc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<uint16_t> o(bytes / 2); long long cyc;
    cudaMemcpy(o.data(), dout, bytes, cudaMemcpyDeviceToHost); cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost);
    long bad = 0;
    for (int k8 = 0; k8 < C / 8; ++k8)
        for (int m = 0; m < pin; ++m)
            for (int c = 0; c < 8; ++c) {
                const int y = iy0 + m / IW, x = ix0 + m % IW;
                const uint16_t want = (y >= 0 && y < H && x >= 0 && x < W) ? h[(((size_t)n * H + y) * W + x) * C + k8 * 8 + c] : 0;
                if (o[((size_t)k8 * pin + m) * 8 + c] != want) ++bad;
            }
    printf("C=%3d box %2dx%2d at (%d,%d) n=%d: %ld mismatches of %d, %lld cycles for %d bytes (%.1f B/clk)\n", C, IH, IW, iy0, ix0, n, bad,
           bytes / 2, cyc, bytes, (double)bytes / cyc);
    cudaFree(d); cudaFree(dout); cudaFree(dc);
    return bad != 0;
}

__global__ void probe3(const __grid_constant__ CUtensorMap tmap, uint8_t* out, int bytes, int c0, int c1, int c2) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t sbar = (uint32_t)__cvta_generic_to_shared(&bar), sdst = (uint32_t)__cvta_generic_to_shared(sm);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sbar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sbar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(sdst), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(c0), "r"(c1), "r"(c2), "r"(sbar) : "memory");
    }
    __syncthreads();
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}" ::"r"(sbar) : "memory");
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = sm[i];
}

static int run3(EncodeFn enc, bool u8, int N, int H, int W, int x0, int y0, int n) {
    const int es = u8 ? 1 : 4, planes = u8 ? 1 : 3, rowe = u8 ? W * 3 : W, bw = u8 ? 224 : 72;
    const size_t total = (size_t)N * 3 * H * W * es;
    std::vector<uint8_t> h(total);
    for (size_t i = 0; i < total; ++i) h[i] = (uint8_t)(1 + (i * 2654435761u >> 9) % 250);
    uint8_t* d; cudaMalloc(&d, total); cudaMemcpy(d, h.data(), total, cudaMemcpyHostToDevice);
    CUtensorMap tm;
    const cuuint64_t dims[3] = {(cuuint64_t)rowe, (cuuint64_t)H, (cuuint64_t)(u8 ? N : N * 3)};
    const cuuint64_t strides[2] = {(cuuint64_t)rowe * es, (cuuint64_t)rowe * es * H};
    const cuuint32_t box[3] = {(cuuint32_t)bw, 35, (cuuint32_t)planes};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&tm, u8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("3-D encode failed: %d (u8=%d)\n", (int)r, (int)u8); return 1; }
    const int bytes = bw * es * 35 * planes;
    uint8_t* dout; cudaMalloc(&dout, bytes);
    cudaFuncSetAttribute(probe3, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    probe3<<<1, 128, bytes>>>(tm, dout, bytes, x0, y0, u8 ? n : n * 3);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("3-D kernel failed (u8=%d): %s\n", (int)u8, cudaGetErrorString(e)); return 1; }
    std::vector<uint8_t> o(bytes);
    cudaMemcpy(o.data(), dout, bytes, cudaMemcpyDeviceToHost);
    long bad = 0;
    for (int p = 0; p < planes; ++p)
        for (int rr = 0; rr < 35; ++rr)
            for (int c = 0; c < bw * es; ++c) {
                const int y = y0 + rr; const long xb = (long)x0 * es + c;
                const uint8_t want = (y >= 0 && y < H && xb >= 0 && xb < (long)rowe * es)
                    ? h[(((size_t)(u8 ? n : n * 3 + p)) * H + y) * rowe * es + xb] : 0;
                if (o[((size_t)p * 35 + rr) * bw * es + c] != want) ++bad;
            }
    printf("3-D %s patch at (%d,%d) n=%d of %dx%d: %ld mismatches of %d bytes\n", u8 ? "uint8" : "fp32", y0, x0, n, H, W, bad, bytes);
    cudaFree(d); cudaFree(dout);
    return bad != 0;
}

int main() {
    EncodeFn enc = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q) != cudaSuccess || !enc) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    int bad = 0;
    bad += run(enc, 2, 64, 128, 64, 10, 18, -1, -1, 1);     // top-left corner tile (padding)
    bad += run(enc, 2, 64, 128, 64, 10, 18, 47, 23, 0);     // interior
    bad += run(enc, 2, 64, 128, 64, 10, 18, 111, 55, 1);    // bottom-right corner
    bad += run(enc, 2, 128, 256, 64, 17, 33, 31, 15, 1);    // stride-2 halo
    bad += run(enc, 2, 32, 64, 96, 10, 18, -1, 7, 0);
    bad += run(enc, 2, 32, 64, 128, 10, 18, 47, -1, 1);
    bad += run(enc, 1, 65, 97, 128, 10, 18, 81, 57, 0);     // odd sizes, partial tile
    // the innermost start coordinate must be a multiple of 16 bytes (x0 = -2 floats is an illegal instruction at run time)
    bad += run3(enc, false, 2, 128, 256, -4, -2, 1);
    bad += run3(enc, false, 2, 128, 256, 188, 94, 0);
    bad += run3(enc, true, 2, 128, 256, -16, -2, 1);
    bad += run3(enc, true, 2, 128, 256, 12 * 48 - 16, 94, 0);
    printf(bad ? "FAILED\n" : "all ok\n");
    return bad;
}
