// fma2_probe.cu -- standalone issue-rate probe for the CUDA-core instructions the bf16 kernels are made of
// (scalar FFMA, packed FFMA2, mixed-precision FHFMA.BF16, F2FP pack, FMNMX, LOP3/SHF/PRMT, IMAD, LDS.128), alone and
// mixed.  Rates come from CUDA-event time at the maximum SM clock (warps are issued by priority, so one CTA's
// clock64 span says nothing).  Not part of the library; built by `make probes`.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

enum { FFMA, FFMA2, FHFMA, F2FP, F2FPRELU, FMNMX, XOR, SHL, PRMT, IMAD, IADD, LDS128, FFMA_XOR, FHFMA_XOR, FFMA_LDS, FHFMA_LDS, HFMA2BF, NMODES };
static const char* kNames[NMODES] = {"FFMA", "FFMA2 (f32x2)", "FHFMA.BF16 (f32 += bf16*bf16)", "F2FP.BF16.PACK_AB", "F2FP.RELU.BF16.PACK_AB",
    "FMNMX", "LOP3 (xor)", "SHL/IMAD.SHL", "PRMT", "IMAD", "IADD3", "LDS.128", "FFMA + LOP3 (1:1)", "FHFMA + LOP3 (1:1)",
    "2 FFMA + LDS.128 (16:8)", "2 FHFMA + LDS.128 (16:8)", "HFMA2.BF16"};

template <int MODE>
__global__ void __launch_bounds__(256) probe(float* out, int iters, float seed) {
    __shared__ uint4 sh[256];
    float a[16];
    uint64_t p[16];
    uint32_t u[16];
    const float x = seed + threadIdx.x * 1e-9f, y = seed * 0.5f;
    uint64_t x2, y2;
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(x2) : "f"(x), "f"(y));
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(y2) : "f"(y), "f"(x));
    const uint32_t xb = __float_as_uint(x), yb = __float_as_uint(y);
    const uint16_t xh = (uint16_t)(xb >> 16), yh = (uint16_t)(yb >> 16);
    sh[threadIdx.x] = make_uint4(xb, yb, xb, yb);
    __syncthreads();
    const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(&sh[threadIdx.x]);
#pragma unroll
    for (int i = 0; i < 16; ++i) { a[i] = i; p[i] = x2 + i; u[i] = threadIdx.x + i; }
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == FFMA || MODE == FFMA_XOR || MODE == FFMA_LDS) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(x), "f"(y));
            if (MODE == FFMA2) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(x2), "l"(y2));
            if (MODE == FHFMA || MODE == FHFMA_XOR || MODE == FHFMA_LDS) asm volatile("fma.rn.f32.bf16 %0, %1, %2, %0;" : "+f"(a[i]) : "h"(xh), "h"(yh));
            if (MODE == F2FP) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(a[(i + 1) & 15]));
            if (MODE == F2FPRELU) asm volatile("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(a[(i + 1) & 15]));
            if (MODE == FMNMX) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(a[(i + 1) & 15]));
            if (MODE == XOR || MODE == FFMA_XOR || MODE == FHFMA_XOR) asm volatile("xor.b32 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) & 15]));
            if (MODE == SHL) asm volatile("shl.b32 %0, %1, 16;" : "=r"(u[i]) : "r"(u[(i + 1) & 15]));
            if (MODE == PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x5410;" : "+r"(u[i]) : "r"(u[(i + 1) & 15]));
            if (MODE == IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(xb), "r"(yb));
            if (MODE == IADD) asm volatile("add.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) & 15]));
            if (MODE == HFMA2BF) asm volatile("fma.rn.bf16x2 %0, %1, %2, %0;" : "+r"(u[i]) : "r"(xb), "r"(yb));
            if ((MODE == LDS128 && (i & 1) == 0) || ((MODE == FFMA_LDS || MODE == FHFMA_LDS) && (i & 1) == 0)) {
                uint32_t r0, r1, r2, r3;
                asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(saddr + ((i * 16) & 0)));
                u[i] ^= r0 ^ r1; u[i + 1] ^= r2 ^ r3;     // 2 LOP3 per LDS: subtract the LOP3 rate when reading the LDS rows
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i] + (float)(p[i] & 0xffff) + (float)u[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static void run(double instr_per_iter) {
    float* out;
    int sms, khz; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const int ctas = sms * 6;
    cudaMalloc(&out, (size_t)ctas * 256 * 4);
    const int iters = 1 << 15;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MODE><<<ctas, 256>>>(out, iters, 1.0f);     // warm-up: clocks up
    cudaEventRecord(e0);
    probe<MODE><<<ctas, 256>>>(out, iters, 1.0f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int nb = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, probe<MODE>, 256, 0);
    const double clk = (double)ms * 1e-3 * khz * 1e3;
    const double warp_instr = (double)ctas * 8 * iters * instr_per_iter;
    printf("%-32s occ %d CTA/SM  %7.3f ms  %.3f listed warp-instr/clk/SMSP  (%.2f clk per instr)\n", kNames[MODE], nb, ms,
           warp_instr / clk / (sms * 4), clk * sms * 4 / warp_instr);
    cudaFree(out);
}

int main() {
    run<FFMA>(16); run<FFMA2>(16); run<FHFMA>(16); run<HFMA2BF>(16); run<F2FP>(16); run<F2FPRELU>(16); run<FMNMX>(16);
    run<XOR>(16); run<SHL>(16); run<PRMT>(16); run<IMAD>(16); run<IADD>(16); run<LDS128>(8 + 16);
    run<FFMA_XOR>(32); run<FHFMA_XOR>(32); run<FFMA_LDS>(16 + 8 + 16); run<FHFMA_LDS>(16 + 8 + 16);
    return cudaDeviceSynchronize() != cudaSuccess;
}
