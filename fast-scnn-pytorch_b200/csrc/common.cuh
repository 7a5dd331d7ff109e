// common.cuh -- shared device helpers for the Fast-SCNN sm_100a kernels.
//
// Activations are NHWC with storage type T (float or __nv_bfloat16); arithmetic is fp32
// unless a kernel says otherwise.  All kernels use 256-thread CTAs and a 128-pixel tile whose
// pointwise contraction runs as a 16x16 grid of register tiles (8 pixels x COUT/16 channels).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

namespace fscnn {

typedef __nv_bfloat16 bf16;

constexpr int kThreads = 256;

__host__ __device__ constexpr int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ constexpr int round_up(int a, int b) { return ceil_div(a, b) * b; }

// ---- storage-type traits ----------------------------------------------------------------
template <typename T>
struct Act;

template <>
struct Act<float> {
    static __device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
    static __device__ __forceinline__ float2 ld2(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
    static __device__ __forceinline__ float ld1(const float* p) { return __ldg(p); }
    static __device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
    static __device__ __forceinline__ void st2(float* p, float2 v) { *reinterpret_cast<float2*>(p) = v; }
    static __device__ __forceinline__ void st1(float* p, float v) { *p = v; }
};

template <>
struct Act<bf16> {
    static __device__ __forceinline__ float4 ld4(const bf16* p) {
        const uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
        const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.x));
        const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.y));
        return make_float4(a.x, a.y, b.x, b.y);
    }
    static __device__ __forceinline__ float2 ld2(const bf16* p) {
        const uint32_t r = __ldg(reinterpret_cast<const uint32_t*>(p));
        return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r));
    }
    static __device__ __forceinline__ float ld1(const bf16* p) { return __bfloat162float(*p); }
    static __device__ __forceinline__ void st4(bf16* p, float4 v) {
        __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
        uint2 r;
        r.x = *reinterpret_cast<uint32_t*>(&a);
        r.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(p) = r;
    }
    static __device__ __forceinline__ void st2(bf16* p, float2 v) {
        *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(v.x, v.y);
    }
    static __device__ __forceinline__ void st1(bf16* p, float v) { *p = __float2bfloat16_rn(v); }
};

__device__ __forceinline__ float relu(float v) { return fmaxf(v, 0.f); }

// ---- register-tile contraction -------------------------------------------------------------
// Output-channel ownership of thread column `tn` (0..15) for a tile of COUT = 16*TN channels:
// channel(q, j) = q*16*VW + tn*VW + j, q < NQ, j < VW, so that each of the NQ shared-memory
// (and global) accesses of a half-warp covers 16*VW contiguous floats.
template <int TN>
struct ColMap {
    static constexpr int VW = (TN % 4 == 0) ? 4 : ((TN % 2 == 0) ? 2 : 1);
    static constexpr int NQ = TN / VW;
    static __device__ __forceinline__ int ch(int tn, int q, int j) { return q * 16 * VW + tn * VW + j; }
};

// Column swizzle of the pixel-major operand tile As[k][128]: pixel column p of row k is stored at
// p ^ swz(k).  It permutes 4-pixel groups inside each 32-pixel group, which keeps the float4 reads
// of the contraction aligned while making the channel-vectorised transposing stores conflict free.
__device__ __forceinline__ int swz(int k) { return ((k >> 2) & 7) << 2; }

// acc[i][.] += sum_k As[k][8*tp + i] * Bs[k][channels of tn];  As rows are LDA floats apart.
template <int KC, int TN, int LDA, int LDB, bool SWZ>
__device__ __forceinline__ void contract_chunk(float (&acc)[8][TN], const float* __restrict__ As,
                                               const float* __restrict__ Bs, int tp, int tn) {
    using CM = ColMap<TN>;
#pragma unroll 8
    for (int k = 0; k < KC; ++k) {
        const int s = SWZ ? swz(k) : 0;
        const float4 a0 = *reinterpret_cast<const float4*>(As + k * LDA + ((8 * tp) ^ s));
        const float4 a1 = *reinterpret_cast<const float4*>(As + k * LDA + ((8 * tp + 4) ^ s));
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        float b[TN];
#pragma unroll
        for (int q = 0; q < CM::NQ; ++q) {
            const float* bp = Bs + k * LDB + q * 16 * CM::VW + tn * CM::VW;
            if (CM::VW == 4) {
                const float4 v = *reinterpret_cast<const float4*>(bp);
                b[q * 4 + 0] = v.x; b[q * 4 + 1] = v.y; b[q * 4 + 2] = v.z; b[q * 4 + 3] = v.w;
            } else if (CM::VW == 2) {
                const float2 v = *reinterpret_cast<const float2*>(bp);
                b[q * 2 + 0] = v.x; b[q * 2 + 1] = v.y;
            } else {
                b[q] = *bp;
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
}

// Copies ROWS x COLS floats (row stride `ld_src` in global memory) into a dense shared tile.
template <int ROWS, int COLS>
__device__ __forceinline__ void load_weight_tile(float* __restrict__ dst, const float* __restrict__ src, int ld_src) {
    static_assert(COLS % 4 == 0, "weight tiles are copied as float4");
    constexpr int V = COLS / 4;
    for (int i = threadIdx.x; i < ROWS * V; i += kThreads) {
        const int r = i / V, c = i % V;
        reinterpret_cast<float4*>(dst)[i] = __ldg(reinterpret_cast<const float4*>(src + (size_t)r * ld_src) + c);
    }
}

// Stores the VW-wide piece q of a thread's output row (after bias / activation).
template <typename T, int VW>
__device__ __forceinline__ void store_vec(T* p, const float* v) {
    if (VW == 4) Act<T>::st4(p, make_float4(v[0], v[1], v[2], v[3]));
    else if (VW == 2) Act<T>::st2(p, make_float2(v[0], v[1]));
    else Act<T>::st1(p, v[0]);
}

// SM count of the current device (persistent kernels size their grid with it)
inline int num_sms() {
    static int cached[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    int& v = cached[dev & 63];
    if (!v && (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0)) v = 148;
    return v;
}

// Whether the stage kernels launched next by this thread get the programmatic-serialization attribute.  The forward sets the hint
// per launch set (api.cu run_stages): on up to 160 Mpixel of input, where a stage's set-up is worth hiding (batch 1: -4.5 % latency;
// 16 / 32 / 64 images per launch: +3.1 / +1.5 / +0.1 % images/s); off for the largest micro-batches, where it measured 0.7 % slower
// (17 326 vs 17 452 images/s at 111 images per launch, twice).  FSCNN_NO_PDL=1 / FSCNN_PDL_ALWAYS=1 force it off / on (A/B switches).
inline bool& pdl_hint() {
    static thread_local bool hint = true;
    return hint;
}
inline bool pdl_enabled() {
    static const bool off = getenv("FSCNN_NO_PDL") != nullptr, always = getenv("FSCNN_PDL_ALWAYS") != nullptr;
    return !off && (always || pdl_hint());
}

// Opt a kernel into `bytes` of dynamic shared memory once per device (the attribute is per context).
template <typename F>
inline cudaError_t ensure_dyn_smem(F* func, size_t bytes, unsigned long long& done_mask) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if ((done_mask >> (dev & 63)) & 1ull) return cudaSuccess;
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) done_mask |= 1ull << (dev & 63);
    return e;
}

// ---- programmatic dependent launch (PDL) --------------------------------------------------------------------------------
// The stage kernels of a forward run back to back on one stream; at small batches a kernel's set-up (barrier init, TMEM
// allocation, weight copies into shared memory: 3-5 us) is a third of its life.  Launched with the programmatic-serialization
// attribute, kernel k+1 may start while kernel k is still running: it does its set-up, then blocks in pdl_wait() until kernel k has
// completed and its writes are visible, and only then touches activations (reads its inputs, and -- because every output depends
// on an input -- writes its outputs).  pdl_launch_dependents() at the top of a kernel lets ITS successor start early in turn.
// Both instructions are no-ops in a kernel launched without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace fscnn
