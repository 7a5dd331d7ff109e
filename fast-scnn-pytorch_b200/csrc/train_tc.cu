// train_tc.cu -- the training path's pointwise contractions (forward and data gradient of nn.Conv2d(cin, cout, 1), reference
// models/fast_scnn.py:73, :107, and the dense 3x3 convolutions through im2col) on the Blackwell tensor cores:
// tcgen05.mma.kind::tf32 with the accumulator in TMEM, behind fscnn_train_set_math(1).
//
//   C[b][m][j] = sum_k A(m, k) * B[b][k][j]        m: output channels (TMEM lanes, 128 per CTA), j: pixels (128 per tile), k: input channels
//   forward       A = W[m][k]          (k contiguous  -> K-major A operand)
//   data gradient A(m, k) = W[k][m]    (m contiguous  -> MN-major A operand)
//   B = x[b] / dy[b]: [k][pixels], pixels contiguous -> MN-major B operand
//
// The tensors are fp32 NCHW as autograd hands them over, so operands are staged with 16-byte cp.async straight into the
// SWIZZLE_NONE canonical layouts (a 16-byte piece = 4 TF32 values = one core-matrix row; out-of-range rows are zero-filled by the
// copy), three K chunks of 32 in flight per CTA:
//   K-major  A : piece (row m, k/4)      at (k/4) * 2048 + m * 16                              LBO 2048 (K pieces), SBO 128 (8-row groups)
//   MN-major A : piece (k, 4 rows m/4)   at (k/8) * 4096 + (m/4) * 128 + (k%8) * 16            LBO 4096 (K groups), SBO 128 (4-row groups)
//   MN-major B : piece (k, 4 pixels j/4) at (k/8) * 4096 + (j/4) * 128 + (k%8) * 16            the same with pixels for rows
// One thread issues four MMAs (K = 8 each) per chunk and commits them to the stage's mbarrier; the four warps read their TMEM lane
// quarters back with tcgen05.ld and store rows of 32 pixels.  Two CTAs per SM (96 KB of stages, 128 TMEM columns each) overlap one
// CTA's epilogue with the other's contraction.  The legacy mma.sync kernels of train.cu (which top out near 77 TFLOP/s on a B200)
// stay as the path for shapes whose rows are not 16-byte aligned (the stem's 27-row im2col matrix).
#include <cstdint>
#include <cstdlib>

#include "kernels.h"
#include "umma.cuh"

namespace fscnn {

namespace {
constexpr int kTcT = 128;                         // 4 warps = the four TMEM lane quarters
constexpr int kTcBM = 128, kTcBN = 128, kTcBK = 32, kTcST = 3;
constexpr int kTcABytes = kTcBM * kTcBK * 4, kTcBBytes = kTcBK * kTcBN * 4, kTcStage = kTcABytes + kTcBBytes;
constexpr size_t kTcSmem = (size_t)kTcST * kTcStage + 1024;      // + room to align the stages to 1024 bytes (swizzle atoms)

__host__ __device__ constexpr uint32_t make_idesc_tf32(int m, int n, bool a_mn, bool b_mn) {
    // c_format F32 (1) at bit 4, a_format / b_format TF32 (2) at bits 7 / 10, a_major / b_major at bits 15 / 16, N >> 3 at 17, M >> 4 at 24
    return (1u << 4) | (2u << 7) | (2u << 10) | (a_mn ? 1u << 15 : 0u) | (b_mn ? 1u << 16 : 0u) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, int accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// MN-major TF32 operands exist only in the SWIZZLE_128B_BASE32B layout (layout type 1; SWIZZLE_NONE with the MN-major bit makes
// kind::tf32 return zeros -- measured): atoms of 4 k rows x 128 bytes (32 values along MN), the 32-byte pieces of a row XORed with
// the row index (Swizzle<2,5,2> on the byte address).  A [32 k][128 mn] chunk is 8 x 4 atoms: K atoms 512 bytes apart (SBO), MN
// atoms 4096 bytes apart (LBO).  `g4` = index of the 16-byte piece along MN (4 values).
__device__ __forceinline__ uint32_t mn_piece_off(int k, int g4) {
    const int r = k & 3, c32 = (g4 >> 1) & 3;
    return (uint32_t)((g4 >> 3) * 4096 + (k >> 2) * 512 + r * 128 + ((c32 ^ r) << 5) + (g4 & 1) * 16);
}
__device__ __forceinline__ uint64_t make_smem_desc_mn32(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    const uint32_t lo = ((smem_addr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
    const uint32_t hi = ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (1u << 29);   // version 1, layout type 1 = SWIZZLE_128B_BASE32B
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ void cp_async_commit_group() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
}  // namespace

struct TcGemmArgs {
    const float *A, *B;
    float* C;
    int M, N, K, nb;                 // output channels, pixels per image, input channels, images
    uint32_t idesc_xor;              // debug
    long long lda;                   // A_KMAJOR: row stride of W (elements); else: stride between k rows of W
    long long sBb, sCb;              // per-image strides of B and C; their row stride is N
};

template <bool A_KMAJOR>
__global__ void __launch_bounds__(kTcT, 2)
gemm_pix_tc_kernel(TcGemmArgs g) {
    extern __shared__ __align__(128) uint8_t tc_smem[];
    __shared__ __align__(8) uint64_t bar_free[kTcST], bar_acc;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t s0 = (smem_u32(tc_smem) + 1023u) & ~1023u;
    const int nchunks = (g.K + kTcBK - 1) / kTcBK;
    const int ny = (g.M + kTcBM - 1) / kTcBM, nx = (g.N + kTcBN - 1) / kTcBN;
    const long long tiles = (long long)g.nb * nx * ny;
    const int my_tiles = (int)((tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
    const int total = my_tiles * nchunks;
    // tiles ordered (image, pixel block, channel block): CTAs running side by side share a B tile through L2
    auto tile_at = [&](int tl, int& b, int& m0, int& j0) {
        const long long t = blockIdx.x + (long long)tl * gridDim.x;
        m0 = (int)(t % ny) * kTcBM;
        j0 = (int)((t / ny) % nx) * kTcBN;
        b = (int)(t / ((long long)ny * nx));
    };
    if (tid == 0) {
        for (int i = 0; i < kTcST; ++i) mbar_init(&bar_free[i], 1);
        mbar_init(&bar_acc, 1);
        fence_mbar_init();
    }
    if (warp == 0) { tmem_alloc(&tmem_base_s, 128); tmem_relinquish(); }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_base_s;

    auto issue = [&](int gc) {
        const int tl = gc / nchunks, chunk = gc - tl * nchunks;
        int b, m0, j0;
        tile_at(tl, b, m0, j0);
        const uint32_t sA = s0 + (gc % kTcST) * kTcStage, sB = sA + kTcABytes;
        const int k0 = chunk * kTcBK;
        const float* Bp = g.B + b * g.sBb;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int i = tid + r * kTcT;
            if (A_KMAJOR) {                  // piece (kq = k / 4, row m): lanes along m -> conflict-free shared-memory writes
                const int m = i & 127, kq = i >> 7;
                const bool ok = m0 + m < g.M && k0 + 4 * kq < g.K;
                cp_async16z(sA + kq * 2048 + m * 16, ok ? g.A + (long long)(m0 + m) * g.lda + k0 + 4 * kq : g.A, ok);
            } else {                         // piece (k, 4 rows mg = m / 4): a warp covers one k row's 128 rows = 512 contiguous bytes
                const int mg = i & 31, k = i >> 5;
                const bool ok = k0 + k < g.K && m0 + 4 * mg < g.M;
                cp_async16z(sA + mn_piece_off(k, mg), ok ? g.A + (long long)(k0 + k) * g.lda + m0 + 4 * mg : g.A, ok);
            }
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int i = tid + r * kTcT;
            const int jg = i & 31, k = i >> 5;
            const bool ok = k0 + k < g.K && j0 + 4 * jg < g.N;
            cp_async16z(sB + mn_piece_off(k, jg), ok ? Bp + (long long)(k0 + k) * g.N + j0 + 4 * jg : Bp, ok);
        }
    };

    const uint32_t idesc = make_idesc_tf32(kTcBM, kTcBN, !A_KMAJOR, true) ^ (g.idesc_xor & ~1u);
#pragma unroll
    for (int s = 0; s < kTcST - 1; ++s) {
        if (s < total) issue(s);
        cp_async_commit_group();
    }
    int chunk = 0, tl = 0;
    for (int gc = 0; gc < total; ++gc) {
        cp_async_wait_group<kTcST - 2>();         // this thread's pieces of chunk gc have landed ...
        fence_async_proxy();                      // ... and are visible to the tensor core (async proxy)
        tc_fence_before_sync();                   // the epilogue's TMEM reads of the previous tile precede the MMAs issued below
        __syncthreads();
        if (tid == 0) {
            tc_fence_after_sync();
            const uint32_t sA = s0 + (gc % kTcST) * kTcStage, sB = sA + kTcABytes;
#pragma unroll
            for (int j = 0; j < kTcBK / 8; ++j) {
                const uint32_t lbo = g.idesc_xor & 1u ? 512u : 4096u, sbo = g.idesc_xor & 1u ? 4096u : 512u;      // debug: bit 0 swaps
                const uint64_t da = A_KMAJOR ? make_smem_desc(sA + j * 2 * 2048, 2048, 128) : make_smem_desc_mn32(sA + j * 1024, lbo, sbo);
                const uint64_t db = make_smem_desc_mn32(sB + j * 1024, lbo, sbo);
                umma_tf32_ss(tmem, da, db, idesc, (chunk | j) != 0);
            }
            umma_commit(&bar_free[gc % kTcST]);   // the stage may be refilled once these MMAs have read it
            if (chunk == nchunks - 1) umma_commit(&bar_acc);
        }
        // refill the stage of chunk gc - 1 with chunk gc + ST - 1 (its MMAs were committed one iteration ago)
        if (gc + kTcST - 1 < total) {
            if (gc >= 1) mbar_wait(&bar_free[(gc - 1) % kTcST], ((gc - 1) / kTcST) & 1);
            issue(gc + kTcST - 1);
        }
        cp_async_commit_group();
        if (++chunk == nchunks) {                 // the tile is complete: read the accumulator back and store it
            int b, m0, j0;
            tile_at(tl, b, m0, j0);
            mbar_wait(&bar_acc, tl & 1);
            tc_fence_after_sync();
            const int m = m0 + warp * 32 + lane;
            float* Cp = g.C + b * g.sCb + (long long)m * g.N + j0;
#pragma unroll 1
            for (int cb = 0; cb < kTcBN / 32; ++cb) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16) + cb * 32, r);
                tmem_ld_wait();
                if (m < g.M) {
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        if (j0 + cb * 32 + 4 * q < g.N)
                            *reinterpret_cast<float4*>(Cp + cb * 32 + 4 * q) = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]),
                                                                                           __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
                }
            }
            chunk = 0;
            ++tl;
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}

static bool tc_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// C[b] = A B[b] for b < nb.  a_kmajor: A = W[M][K] (row stride lda); else A(m, k) = W[k * lda + m].  Returns cudaErrorNotSupported when
// the shape cannot take the 16-byte copies (the caller then uses the mma.sync kernel).
cudaError_t launch_gemm_pix_tc(const float* A, const float* B, float* C, int M, int N, int K, int nb, long long lda, bool a_kmajor,
                               cudaStream_t s) {
    if (!tc_aligned16(A) || !tc_aligned16(B) || !tc_aligned16(C) || (N & 3) || (lda & 3) || (a_kmajor ? (K & 3) : (M & 3)))
        return cudaErrorNotSupported;
    static unsigned long long done_k = 0, done_m = 0;
    cudaError_t e = a_kmajor ? ensure_dyn_smem(gemm_pix_tc_kernel<true>, kTcSmem, done_k) : ensure_dyn_smem(gemm_pix_tc_kernel<false>, kTcSmem, done_m);
    if (e != cudaSuccess) return e;
    TcGemmArgs g{A, B, C, M, N, K, nb, getenv("FSCNN_TC_IDESC_XOR") ? (uint32_t)strtoul(getenv("FSCNN_TC_IDESC_XOR"), nullptr, 0) : 0u, lda, (long long)K * N, (long long)M * N};
    const long long tiles = (long long)nb * ((N + kTcBN - 1) / kTcBN) * ((M + kTcBM - 1) / kTcBM);
    const int ctas = (int)(tiles < 2ll * num_sms() ? tiles : 2ll * num_sms());
    if (a_kmajor) gemm_pix_tc_kernel<true><<<ctas, kTcT, kTcSmem, s>>>(g);
    else gemm_pix_tc_kernel<false><<<ctas, kTcT, kTcSmem, s>>>(g);
    return cudaGetLastError();
}

}  // namespace fscnn
