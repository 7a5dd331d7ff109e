// train_tc.cu -- the training path's pointwise contractions (forward and data gradient of nn.Conv2d(cin, cout, 1), reference
// models/fast_scnn.py:73, :107, and the dense 3x3 convolutions through im2col) on the Blackwell tensor cores:
// tcgen05.mma.kind::tf32 with the accumulator in TMEM, behind fscnn_train_set_math(1).
//
//   C[b][m][j] = sum_k A(m, k) * B[b][k][j]        m: output channels (TMEM lanes, 128 per CTA), j: pixels (128 per tile), k: input channels
//   forward       A = W[m][k]          (k contiguous  -> K-major A operand)
//   data gradient A(m, k) = W[k][m]    (m contiguous  -> MN-major A operand)
//   B = x[b] / dy[b]: [k][pixels], pixels contiguous -> MN-major B operand
//
// The tensors are fp32 NCHW as autograd hands them over, so operands are staged with 16-byte cp.async (a piece = 4 TF32 values;
// out-of-range rows / columns / K tails are zero-filled by the copy) straight into the layouts the tensor core reads:
//   K-major  A : SWIZZLE_NONE core matrices (8 rows x 16 bytes): piece (row m, k/4) at (k/4) * 2064 + m * 16; LBO 2064, SBO 128
//   MN-major A / B : MN-major TF32 operands exist only in the SWIZZLE_128B_BASE32B layout (with SWIZZLE_NONE kind::tf32 returns
//                zeros -- measured; CUTLASS says the same): atoms of 4 k rows x 128 bytes, 32-byte pieces XORed with the row index
// A CTA (4 warps) walks output tiles of 128 channels x 256 (or 128) pixels with 4 (3) K chunks of 16 (32) in flight; one thread issues
// the chunk's MMAs (K = 8 each) and commits them to the stage's mbarrier; the accumulator is read back with tcgen05.ld, transposed
// through the stage the tile's last chunk vacated and stored as rows of 128 contiguous bytes.  Two CTAs per SM overlap one CTA's
// epilogue with the other's contraction.  Measured on a B200 (config 5, 96 x 96 layers): 3.2-3.7 TB/s of operand + result traffic
// against 1.9-2.7 for the legacy mma.sync kernels of train.cu, which stay as the path for rows that are not 16-byte aligned (the
// stem's 27-row im2col matrix) and for the weight gradient.  What the measurements said on the way: storing the accumulator
// straight from the tcgen05.ld registers (32 rows x 16 bytes per instruction) cost 40-55 % of the kernel; 512-byte operand rows
// (128-pixel tiles) reach 2.7 TB/s, 1 KB rows 3.2 -- DRAM page locality, not the tensor core, sets the pace.
#include <cstdint>
#include <cstdlib>

#include "kernels.h"
#include "umma.cuh"

namespace fscnn {

namespace {
constexpr int kTcT = 128;                         // 4 warps = the four TMEM lane quarters
constexpr int kTcBM = 128;
constexpr int kTcALbo = 2048 + 16;                  // K-major A: K pieces 2064 bytes apart, so the pieces of a row land in different bank groups
// Two tilings: 256 pixels x 16 input channels per stage, 4 stages (1 KB contiguous per operand row: DRAM pages are used twice as
// well as with 512-byte rows -- 3.2 instead of 2.7 TB/s on the 96 x 96 layers) and 128 x 32, 3 stages, for the small planes (24 x 24 =
// 576 pixels, where a 256-wide tile wastes a quarter of the work).  Either way a stage holds 16 KB of B and <= 16.5 KB of A.
template <int BN_>
struct TcCfg {
    static constexpr int BN = BN_, BK = BN_ == 256 ? 16 : 32, ST = BN_ == 256 ? 4 : 3;
    static constexpr int BBytes = BK * BN * 4;                                  // 16 KB, 1024-byte aligned swizzle atoms
    static constexpr int ABytes = (BK / 4) * kTcALbo;                           // >= the MN-major form (BK x 128 x 4 bytes)
    static constexpr int Stage = (BBytes + ABytes + 1023) / 1024 * 1024;
    static constexpr int MnLbo = (BK / 4) * 512;                                // MN atoms (32 values) one column of K atoms apart
    static constexpr size_t Smem = (size_t)ST * Stage + 1024;                   // + room to align the stages to 1024 bytes
};

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
__device__ __forceinline__ uint32_t mn_piece_off(int k, int g4, int mn_lbo) {
    const int r = k & 3, c32 = (g4 >> 1) & 3;
    return (uint32_t)((g4 >> 3) * mn_lbo + (k >> 2) * 512 + r * 128 + ((c32 ^ r) << 5) + (g4 & 1) * 16);
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
    uint32_t dbg;                    // bit 1: skip the epilogue's global stores (timing experiments); 0 in production
    long long lda;                   // A_KMAJOR: row stride of W (elements); else: stride between k rows of W
    long long sBb, sCb;              // per-image strides of B and C; their row stride is N
};

template <bool A_KMAJOR, int BN>
__global__ void __launch_bounds__(kTcT, 2)
gemm_pix_tc_kernel(TcGemmArgs g) {
    using Cf = TcCfg<BN>;
    constexpr int kTcBN = Cf::BN, kTcBK = Cf::BK, kTcST = Cf::ST, kTcStage = Cf::Stage, kTcBBytes = Cf::BBytes, kTcMnLbo = Cf::MnLbo;
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
    if (warp == 0) { tmem_alloc(&tmem_base_s, kTcBN); tmem_relinquish(); }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_base_s;

    auto issue = [&](int gc) {
        const int tl = gc / nchunks, chunk = gc - tl * nchunks;
        int b, m0, j0;
        tile_at(tl, b, m0, j0);
        const uint32_t sB = s0 + (gc % kTcST) * kTcStage, sA = sB + kTcBBytes;
        const int k0 = chunk * kTcBK;
        const float* Bp = g.B + b * g.sBb;
#pragma unroll
        for (int r = 0; r < kTcBM * kTcBK / 4 / kTcT; ++r) {
            const int i = tid + r * kTcT;
            if (A_KMAJOR) {                  // piece (row m, kq = k / 4): BK / 4 lanes read the contiguous bytes of a row's chunk
                const int kq = i & (kTcBK / 4 - 1), m = i / (kTcBK / 4);
                const bool ok = m0 + m < g.M && k0 + 4 * kq < g.K;
                cp_async16z(sA + kq * kTcALbo + m * 16, ok ? g.A + (long long)(m0 + m) * g.lda + k0 + 4 * kq : g.A, ok);
            } else {                         // piece (k, 4 rows mg = m / 4): a warp covers one k row's 128 rows = 512 contiguous bytes
                const int mg = i & 31, k = i >> 5;
                const bool ok = k0 + k < g.K && m0 + 4 * mg < g.M;
                cp_async16z(sA + mn_piece_off(k, mg, kTcMnLbo), ok ? g.A + (long long)(k0 + k) * g.lda + m0 + 4 * mg : g.A, ok);
            }
        }
#pragma unroll
        for (int r = 0; r < kTcBN * kTcBK / 4 / kTcT; ++r) {
            const int i = tid + r * kTcT;
            const int jg = i & (kTcBN / 4 - 1), k = i / (kTcBN / 4);
            const bool ok = k0 + k < g.K && j0 + 4 * jg < g.N;
            cp_async16z(sB + mn_piece_off(k, jg, kTcMnLbo), ok ? Bp + (long long)(k0 + k) * g.N + j0 + 4 * jg : Bp, ok);
        }
    };

    constexpr uint32_t idesc = make_idesc_tf32(kTcBM, kTcBN, !A_KMAJOR, true);
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
            const uint32_t sB = s0 + (gc % kTcST) * kTcStage, sA = sB + kTcBBytes;
#pragma unroll
            for (int j = 0; j < kTcBK / 8; ++j) {
                const uint32_t lbo = kTcMnLbo, sbo = 512u;
                const uint64_t da = A_KMAJOR ? make_smem_desc(sA + j * 2 * kTcALbo, kTcALbo, 128) : make_smem_desc_mn32(sA + j * 1024, lbo, sbo);
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
            // A lane holds one output row (its TMEM lane) x 32 pixels: stored directly, a warp instruction would touch 32 rows x 16
            // bytes.  The warp transposes each 32 x 32 block through a 4 KB slab (16-byte pieces XORed with the row, conflict-free)
            // inside the stage the tile's last chunk just vacated (its MMAs have completed: bar_acc), and writes 4 rows x 128
            // contiguous bytes per instruction.
            const uint32_t slab = s0 + (gc % kTcST) * kTcStage + warp * 4096;
            float* Cb = g.C + b * g.sCb + (long long)(m0 + warp * 32) * g.N + j0;
#pragma unroll 1
            for (int cb = 0; cb < kTcBN / 32; ++cb) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16) + cb * 32, r);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 8; ++q) sts128(slab + lane * 128 + ((q ^ (lane & 7)) << 4), r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
                __syncwarp();
                if (!(g.dbg & 2u)) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int row = i * 4 + (lane >> 3), q = lane & 7;
                        const uint4 v = lds128(slab + row * 128 + ((q ^ (row & 7)) << 4));
                        if (m0 + warp * 32 + row < g.M && j0 + cb * 32 + 4 * q < g.N)
                            *reinterpret_cast<uint4*>(Cb + (long long)row * g.N + cb * 32 + 4 * q) = v;
                    }
                }
                __syncwarp();
            }
            chunk = 0;
            ++tl;
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, kTcBN);
}

static bool tc_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// C[b] = A B[b] for b < nb.  a_kmajor: A = W[M][K] (row stride lda); else A(m, k) = W[k * lda + m].  Returns cudaErrorNotSupported when
// the shape cannot take the 16-byte copies (the caller then uses the mma.sync kernel).
cudaError_t launch_gemm_pix_tc(const float* A, const float* B, float* C, int M, int N, int K, int nb, long long lda, bool a_kmajor,
                               cudaStream_t s) {
    if (!tc_aligned16(A) || !tc_aligned16(B) || !tc_aligned16(C) || (N & 3) || (lda & 3) || (a_kmajor ? (K & 3) : (M & 3)))
        return cudaErrorNotSupported;
    TcGemmArgs g{A, B, C, M, N, K, nb, 0u, lda, (long long)K * N, (long long)M * N};
    const bool wide = N >= 2048;
    const int bn = wide ? 256 : 128;
    const long long tiles = (long long)nb * ((N + bn - 1) / bn) * ((M + kTcBM - 1) / kTcBM);
    const int ctas = (int)(tiles < 2ll * num_sms() ? tiles : 2ll * num_sms());
    static unsigned long long done[4] = {0, 0, 0, 0};
    cudaError_t e;
    if (a_kmajor && wide) {
        if ((e = ensure_dyn_smem(gemm_pix_tc_kernel<true, 256>, TcCfg<256>::Smem, done[0])) != cudaSuccess) return e;
        gemm_pix_tc_kernel<true, 256><<<ctas, kTcT, TcCfg<256>::Smem, s>>>(g);
    } else if (a_kmajor) {
        if ((e = ensure_dyn_smem(gemm_pix_tc_kernel<true, 128>, TcCfg<128>::Smem, done[1])) != cudaSuccess) return e;
        gemm_pix_tc_kernel<true, 128><<<ctas, kTcT, TcCfg<128>::Smem, s>>>(g);
    } else if (wide) {
        if ((e = ensure_dyn_smem(gemm_pix_tc_kernel<false, 256>, TcCfg<256>::Smem, done[2])) != cudaSuccess) return e;
        gemm_pix_tc_kernel<false, 256><<<ctas, kTcT, TcCfg<256>::Smem, s>>>(g);
    } else {
        if ((e = ensure_dyn_smem(gemm_pix_tc_kernel<false, 128>, TcCfg<128>::Smem, done[3])) != cudaSuccess) return e;
        gemm_pix_tc_kernel<false, 128><<<ctas, kTcT, TcCfg<128>::Smem, s>>>(g);
    }
    return cudaGetLastError();
}

}  // namespace fscnn
