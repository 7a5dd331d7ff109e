// head.cu -- the tail of the forward path:
//   up_logits  : F.interpolate(logits, size, 'bilinear', align_corners=True) (reference
//                models/fast_scnn.py:40 and :44) -> NCHW fp32, API-parity output of forward();
//   up_argmax  : the same resize fused with torch.argmax(outputs[0], 1) (eval.py:45, demo.py:48) and,
//                when labels are given, with SegmentationMetric's counting (utils/metric.py:73-105),
//                so the full-resolution logits (159 MB / image at 1024x2048) never exist;
//   confusion  : metric counting for class maps that already exist on the device.
//
// One CTA covers 64 output rows x 128 columns; a thread owns 8 rows x 4 adjacent columns.  The
// low-resolution logits the CTA needs (<= 10 x 18 pixels, because the resize ratio is always < 1/8)
// are staged class-major in shared memory, so every class costs 9 conflict-free LDS per thread for
// 32 output pixels.  Interpolation order matches ATen's upsample_bilinear2d: horizontal lerp first,
// then vertical.  argmax keeps the FIRST maximal class and treats NaN as maximal, like torch.
#include <cstdlib>

#include "kernels.h"
#include "tma_host.h"
#include "umma.cuh"

#include "../../include/fscnn_b200.h"

namespace fscnn {

constexpr int kTR = 12, kTC = 20;        // staged low-res tile (<= 10 x 18 needed; the 3rd tap column/row may overhang)
constexpr int kHistMaxBins = 4096;       // (nc+1)^2 above this -> histogram goes straight to global atomics

template <int DT>
struct Lab;
template <> struct Lab<FSCNN_U8> { typedef unsigned char type; };
template <> struct Lab<FSCNN_I32> { typedef int type; };
template <> struct Lab<FSCNN_I64> { typedef long long type; };

__device__ __forceinline__ long long load_label(const void* p, int dtype, size_t i) {
    if (dtype == FSCNN_U8) return (long long)__ldg(reinterpret_cast<const unsigned char*>(p) + i);
    if (dtype == FSCNN_I32) return (long long)__ldg(reinterpret_cast<const int*>(p) + i);
    return __ldg(reinterpret_cast<const long long*>(p) + i);
}

// four consecutive labels starting at element i (vector loads when the row pitch keeps them aligned)
__device__ __forceinline__ void load_labels4(const void* p, int dtype, size_t i, bool aligned, int remaining, long long (&out)[4]) {
    if (aligned) {
        if (dtype == FSCNN_U8) {
            const uchar4 v = __ldg(reinterpret_cast<const uchar4*>(reinterpret_cast<const unsigned char*>(p) + i));
            out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
        } else if (dtype == FSCNN_I32) {
            const int4 v = __ldg(reinterpret_cast<const int4*>(reinterpret_cast<const int*>(p) + i));
            out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
        } else {
            const longlong2 a = __ldg(reinterpret_cast<const longlong2*>(reinterpret_cast<const long long*>(p) + i));
            const longlong2 b = __ldg(reinterpret_cast<const longlong2*>(reinterpret_cast<const long long*>(p) + i) + 1);
            out[0] = a.x; out[1] = a.y; out[2] = b.x; out[3] = b.y;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) out[j] = j < remaining ? load_label(p, dtype, i + j) : -1;
    }
}

// Adds one (label, pred) observation to the confusion accumulator held in `hist` (shared, uint32) or,
// when hist == nullptr, directly to the global int64 accumulator.
struct ConfSink {
    unsigned int* hist;
    unsigned long long* conf;
    int nc;
    __device__ __forceinline__ void add_key(int bin, unsigned int count) const {
        if (hist) atomicAdd(hist + bin, count);
        else atomicAdd(conf + bin, (unsigned long long)count);
    }
    __device__ __forceinline__ void add(int row, int col, unsigned int count) const {
        const int bin = row * (nc + 1) + col;
        if (hist) atomicAdd(hist + bin, count);
        else atomicAdd(conf + bin, (unsigned long long)count);
    }
};

// ---- exact class pruning by pairwise dominance --------------------------------------------------------------
// A low-resolution CELL is the 2x2 group of staged taps (R..R+1) x (Q..Q+1); every output pixel interpolates inside
// exactly one cell.  The interpolation (horizontal fma/mul, then vertical fma/mul, all weights >= 0, round-to-nearest)
// is a monotone non-decreasing function of each tap, so if class c is >= class d at all four corners of a cell, the
// COMPUTED value of c is >= the computed value of d at every pixel of the cell.  With c < d that already means d can
// never be torch.argmax's answer there (first maximal index wins); with c > d the four gaps must exceed the rounding
// bound of the two lerps (< 4 ulp of the largest staged |logit|; 2e-6 x that leaves 8x headroom) so that c is strictly
// larger.  Per cell, the champions (first maximal class) of corner (R,Q) and of corner (R+1,Q+1) each eliminate the
// classes they dominate; what is left is a bit mask of the classes that can still win somewhere in the cell.  Constant
// regions keep ONE class (no interpolation at all), a boundary between two regions keeps two, all-tied logits keep
// class 0, and only logits where no class dominates another anywhere (e.g. class order flipping between neighbouring
// taps) keep everything: the mask is bit-identical to the exhaustive loop in every case, only the time depends on the data.
#ifndef FSCNN_TAIL_PF
#define FSCNN_TAIL_PF 1
#endif
#ifndef FSCNN_TAIL_NW
#define FSCNN_TAIL_NW 8
#endif
constexpr int kANW = FSCNN_TAIL_NW;              // warps per CTA of the argmax kernel: the CTA covers 8 * kANW rows x 128 columns
constexpr int kATH = kANW * 32;
constexpr int kATR = kANW + 4;                   // staged low-res rows: 8 * kANW / 7.3 + 3 taps
constexpr int kACS = kATR * kTC;                 // class stride of the staged tile
constexpr int kACR = kATR - 1, kCC = kTC - 1;    // cells per staged tile
constexpr int kCR = kTR - 1;                     // cell rows of the TMA kernel's 12-row box

// The staged logits come in two layouts: class-major [class][row][col] (PM = false: the staging loop of the direct kernel
// transposes while it copies) and pixel-major [row][col][padded class] (PM = true: what a TMA box of the NHWC logits tensor
// looks like).  `off` is the pixel index row * kTC + col in both.
template <bool PM>
__device__ __forceinline__ float ldl(const float* __restrict__ L, int c, int off, int ncp) {
    return PM ? L[off * ncp + c] : L[c * kACS + off];
}

// o0..o3: pixel offsets of the cell's corners (row, col), (row, col+1), (row+1, col), (row+1, col+1), clamped at the image border
template <bool PM>
__device__ __forceinline__ unsigned int dominated_by(const float* __restrict__ L, int nc, int ncp, int o0, int o1, int o2, int o3,
                                                     int ch, float margin) {
    const float p0 = ldl<PM>(L, ch, o0, ncp), p1 = ldl<PM>(L, ch, o1, ncp), p2 = ldl<PM>(L, ch, o2, ncp), p3 = ldl<PM>(L, ch, o3, ncp);
    unsigned int elim = 0u;
#pragma unroll 4
    for (int d = 0; d < nc; ++d) {
        const float m = fminf(fminf(p0 - ldl<PM>(L, d, o0, ncp), p1 - ldl<PM>(L, d, o1, ncp)),
                              fminf(p2 - ldl<PM>(L, d, o2, ncp), p3 - ldl<PM>(L, d, o3, ncp)));
        const bool e = ch < d ? (m >= 0.f) : (m > margin);      // d == ch: m == 0 > margin is false
        elim |= (e ? 1u : 0u) << d;
    }
    return elim;
}

template <bool PM>
__device__ __forceinline__ unsigned int cell_survivors(const float* __restrict__ L, int nc, int ncp, int o0, int o1, int o2, int o3,
                                                       float margin) {
    int ch0 = 0, ch3 = 0;
    float b0 = ldl<PM>(L, 0, o0, ncp), b3 = ldl<PM>(L, 0, o3, ncp);
#pragma unroll 4
    for (int c = 1; c < nc; ++c) {
        const float v0 = ldl<PM>(L, c, o0, ncp), v3 = ldl<PM>(L, c, o3, ncp);
        if (v0 > b0) { b0 = v0; ch0 = c; }
        if (v3 > b3) { b3 = v3; ch3 = c; }
    }
    unsigned int elim = dominated_by<PM>(L, nc, ncp, o0, o1, o2, o3, ch0, margin);
    if (ch3 != ch0) elim |= dominated_by<PM>(L, nc, ncp, o0, o1, o2, o3, ch3, margin);
    return ~elim & (nc >= 32 ? 0xffffffffu : ((1u << nc) - 1u));
}

// One 4-row half of a thread's 8x4 output block: finite-logit argmax over the candidate classes `cand` (ascending, so
// `v > best` keeps the first maximal class).  Rows i < KH interpolate between the horizontally-lerped staged rows 0/1,
// rows i >= KH between rows 1/2 (KH is warp-uniform and a template constant, so the vertical step is one FMUL + one FFMA
// per pixel and unused staged rows are never read).  The winning class indices are kept packed, one byte per pixel, one
// register per row (PRMT replaces the winner's byte), which is also the uint8 mask word the thread stores.
template <int KH, bool PM>
__device__ __forceinline__ void argmax_half_fast(const float* __restrict__ Ls, int ncp, unsigned int cand, const int (&ro)[3], const int (&co)[3],
                                                 const float (&wx)[4][3], const float (&wa)[4], const float (&wb)[4], unsigned int (&idx)[4]) {
    float best[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        idx[i] = 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j) best[i][j] = -INFINITY;
    }
    while (cand) {
        const unsigned int c = __ffs(cand) - 1;
        cand &= cand - 1u;
        float hrow[3][4];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            if ((KH == 4 && r == 2) || (KH == 0 && r == 0)) continue;      // staged rows this half never uses
            const float v0 = ldl<PM>(Ls, c, ro[r] + co[0], ncp), v1 = ldl<PM>(Ls, c, ro[r] + co[1], ncp), v2 = ldl<PM>(Ls, c, ro[r] + co[2], ncp);
#pragma unroll
            for (int j = 0; j < 4; ++j) hrow[r][j] = fmaf(wx[j][2], v2, fmaf(wx[j][1], v1, wx[j][0] * v0));
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float val = fmaf(wb[i], i < KH ? hrow[1][j] : hrow[2][j], wa[i] * (i < KH ? hrow[0][j] : hrow[1][j]));
                if (val > best[i][j]) {
                    best[i][j] = val;
                    idx[i] = __byte_perm(idx[i], c, j == 0 ? 0x3214 : (j == 1 ? 0x3240 : (j == 2 ? 0x3410 : 0x4210)));
                }
            }
        }
    }
}

// The same half with torch.argmax's full semantics (NaN is maximal, the first one wins) over all classes, reading exactly
// the 2 x 2 taps ATen reads per pixel -- a non-finite value under a zero weight of the 3-tap form would poison pixels torch
// keeps finite.  hx = 1 - lx > 0 always, so a zero first weight marks the columns that use the second tap pair; sy marks
// the rows that do.  Used when a staged logit is NaN / Inf or there are more classes than the candidate mask holds.
template <bool PM>
__device__ __forceinline__ void argmax_half_generic(const float* __restrict__ Ls, int nc, int ncp, const int (&ro)[3], const int (&co)[3],
                                                    const float (&wx)[4][3], const float (&wa)[4], const float (&wb)[4],
                                                    unsigned int sy_mask, unsigned int (&idx)[4]) {
    float best[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        idx[i] = 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j) best[i][j] = 0.f;
    }
    for (int c = 0; c < nc; ++c) {
        float hrow[3][4];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const float v0 = ldl<PM>(Ls, c, ro[r] + co[0], ncp), v1 = ldl<PM>(Ls, c, ro[r] + co[1], ncp), v2 = ldl<PM>(Ls, c, ro[r] + co[2], ncp);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const bool sx = wx[j][0] == 0.f;
                hrow[r][j] = fmaf(sx ? wx[j][2] : wx[j][1], sx ? v2 : v1, (sx ? wx[j][1] : wx[j][0]) * (sx ? v1 : v0));
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const bool sy = (sy_mask >> i) & 1u;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float val = fmaf(wb[i], sy ? hrow[2][j] : hrow[1][j], wa[i] * (sy ? hrow[1][j] : hrow[0][j]));
                // first max wins; NaN beats everything except an earlier NaN
                const bool upd = (c == 0) || (!(val <= best[i][j]) && (best[i][j] == best[i][j]));
                if (upd) {
                    best[i][j] = val;
                    idx[i] = __byte_perm(idx[i], (unsigned int)c, j == 0 ? 0x3214 : (j == 1 ? 0x3240 : (j == 2 ? 0x3410 : 0x4210)));
                }
            }
        }
    }
}

// predicated shared-memory reduction: no branch, so 32 lanes with different run lengths stay converged
__device__ __forceinline__ void red_shared_if(bool p, uint32_t addr, unsigned int v) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p red.shared.add.u32 [%0], %1;\n\t}" ::"r"(addr), "r"(v), "r"((int)p) : "memory");
}

// MODE 0: write NCHW fp32 logits (forward()'s API-parity output).
__global__ void __launch_bounds__(kThreads, 2)
upsample_logits_kernel(const float* __restrict__ low, int nc, int ncp, float* __restrict__ out_logits, int hl, int wl, int H, int W) {
    extern __shared__ __align__(16) float dynsm[];
    float* Ls = dynsm;                                            // [nc][kTR][kTC]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = blockIdx.z;
    const int yb = blockIdx.y * 64, xb = blockIdx.x * 128;
    const float scy = H > 1 ? (float)(hl - 1) / (float)(H - 1) : 0.f;
    const float scx = W > 1 ? (float)(wl - 1) / (float)(W - 1) : 0.f;
    const int rb = min((int)(scy * (float)yb), hl - 1);          // first staged low-res row / column
    const int cb = min((int)(scx * (float)xb), wl - 1);
    {
        const int nv = ncp >> 2;
        for (int i = tid; i < kTR * kTC * nv; i += kThreads) {
            const int v = i % nv, px = i / nv;
            const int r = px / kTC, q = px % kTC;
            const int rr = min(rb + r, hl - 1), qq = min(cb + q, wl - 1);
            const float4 t = __ldg(reinterpret_cast<const float4*>(low + (((size_t)n * hl + rr) * wl + qq) * ncp) + v);
            const int c = 4 * v;
            float* dst = Ls + (c * kTR + r) * kTC + q;
            dst[0] = t.x;
            if (c + 1 < nc) dst[kTR * kTC] = t.y;
            if (c + 2 < nc) dst[2 * kTR * kTC] = t.z;
            if (c + 3 < nc) dst[3 * kTR * kTC] = t.w;
        }
    }
    __syncthreads();
    const int x0 = xb + lane * 4, y0 = yb + warp * 8;
    if (x0 >= W || y0 >= H) return;
    // Separable weights over the 3 x 3 staged taps of the thread's 8 x 4 block; every pixel reads exactly the 2 x 2 taps ATen
    // reads (a non-finite logit under a zero weight would poison pixels torch keeps finite): hx = 1 - lx > 0 always, so
    // `sx` / `sy` pick the tap pair explicitly.  Horizontal lerp first, then vertical, like ATen.
    float hx[4], lx[4], hy[8], ly[8];
    bool sx[4], sy[8];
    const int c0 = min((int)(scx * (float)x0), wl - 1), r0 = min((int)(scy * (float)y0), hl - 1);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float fx = scx * (float)(x0 + j);
        const int q = min((int)fx, wl - 1);
        lx[j] = fx - (float)q; hx[j] = 1.f - lx[j];
        sx[j] = (q - c0) != 0;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float fy = scy * (float)(y0 + i);
        const int r = min((int)fy, hl - 1);
        ly[i] = fy - (float)r; hy[i] = 1.f - ly[i];
        sy[i] = (r - r0) != 0;
    }
    const int base = (r0 - rb) * kTC + (c0 - cb);     // the staged tile holds clamped copies beyond the border (ATen's x1 = min(x0 + 1, w - 1))
    for (int c = 0; c < nc; ++c) {
        const float* lc = Ls + c * kTR * kTC + base;
        float hrow[3][4];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const float v0 = lc[r * kTC], v1 = lc[r * kTC + 1], v2 = lc[r * kTC + 2];
#pragma unroll
            for (int j = 0; j < 4; ++j) hrow[r][j] = fmaf(lx[j], sx[j] ? v2 : v1, hx[j] * (sx[j] ? v1 : v0));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (y0 + i >= H) continue;
            float val[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) val[j] = fmaf(ly[i], sy[i] ? hrow[2][j] : hrow[1][j], hy[i] * (sy[i] ? hrow[1][j] : hrow[0][j]));
            float* o = out_logits + (((size_t)n * nc + c) * H + (y0 + i)) * W + x0;
            if ((W & 3) == 0) {
                __stcs(reinterpret_cast<float4*>(o), make_float4(val[0], val[1], val[2], val[3]));
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (x0 + j < W) o[j] = val[j];
            }
        }
    }
}

// Four consecutive labels of row `i` starting at element `e`, reduced to a row code of the confusion accumulator:
// -1 = unlabeled (label < 0; metric.py:79/:96, label+1 > 0 marks a labeled pixel), nc = label >= nclass (overflow row).
// LDT: FSCNN_U8 / FSCNN_I32 / FSCNN_I64.  `aligned`: vector loads (the row pitch keeps 4-element groups aligned).
template <int LDT>
__device__ __forceinline__ void load_codes4(const void* __restrict__ p, size_t e, bool aligned, int remaining, int nc, int (&code)[4]) {
    if (LDT == FSCNN_U8) {
        const unsigned char* q = reinterpret_cast<const unsigned char*>(p) + e;
        unsigned int v[4];
        if (aligned) {
            const uchar4 t = __ldg(reinterpret_cast<const uchar4*>(q));
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = j < remaining ? (unsigned int)__ldg(q + j) : 0xffffffffu;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) code[j] = v[j] == 0xffffffffu ? -1 : (int)min(v[j], (unsigned int)nc);
    } else if (LDT == FSCNN_I32) {
        const int* q = reinterpret_cast<const int*>(p) + e;
        int v[4];
        if (aligned) {
            const int4 t = __ldg(reinterpret_cast<const int4*>(q));
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = j < remaining ? __ldg(q + j) : -1;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) code[j] = v[j] < 0 ? -1 : min(v[j], nc);
    } else {
        const long long* q = reinterpret_cast<const long long*>(p) + e;
        int lo[4], hi[4];
        if (aligned) {
            const int4 a = __ldg(reinterpret_cast<const int4*>(q)), b = __ldg(reinterpret_cast<const int4*>(q) + 1);
            lo[0] = a.x; hi[0] = a.y; lo[1] = a.z; hi[1] = a.w; lo[2] = b.x; hi[2] = b.y; lo[3] = b.z; hi[3] = b.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const long long t = j < remaining ? __ldg(q + j) : -1ll;
                lo[j] = (int)(t & 0xffffffffll); hi[j] = (int)(t >> 32);
            }
        }
        // sign from the high word; any value that does not fit [0, nc) in the low word alone is >= nclass
#pragma unroll
        for (int j = 0; j < 4; ++j) code[j] = hi[j] < 0 ? -1 : ((hi[j] == 0 && (unsigned int)lo[j] < (unsigned int)nc) ? lo[j] : nc);
    }
}

// argmax mask (+ optional confusion counts) straight from the low-resolution logits.  LDT: label dtype, or -1 = no labels.
template <int LDT>
__global__ void __launch_bounds__(kATH, 768 / kATH)
upsample_argmax_kernel(const float* __restrict__ low, int nc, int ncp, void* __restrict__ mask, int mask_dtype,
                       const void* __restrict__ labels, unsigned long long* __restrict__ conf,
                       int hl, int wl, int H, int W, int use_smem_hist, int prune) {
    extern __shared__ __align__(16) float dynsm[];
    float* Ls = dynsm;                                            // [nc][kATR][kTC]
    unsigned int* hist = reinterpret_cast<unsigned int*>(dynsm + nc * kACS);
    __shared__ unsigned int blk_labeled, blk_correct, tile_amax;
    __shared__ unsigned int cellmask[kACR * kCC];
    constexpr bool do_hist = LDT >= 0;
    constexpr int esz = LDT == FSCNN_U8 ? 1 : (LDT == FSCNN_I32 ? 4 : 8);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = blockIdx.z;
    const int yb = blockIdx.y * (8 * kANW), xb = blockIdx.x * 128;
    const float scy = H > 1 ? (float)(hl - 1) / (float)(H - 1) : 0.f;
    const float scx = W > 1 ? (float)(wl - 1) / (float)(W - 1) : 0.f;
    const int rb = min((int)(scy * (float)yb), hl - 1);          // first staged low-res row / column
    const int cb = min((int)(scx * (float)xb), wl - 1);
    const int x0 = xb + lane * 4, y0 = yb + warp * 8;
    const bool live = (x0 < W) && (y0 < H);

    if (do_hist && live) {   // pull this thread's label rows towards L1 while the logits are staged and the classes compared
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (y0 + i < H) {
#if FSCNN_TAIL_PF == 1
                asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const char*>(labels) + (((size_t)n * H + (y0 + i)) * W + x0) * esz));
#elif FSCNN_TAIL_PF == 2
                asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(labels) + (((size_t)n * H + (y0 + i)) * W + x0) * esz));
#elif FSCNN_TAIL_PF == 3
                asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(reinterpret_cast<const char*>(labels) + (((size_t)n * H + (y0 + i)) * W + x0) * esz));
#endif
            }
    }
    // stage the low-res tile class-major: one float4 (4 classes of one pixel) per item, independent loads in flight
    int nonfinite = 0;
    float amax = 0.f;
    if (tid == 0) { tile_amax = 0u; blk_labeled = 0u; blk_correct = 0u; }
    {
        const int nv = ncp >> 2;
        for (int i = tid; i < kACS * nv; i += kATH) {
            const int px = i / nv, v = i - px * nv;
            const int r = px / kTC, q = px - r * kTC;
            const int rr = min(rb + r, hl - 1), qq = min(cb + q, wl - 1);
            const float4 t = __ldg(reinterpret_cast<const float4*>(low + (((size_t)n * hl + rr) * wl + qq) * ncp) + v);
            const int c = 4 * v;
            nonfinite |= !(fabsf(t.x) <= 3.4e38f) | !(fabsf(t.y) <= 3.4e38f) | !(fabsf(t.z) <= 3.4e38f) | !(fabsf(t.w) <= 3.4e38f);
            amax = fmaxf(fmaxf(amax, fmaxf(fabsf(t.x), fabsf(t.y))), fmaxf(fabsf(t.z), fabsf(t.w)));
            float* dst = Ls + c * kACS + px;
            dst[0] = t.x;
            if (c + 1 < nc) dst[kACS] = t.y;
            if (c + 2 < nc) dst[2 * kACS] = t.z;
            if (c + 3 < nc) dst[3 * kACS] = t.w;
        }
    }
    if (do_hist && use_smem_hist)
        for (int i = tid; i < (nc + 1) * (nc + 1); i += kATH) hist[i] = 0u;
    // NaN / Inf among the staged logits (padding classes are finite zeros) selects the exact-semantics generic loop; so do
    // more classes than the candidate bit mask holds
    const bool slow = (__syncthreads_or(nonfinite) != 0) || nc > 32;
    if (!slow) {
        // the tile's largest |logit| (non-negative floats order like their bit patterns), then one cell per thread over the
        // cells this tile's pixels interpolate in
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        if (lane == 0) atomicMax(&tile_amax, __float_as_uint(amax));
        __syncthreads();
        const float margin = 2e-6f * __uint_as_float(tile_amax) + 1e-30f;
        const int rn = min(min((int)(scy * (float)min(yb + 8 * kANW - 1, H - 1)), hl - 1) - rb + 1, kACR);
        const int qn = min(min((int)(scx * (float)min(xb + 127, W - 1)), wl - 1) - cb + 1, kCC);
        if (tid < rn * qn) {      // prune == 0 (tests, worst-case timing): every class stays a candidate everywhere
            const int R = tid / qn, Q = tid - R * qn;
            const int o = R * kTC + Q;      // the staged tile holds clamped copies, so the +1 neighbours are always in range
            cellmask[R * kCC + Q] = prune ? cell_survivors<false>(Ls, nc, ncp, o, o + 1, o + kTC, o + kTC + 1, margin)
                                          : (nc >= 32 ? 0xffffffffu : ((1u << nc) - 1u));
        }
        __syncthreads();
    }

    unsigned int labeled = 0, correct = 0;
    if (live) {
        // Separable interpolation weights, class-invariant: column j mixes the three staged columns c0..c0+2 with
        // (hx, lx, 0) or (0, hx, lx).  A zero weight adds an exact 0, so for finite logits each pixel is fma(lx, b, hx*a)
        // of its own two taps -- horizontal first, then vertical, like ATen.
        float wx[4][3];
        const int c0 = min((int)(scx * (float)x0), wl - 1);
        bool second_col = false;      // some column of this thread interpolates inside the cell to the right of c0's
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float fx = scx * (float)(x0 + j);
            const int q = min((int)fx, wl - 1);
            const float lx = fx - (float)q, hx = 1.f - lx;
            const bool s = (q - c0) != 0;
            second_col |= s;
            wx[j][0] = s ? 0.f : hx; wx[j][1] = s ? hx : lx; wx[j][2] = s ? lx : 0.f;
        }
        const int r0 = min((int)(scy * (float)y0), hl - 1);
        // tile-relative offsets of the 3 rows / 3 columns; the staged tile holds clamped copies beyond the image border,
        // which is ATen's x1 = x0 + (x0 < w-1)
        int ro[3], co[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            ro[k] = (r0 + k - rb) * kTC;
            co[k] = c0 + k - cb;
        }
        const unsigned int* cm = cellmask + (r0 - rb) * kCC + (c0 - cb);
        const uint32_t hist_s = (uint32_t)__cvta_generic_to_shared(hist);
        const int stride = nc + 1;
        int run_key = -1;            // run-length state of the metric counting, carried across both halves
        unsigned int run = 0;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {      // rows 0..3, then 4..7 of the thread's 8 x 4 block (one copy of the code)
            const int yh = y0 + 4 * half;
            float wa[4], wb[4];
            int kh = 0;               // rows of this half whose taps are staged rows (r0, r0+1); the rest use (r0+1, r0+2)
            unsigned int sy_mask = 0u;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float fy = scy * (float)(yh + i);
                const int r = min((int)fy, hl - 1);
                const float ly = fy - (float)r;
                wa[i] = 1.f - ly; wb[i] = ly;
                const bool s = (r - r0) != 0;
                kh += s ? 0 : 1;
                sy_mask |= (s ? 1u : 0u) << i;
            }
            unsigned int idx[4];      // winning class of 4 rows x 4 columns, one byte per pixel
            if (!slow) {
                // candidate classes of this half = union over the (at most 2 x 2) cells its pixels interpolate in; kh is the
                // same for every lane of the warp (they share y0)
                unsigned int cand = 0u;
                if (kh > 0) cand |= cm[0] | (second_col ? cm[1] : 0u);
                if (kh < 4) cand |= cm[kCC] | (second_col ? cm[kCC + 1] : 0u);
                if ((cand & (cand - 1u)) == 0u) {      // one class dominates: constant fill, no interpolation
                    const unsigned int fill = (unsigned int)(__ffs(cand) - 1) * 0x01010101u;
#pragma unroll
                    for (int i = 0; i < 4; ++i) idx[i] = fill;
                } else {
                    switch (kh) {
                        case 0: argmax_half_fast<0, false>(Ls, ncp, cand, ro, co, wx, wa, wb, idx); break;
                        case 1: argmax_half_fast<1, false>(Ls, ncp, cand, ro, co, wx, wa, wb, idx); break;
                        case 2: argmax_half_fast<2, false>(Ls, ncp, cand, ro, co, wx, wa, wb, idx); break;
                        case 3: argmax_half_fast<3, false>(Ls, ncp, cand, ro, co, wx, wa, wb, idx); break;
                        default: argmax_half_fast<4, false>(Ls, ncp, cand, ro, co, wx, wa, wb, idx); break;
                    }
                }
            } else {
                argmax_half_generic<false>(Ls, nc, ncp, ro, co, wx, wa, wb, sy_mask, idx);
            }

            // ---- write the mask ----
            if (mask) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (yh + i < H) {
                        const size_t off = ((size_t)n * H + (yh + i)) * W + x0;
                        const int b0 = idx[i] & 255u, b1 = (idx[i] >> 8) & 255u, b2 = (idx[i] >> 16) & 255u, b3 = idx[i] >> 24;
                        if (mask_dtype == FSCNN_U8) {
                            unsigned char* m = reinterpret_cast<unsigned char*>(mask) + off;
                            if ((W & 3) == 0) {
                                *reinterpret_cast<unsigned int*>(m) = idx[i];
                            } else {
                                if (x0 < W) m[0] = (unsigned char)b0;
                                if (x0 + 1 < W) m[1] = (unsigned char)b1;
                                if (x0 + 2 < W) m[2] = (unsigned char)b2;
                                if (x0 + 3 < W) m[3] = (unsigned char)b3;
                            }
                        } else if (mask_dtype == FSCNN_I32) {
                            int* m = reinterpret_cast<int*>(mask) + off;
                            if ((W & 3) == 0) {
                                *reinterpret_cast<int4*>(m) = make_int4(b0, b1, b2, b3);
                            } else {
                                if (x0 < W) m[0] = b0;
                                if (x0 + 1 < W) m[1] = b1;
                                if (x0 + 2 < W) m[2] = b2;
                                if (x0 + 3 < W) m[3] = b3;
                            }
                        } else {
                            long long* m = reinterpret_cast<long long*>(mask) + off;
                            if ((W & 3) == 0) {
                                *reinterpret_cast<longlong2*>(m) = make_longlong2(b0, b1);
                                *reinterpret_cast<longlong2*>(m + 2) = make_longlong2(b2, b3);
                            } else {
                                if (x0 < W) m[0] = b0;
                                if (x0 + 1 < W) m[1] = b1;
                                if (x0 + 2 < W) m[2] = b2;
                                if (x0 + 3 < W) m[3] = b3;
                            }
                        }
                    }
                }
            }

            // ---- SegmentationMetric counting: run-length aggregated shared-memory reductions, branch free ----
            // A thread walks its pixels row by row and issues ONE predicated red.shared per run of equal (label, prediction)
            // pairs: real label maps give a couple per thread, uniform-random labels one per pixel; no lane ever branches.
            if (do_hist) {
                int code[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {       // all four row loads in flight before the first one is consumed
#pragma unroll
                    for (int j = 0; j < 4; ++j) code[i][j] = -1;
                    if (yh + i < H) load_codes4<LDT < 0 ? 0 : LDT>(labels, ((size_t)n * H + (yh + i)) * W + x0, (W & 3) == 0, W - x0, nc, code[i]);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int row = code[i][j];
                        const bool ok = row >= 0;
                        const int col = (int)((idx[i] >> (8 * j)) & 255u);
                        const int key = row * stride + col;
                        const bool brk = ok && key != run_key;      // a new run starts at this pixel (run_key starts at -1)
                        if (use_smem_hist) {
                            // the run that ends here; before the first labeled pixel that is `0 pixels of bin 0`, which adds nothing
                            red_shared_if(brk, hist_s + 4u * (uint32_t)max(run_key, 0), run);
                        } else {
                            labeled += ok ? 1u : 0u;
                            correct += (row == col) ? 1u : 0u;                      // col >= 0, so an unlabeled pixel never matches
                            if (brk && run != 0u) atomicAdd(conf + run_key, (unsigned long long)run);
                        }
                        run = brk ? 1u : run + (ok ? 1u : 0u);
                        run_key = brk ? key : run_key;
                    }
                }
            }
        }
        if (do_hist) {
            if (use_smem_hist) red_shared_if(run != 0u, hist_s + 4u * (uint32_t)max(run_key, 0), run);
            else if (run) atomicAdd(conf + run_key, (unsigned long long)run);
        }
    }
    if (!do_hist) return;
    __syncthreads();
    const int nb = (nc + 1) * (nc + 1);
    if (use_smem_hist) {
        // the tile's bins go to the global accumulator; `labeled` is their sum and `correct` the sum of the diagonal
        // (bin = row * (nc + 1) + col, row == col < nc  <=>  bin % (nc + 2) == 0 and bin < nc * (nc + 2))
        for (int i = tid; i < nb; i += kATH) {
            const unsigned int v = hist[i];
            if (v) atomicAdd(conf + i, (unsigned long long)v);
            labeled += v;
            correct += (i % (nc + 2) == 0 && i < nc * (nc + 2)) ? v : 0u;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        labeled += __shfl_xor_sync(0xffffffffu, labeled, o);
        correct += __shfl_xor_sync(0xffffffffu, correct, o);
    }
    if (lane == 0) { atomicAdd(&blk_labeled, labeled); atomicAdd(&blk_correct, correct); }
    __syncthreads();
    if (tid == 0) {
        if (blk_labeled) atomicAdd(conf + nb, (unsigned long long)blk_labeled);
        if (blk_correct) atomicAdd(conf + nb + 1, (unsigned long long)blk_correct);
    }
}

// ---- the same stage as a PERSISTENT kernel whose inputs arrive by TMA, one tile ahead -----------------------------------
// ncu on the direct kernel above (profiles/r02_tail_kernel.md): a fifth of its time is warps waiting for the label loads they
// issued a moment earlier and another tenth for the logits of the tile; nothing overlaps them because every CTA does
// load -> compare -> load -> count once and exits.  Here a CTA walks tiles (tile = blockIdx.x + k * gridDim.x) and the copy engine
// runs ahead of it: ONE cp.async.bulk.tensor brings the 12 x 20 x padded-classes box of low-resolution logits of the next tile
// into a raw buffer (pixel-major, the NHWC tensor's own layout; outside the tensor the engine zero-fills) as soon as the
// current tile has been transposed out of it into the class-major, border-clamped tile the comparison code wants, and the
// 64 x 128 label box of the next tile is requested as soon as this tile's labels have been counted.  The confusion bins stay
// in shared memory for the CTA's whole life and reach the global accumulator once.  Needs 16-byte aligned label rows
// (W * element size a multiple of 16); other shapes keep the direct kernel.
template <int LDT>
__global__ void __launch_bounds__(kThreads, LDT == FSCNN_I64 ? 2 : 3)
upsample_argmax_tma_kernel(const __grid_constant__ CUtensorMap lmap, const __grid_constant__ CUtensorMap labmap, int nc, int ncp,
                           void* __restrict__ mask, int mask_dtype, unsigned long long* __restrict__ conf, int hl, int wl, int H, int W,
                           int tiles_x, int tiles_y, int ntiles, int prune) {
    constexpr bool do_hist = LDT >= 0;
    constexpr int esz = LDT == FSCNN_U8 ? 1 : (LDT == FSCNN_I32 ? 4 : 8);
    extern __shared__ __align__(128) unsigned char tsm[];
    const int lbytes = kTR * kTC * ncp * 4;                        // one logits box
    const int lstride = (lbytes + 127) & ~127;
    const float* raw = reinterpret_cast<const float*>(tsm);        // [12][20][ncp], as the copy engine delivers it
    float* Ls = reinterpret_cast<float*>(tsm + lstride);           // [nc][12][20], clamped at the image border
    unsigned char* labs = tsm + 2 * lstride;                       // [64][128] labels of the current tile
    unsigned int* hist = reinterpret_cast<unsigned int*>(labs + (do_hist ? 64 * 128 * esz : 0));
    __shared__ __align__(8) uint64_t bar_l, bar_lab;
    __shared__ unsigned int tile_amax[2], blk_labeled, blk_correct;
    __shared__ unsigned int cellmask[kCR * kCC];
    static_assert(kTR == kATR, "the TMA kernel shares the direct kernel's class-major tile");

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float scy = H > 1 ? (float)(hl - 1) / (float)(H - 1) : 0.f;
    const float scx = W > 1 ? (float)(wl - 1) / (float)(W - 1) : 0.f;
    const int nb = (nc + 1) * (nc + 1);
    auto tile_coords = [&](int tile, int& n, int& yb, int& xb) {
        const int tx = tile % tiles_x, r = tile / tiles_x;
        n = r / tiles_y; yb = (r % tiles_y) * 64; xb = tx * 128;
    };
    auto issue_logits = [&](int tile) {
        int n, yb, xb;
        tile_coords(tile, n, yb, xb);
        const int rb = min((int)(scy * (float)yb), hl - 1), cb = min((int)(scx * (float)xb), wl - 1);
        mbar_arrive_expect_tx(&bar_l, (uint32_t)lbytes);
        tma_load_4d(smem_u32(raw), &lmap, 0, cb, rb, n, &bar_l);
    };
    auto issue_labels = [&](int tile) {
        int n, yb, xb;
        tile_coords(tile, n, yb, xb);
        mbar_arrive_expect_tx(&bar_lab, 64u * 128u * esz);
        tma_load_3d(smem_u32(labs), &labmap, xb, yb, n, &bar_lab);
    };

    pdl_launch_dependents();
    if (tid == 0) {
        mbar_init(&bar_l, 1); mbar_init(&bar_lab, 1);
        fence_mbar_init();
        tile_amax[0] = 0u; tile_amax[1] = 0u; blk_labeled = 0u; blk_correct = 0u;
    }
    if (do_hist)
        for (int i = tid; i < nb; i += kThreads) hist[i] = 0u;
    __syncthreads();
    pdl_wait();      // the low-resolution logits are the previous stage's output; mask / confusion writes follow them
    if (tid == 0) {
        tma_prefetch_desc(&lmap);
        issue_logits(blockIdx.x);
        if (do_hist) { tma_prefetch_desc(&labmap); issue_labels(blockIdx.x); }
    }

    int it = 0;
#pragma unroll 1
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        int n, yb, xb;
        tile_coords(tile, n, yb, xb);
        const int rb = min((int)(scy * (float)yb), hl - 1);          // first staged low-res row / column
        const int cb = min((int)(scx * (float)xb), wl - 1);
        const int rmax = hl - 1 - rb, qmax = wl - 1 - cb;            // tile-relative clamps of the staged box
        mbar_wait(&bar_l, it & 1);
        // transpose the box into the class-major tile (clamped copies beyond the image border = ATen's x1 = min(x0 + 1, w - 1)),
        // checking for NaN / Inf (-> the exact-semantics generic loop) and taking the largest |logit| (-> the margin) on the way.
        // Items run pixel-fastest: conflict-free stores, 16-byte loads 4 * ncp bytes apart.
        int nonfinite = 0;
        float amax = 0.f;
        {
            const int nv = ncp >> 2;
            for (int i = tid; i < kTR * kTC * nv; i += kThreads) {
                const int v = i / (kTR * kTC), px = i - v * (kTR * kTC);
                const int r = px / kTC, q = px - r * kTC;
                const float4 t = *reinterpret_cast<const float4*>(raw + (min(r, rmax) * kTC + min(q, qmax)) * ncp + 4 * v);
                const int c = 4 * v;
                nonfinite |= !(fabsf(t.x) <= 3.4e38f) | !(fabsf(t.y) <= 3.4e38f) | !(fabsf(t.z) <= 3.4e38f) | !(fabsf(t.w) <= 3.4e38f);
                amax = fmaxf(fmaxf(amax, fmaxf(fabsf(t.x), fabsf(t.y))), fmaxf(fabsf(t.z), fabsf(t.w)));
                float* dst = Ls + c * kACS + px;
                dst[0] = t.x;
                if (c + 1 < nc) dst[kACS] = t.y;
                if (c + 2 < nc) dst[2 * kACS] = t.z;
                if (c + 3 < nc) dst[3 * kACS] = t.w;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        if (lane == 0) atomicMax(&tile_amax[it & 1], __float_as_uint(amax));
        if (tid == 0) tile_amax[(it + 1) & 1] = 0u;
        const bool slow = (__syncthreads_or(nonfinite) != 0) || nc > 32;
        // the raw box has been consumed: the copy engine may fetch the next tile's while this one is compared and counted
        if (tid == 0 && tile + (int)gridDim.x < ntiles) issue_logits(tile + gridDim.x);
        if (!slow) {
            const float margin = 2e-6f * __uint_as_float(tile_amax[it & 1]) + 1e-30f;
            const int rn = min(min((int)(scy * (float)min(yb + 63, H - 1)), hl - 1) - rb + 1, kCR);
            const int qn = min(min((int)(scx * (float)min(xb + 127, W - 1)), wl - 1) - cb + 1, kCC);
            if (tid < rn * qn) {
                const int R = tid / qn, Q = tid - R * qn;
                const int o = R * kTC + Q;
                cellmask[R * kCC + Q] = prune ? cell_survivors<false>(Ls, nc, ncp, o, o + 1, o + kTC, o + kTC + 1, margin)
                                              : (nc >= 32 ? 0xffffffffu : ((1u << nc) - 1u));
            }
            __syncthreads();
        }

        const int x0 = xb + lane * 4, y0 = yb + warp * 8;
        const bool live = (x0 < W) && (y0 < H);
        if (live) {
            float wx[4][3];
            const int c0 = min((int)(scx * (float)x0), wl - 1);
            bool second_col = false;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float fx = scx * (float)(x0 + j);
                const int q = min((int)fx, wl - 1);
                const float lx = fx - (float)q, hx = 1.f - lx;
                const bool s = (q - c0) != 0;
                second_col |= s;
                wx[j][0] = s ? 0.f : hx; wx[j][1] = s ? hx : lx; wx[j][2] = s ? lx : 0.f;
            }
            const int r0 = min((int)(scy * (float)y0), hl - 1);
            int ro[3], co[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                ro[k] = (r0 + k - rb) * kTC;
                co[k] = c0 + k - cb;
            }
            const unsigned int* cm = cellmask + (r0 - rb) * kCC + (c0 - cb);
            const uint32_t hist_s = smem_u32(hist);
            const int stride = nc + 1;
            int run_key = -1;
            unsigned int run = 0;
            bool labels_ready = false;
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                const int yh = y0 + 4 * half;
                float wa[4], wb[4];
                int kh = 0;
                unsigned int sy_mask = 0u;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float fy = scy * (float)(yh + i);
                    const int r = min((int)fy, hl - 1);
                    const float ly = fy - (float)r;
                    wa[i] = 1.f - ly; wb[i] = ly;
                    const bool s = (r - r0) != 0;
                    kh += s ? 0 : 1;
                    sy_mask |= (s ? 1u : 0u) << i;
                }
                unsigned int idx[4];
                if (!slow) {
                    unsigned int cand = 0u;
                    if (kh > 0) cand |= cm[0] | (second_col ? cm[1] : 0u);
                    if (kh < 4) cand |= cm[kCC] | (second_col ? cm[kCC + 1] : 0u);
                    if ((cand & (cand - 1u)) == 0u) {
                        const unsigned int fill = (unsigned int)(__ffs(cand) - 1) * 0x01010101u;
#pragma unroll
                        for (int i = 0; i < 4; ++i) idx[i] = fill;
                    } else {
                        switch (kh) {
                            case 0: argmax_half_fast<0, false>(Ls, ncp, cand, ro, co, wx, wa, wb, idx); break;
                            case 1: argmax_half_fast<1, false>(Ls, ncp, cand, ro, co, wx, wa, wb, idx); break;
                            case 2: argmax_half_fast<2, false>(Ls, ncp, cand, ro, co, wx, wa, wb, idx); break;
                            case 3: argmax_half_fast<3, false>(Ls, ncp, cand, ro, co, wx, wa, wb, idx); break;
                            default: argmax_half_fast<4, false>(Ls, ncp, cand, ro, co, wx, wa, wb, idx); break;
                        }
                    }
                } else {
                    argmax_half_generic<false>(Ls, nc, ncp, ro, co, wx, wa, wb, sy_mask, idx);
                }
                if (mask) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if (yh + i < H) {
                            const size_t off = ((size_t)n * H + (yh + i)) * W + x0;
                            const int b0 = idx[i] & 255u, b1 = (idx[i] >> 8) & 255u, b2 = (idx[i] >> 16) & 255u, b3 = idx[i] >> 24;
                            if (mask_dtype == FSCNN_U8) {
                                unsigned char* m = reinterpret_cast<unsigned char*>(mask) + off;
                                if ((W & 3) == 0) {
                                    *reinterpret_cast<unsigned int*>(m) = idx[i];
                                } else {
                                    if (x0 < W) m[0] = (unsigned char)b0;
                                    if (x0 + 1 < W) m[1] = (unsigned char)b1;
                                    if (x0 + 2 < W) m[2] = (unsigned char)b2;
                                    if (x0 + 3 < W) m[3] = (unsigned char)b3;
                                }
                            } else if (mask_dtype == FSCNN_I32) {
                                int* m = reinterpret_cast<int*>(mask) + off;
                                if ((W & 3) == 0) {
                                    *reinterpret_cast<int4*>(m) = make_int4(b0, b1, b2, b3);
                                } else {
                                    if (x0 < W) m[0] = b0;
                                    if (x0 + 1 < W) m[1] = b1;
                                    if (x0 + 2 < W) m[2] = b2;
                                    if (x0 + 3 < W) m[3] = b3;
                                }
                            } else {
                                long long* m = reinterpret_cast<long long*>(mask) + off;
                                if ((W & 3) == 0) {
                                    *reinterpret_cast<longlong2*>(m) = make_longlong2(b0, b1);
                                    *reinterpret_cast<longlong2*>(m + 2) = make_longlong2(b2, b3);
                                } else {
                                    if (x0 < W) m[0] = b0;
                                    if (x0 + 1 < W) m[1] = b1;
                                    if (x0 + 2 < W) m[2] = b2;
                                    if (x0 + 3 < W) m[3] = b3;
                                }
                            }
                        }
                    }
                }
                if (do_hist) {
                    if (!labels_ready) { mbar_wait(&bar_lab, it & 1); labels_ready = true; }
                    int code[4][4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) code[i][j] = -1;
                        // the label box sits in shared memory as [64 rows][128 columns]; rows / columns outside the image were
                        // zero-filled by the copy engine and are skipped here
                        if (yh + i < H) {
                            const unsigned char* lp = labs + ((size_t)(warp * 8 + 4 * half + i) * 128 + lane * 4) * esz;
                            if (LDT == FSCNN_U8) {
                                const unsigned int v = *reinterpret_cast<const unsigned int*>(lp);
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const unsigned int t = (v >> (8 * j)) & 255u;
                                    code[i][j] = (int)min(t, (unsigned int)nc);
                                }
                            } else if (LDT == FSCNN_I32) {
                                const int4 v = *reinterpret_cast<const int4*>(lp);
                                const int t[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                                for (int j = 0; j < 4; ++j) code[i][j] = t[j] < 0 ? -1 : min(t[j], nc);
                            } else {
                                const int4 a = *reinterpret_cast<const int4*>(lp), b = *reinterpret_cast<const int4*>(lp + 16);
                                const int lo[4] = {a.x, a.z, b.x, b.z}, hi[4] = {a.y, a.w, b.y, b.w};
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    code[i][j] = hi[j] < 0 ? -1 : ((hi[j] == 0 && (unsigned int)lo[j] < (unsigned int)nc) ? lo[j] : nc);
                            }
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                if (x0 + j >= W) code[i][j] = -1;
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int row = code[i][j];
                            const bool ok = row >= 0;
                            const int col = (int)((idx[i] >> (8 * j)) & 255u);
                            const int key = row * stride + col;
                            const bool brk = ok && key != run_key;
                            red_shared_if(brk, hist_s + 4u * (uint32_t)max(run_key, 0), run);
                            run = brk ? 1u : run + (ok ? 1u : 0u);
                            run_key = brk ? key : run_key;
                        }
                    }
                }
            }
            if (do_hist) red_shared_if(run != 0u, hist_s + 4u * (uint32_t)max(run_key, 0), run);
        } else if (do_hist) {
            mbar_wait(&bar_lab, it & 1);      // keep every thread's view of the barrier phase in step
        }
        __syncthreads();      // labels counted; the class-major tile and the cell masks are free
        if (do_hist && tid == 0 && tile + (int)gridDim.x < ntiles) issue_labels(tile + gridDim.x);
    }
    if (!do_hist) return;
    // the CTA's bins go to the global accumulator once; `labeled` is their sum and `correct` the sum of the diagonal
    unsigned int labeled = 0, correct = 0;
    for (int i = tid; i < nb; i += kThreads) {
        const unsigned int v = hist[i];
        if (v) atomicAdd(conf + i, (unsigned long long)v);
        labeled += v;
        correct += (i % (nc + 2) == 0 && i < nc * (nc + 2)) ? v : 0u;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        labeled += __shfl_xor_sync(0xffffffffu, labeled, o);
        correct += __shfl_xor_sync(0xffffffffu, correct, o);
    }
    if (lane == 0) { atomicAdd(&blk_labeled, labeled); atomicAdd(&blk_correct, correct); }
    __syncthreads();
    if (tid == 0) {
        if (blk_labeled) atomicAdd(conf + nb, (unsigned long long)blk_labeled);
        if (blk_correct) atomicAdd(conf + nb + 1, (unsigned long long)blk_correct);
    }
}

static size_t up_smem_bytes(int nc, bool hist, int rows = kTR) {
    size_t b = (size_t)nc * rows * kTC * sizeof(float);
    if (hist) b += (size_t)(nc + 1) * (nc + 1) * sizeof(unsigned int);
    return b;
}

cudaError_t launch_up_logits(const float* low, int nc, int ncp, float* out, int n, int hl, int wl, int h, int w,
                             cudaStream_t s) {
    if ((double)(hl - 1) * 7.0 > (double)(h - 1) || (double)(wl - 1) * 7.0 > (double)(w - 1)) return cudaErrorInvalidValue;
    static unsigned long long configured = 0;
    static size_t configured_bytes = 48 * 1024;
    const size_t smem = up_smem_bytes(nc, false);
    if (smem > configured_bytes) { configured = 0; configured_bytes = smem; }
    if (smem > 48 * 1024) {
        cudaError_t e = ensure_dyn_smem(upsample_logits_kernel, configured_bytes, configured);
        if (e != cudaSuccess) return e;
    }
    dim3 grid(ceil_div(w, 128), ceil_div(h, 64), n);
    upsample_logits_kernel<<<grid, kThreads, smem, s>>>(low, nc, ncp, out, hl, wl, h, w);
    return cudaGetLastError();
}

template <int LDT>
static cudaError_t run_up_argmax(const float* low, int nc, int ncp, void* mask, int mask_dtype, const void* labels,
                                 unsigned long long* conf, int n, int hl, int wl, int h, int w, cudaStream_t s, bool prune) {
    const int smem_hist = (LDT >= 0) && ((nc + 1) * (nc + 1) <= kHistMaxBins);
    static unsigned long long configured = 0;
    static size_t configured_bytes = 48 * 1024;
    const size_t smem = up_smem_bytes(nc, smem_hist, kATR);
    if (smem > configured_bytes) { configured = 0; configured_bytes = smem; }
    if (smem > 48 * 1024) {
        cudaError_t e = ensure_dyn_smem(upsample_argmax_kernel<LDT>, configured_bytes, configured);
        if (e != cudaSuccess) return e;
    }
    dim3 grid(ceil_div(w, 128), ceil_div(h, 8 * kANW), n);
    upsample_argmax_kernel<LDT><<<grid, kATH, smem, s>>>(low, nc, ncp, mask, mask_dtype, labels, conf, hl, wl, h, w, smem_hist,
                                                             prune ? 1 : 0);
    return cudaGetLastError();
}

// the persistent TMA kernel when the shapes allow it: <= 32 classes worth of candidate mask is not required (the generic loop
// runs inside it too), but the shared-memory bins are, and the label rows must be 16-byte aligned for the copy engine
template <int LDT>
static cudaError_t run_up_argmax_tma(const float* low, int nc, int ncp, void* mask, int mask_dtype, const void* labels,
                                     unsigned long long* conf, int n, int hl, int wl, int h, int w, cudaStream_t s, bool prune, bool* used) {
    constexpr int esz = LDT == FSCNN_U8 ? 1 : (LDT == FSCNN_I32 ? 4 : 8);
    *used = false;
    if (kTR != 12 || (LDT >= 0 && (nc + 1) * (nc + 1) > kHistMaxBins)) return cudaSuccess;
    if (LDT >= 0 && ((reinterpret_cast<uintptr_t>(labels) & 15) || ((size_t)w * esz) % 16 || ((size_t)h * w * esz) % 16)) return cudaSuccess;
    if ((reinterpret_cast<uintptr_t>(low) & 15) || (ncp & 3)) return cudaSuccess;
    const size_t lbytes = ((size_t)kTR * kTC * ncp * 4 + 127) & ~(size_t)127;
    const size_t smem = 2 * lbytes + (LDT >= 0 ? (size_t)64 * 128 * esz + (size_t)(nc + 1) * (nc + 1) * 4 : 0) + 128;   // raw box + class-major tile
    if (smem > 200 * 1024) return cudaSuccess;
    CUtensorMap lmap, labmap;
    {
        const cuuint64_t dims[4] = {(cuuint64_t)ncp, (cuuint64_t)wl, (cuuint64_t)hl, (cuuint64_t)n};
        const cuuint64_t strides[3] = {(cuuint64_t)ncp * 4, (cuuint64_t)wl * ncp * 4, (cuuint64_t)hl * wl * ncp * 4};
        const cuuint32_t box[4] = {(cuuint32_t)ncp, (cuuint32_t)kTC, (cuuint32_t)kTR, 1};
        if (ncp > 256 || make_tiled_map(&lmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, low, dims, strides, box) != cudaSuccess) return cudaSuccess;
    }
    labmap = lmap;
    if (LDT >= 0) {
        const cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
        const cuuint64_t strides[2] = {(cuuint64_t)w * esz, (cuuint64_t)h * w * esz};
        const cuuint32_t box[3] = {128, 64, 1};
        const CUtensorMapDataType dt = LDT == FSCNN_U8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : (LDT == FSCNN_I32 ? CU_TENSOR_MAP_DATA_TYPE_INT32 : CU_TENSOR_MAP_DATA_TYPE_INT64);
        if (make_tiled_map(&labmap, dt, 3, labels, dims, strides, box) != cudaSuccess) return cudaSuccess;
    }
    static unsigned long long configured = 0;
    static size_t configured_bytes = 0;
    static int ctas_per_sm[64] = {};
    if (smem > configured_bytes) { configured = 0; configured_bytes = smem; }
    cudaError_t e = ensure_dyn_smem(upsample_argmax_tma_kernel<LDT>, configured_bytes, configured);
    if (e != cudaSuccess) return e;
    int dev = 0;
    cudaGetDevice(&dev);
    int& occ = ctas_per_sm[dev & 63];
    if (occ <= 0 && (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, upsample_argmax_tma_kernel<LDT>, kThreads, configured_bytes) != cudaSuccess || occ <= 0))
        occ = 1;
    const int tiles_x = ceil_div(w, 128), tiles_y = ceil_div(h, 64), ntiles = tiles_x * tiles_y * n;
    const int grid = min(ntiles, num_sms() * occ);
    *used = true;
    return launch_pdl(upsample_argmax_tma_kernel<LDT>, grid, kThreads, smem, s, lmap, labmap, nc, ncp, mask, mask_dtype, conf, hl, wl, h, w, tiles_x,
                      tiles_y, ntiles, prune ? 1 : 0);
}

static int g_tail_tma = -1;   // FSCNN_TAIL_TMA=0 keeps the direct kernel (A/B timing)

cudaError_t launch_up_argmax(const float* low, int nc, int ncp, void* mask, int mask_dtype, const void* labels,
                             int label_dtype, unsigned long long* conf, int n, int hl, int wl, int h, int w,
                             cudaStream_t s, bool prune) {
    // a 128-pixel tile row must interpolate inside at most 18 low-res cells (17 + the one to the right): ratio < 17/124
    if ((double)(hl - 1) * 7.3 > (double)(h - 1) || (double)(wl - 1) * 7.3 > (double)(w - 1)) return cudaErrorInvalidValue;
    if (labels && label_dtype != FSCNN_U8 && label_dtype != FSCNN_I32 && label_dtype != FSCNN_I64) return cudaErrorInvalidValue;
    if (g_tail_tma < 0) { const char* v = getenv("FSCNN_TAIL_TMA"); g_tail_tma = (v && v[0] == '0') ? 0 : 1; }
    if (g_tail_tma) {
        bool used = false;
        cudaError_t e;
        if (!labels) e = run_up_argmax_tma<-1>(low, nc, ncp, mask, mask_dtype, nullptr, nullptr, n, hl, wl, h, w, s, prune, &used);
        else if (label_dtype == FSCNN_U8) e = run_up_argmax_tma<FSCNN_U8>(low, nc, ncp, mask, mask_dtype, labels, conf, n, hl, wl, h, w, s, prune, &used);
        else if (label_dtype == FSCNN_I32) e = run_up_argmax_tma<FSCNN_I32>(low, nc, ncp, mask, mask_dtype, labels, conf, n, hl, wl, h, w, s, prune, &used);
        else e = run_up_argmax_tma<FSCNN_I64>(low, nc, ncp, mask, mask_dtype, labels, conf, n, hl, wl, h, w, s, prune, &used);
        if (e != cudaSuccess || used) return e;
    }
    if (!labels) return run_up_argmax<-1>(low, nc, ncp, mask, mask_dtype, nullptr, nullptr, n, hl, wl, h, w, s, prune);
    if (label_dtype == FSCNN_U8) return run_up_argmax<FSCNN_U8>(low, nc, ncp, mask, mask_dtype, labels, conf, n, hl, wl, h, w, s, prune);
    if (label_dtype == FSCNN_I32) return run_up_argmax<FSCNN_I32>(low, nc, ncp, mask, mask_dtype, labels, conf, n, hl, wl, h, w, s, prune);
    return run_up_argmax<FSCNN_I64>(low, nc, ncp, mask, mask_dtype, labels, conf, n, hl, wl, h, w, s, prune);
}

// ---- SegmentationMetric counting on existing class maps (utils/metric.py:73-105) ----
__global__ void __launch_bounds__(kThreads)
confusion_kernel(const void* __restrict__ pred, int pred_dtype, const void* __restrict__ label, int label_dtype,
                 long long npix, int nc, unsigned long long* __restrict__ conf, int use_smem_hist) {
    extern __shared__ unsigned int hist_dyn[];
    __shared__ unsigned int blk_labeled, blk_correct;
    const int tid = threadIdx.x;
    const int nb = (nc + 1) * (nc + 1);
    if (use_smem_hist)
        for (int i = tid; i < nb; i += kThreads) hist_dyn[i] = 0u;
    if (tid == 0) { blk_labeled = 0u; blk_correct = 0u; }
    __syncthreads();
    const ConfSink sink{use_smem_hist ? hist_dyn : nullptr, conf, nc};
    unsigned int labeled = 0, correct = 0;
    // each thread walks a contiguous run of 16 pixels so that equal neighbours aggregate
    const long long chunk = 16;
    for (long long base = ((long long)blockIdx.x * kThreads + tid) * chunk; base < npix;
         base += (long long)gridDim.x * kThreads * chunk) {
        int run_row = -1, run_col = 0;
        unsigned int run = 0;
        const long long end = base + chunk < npix ? base + chunk : npix;
        for (long long i = base; i < end; ++i) {
            const long long lab = load_label(label, label_dtype, (size_t)i);
            if (lab < 0) continue;
            const long long pr = load_label(pred, pred_dtype, (size_t)i);
            const int row = lab < nc ? (int)lab : nc;
            const int col = (pr >= 0 && pr < nc) ? (int)pr : nc;   // preds outside the class range: overflow column
            labeled += 1;
            correct += (lab == pr);
            if (row == run_row && col == run_col) { run += 1; continue; }
            if (run) sink.add(run_row, run_col, run);
            run_row = row; run_col = col; run = 1;
        }
        if (run) sink.add(run_row, run_col, run);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        labeled += __shfl_xor_sync(0xffffffffu, labeled, o);
        correct += __shfl_xor_sync(0xffffffffu, correct, o);
    }
    if ((tid & 31) == 0) { atomicAdd(&blk_labeled, labeled); atomicAdd(&blk_correct, correct); }
    __syncthreads();
    if (use_smem_hist)
        for (int i = tid; i < nb; i += kThreads) {
            const unsigned int v = hist_dyn[i];
            if (v) atomicAdd(conf + i, (unsigned long long)v);
        }
    if (tid == 0) {
        if (blk_labeled) atomicAdd(conf + nb, (unsigned long long)blk_labeled);
        if (blk_correct) atomicAdd(conf + nb + 1, (unsigned long long)blk_correct);
    }
}

cudaError_t launch_confusion(const void* pred, int pred_dtype, const void* label, int label_dtype, long long npix, int nc,
                             unsigned long long* conf, cudaStream_t s) {
    if (npix <= 0) return cudaSuccess;
    const int nb = (nc + 1) * (nc + 1);
    const int smem_hist = nb <= kHistMaxBins;
    long long blocks = (npix + (long long)kThreads * 16 - 1) / ((long long)kThreads * 16);
    if (blocks > 148 * 8) blocks = 148 * 8;   // grid-stride beyond 8 CTAs per SM
    confusion_kernel<<<(unsigned)blocks, kThreads, smem_hist ? nb * sizeof(unsigned int) : 0, s>>>(
        pred, pred_dtype, label, label_dtype, npix, nc, conf, smem_hist);
    return cudaGetLastError();
}


// ---- palette rendering of a class map (the step right after the path: reference utils/visualize.py:7-36 puts a
//      palette on the uint8 mask with PIL on the host; here rgb[p] = palette[mask[p]] in one pass on the device) ----
struct Palette { unsigned char rgb[768]; };

__global__ void __launch_bounds__(kThreads)
colorize_kernel(const void* __restrict__ mask, int dtype, long long npix, Palette pal, unsigned char* __restrict__ rgb) {
    __shared__ unsigned char ps[768];
    for (int i = threadIdx.x; i < 768; i += kThreads) ps[i] = pal.rgb[i];
    __syncthreads();
    for (long long q = ((long long)blockIdx.x * kThreads + threadIdx.x) * 4; q < npix; q += (long long)gridDim.x * kThreads * 4) {
        unsigned char o[12];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long v = q + j < npix ? load_label(mask, dtype, (size_t)(q + j)) : 0;
            const int c = (int)(v & 255);
            o[3 * j] = ps[3 * c]; o[3 * j + 1] = ps[3 * c + 1]; o[3 * j + 2] = ps[3 * c + 2];
        }
        if (q + 4 <= npix) {   // 12 bytes at a 4-byte aligned offset (q is a multiple of 4)
            uint32_t* dst = reinterpret_cast<uint32_t*>(rgb + 3 * q);
            dst[0] = o[0] | (o[1] << 8) | (o[2] << 16) | ((uint32_t)o[3] << 24);
            dst[1] = o[4] | (o[5] << 8) | (o[6] << 16) | ((uint32_t)o[7] << 24);
            dst[2] = o[8] | (o[9] << 8) | (o[10] << 16) | ((uint32_t)o[11] << 24);
        } else {
            for (int j = 0; q + j < npix; ++j) { rgb[3 * (q + j)] = o[3 * j]; rgb[3 * (q + j) + 1] = o[3 * j + 1]; rgb[3 * (q + j) + 2] = o[3 * j + 2]; }
        }
    }
}

cudaError_t launch_colorize(const void* mask, int dtype, long long npix, const unsigned char* palette768, unsigned char* rgb,
                            cudaStream_t s) {
    if (npix <= 0) return cudaSuccess;
    Palette pal;
    for (int i = 0; i < 768; ++i) pal.rgb[i] = palette768[i];
    long long blocks = (npix + kThreads * 4 - 1) / (kThreads * 4);
    if (blocks > 148 * 16) blocks = 148 * 16;
    colorize_kernel<<<(unsigned)blocks, kThreads, 0, s>>>(mask, dtype, npix, pal, rgb);
    return cudaGetLastError();
}


// ---- overlay of a class map on the frame it was predicted from (reference demo_tusimple.py:87-104 create_overlay, the host
//      numpy version: overlay[m] = (1 - alpha) * image[m] + alpha * colour[m], truncated to uint8, for the pixels m of the drawn
//      classes).  One pass on the device; float64 arithmetic like numpy's so that the truncation lands on the same side. ----
struct OverlayTab { unsigned char rgb[768]; unsigned int draw[8]; };      // palette + bit c set = class c is drawn

__global__ void __launch_bounds__(kThreads)
overlay_kernel(const unsigned char* __restrict__ image, const void* __restrict__ mask, int dtype, long long npix, OverlayTab tab,
               double alpha, unsigned char* __restrict__ out) {
    __shared__ unsigned char ps[768];
    __shared__ unsigned int draw[8];
    for (int i = threadIdx.x; i < 768; i += kThreads) ps[i] = tab.rgb[i];
    if (threadIdx.x < 8) draw[threadIdx.x] = tab.draw[threadIdx.x];
    __syncthreads();
    for (long long q = (long long)blockIdx.x * kThreads + threadIdx.x; q < npix; q += (long long)gridDim.x * kThreads) {
        const long long v = load_label(mask, dtype, (size_t)q);
        const bool on = v >= 0 && v < 256 && ((draw[v >> 5] >> (v & 31)) & 1u);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const unsigned char px = image[3 * q + c];
            out[3 * q + c] = on ? (unsigned char)((1.0 - alpha) * (double)px + alpha * (double)ps[3 * (int)v + c]) : px;
        }
    }
}

cudaError_t launch_overlay(const unsigned char* image, const void* mask, int dtype, long long npix, const unsigned char* palette768,
                           const unsigned int* draw8, double alpha, unsigned char* out, cudaStream_t s) {
    if (npix <= 0) return cudaSuccess;
    OverlayTab tab;
    for (int i = 0; i < 768; ++i) tab.rgb[i] = palette768[i];
    for (int i = 0; i < 8; ++i) tab.draw[i] = draw8[i];
    long long blocks = (npix + kThreads - 1) / kThreads;
    if (blocks > 148 * 16) blocks = 148 * 16;
    overlay_kernel<<<(unsigned)blocks, kThreads, 0, s>>>(image, mask, dtype, npix, tab, alpha, out);
    return cudaGetLastError();
}

}  // namespace fscnn
