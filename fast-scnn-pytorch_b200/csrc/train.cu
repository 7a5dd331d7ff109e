// train.cu -- the training step (BASELINE config 5): every layer of the network in TRAINING mode and the
// OHEM loss, forward and backward, as CUDA kernels behind the C ABI (fscnn_train_* in include/fscnn_b200.h).
//
//   depthwise 3x3 (pad 1, stride 1/2, groups = C)   nn.Conv2d(c, c, 3, s, 1, groups=c)      models/fast_scnn.py:70, :86
//   pointwise 1x1 / dense 3x3 through im2col        nn.Conv2d(cin, cout, 1) / (.., 3, ..)   models/fast_scnn.py:73, :107, :24-31
//   BatchNorm2d, batch statistics (+ ReLU)           nn.BatchNorm2d in train mode            models/fast_scnn.py:71-75
//   SoftmaxCrossEntropyOHEMLoss (+ the final resize) utils/loss.py:127-182 (host numpy argsort -> device radix select)
//   bias, bilinear resize, adaptive pooling, dropout, add + ReLU, SGD
//
// Tensors keep PyTorch's layout (NCHW fp32, contiguous) because autograd hands them over that way.  Reductions that feed parameters
// are fp64: per-CTA partials + one finalising pass (BatchNorm, pointwise weight gradient), or double atomics rounded to float once
// (depthwise weight gradient); the fused loss backward adds into the low-resolution gradient with float atomics.  The contractions
// are fp32 FMA by default; fscnn_train_set_math(1) moves them to the tensor cores with TF32 operands (train_tc.cu: tcgen05; here: the
// mma.sync kernels for the weight gradient and unaligned shapes).  Parity: tests/test_gpu_train.py against torch.autograd of the
// unmodified reference modules.
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <initializer_list>

#include "kernels.h"

namespace fscnn {

namespace {
constexpr int kT = 256;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sumf(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// sums NV values per thread over the CTA; result valid in thread 0
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* sm /* [NV][8] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        v[i] = warp_sum(v[i]);
        if (lane == 0) sm[i * 8 + warp] = v[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double s = 0.0;
            for (int w = 0; w < kT / 32; ++w) s += sm[i * 8 + w];
            v[i] = s;
        }
    }
}
}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
// depthwise 3x3, pad 1
// ---------------------------------------------------------------------------------------------------------------------
// All depthwise kernels share one shape: a warp owns 32 adjacent columns of one (image, channel) plane and walks kDwRows rows of
// them with the 3 x 3 window in registers.  A row costs one coalesced load per lane; the left / right neighbours come from the
// adjacent lanes by shuffle (lanes 0 / 31 load theirs), so every input element is read once per row block and no index is divided
// per element.  Rows / columns outside the image are zeros, which leaves the fmaf chain of a border pixel exact.
constexpr int kDwRows = 32;
constexpr unsigned int kFull = 0xffffffffu;
struct Row3 { float l, m, r; };

// stride 1: the values at columns ox - 1, ox, ox + 1 of row iy.  Branch-free (out-of-range rows read nothing and give zeros through
// predicated loads), so the loads of an unrolled walk are all issued before the first one is needed.
__device__ __forceinline__ Row3 dw_row1(const float* __restrict__ p, int iy, int H, int W, int ox, int lane) {
    const bool rok = (unsigned)iy < (unsigned)H;
    const float* r = p + (long long)(rok ? iy : 0) * W;
    Row3 v;
    v.m = (rok && ox < W) ? __ldg(r + ox) : 0.f;
    float edge = 0.f;                    // lane 0: column ox - 1, lane 31: column ox + 1
    const int ex = lane == 0 ? ox - 1 : ox + 1;
    if ((lane == 0 || lane == 31) && rok && ex >= 0 && ex < W) edge = __ldg(r + ex);
    v.l = __shfl_up_sync(kFull, v.m, 1);
    v.r = __shfl_down_sync(kFull, v.m, 1);
    if (lane == 0) v.l = edge;
    if (lane == 31) v.r = edge;
    return v;
}
// stride 2: the values at columns 2 ox - 1, 2 ox, 2 ox + 1 of row iy, in two phases so that a kernel can request several rows
// before the first shuffle needs one of them: raw2_load issues the (predicated) loads, raw2_row builds the triple
struct Raw2 { float e, o, edge; };
template <bool VEC2>
__device__ __forceinline__ Raw2 raw2_load(const float* __restrict__ p, int iy, int H, int W, int ox, int lane, bool want) {
    const bool rok = want && (unsigned)iy < (unsigned)H;
    const float* r = p + (long long)(rok ? iy : 0) * W;
    const int ix = 2 * ox;
    Raw2 v;
    if (VEC2) {                          // W is even, so ix < W implies ix + 1 < W
        const float2 t = (rok && ix < W) ? __ldg(reinterpret_cast<const float2*>(r + ix)) : make_float2(0.f, 0.f);
        v.e = t.x; v.o = t.y;
    } else {
        v.e = (rok && ix < W) ? __ldg(r + ix) : 0.f;
        v.o = (rok && ix + 1 < W) ? __ldg(r + ix + 1) : 0.f;
    }
    v.edge = (lane == 0 && rok && ix > 0 && ix - 1 < W) ? __ldg(r + ix - 1) : 0.f;
    return v;
}
__device__ __forceinline__ Row3 raw2_row(const Raw2& v, int lane) {
    Row3 o;
    o.m = v.e; o.r = v.o;
    o.l = __shfl_up_sync(kFull, v.o, 1);
    if (lane == 0) o.l = v.edge;
    return o;
}
constexpr int kDw2U = 4;      // output rows per batch of loads in the stride-2 kernels
__device__ __forceinline__ float dot9(const Row3& a, const Row3& b, const Row3& c, const float (&k)[9]) {
    float acc = 0.f;
    acc = fmaf(a.l, k[0], acc); acc = fmaf(a.m, k[1], acc); acc = fmaf(a.r, k[2], acc);
    acc = fmaf(b.l, k[3], acc); acc = fmaf(b.m, k[4], acc); acc = fmaf(b.r, k[5], acc);
    acc = fmaf(c.l, k[6], acc); acc = fmaf(c.m, k[7], acc); acc = fmaf(c.r, k[8], acc);
    return acc;
}
__device__ __forceinline__ void outer9(float g, const Row3& a, const Row3& b, const Row3& c, float (&acc)[9]) {
    acc[0] = fmaf(g, a.l, acc[0]); acc[1] = fmaf(g, a.m, acc[1]); acc[2] = fmaf(g, a.r, acc[2]);
    acc[3] = fmaf(g, b.l, acc[3]); acc[4] = fmaf(g, b.m, acc[4]); acc[5] = fmaf(g, b.r, acc[5]);
    acc[6] = fmaf(g, c.l, acc[6]); acc[7] = fmaf(g, c.m, acc[7]); acc[8] = fmaf(g, c.r, acc[8]);
}
// a warp's 9 partial sums -> dwacc[c][9] (double, zeroed by the launcher)
__device__ __forceinline__ void dw_flush(const float (&acc)[9], double* __restrict__ dwacc, int c, int lane) {
    double mine = 0.0;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const double v = warp_sum((double)acc[t]);
        if (lane == t) mine = v;
    }
    if (lane < 9 && mine != 0.0) atomicAdd(dwacc + c * 9 + lane, mine);
}
struct DwTasks {      // tasks ordered (plane, row block, column block)
    int RB, CB;
    long long total;
    __device__ __forceinline__ void at(long long t, long long& plane, int& y0, int& x0) const {
        x0 = (int)(t % CB) * 32;
        y0 = (int)((t / CB) % RB) * kDwRows;
        plane = t / ((long long)CB * RB);
    }
};

// y = depthwise(x) over output rows / columns (Ho, Wo); flip: the taps are read back to front (the stride-1 data gradient is the
// same correlation of dy with the flipped kernel)
template <int STRIDE, bool VEC2>
__global__ void __launch_bounds__(kT)
dw_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y, int C, int H, int W, int Ho, int Wo,
              DwTasks tk, int flip) {
    const int lane = threadIdx.x & 31;
    for (long long t = (long long)blockIdx.x * (kT / 32) + (threadIdx.x >> 5); t < tk.total; t += (long long)gridDim.x * (kT / 32)) {
        long long plane;
        int oy0, ox0;
        tk.at(t, plane, oy0, ox0);
        const int c = (int)(plane % C), ox = ox0 + lane, oy1 = min(Ho, oy0 + kDwRows);
        float k[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) k[i] = __ldg(w + c * 9 + (flip ? 8 - i : i));
        const float* xp = x + plane * H * W;
        float* yp = y + plane * Ho * Wo;
        if (STRIDE == 1) {
            Row3 a = dw_row1(xp, oy0 - 1, H, W, ox, lane), b = dw_row1(xp, oy0, H, W, ox, lane);
#pragma unroll 4
            for (int oy = oy0; oy < oy1; ++oy) {
                const Row3 cc = dw_row1(xp, oy + 1, H, W, ox, lane);
                const float o = dot9(a, b, cc, k);
                if (ox < Wo) yp[(long long)oy * Wo + ox] = o;
                a = b; b = cc;
            }
        } else {
            Row3 a = raw2_row(raw2_load<VEC2>(xp, 2 * oy0 - 1, H, W, ox, lane, true), lane);
            for (int oyb = oy0; oyb < oy1; oyb += kDw2U) {
                Raw2 raw[2 * kDw2U];
#pragma unroll
                for (int i = 0; i < 2 * kDw2U; ++i) raw[i] = raw2_load<VEC2>(xp, 2 * oyb + i, H, W, ox, lane, oyb + (i >> 1) < oy1);
#pragma unroll
                for (int u = 0; u < kDw2U; ++u) {
                    const Row3 b = raw2_row(raw[2 * u], lane), cc = raw2_row(raw[2 * u + 1], lane);
                    const float o = dot9(a, b, cc, k);
                    if (oyb + u < oy1 && ox < Wo) yp[(long long)(oyb + u) * Wo + ox] = o;
                    a = cc;
                }
            }
        }
    }
}

// stride 1 backward, both gradients from one pass over x and dy:
//   dx[iy,ix] = sum dy[iy+1-ky, ix+1-kx] w[ky,kx]   (DX)      dw[c,ky,kx] += dy[oy,ox] x[oy+ky-1, ox+kx-1]   (DWG)
template <bool DX, bool DWG>
__global__ void __launch_bounds__(kT)
dw_bwd1_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ dy, float* __restrict__ dx,
               double* __restrict__ dwacc, int C, int H, int W, DwTasks tk) {
    const int lane = threadIdx.x & 31;
    for (long long t = (long long)blockIdx.x * (kT / 32) + (threadIdx.x >> 5); t < tk.total; t += (long long)gridDim.x * (kT / 32)) {
        long long plane;
        int oy0, ox0;
        tk.at(t, plane, oy0, ox0);
        const int c = (int)(plane % C), ox = ox0 + lane, oy1 = min(H, oy0 + kDwRows);
        float k[9], acc[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) { k[i] = DX ? __ldg(w + c * 9 + 8 - i) : 0.f; acc[i] = 0.f; }
        const float* gp = dy + plane * H * W;
        const float* xp = x + plane * H * W;
        Row3 ga = dw_row1(gp, oy0 - 1, H, W, ox, lane), gb = dw_row1(gp, oy0, H, W, ox, lane);
        Row3 xa{0.f, 0.f, 0.f}, xb{0.f, 0.f, 0.f};
        if (DWG) { xa = dw_row1(xp, oy0 - 1, H, W, ox, lane); xb = dw_row1(xp, oy0, H, W, ox, lane); }
#pragma unroll 4
        for (int oy = oy0; oy < oy1; ++oy) {
            Row3 gc{0.f, 0.f, 0.f}, xc{0.f, 0.f, 0.f};
            if (DX) gc = dw_row1(gp, oy + 1, H, W, ox, lane);
            else gc.m = (oy + 1 < H && ox < W) ? __ldg(gp + (long long)min(oy + 1, H - 1) * W + ox) : 0.f;
            if (DWG) xc = dw_row1(xp, oy + 1, H, W, ox, lane);
            if (DX && ox < W) dx[plane * H * W + (long long)oy * W + ox] = dot9(ga, gb, gc, k);
            if (DWG) outer9(gb.m, xa, xb, xc, acc);
            ga = gb; gb = gc; xa = xb; xb = xc;
        }
        if (DWG) dw_flush(acc, dwacc, c, lane);
    }
}

// Stride 1, width a multiple of 4 and at most 128, 16-byte aligned planes: a lane owns FOUR adjacent columns (one float4 per row), the
// W / 4 lanes of a group span the whole row, so the only neighbours a lane needs are its two adjacent lanes' edge values (zero
// padding at the group's ends), and a warp runs 32 / (W / 4) groups -- row blocks of possibly different planes -- side by side.
// ~2.4x fewer instructions per pixel than the one-column walk.  `in` is convolved with k (read back to front when flip) into `out`
// when OUT; DWG accumulates dw[c] += in_centre x window(xin), reduced per group through a shared-memory slab.
struct Row6 { float l, a, b, c, d, r; };
// two phases, so that a kernel can request several rows before the first shuffle needs one of them
__device__ __forceinline__ float4 raw4_load(const float* __restrict__ p, int iy, int H, int W, int col, bool want) {
    const bool ok = want && (unsigned)iy < (unsigned)H;
    return ok ? __ldg(reinterpret_cast<const float4*>(p + (long long)(ok ? iy : 0) * W + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
}
__device__ __forceinline__ Row6 raw4_row(const float4& v, bool first, bool last) {
    Row6 o{0.f, v.x, v.y, v.z, v.w, 0.f};
    const float l = __shfl_up_sync(kFull, v.w, 1), r = __shfl_down_sync(kFull, v.x, 1);
    o.l = first ? 0.f : l;
    o.r = last ? 0.f : r;
    return o;
}
__device__ __forceinline__ Row6 dw_row4(const float* __restrict__ p, int iy, int H, int W, int col, bool active, bool first, bool last) {
    return raw4_row(raw4_load(p, iy, H, W, col, active), first, last);
}
__device__ __forceinline__ float dot3x3(float a0, float a1, float a2, float b0, float b1, float b2, float c0, float c1, float c2,
                                        const float (&k)[9]) {
    float acc = 0.f;
    acc = fmaf(a0, k[0], acc); acc = fmaf(a1, k[1], acc); acc = fmaf(a2, k[2], acc);
    acc = fmaf(b0, k[3], acc); acc = fmaf(b1, k[4], acc); acc = fmaf(b2, k[5], acc);
    acc = fmaf(c0, k[6], acc); acc = fmaf(c1, k[7], acc); acc = fmaf(c2, k[8], acc);
    return acc;
}
__device__ __forceinline__ void outer3x3(float g, float a0, float a1, float a2, float b0, float b1, float b2, float c0, float c1, float c2,
                                         float (&acc)[9]) {
    acc[0] = fmaf(g, a0, acc[0]); acc[1] = fmaf(g, a1, acc[1]); acc[2] = fmaf(g, a2, acc[2]);
    acc[3] = fmaf(g, b0, acc[3]); acc[4] = fmaf(g, b1, acc[4]); acc[5] = fmaf(g, b2, acc[5]);
    acc[6] = fmaf(g, c0, acc[6]); acc[7] = fmaf(g, c1, acc[7]); acc[8] = fmaf(g, c2, acc[8]);
}

template <bool OUT, bool DWG>
__global__ void __launch_bounds__(kT)
dw_s1v4_kernel(const float* __restrict__ in, const float* __restrict__ xin, const float* __restrict__ w, float* __restrict__ out,
               double* __restrict__ dwacc, int C, int H, int W, int RB, long long tasks, int flip) {
    __shared__ float slab[kT / 32][32][9];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int L = W >> 2, G = 32 / L, grp = lane / L, j = lane - grp * L;
    const bool lane_ok = grp < G, first = j == 0, last = j == L - 1;
    const long long wtasks = (tasks + G - 1) / G;
    for (long long wt = (long long)blockIdx.x * (kT / 32) + warp; wt < wtasks; wt += (long long)gridDim.x * (kT / 32)) {
        const long long t = wt * G + grp;
        const bool active = lane_ok && t < tasks;
        const long long plane = active ? t / RB : 0;
        const int oy0 = active ? (int)(t - plane * RB) * kDwRows : 0, oy1 = active ? min(H, oy0 + kDwRows) : 0;
        const int c = (int)(plane % C), col = 4 * j;
        float k[9], acc[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) { k[i] = OUT ? __ldg(w + c * 9 + (flip ? 8 - i : i)) : 0.f; acc[i] = 0.f; }
        const float* ip = in + plane * H * W;
        const float* xp = DWG ? xin + plane * H * W : nullptr;
        float* op = OUT ? out + plane * H * W : nullptr;
        Row6 ga = dw_row4(ip, oy0 - 1, H, W, col, active, first, last), gb = dw_row4(ip, oy0, H, W, col, active, first, last);
        Row6 xa{0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, xb = xa;
        if (DWG) { xa = dw_row4(xp, oy0 - 1, H, W, col, active, first, last); xb = dw_row4(xp, oy0, H, W, col, active, first, last); }
        constexpr int U = 4;
        for (int rb = 0; rb < kDwRows; rb += U) {
            float4 rg[U], rx[U];                                // rows oy + 1 of the next U outputs, requested together
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int oy = oy0 + rb + u;
                rg[u] = raw4_load(ip, oy + 1, H, W, col, active && oy < oy1);
                if (DWG) rx[u] = raw4_load(xp, oy + 1, H, W, col, active && oy < oy1);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int oy = oy0 + rb + u;
                const bool row_ok = oy < oy1;                   // per group
                Row6 gc{0.f, rg[u].x, rg[u].y, rg[u].z, rg[u].w, 0.f}, xc{0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (OUT) gc = raw4_row(rg[u], first, last);
                if (DWG) xc = raw4_row(rx[u], first, last);
                if (OUT && row_ok) {
                    float4 o;
                    o.x = dot3x3(ga.l, ga.a, ga.b, gb.l, gb.a, gb.b, gc.l, gc.a, gc.b, k);
                    o.y = dot3x3(ga.a, ga.b, ga.c, gb.a, gb.b, gb.c, gc.a, gc.b, gc.c, k);
                    o.z = dot3x3(ga.b, ga.c, ga.d, gb.b, gb.c, gb.d, gc.b, gc.c, gc.d, k);
                    o.w = dot3x3(ga.c, ga.d, ga.r, gb.c, gb.d, gb.r, gc.c, gc.d, gc.r, k);
                    *reinterpret_cast<float4*>(op + (long long)oy * W + col) = o;
                }
                if (DWG && row_ok) {
                    outer3x3(gb.a, xa.l, xa.a, xa.b, xb.l, xb.a, xb.b, xc.l, xc.a, xc.b, acc);
                    outer3x3(gb.b, xa.a, xa.b, xa.c, xb.a, xb.b, xb.c, xc.a, xc.b, xc.c, acc);
                    outer3x3(gb.c, xa.b, xa.c, xa.d, xb.b, xb.c, xb.d, xc.b, xc.c, xc.d, acc);
                    outer3x3(gb.d, xa.c, xa.d, xa.r, xb.c, xb.d, xb.r, xc.c, xc.d, xc.r, acc);
                }
                ga = gb; gb = gc; xa = xb; xb = xc;
            }
        }
        if (DWG) {      // per-group sum through the warp's slab, then 9 double atomics by the group's first lane
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 9; ++i) slab[warp][lane][i] = acc[i];
            __syncwarp();
            if (active && first) {
                double tot[9];
#pragma unroll
                for (int i = 0; i < 9; ++i) tot[i] = 0.0;
                for (int q = 0; q < L; ++q)
#pragma unroll
                    for (int i = 0; i < 9; ++i) tot[i] += (double)slab[warp][lane + q][i];
#pragma unroll
                for (int i = 0; i < 9; ++i)
                    if (tot[i] != 0.0) atomicAdd(dwacc + c * 9 + i, tot[i]);
            }
        }
    }
}

// stride 2 data gradient, lane = dy column ox -> dx columns 2 ox and 2 ox + 1, dy row oy -> dx rows 2 oy and 2 oy + 1:
//   dx[2oy  ][2ox] = d w11                      dx[2oy  ][2ox+1] = dn w10 + d w12
//   dx[2oy+1][2ox] = nd w01 + d w21             dx[2oy+1][2ox+1] = ndn w00 + nd w02 + dn w20 + d w22
// with d = dy[oy][ox], dn = dy[oy][ox+1], nd / ndn the same in row oy + 1 (zeros outside)
template <bool VEC2>
__global__ void __launch_bounds__(kT)
dw_bwd2_data_kernel(const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx, int C, int H, int W, int Ho,
                    int Wo, DwTasks tk) {
    const int lane = threadIdx.x & 31;
    for (long long t = (long long)blockIdx.x * (kT / 32) + (threadIdx.x >> 5); t < tk.total; t += (long long)gridDim.x * (kT / 32)) {
        long long plane;
        int oy0, ox0;
        tk.at(t, plane, oy0, ox0);
        const int c = (int)(plane % C), ox = ox0 + lane, oy1 = min(Ho, oy0 + kDwRows);
        float k[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) k[i] = __ldg(w + c * 9 + i);
        const float* gp = dy + plane * Ho * Wo;
        float* dp = dx + plane * H * W;
        // phase 1 of a dy row: its value and lane 31's right neighbour; phase 2: the neighbour of every lane
        auto load = [&](int oy, bool want, float& d, float& edge) {
            const bool rok = want && oy < Ho;
            const float* r = gp + (long long)(rok ? oy : 0) * Wo;
            d = (rok && ox < Wo) ? __ldg(r + ox) : 0.f;
            edge = (lane == 31 && rok && ox + 1 < Wo) ? __ldg(r + ox + 1) : 0.f;
        };
        auto right = [&](float d, float edge) {
            const float dn = __shfl_down_sync(kFull, d, 1);
            return lane == 31 ? edge : dn;
        };
        float d, dn;
        {
            float e0;
            load(oy0, true, d, e0);
            dn = right(d, e0);
        }
        for (int oyb = oy0; oyb < oy1; oyb += kDw2U) {
            float rd[kDw2U], re[kDw2U];
#pragma unroll
            for (int u = 0; u < kDw2U; ++u) load(oyb + u + 1, oyb + u < oy1, rd[u], re[u]);
#pragma unroll
            for (int u = 0; u < kDw2U; ++u) {
                const int oy = oyb + u;
                const float nd = rd[u], ndn = right(rd[u], re[u]);
                const int iy = 2 * oy, ix = 2 * ox;
                const float e0 = d * k[4], e1 = fmaf(dn, k[3], d * k[5]);
                const float o0 = fmaf(nd, k[1], d * k[7]), o1 = fmaf(ndn, k[0], fmaf(nd, k[2], fmaf(dn, k[6], d * k[8])));
                if (oy < oy1) {
                    if (VEC2) {
                        if (ix < W) {
                            *reinterpret_cast<float2*>(dp + (long long)iy * W + ix) = make_float2(e0, e1);
                            if (iy + 1 < H) *reinterpret_cast<float2*>(dp + (long long)(iy + 1) * W + ix) = make_float2(o0, o1);
                        }
                    } else {
                        if (ix < W) dp[(long long)iy * W + ix] = e0;
                        if (ix + 1 < W) dp[(long long)iy * W + ix + 1] = e1;
                        if (iy + 1 < H) {
                            if (ix < W) dp[(long long)(iy + 1) * W + ix] = o0;
                            if (ix + 1 < W) dp[(long long)(iy + 1) * W + ix + 1] = o1;
                        }
                    }
                }
                d = nd; dn = ndn;
            }
        }
    }
}

// stride 2 weight gradient: the forward's walk with dy[oy][ox] against the 3 x 3 input window
template <bool VEC2>
__global__ void __launch_bounds__(kT)
dw_bwd2_weight_kernel(const float* __restrict__ x, const float* __restrict__ dy, double* __restrict__ dwacc, int C, int H, int W,
                      int Ho, int Wo, DwTasks tk) {
    const int lane = threadIdx.x & 31;
    for (long long t = (long long)blockIdx.x * (kT / 32) + (threadIdx.x >> 5); t < tk.total; t += (long long)gridDim.x * (kT / 32)) {
        long long plane;
        int oy0, ox0;
        tk.at(t, plane, oy0, ox0);
        const int c = (int)(plane % C), ox = ox0 + lane, oy1 = min(Ho, oy0 + kDwRows);
        float acc[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) acc[i] = 0.f;
        const float* xp = x + plane * H * W;
        const float* gp = dy + plane * Ho * Wo;
        Row3 a = raw2_row(raw2_load<VEC2>(xp, 2 * oy0 - 1, H, W, ox, lane, true), lane);
        for (int oyb = oy0; oyb < oy1; oyb += kDw2U) {
            Raw2 raw[2 * kDw2U];
            float g[kDw2U];
#pragma unroll
            for (int i = 0; i < 2 * kDw2U; ++i) raw[i] = raw2_load<VEC2>(xp, 2 * oyb + i, H, W, ox, lane, oyb + (i >> 1) < oy1);
#pragma unroll
            for (int u = 0; u < kDw2U; ++u) g[u] = (oyb + u < oy1 && ox < Wo) ? __ldg(gp + (long long)(oyb + u) * Wo + ox) : 0.f;
#pragma unroll
            for (int u = 0; u < kDw2U; ++u) {
                const Row3 b = raw2_row(raw[2 * u], lane), cc = raw2_row(raw[2 * u + 1], lane);
                outer9(g[u], a, b, cc, acc);
                a = cc;
            }
        }
        dw_flush(acc, dwacc, c, lane);
    }
}

__global__ void __launch_bounds__(kT)
double_to_float_kernel(const double* __restrict__ in, float* __restrict__ out, int count) {
    const int i = blockIdx.x * kT + threadIdx.x;
    if (i < count) out[i] = (float)in[i];
}

// out[i] = sum_s partial[s][i]
__global__ void __launch_bounds__(kT)
reduce_partials_kernel(const double* __restrict__ partial, float* __restrict__ out, int count, int S) {
    const int i = blockIdx.x * kT + threadIdx.x;
    if (i >= count) return;
    double s = 0.0;
    for (int k = 0; k < S; ++k) s += partial[(long long)k * count + i];
    out[i] = (float)s;
}

// ---------------------------------------------------------------------------------------------------------------------
// pointwise 1x1 as a tiled fp32 GEMM with generic operand strides:
//   C[b][m][j] (+)= sum_k A(b, m, k) * B(b, k, j),  A(b,m,k) = A[b*sAb + m*sAm + k*sAk],  B(b,k,j) = B[b*sBb + k*sBk + j*sBj]
//   forward      : A = W[cout][cin] (sAb 0),            B = x[n] ([cin][hw]),        C = y[n]
//   data gradient: A = W^T (sAm 1, sAk cin),            B = dy[n] ([cout][hw]),      C = dx[n]
//   weight grad. : A = dy[n] ([cout][hw]),              B = x[n]^T (sBk 1, sBj hw),  C = partial[b] ([cout][cin]), then reduced
// Two kernels, one per shape class; `splitk` CTAs along grid.z share one batch item's K range in the weight gradient.
// ---------------------------------------------------------------------------------------------------------------------
struct GemmArgs {
    const float *A, *B;
    float* C;
    int M, N, K;
    long long sAb, sAm, sAk, sBb, sBk, sBj, sCb;
    int ldc, splitk;
};

// The weight-gradient shape: a small output (M = cout, N = cin) and a very long K (the pixels of an image slice), both operands
// with unit stride along K.  64 x 64 tile, K chunks of 32 read as float4 along K (128 contiguous bytes per row) and prefetched into
// registers while the previous chunk is multiplied; `splitk` CTAs along grid.z share one image's pixels.
__global__ void __launch_bounds__(kT)
gemm_wgrad_kernel(GemmArgs g) {
    constexpr int BM = 64, BN = 64, BK = 32;
    __shared__ __align__(16) float As[BK][BM + 4], Bs[BK][BN + 4];
    const int b = blockIdx.z / g.splitk, sk = blockIdx.z % g.splitk;
    const int m0 = blockIdx.y * BM, j0 = blockIdx.x * BN;
    const int tid = threadIdx.x, tm = tid / 16, tj = tid % 16;
    const float* A = g.A + b * g.sAb;
    const float* B = g.B + b * g.sBb;
    const int kper = ((g.K + g.splitk - 1) / g.splitk + BK - 1) / BK * BK;
    const int k_begin = sk * kper, k_end = min(g.K, k_begin + kper);
    const bool avec = ((g.sAm & 3) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
    const bool bvec = ((g.sBj & 3) == 0) && ((reinterpret_cast<uintptr_t>(B) & 15) == 0);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float4 ra[2], rb[2];
    auto load4 = [&](const float* base, long long row_stride, int row, int row_limit, int k, bool vec) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < row_limit) {
            const float* src = base + (long long)row * row_stride + k;
            if (vec && k + 3 < k_end) {
                v = __ldg(reinterpret_cast<const float4*>(src));
            } else {
                float t[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) t[q] = k + q < k_end ? __ldg(src + q) : 0.f;
                v = make_float4(t[0], t[1], t[2], t[3]);
            }
        }
        return v;
    };
    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {      // 64 rows x 8 float4 per operand
            const int i = tid + r * kT, row = i / 8, k4 = (i % 8) * 4;
            ra[r] = load4(A, g.sAm, m0 + row, g.M, k0 + k4, avec && ((k0 + k4) & 3) == 0);
            rb[r] = load4(B, g.sBj, j0 + row, g.N, k0 + k4, bvec && ((k0 + k4) & 3) == 0);
        }
    };
    auto store_tiles = [&]() {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int i = tid + r * kT, row = i / 8, k4 = (i % 8) * 4;
            As[k4][row] = ra[r].x; As[k4 + 1][row] = ra[r].y; As[k4 + 2][row] = ra[r].z; As[k4 + 3][row] = ra[r].w;
            Bs[k4][row] = rb[r].x; Bs[k4 + 1][row] = rb[r].y; Bs[k4 + 2][row] = rb[r].z; Bs[k4 + 3][row] = rb[r].w;
        }
    };
    if (k_begin < k_end) load_tiles(k_begin);
    for (int k0 = k_begin; k0 < k_end; k0 += BK) {
        store_tiles();
        __syncthreads();
        if (k0 + BK < k_end) load_tiles(k0 + BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&As[k][tm * 4]);
            const float4 bb = *reinterpret_cast<const float4*>(&Bs[k][tj * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    float* C = g.C + (long long)blockIdx.z * g.sCb;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + tm * 4 + i;
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int jj = j0 + tj * 4 + j;
            if (jj < g.N) C[(long long)m * g.ldc + jj] = acc[i][j];
        }
    }
}

// The forward / data-gradient shape of the same contraction: few output channels (M <= 768), many pixels per image (N = H*W,
// unit stride in B and C), K = input channels.  64 x 128 tile, K chunks of 16, 256 threads x (4 channels x 8 pixels); the
// thread's pixels are two groups of 4, 64 apart, so that a warp's 16-byte shared-memory reads are contiguous; the next
// chunk's global loads are issued before the current chunk is multiplied.
__global__ void __launch_bounds__(kT)
gemm_pix_kernel(GemmArgs g) {
    constexpr int BM = 64, BN = 128, BK = 16;
    __shared__ __align__(16) float As[BK][BM + 4], Bs[BK][BN + 4];
    const int b = blockIdx.z;
    const int m0 = blockIdx.y * BM, j0 = blockIdx.x * BN;
    const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
    const float* A = g.A + b * g.sAb;
    const float* B = g.B + b * g.sBb;
    const bool vec = ((g.sBk & 3) == 0) && ((reinterpret_cast<uintptr_t>(B) & 15) == 0);
    const bool a_k_fast = g.sAk == 1;
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    float ra[4];
    float4 rb[2];
    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {      // A: 64 x 16
            const int i = tid + r * kT;
            const int k = a_k_fast ? i % BK : i / BM, m = a_k_fast ? i / BK : i % BM;
            ra[r] = (m0 + m < g.M && k0 + k < g.K) ? __ldg(A + (long long)(m0 + m) * g.sAm + (long long)(k0 + k) * g.sAk) : 0.f;
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {      // B: 16 x 128 as float4
            const int i = tid + r * kT, k = i / 32, j = (i % 32) * 4;
            const float* src = B + (long long)(k0 + k) * g.sBk + j0 + j;
            if (k0 + k < g.K && vec && j0 + j + 3 < g.N) {
                rb[r] = __ldg(reinterpret_cast<const float4*>(src));
            } else {
                float t[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) t[q] = (k0 + k < g.K && j0 + j + q < g.N) ? __ldg(src + q) : 0.f;
                rb[r] = make_float4(t[0], t[1], t[2], t[3]);
            }
        }
    };
    auto store_tiles = [&]() {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int i = tid + r * kT;
            const int k = a_k_fast ? i % BK : i / BM, m = a_k_fast ? i / BK : i % BM;
            As[k][m] = ra[r];
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int i = tid + r * kT, k = i / 32, j = (i % 32) * 4;
            *reinterpret_cast<float4*>(&Bs[k][j]) = rb[r];
        }
    };
    load_tiles(0);
    for (int k0 = 0; k0 < g.K; k0 += BK) {
        store_tiles();
        __syncthreads();
        if (k0 + BK < g.K) load_tiles(k0 + BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    float* C = g.C + (long long)b * g.sCb;
    const bool cvec = ((g.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= g.M) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int jj = j0 + h * 64 + tx * 4;
            float* dst = C + (long long)m * g.ldc + jj;
            if (cvec && jj + 3 < g.N) {
                *reinterpret_cast<float4*>(dst) = make_float4(acc[i][4 * h], acc[i][4 * h + 1], acc[i][4 * h + 2], acc[i][4 * h + 3]);
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (jj + q < g.N) dst[q] = acc[i][4 * h + q];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// The same two contractions on the tensor cores (mma.sync m16n8k8, TF32 operands rounded to nearest, fp32 accumulators): 10 mantissa
// bits per operand, which is what cuDNN convolutions do by default under torch.backends.cudnn.allow_tf32 and about what the
// reference's fp16 autocast keeps; opt-in through fscnn_train_set_math(1).  TERMS = 3 is the split form (x = hi + lo, products
// lo_a hi_b + hi_a lo_b + hi_a hi_b, ~21 mantissa bits): measured on a B200 it is no faster than the fp32 FMA kernels (4.25 vs 4.6 ms
// for the forward / data-gradient GEMMs of a config-5 step) and 3-8x less accurate, so it is not reachable from the ABI.
// Shared-memory tiles use pitches of 8 (k-major tiles) or 4 (k-fastest tiles) modulo 32 words, so the per-lane fragment loads of
// a warp hit 32 distinct banks.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t tf32_hi(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
template <int TERMS>
__device__ __forceinline__ void tf32_split(float x, uint32_t& hi, uint32_t& lo) {
    hi = tf32_hi(x);
    lo = TERMS == 3 ? tf32_hi(x - __uint_as_float(hi)) : 0u;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int TERMS>
__device__ __forceinline__ void mma_split(float (&d)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], const uint32_t (&bh)[2],
                                          const uint32_t (&bl)[2]) {
    if (TERMS == 3) { mma_tf32(d, al, bh); mma_tf32(d, ah, bl); }
    mma_tf32(d, ah, bh);
}

// forward / data gradient: C[b][m][j] = sum_k A(m,k) B[b][k][j]; 64 x 128 tile, K chunks of 16, 8 warps as 2 x 4, warp tile 32 x 32
template <int TERMS>
__global__ void __launch_bounds__(kT)
gemm_pix_mma_kernel(GemmArgs g) {
    constexpr int BM = 64, BN = 128, BK = 16, PA = BM + 8, PB = BN + 8;
    __shared__ __align__(16) float As[BK][PA], Bs[BK][PB];
    const int b = blockIdx.z;
    const int m0 = blockIdx.y * BM, j0 = blockIdx.x * BN;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gq = lane >> 2, tq = lane & 3;
    const int wm = (warp >> 2) * 32, wn = (warp & 3) * 32;
    const float* A = g.A + b * g.sAb;
    const float* B = g.B + b * g.sBb;
    const bool vec = ((g.sBk & 3) == 0) && ((reinterpret_cast<uintptr_t>(B) & 15) == 0);
    const bool a_k_fast = g.sAk == 1;
    float acc[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[i][j][q] = 0.f;
    float ra[4];
    float4 rb[2];
    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {      // A: 64 x 16
            const int i = tid + r * kT;
            const int k = a_k_fast ? i % BK : i / BM, m = a_k_fast ? i / BK : i % BM;
            ra[r] = (m0 + m < g.M && k0 + k < g.K) ? __ldg(A + (long long)(m0 + m) * g.sAm + (long long)(k0 + k) * g.sAk) : 0.f;
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {      // B: 16 x 128 as float4
            const int i = tid + r * kT, k = i / 32, j = (i % 32) * 4;
            const float* src = B + (long long)(k0 + k) * g.sBk + j0 + j;
            if (k0 + k < g.K && vec && j0 + j + 3 < g.N) {
                rb[r] = __ldg(reinterpret_cast<const float4*>(src));
            } else {
                float t[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) t[q] = (k0 + k < g.K && j0 + j + q < g.N) ? __ldg(src + q) : 0.f;
                rb[r] = make_float4(t[0], t[1], t[2], t[3]);
            }
        }
    };
    auto store_tiles = [&]() {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int i = tid + r * kT;
            const int k = a_k_fast ? i % BK : i / BM, m = a_k_fast ? i / BK : i % BM;
            As[k][m] = ra[r];
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int i = tid + r * kT, k = i / 32, j = (i % 32) * 4;
            *reinterpret_cast<float4*>(&Bs[k][j]) = rb[r];
        }
    };
    load_tiles(0);
    for (int k0 = 0; k0 < g.K; k0 += BK) {
        store_tiles();
        __syncthreads();
        if (k0 + BK < g.K) load_tiles(k0 + BK);
#pragma unroll
        for (int ks = 0; ks < BK; ks += 8) {
            uint32_t ah[2][4], al[2][4], bh[4][2], bl[4][2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int m = wm + i * 16 + gq;
                tf32_split<TERMS>(As[ks + tq][m], ah[i][0], al[i][0]);
                tf32_split<TERMS>(As[ks + tq][m + 8], ah[i][1], al[i][1]);
                tf32_split<TERMS>(As[ks + tq + 4][m], ah[i][2], al[i][2]);
                tf32_split<TERMS>(As[ks + tq + 4][m + 8], ah[i][3], al[i][3]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = wn + j * 8 + gq;
                tf32_split<TERMS>(Bs[ks + tq][n], bh[j][0], bl[j][0]);
                tf32_split<TERMS>(Bs[ks + tq + 4][n], bh[j][1], bl[j][1]);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) mma_split<TERMS>(acc[i][j], ah[i], al[i], bh[j], bl[j]);
        }
        __syncthreads();
    }
    float* C = g.C + (long long)b * g.sCb;
    const bool cvec = ((g.ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(C) & 7) == 0);
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int m = m0 + wm + i * 16 + gq + h * 8;
            if (m >= g.M) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int jj = j0 + wn + j * 8 + 2 * tq;
                float* dst = C + (long long)m * g.ldc + jj;
                if (cvec && jj + 1 < g.N) {
                    *reinterpret_cast<float2*>(dst) = make_float2(acc[i][j][2 * h], acc[i][j][2 * h + 1]);
                } else {
                    if (jj < g.N) dst[0] = acc[i][j][2 * h];
                    if (jj + 1 < g.N) dst[1] = acc[i][j][2 * h + 1];
                }
            }
        }
}

// weight gradient: C[part][m][j] = sum over this part's k of A[b][m][k] B[b][j][k] (both operands k-fastest); 64 x 64 tile, K chunks
// of 32, 8 warps as 2 x 4, warp tile 32 x 16
template <int TERMS>
__global__ void __launch_bounds__(kT)
gemm_wgrad_mma_kernel(GemmArgs g) {
    constexpr int BM = 64, BN = 64, BK = 32, P = BK + 4;
    __shared__ __align__(16) float As[BM][P], Bs[BN][P];
    const int b = blockIdx.z / g.splitk, sk = blockIdx.z % g.splitk;
    const int m0 = blockIdx.y * BM, j0 = blockIdx.x * BN;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gq = lane >> 2, tq = lane & 3;
    const int wm = (warp >> 2) * 32, wn = (warp & 3) * 16;
    const float* A = g.A + b * g.sAb;
    const float* B = g.B + b * g.sBb;
    const int kper = ((g.K + g.splitk - 1) / g.splitk + BK - 1) / BK * BK;
    const int k_begin = sk * kper, k_end = min(g.K, k_begin + kper);
    const bool avec = ((g.sAm & 3) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
    const bool bvec = ((g.sBj & 3) == 0) && ((reinterpret_cast<uintptr_t>(B) & 15) == 0);
    float acc[2][2][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[i][j][q] = 0.f;
    float4 ra[2], rb[2];
    auto load4 = [&](const float* base, long long row_stride, int row, int row_limit, int k, bool vec) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < row_limit) {
            const float* src = base + (long long)row * row_stride + k;
            if (vec && k + 3 < k_end) {
                v = __ldg(reinterpret_cast<const float4*>(src));
            } else {
                float t[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) t[q] = k + q < k_end ? __ldg(src + q) : 0.f;
                v = make_float4(t[0], t[1], t[2], t[3]);
            }
        }
        return v;
    };
    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {      // 64 rows x 8 float4 per operand
            const int i = tid + r * kT, row = i / 8, k4 = (i % 8) * 4;
            ra[r] = load4(A, g.sAm, m0 + row, g.M, k0 + k4, avec && ((k0 + k4) & 3) == 0);
            rb[r] = load4(B, g.sBj, j0 + row, g.N, k0 + k4, bvec && ((k0 + k4) & 3) == 0);
        }
    };
    auto store_tiles = [&]() {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int i = tid + r * kT, row = i / 8, k4 = (i % 8) * 4;
            *reinterpret_cast<float4*>(&As[row][k4]) = ra[r];
            *reinterpret_cast<float4*>(&Bs[row][k4]) = rb[r];
        }
    };
    if (k_begin < k_end) load_tiles(k_begin);
    for (int k0 = k_begin; k0 < k_end; k0 += BK) {
        store_tiles();
        __syncthreads();
        if (k0 + BK < k_end) load_tiles(k0 + BK);
#pragma unroll
        for (int ks = 0; ks < BK; ks += 8) {
            uint32_t ah[2][4], al[2][4], bh[2][2], bl[2][2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int m = wm + i * 16 + gq;
                tf32_split<TERMS>(As[m][ks + tq], ah[i][0], al[i][0]);
                tf32_split<TERMS>(As[m + 8][ks + tq], ah[i][1], al[i][1]);
                tf32_split<TERMS>(As[m][ks + tq + 4], ah[i][2], al[i][2]);
                tf32_split<TERMS>(As[m + 8][ks + tq + 4], ah[i][3], al[i][3]);
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int n = wn + j * 8 + gq;
                tf32_split<TERMS>(Bs[n][ks + tq], bh[j][0], bl[j][0]);
                tf32_split<TERMS>(Bs[n][ks + tq + 4], bh[j][1], bl[j][1]);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) mma_split<TERMS>(acc[i][j], ah[i], al[i], bh[j], bl[j]);
        }
        __syncthreads();
    }
    float* C = g.C + (long long)blockIdx.z * g.sCb;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int m = m0 + wm + i * 16 + gq + h * 8;
            if (m >= g.M) continue;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int jj = j0 + wn + j * 8 + 2 * tq;
                if (jj < g.N) C[(long long)m * g.ldc + jj] = acc[i][j][2 * h];
                if (jj + 1 < g.N) C[(long long)m * g.ldc + jj + 1] = acc[i][j][2 * h + 1];
            }
        }
}

__global__ void __launch_bounds__(kT)
reduce_partials_f_kernel(const float* __restrict__ partial, float* __restrict__ out, int count, int S) {
    const int i = blockIdx.x * kT + threadIdx.x;
    if (i >= count) return;
    double s = 0.0;
    for (int k = 0; k < S; ++k) s += (double)partial[(long long)k * count + i];
    out[i] = (float)s;
}

// ---------------------------------------------------------------------------------------------------------------------
// BatchNorm2d with batch statistics (train mode) + optional ReLU
// ---------------------------------------------------------------------------------------------------------------------
// The normalised value, one rounding sequence shared by the forward and by the backward's ReLU mask (y > 0 <=> bn_value > 0), so the
// backward never reads y.
__device__ __forceinline__ float bn_value(float x, float mu, float rs, float gamma, float beta) {
    return __fmaf_rn(__fmul_rn(__fsub_rn(x, mu), rs), gamma, beta);
}

// The per-channel reductions walk channel c's N planes of HW values; a CTA (c, s) owns the s-th slice of every plane.  VEC: HW is a
// multiple of 4 and the tensors are 16-byte aligned -> the slice is walked as float4 with ONE flattened index over (image, vector)
// (a 32-bit division per 16 bytes); otherwise nested scalar loops.
struct ChanSlice { int lo, hi; };       // element range inside a plane
__device__ __forceinline__ ChanSlice chan_slice(int HW, int S, int s, int gran) {
    const int units = (HW + gran - 1) / gran, per = (units + S - 1) / S;
    ChanSlice r{min(HW, s * per * gran), min(HW, (s + 1) * per * gran)};
    return r;
}

// grid (C, S): partial[s][c][2] = {sum x, sum x^2} over this CTA's share of the N*HW values of channel c
template <bool VEC>
__global__ void __launch_bounds__(kT)
bn_stats_kernel(const float* __restrict__ x, double* __restrict__ partial, int N, int C, int HW) {
    __shared__ double sm[2 * 8];
    const int c = blockIdx.x, S = gridDim.y;
    const ChanSlice sl = chan_slice(HW, S, blockIdx.y, VEC ? 4 : 1);
    double v[2] = {0.0, 0.0};
    if (VEC) {
        const int nv = (sl.hi - sl.lo) >> 2, total = N * nv;
#pragma unroll 4
        for (int i = threadIdx.x; i < total; i += kT) {
            const int n = i / nv, q = i - n * nv;
            const float4 t = __ldg(reinterpret_cast<const float4*>(x + ((long long)n * C + c) * HW + sl.lo) + q);
            const float s1 = (t.x + t.y) + (t.z + t.w);
            const float s2 = fmaf(t.x, t.x, t.y * t.y) + fmaf(t.z, t.z, t.w * t.w);
            v[0] += (double)s1; v[1] += (double)s2;
        }
    } else {
        for (int n = 0; n < N; ++n) {
            const float* xp = x + ((long long)n * C + c) * HW;
#pragma unroll 4
            for (int p = sl.lo + threadIdx.x; p < sl.hi; p += kT) {
                const double t = (double)__ldg(xp + p);
                v[0] += t; v[1] += t * t;
            }
        }
    }
    block_sum<2>(v, sm);
    if (threadIdx.x == 0) { partial[((long long)blockIdx.y * C + c) * 2] = v[0]; partial[((long long)blockIdx.y * C + c) * 2 + 1] = v[1]; }
}

// mean / biased variance -> save_mean, save_rstd; running stats with momentum and the UNBIASED variance (PyTorch semantics)
__global__ void __launch_bounds__(kT)
bn_finalize_kernel(const double* __restrict__ partial, int C, int S, long long count, float eps, float momentum,
                   float* __restrict__ save_mean, float* __restrict__ save_rstd, float* __restrict__ running_mean,
                   float* __restrict__ running_var) {
    const int c = blockIdx.x * kT + threadIdx.x;
    if (c >= C) return;
    double s = 0.0, ss = 0.0;
    for (int k = 0; k < S; ++k) { s += partial[((long long)k * C + c) * 2]; ss += partial[((long long)k * C + c) * 2 + 1]; }
    const double mean = s / (double)count;
    double var = ss / (double)count - mean * mean;
    if (var < 0.0) var = 0.0;
    save_mean[c] = (float)mean;
    save_rstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) {
        const double unbiased = count > 1 ? var * (double)count / (double)(count - 1) : var;
        running_mean[c] = (float)((1.0 - momentum) * (double)running_mean[c] + momentum * mean);
        running_var[c] = (float)((1.0 - momentum) * (double)running_var[c] + momentum * unbiased);
    }
}

// elementwise passes: grid (planes = N*C, chunks); a CTA owns kT * 16 consecutive elements of one plane, so the channel's constants
// are loaded once and no index is divided
constexpr int kEltChunk = kT * 16;

template <bool VEC>
__global__ void __launch_bounds__(kT)
bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ rstd,
                const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ y, int C, int HW, int relu) {
    const int c = blockIdx.x % C;
    const float mu = __ldg(mean + c), rs = __ldg(rstd + c), ga = __ldg(gamma + c), be = __ldg(beta + c);
    const long long base = (long long)blockIdx.x * HW;
    const int lo = blockIdx.y * kEltChunk, hi = min(HW, lo + kEltChunk);
    const float floor_ = relu ? 0.f : -FLT_MAX;
    if (VEC) {
        const float4* xp = reinterpret_cast<const float4*>(x + base);
        float4* yp = reinterpret_cast<float4*>(y + base);
#pragma unroll 4
        for (int q = (lo >> 2) + threadIdx.x; q < (hi >> 2); q += kT) {
            const float4 t = __ldg(xp + q);
            float4 o;
            o.x = fmaxf(bn_value(t.x, mu, rs, ga, be), floor_); o.y = fmaxf(bn_value(t.y, mu, rs, ga, be), floor_);
            o.z = fmaxf(bn_value(t.z, mu, rs, ga, be), floor_); o.w = fmaxf(bn_value(t.w, mu, rs, ga, be), floor_);
            yp[q] = o;
        }
    } else {
#pragma unroll 4
        for (int p = lo + threadIdx.x; p < hi; p += kT) y[base + p] = fmaxf(bn_value(__ldg(x + base + p), mu, rs, ga, be), floor_);
    }
}

// grid (C, S): partial[s][c][2] = {sum g, sum g * xhat}, g = dy masked by the ReLU (bn_value > 0) when relu
template <bool VEC>
__global__ void __launch_bounds__(kT)
bn_bwd_reduce_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ mean,
                     const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                     double* __restrict__ partial, int N, int C, int HW, int relu) {
    __shared__ double sm[2 * 8];
    const int c = blockIdx.x, S = gridDim.y;
    const float mu = __ldg(mean + c), rs = __ldg(rstd + c), ga = __ldg(gamma + c), be = __ldg(beta + c);
    const ChanSlice sl = chan_slice(HW, S, blockIdx.y, VEC ? 4 : 1);
    double v[2] = {0.0, 0.0};
    auto one = [&](float xv, float g, float& s1, float& s2) {
        if (relu && !(bn_value(xv, mu, rs, ga, be) > 0.f)) g = 0.f;
        s1 += g;
        s2 = fmaf(g, (xv - mu) * rs, s2);
    };
    if (VEC) {
        const int nv = (sl.hi - sl.lo) >> 2, total = N * nv;
#pragma unroll 2
        for (int i = threadIdx.x; i < total; i += kT) {
            const int n = i / nv, q = i - n * nv;
            const long long e = ((long long)n * C + c) * HW + sl.lo;
            const float4 t = __ldg(reinterpret_cast<const float4*>(x + e) + q), g = __ldg(reinterpret_cast<const float4*>(dy + e) + q);
            float s1 = 0.f, s2 = 0.f;
            one(t.x, g.x, s1, s2); one(t.y, g.y, s1, s2); one(t.z, g.z, s1, s2); one(t.w, g.w, s1, s2);
            v[0] += (double)s1; v[1] += (double)s2;
        }
    } else {
        for (int n = 0; n < N; ++n) {
            const long long e = ((long long)n * C + c) * HW;
#pragma unroll 4
            for (int p = sl.lo + threadIdx.x; p < sl.hi; p += kT) {
                float s1 = 0.f, s2 = 0.f;
                one(__ldg(x + e + p), __ldg(dy + e + p), s1, s2);
                v[0] += (double)s1; v[1] += (double)s2;
            }
        }
    }
    block_sum<2>(v, sm);
    if (threadIdx.x == 0) { partial[((long long)blockIdx.y * C + c) * 2] = v[0]; partial[((long long)blockIdx.y * C + c) * 2 + 1] = v[1]; }
}

// dbeta = sum g, dgamma = sum g*xhat; dx = gamma * rstd * (g - dbeta/m - xhat * dgamma/m)
__global__ void __launch_bounds__(kT)
bn_bwd_finalize_kernel(const double* __restrict__ partial, int C, int S, float* __restrict__ dgamma, float* __restrict__ dbeta) {
    const int c = blockIdx.x * kT + threadIdx.x;
    if (c >= C) return;
    double s = 0.0, ss = 0.0;
    for (int k = 0; k < S; ++k) { s += partial[((long long)k * C + c) * 2]; ss += partial[((long long)k * C + c) * 2 + 1]; }
    dbeta[c] = (float)s;
    dgamma[c] = (float)ss;
}

template <bool VEC>
__global__ void __launch_bounds__(kT)
bn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ mean,
                    const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                    const float* __restrict__ dgamma, const float* __restrict__ dbeta, float* __restrict__ dx, int C, int HW,
                    int relu, float inv_count) {
    const int c = blockIdx.x % C;
    const float mu = __ldg(mean + c), rs = __ldg(rstd + c), ga = __ldg(gamma + c), be = __ldg(beta + c);
    const float k0 = ga * rs, k1 = __ldg(dbeta + c) * inv_count, k2 = __ldg(dgamma + c) * inv_count;
    const long long base = (long long)blockIdx.x * HW;
    const int lo = blockIdx.y * kEltChunk, hi = min(HW, lo + kEltChunk);
    auto one = [&](float xv, float g) {
        if (relu && !(bn_value(xv, mu, rs, ga, be) > 0.f)) g = 0.f;
        return k0 * (g - k1 - (xv - mu) * rs * k2);
    };
    if (VEC) {
        const float4* xp = reinterpret_cast<const float4*>(x + base);
        const float4* gp = reinterpret_cast<const float4*>(dy + base);
        float4* op = reinterpret_cast<float4*>(dx + base);
#pragma unroll 2
        for (int q = (lo >> 2) + threadIdx.x; q < (hi >> 2); q += kT) {
            const float4 t = __ldg(xp + q), g = __ldg(gp + q);
            float4 o;
            o.x = one(t.x, g.x); o.y = one(t.y, g.y); o.z = one(t.z, g.z); o.w = one(t.w, g.w);
            op[q] = o;
        }
    } else {
#pragma unroll 4
        for (int p = lo + threadIdx.x; p < hi; p += kT) dx[base + p] = one(__ldg(x + base + p), __ldg(dy + base + p));
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// SoftmaxCrossEntropyOHEMLoss (utils/loss.py:143-182) on the device
//   1. prob[p] = softmax(logits[:, p])[label[p]] for valid pixels (label != ignore), +inf-like sentinel otherwise
//   2. keep-all if min_kept >= num_valid; else threshold = thresh, and if the min(num_valid, min_kept)-th smallest prob
//      exceeds thresh the threshold becomes that value (loss.py:166-171): an exact order statistic by 4-pass radix select
//      on the float bit patterns (probabilities are non-negative, so they order like unsigned integers)
//   3. weighted cross entropy over the kept pixels, mean-reduced by the kept weights (nn.CrossEntropyLoss(weight, ignore))
// ---------------------------------------------------------------------------------------------------------------------
// state (device, int64/uint32 words): [0] num_valid, [1] selected prefix, [2] remaining rank, [3] threshold bits, [4] keep_all
__global__ void __launch_bounds__(kT)
ohem_prob_kernel(const float* __restrict__ logits, const long long* __restrict__ label, float* __restrict__ prob, int C, int HW,
                 long long npix, long long ignore, unsigned long long* __restrict__ state) {
    unsigned int valid = 0;
    for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < npix; i += (long long)gridDim.x * kT) {
        const long long n = i / HW, p = i % HW;
        const long long lab = label[i];
        float out = __uint_as_float(0x7f800000u);      // +inf: never selected, never kept
        if (lab != ignore) {
            valid += 1;
            const float* lp = logits + n * C * HW + p;
            float mx = -FLT_MAX;
            for (int c = 0; c < C; ++c) mx = fmaxf(mx, __ldg(lp + (long long)c * HW));
            float sum = 0.f, el = 0.f;
            for (int c = 0; c < C; ++c) {           // same order as numpy's axis-0 sum (loss.py:153-155)
                const float e = expf(__ldg(lp + (long long)c * HW) - mx);
                sum += e;
                if (c == lab) el = e;
            }
            out = el / sum;
        }
        prob[i] = out;
    }
    valid = (unsigned int)warp_sumf((float)valid) ;   // counts per warp stay below 2^24
    if ((threadIdx.x & 31) == 0 && valid) atomicAdd(state, (unsigned long long)valid);
}

// one radix pass: histogram of byte `shift/8` over the values whose higher bytes equal the selected prefix
__global__ void __launch_bounds__(kT)
ohem_hist_kernel(const float* __restrict__ prob, long long npix, int shift, const unsigned long long* __restrict__ state,
                 unsigned int* __restrict__ hist) {
    __shared__ unsigned int h[256];
    h[threadIdx.x] = 0u;
    __syncthreads();
    const unsigned int prefix = (unsigned int)state[1];
    const unsigned int mask = shift == 24 ? 0u : (0xffffffffu << (shift + 8));
    for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < npix; i += (long long)gridDim.x * kT) {
        const unsigned int b = __float_as_uint(prob[i]);
        if (b < 0x7f800000u && (b & mask) == (prefix & mask)) atomicAdd(&h[(b >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&hist[threadIdx.x], h[threadIdx.x]);
}

// picks the bucket that holds the wanted rank, narrows prefix / rank; after the last pass fixes the threshold.
// One CTA of 256 threads: thread b owns bucket b, an inclusive scan over the counts finds the bucket in a few steps.
__global__ void __launch_bounds__(256)
ohem_select_kernel(unsigned long long* __restrict__ state, unsigned int* __restrict__ hist, int shift, int min_kept, float thresh) {
    __shared__ unsigned long long scan[256];
    __shared__ unsigned long long rank_s;
    __shared__ unsigned int prefix_s;
    const int b = threadIdx.x;
    if (b == 0) {
        if (shift == 24) {      // first pass: decide the mode
            const unsigned long long nv = state[0];
            state[4] = ((unsigned long long)min_kept >= nv) ? 1ull : 0ull;
            state[1] = 0ull;
            const unsigned long long k = nv < (unsigned long long)min_kept ? nv : (unsigned long long)min_kept;
            state[2] = k > 0 ? k - 1 : 0;      // 0-based rank of the order statistic (loss.py:168)
            state[3] = (unsigned long long)__float_as_uint(thresh);
        }
        rank_s = state[2];
        prefix_s = (unsigned int)state[1];
    }
    const unsigned int cnt = hist[b];
    hist[b] = 0u;
    scan[b] = cnt;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {      // Hillis-Steele inclusive scan
        const unsigned long long v = b >= o ? scan[b - o] : 0ull;
        __syncthreads();
        scan[b] += v;
        __syncthreads();
    }
    const unsigned long long rank = rank_s, incl = scan[b], excl = incl - cnt;
    if (rank >= excl && rank < incl) {       // exactly one bucket holds the rank (none when there are no valid pixels)
        const unsigned int prefix = prefix_s | ((unsigned int)b << shift);
        state[1] = prefix;
        state[2] = rank - excl;
        if (shift == 0 && min_kept > 0 && state[0] > 0) {
            const float kth = __uint_as_float(prefix);
            if (kth > thresh) state[3] = (unsigned long long)prefix;      // loss.py:170-171
        }
    }
}

// per pixel: kept = valid && (keep_all || prob <= threshold); partial sums {sum w*nll, sum w} (double) per CTA
__global__ void __launch_bounds__(kT)
ohem_loss_kernel(const float* __restrict__ logits, const long long* __restrict__ label, const float* __restrict__ prob,
                 const float* __restrict__ weight, int C, int HW, long long npix, long long ignore,
                 const unsigned long long* __restrict__ state, double* __restrict__ partial) {
    __shared__ double sm[2 * 8];
    const bool keep_all = state[4] != 0ull;
    const float thr = __uint_as_float((unsigned int)state[3]);
    double v[2] = {0.0, 0.0};
    for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < npix; i += (long long)gridDim.x * kT) {
        const long long lab = label[i];
        if (lab == ignore || !(keep_all || prob[i] <= thr)) continue;
        const long long n = i / HW, p = i % HW;
        const float* lp = logits + n * C * HW + p;
        float mx = -FLT_MAX;
        for (int c = 0; c < C; ++c) mx = fmaxf(mx, __ldg(lp + (long long)c * HW));
        float sum = 0.f;
        for (int c = 0; c < C; ++c) sum += expf(__ldg(lp + (long long)c * HW) - mx);
        const float nll = -((__ldg(lp + lab * HW) - mx) - logf(sum));
        const float w = weight ? __ldg(weight + lab) : 1.f;
        v[0] += (double)w * (double)nll;
        v[1] += (double)w;
    }
    block_sum<2>(v, sm);
    if (threadIdx.x == 0) { partial[blockIdx.x * 2] = v[0]; partial[blockIdx.x * 2 + 1] = v[1]; }
}

// out[0] = loss = sum w*nll / sum w, out[1] = sum w, out[2] = kept-mode threshold (for inspection); one CTA sums the partials
__global__ void __launch_bounds__(kT)
ohem_finalize_kernel(const double* __restrict__ partial, int nblocks, const unsigned long long* __restrict__ state,
                     float* __restrict__ out) {
    __shared__ double sm[2 * 8];
    double v[2] = {0.0, 0.0};
    for (int i = threadIdx.x; i < nblocks; i += kT) { v[0] += partial[2 * i]; v[1] += partial[2 * i + 1]; }
    block_sum<2>(v, sm);
    if (threadIdx.x != 0) return;
    out[0] = (float)(v[0] / v[1]);     // 0/0 = NaN when nothing is kept, like torch's mean over an empty selection
    out[1] = (float)v[1];
    out[2] = __uint_as_float((unsigned int)state[3]);
}

// dlogits[n,c,p] = gout * w[label]/sum_w * (softmax_c - [c == label]) on kept pixels, 0 elsewhere
__global__ void __launch_bounds__(kT)
ohem_grad_kernel(const float* __restrict__ logits, const long long* __restrict__ label, const float* __restrict__ prob,
                 const float* __restrict__ weight, int C, int HW, long long npix, long long ignore,
                 const unsigned long long* __restrict__ state, const float* __restrict__ loss_out, const float* __restrict__ gout,
                 float* __restrict__ dlogits) {
    const bool keep_all = state[4] != 0ull;
    const float thr = __uint_as_float((unsigned int)state[3]);
    const float scale = __ldg(gout) / __ldg(loss_out + 1);
    for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < npix; i += (long long)gridDim.x * kT) {
        const long long n = i / HW, p = i % HW;
        const long long lab = label[i];
        const float* lp = logits + n * C * HW + p;
        float* dp = dlogits + n * C * HW + p;
        if (lab == ignore || !(keep_all || prob[i] <= thr)) {
            for (int c = 0; c < C; ++c) dp[(long long)c * HW] = 0.f;
            continue;
        }
        float mx = -FLT_MAX;
        for (int c = 0; c < C; ++c) mx = fmaxf(mx, __ldg(lp + (long long)c * HW));
        float sum = 0.f;
        for (int c = 0; c < C; ++c) sum += expf(__ldg(lp + (long long)c * HW) - mx);
        const float w = (weight ? __ldg(weight + lab) : 1.f) * scale, inv = 1.f / sum;
        for (int c = 0; c < C; ++c) {
            const float sm_c = expf(__ldg(lp + (long long)c * HW) - mx) * inv;
            dp[(long long)c * HW] = w * (sm_c - (c == lab ? 1.f : 0.f));
        }
    }
}

// F.interpolate(mode='bilinear', align_corners=True): the arithmetic of the eval path's kernels
__device__ __forceinline__ void ac_coord(int o, float sc, int in, int& i0, int& i1, float& l) {
    const float f = sc * (float)o;
    i0 = min((int)f, in - 1);
    i1 = min(i0 + 1, in - 1);
    l = f - (float)i0;
}

// ---------------------------------------------------------------------------------------------------------------------
// The loss of a head straight from its LOW-RESOLUTION logits: F.interpolate(x, size, 'bilinear', align_corners=True)
// (fast_scnn.py:40, :44) composed with the OHEM cross entropy.  The full-resolution logits (16 x 19 x 768 x 768 fp32 = 717 MB per
// head at BASELINE config 5) and their gradient never exist: every kernel re-interpolates the 19 values of a pixel from the
// L2-resident low-resolution tensor with exactly bilinear_fwd_kernel's arithmetic, so the selection and the loss are bit-identical
// to the unfused pair.  The backward scatters d loss / d logit through the resize's transpose: a CTA owns 64 rows x 32 columns of
// pixels, reduces each class's contribution over the lanes that share a low-resolution column (segmented shuffle), accumulates into
// a shared-memory tile of the low-resolution gradient and flushes it with float atomics (summation order is not deterministic).
// ---------------------------------------------------------------------------------------------------------------------
struct UpGeom { int C, hl, wl, H, W; float scy, scx; };

__device__ __forceinline__ float up_value(const float* __restrict__ base, int wl, int y0, int y1, int x0, int x1, float ly, float lx) {
    const float top = fmaf(lx, __ldg(base + y0 * wl + x1), (1.f - lx) * __ldg(base + y0 * wl + x0));
    const float bot = fmaf(lx, __ldg(base + y1 * wl + x1), (1.f - lx) * __ldg(base + y1 * wl + x0));
    return fmaf(ly, bot, (1.f - ly) * top);
}

// CT > 0: class count known at compile time -> the pixel's interpolated logits live in registers and are computed once;
// CT == 0: any class count, the values are re-interpolated in every pass.
template <int CT>
struct UpVals {
    float v[CT > 0 ? CT : 1];
    __device__ __forceinline__ void load(const float* __restrict__ lp, const UpGeom& g, int y0, int y1, int x0, int x1, float ly, float lx) {
        if (CT > 0) {
#pragma unroll
            for (int c = 0; c < CT; ++c) v[c] = up_value(lp + (long long)c * g.hl * g.wl, g.wl, y0, y1, x0, x1, ly, lx);
        }
    }
    __device__ __forceinline__ float get(int c, const float* __restrict__ lp, const UpGeom& g, int y0, int y1, int x0, int x1, float ly,
                                         float lx) const {
        return CT > 0 ? v[c] : up_value(lp + (long long)c * g.hl * g.wl, g.wl, y0, y1, x0, x1, ly, lx);
    }
};
#define UP_FOR_CLASSES(c) _Pragma("unroll") for (int c = 0; c < (CT > 0 ? CT : g.C); ++c)

template <int CT>
__global__ void __launch_bounds__(kT)
ohem_up_prob_kernel(const float* __restrict__ low, const long long* __restrict__ label, float* __restrict__ prob, UpGeom g, long long npix,
                    long long ignore, unsigned long long* __restrict__ state) {
    unsigned int valid = 0;
    const long long HW = (long long)g.H * g.W;
    for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < npix; i += (long long)gridDim.x * kT) {
        const long long lab = label[i];
        float out = __uint_as_float(0x7f800000u);
        if (lab != ignore) {
            valid += 1;
            const long long n = i / HW;
            const int y = (int)((i % HW) / g.W), x = (int)(i % g.W);
            int y0, y1, x0, x1;
            float ly, lx;
            ac_coord(y, g.scy, g.hl, y0, y1, ly);
            ac_coord(x, g.scx, g.wl, x0, x1, lx);
            const float* lp = low + n * g.C * g.hl * g.wl;
            UpVals<CT> u;
            u.load(lp, g, y0, y1, x0, x1, ly, lx);
            float mx = -FLT_MAX;
            UP_FOR_CLASSES(c) mx = fmaxf(mx, u.get(c, lp, g, y0, y1, x0, x1, ly, lx));
            float sum = 0.f, el = 0.f;
            UP_FOR_CLASSES(c) {
                const float e = expf(u.get(c, lp, g, y0, y1, x0, x1, ly, lx) - mx);
                sum += e;
                if (c == lab) el = e;
            }
            out = el / sum;
        }
        prob[i] = out;
    }
    valid = (unsigned int)warp_sumf((float)valid);
    if ((threadIdx.x & 31) == 0 && valid) atomicAdd(state, (unsigned long long)valid);
}

template <int CT>
__global__ void __launch_bounds__(kT)
ohem_up_loss_kernel(const float* __restrict__ low, const long long* __restrict__ label, const float* __restrict__ prob,
                    const float* __restrict__ weight, UpGeom g, long long npix, long long ignore,
                    const unsigned long long* __restrict__ state, double* __restrict__ partial) {
    __shared__ double sm[2 * 8];
    const bool keep_all = state[4] != 0ull;
    const float thr = __uint_as_float((unsigned int)state[3]);
    const long long HW = (long long)g.H * g.W;
    double v[2] = {0.0, 0.0};
    for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < npix; i += (long long)gridDim.x * kT) {
        const long long lab = label[i];
        if (lab == ignore || !(keep_all || prob[i] <= thr)) continue;
        const long long n = i / HW;
        const int y = (int)((i % HW) / g.W), x = (int)(i % g.W);
        int y0, y1, x0, x1;
        float ly, lx;
        ac_coord(y, g.scy, g.hl, y0, y1, ly);
        ac_coord(x, g.scx, g.wl, x0, x1, lx);
        const float* lp = low + n * g.C * g.hl * g.wl;
        UpVals<CT> u;
        u.load(lp, g, y0, y1, x0, x1, ly, lx);
        float mx = -FLT_MAX;
        UP_FOR_CLASSES(c) mx = fmaxf(mx, u.get(c, lp, g, y0, y1, x0, x1, ly, lx));
        float sum = 0.f, vl = 0.f;
        UP_FOR_CLASSES(c) {
            const float t = u.get(c, lp, g, y0, y1, x0, x1, ly, lx);
            sum += expf(t - mx);
            if (c == lab) vl = t;
        }
        const float nll = -((vl - mx) - logf(sum));
        const float w = weight ? __ldg(weight + lab) : 1.f;
        v[0] += (double)w * (double)nll;
        v[1] += (double)w;
    }
    block_sum<2>(v, sm);
    if (threadIdx.x == 0) { partial[blockIdx.x * 2] = v[0]; partial[blockIdx.x * 2 + 1] = v[1]; }
}

// grid (ceil(W/32), ceil(H/64), N); block 256 = 8 warps; warp w walks rows 8w .. 8w+7 of the tile, lane = column
constexpr int kUpTR = 12, kUpTC = 8;      // low-resolution rows / columns a 64 x 32 pixel tile can touch at ratios <= 1/7
template <int CT>
__global__ void __launch_bounds__(kT)
ohem_up_grad_kernel(const float* __restrict__ low, const long long* __restrict__ label, const float* __restrict__ prob,
                    const float* __restrict__ weight, UpGeom g, long long ignore, const unsigned long long* __restrict__ state,
                    const float* __restrict__ loss_out, const float* __restrict__ gout, float* __restrict__ dlow) {
    extern __shared__ float acc[];       // [C][kUpTR][kUpTC]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = blockIdx.z, yb = blockIdx.y * 64, xb = blockIdx.x * 32;
    for (int i = tid; i < g.C * kUpTR * kUpTC; i += kT) acc[i] = 0.f;
    __syncthreads();
    const bool keep_all = state[4] != 0ull;
    const float thr = __uint_as_float((unsigned int)state[3]);
    const float scale = __ldg(gout) / __ldg(loss_out + 1);
    const int ry0 = min((int)(g.scy * (float)yb), g.hl - 1), rx0 = min((int)(g.scx * (float)xb), g.wl - 1);
    const float* lp = low + (long long)n * g.C * g.hl * g.wl;
    const int x = xb + lane;
    int x0 = 0, x1 = 0;
    float lx = 0.f;
    if (x < g.W) ac_coord(x, g.scx, g.wl, x0, x1, lx);
    // lanes that share x0 are contiguous: segment heads and the lanes each reduction step may add, once per thread
    const int prev_x0 = __shfl_up_sync(0xffffffffu, x0, 1);
    const bool head = lane == 0 || prev_x0 != x0;
    unsigned int same = 0u;            // bit k: lane + 2^k is in this lane's segment
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const int ux0 = __shfl_down_sync(0xffffffffu, x0, 1 << k);
        if (lane + (1 << k) < 32 && ux0 == x0) same |= 1u << k;
    }
    for (int rr = 0; rr < 8; ++rr) {
        const int y = yb + warp * 8 + rr;
        if (y >= g.H) break;                                   // warp-uniform
        const long long i = ((long long)n * g.H + y) * g.W + x;
        long long lab = ignore;
        bool kept = false;
        if (x < g.W) {
            lab = label[i];
            kept = lab != ignore && (keep_all || prob[i] <= thr);
        }
        if (__ballot_sync(0xffffffffu, kept) == 0u) continue;  // nothing kept in this row segment
        int y0, y1;
        float ly;
        ac_coord(y, g.scy, g.hl, y0, y1, ly);
        UpVals<CT> u;
        float mx = -FLT_MAX, sum = 1.f;
        if (kept) {
            u.load(lp, g, y0, y1, x0, x1, ly, lx);
            UP_FOR_CLASSES(c) mx = fmaxf(mx, u.get(c, lp, g, y0, y1, x0, x1, ly, lx));
            sum = 0.f;
            UP_FOR_CLASSES(c) sum += expf(u.get(c, lp, g, y0, y1, x0, x1, ly, lx) - mx);
        } else if (CT > 0) {
#pragma unroll
            for (int c = 0; c < (CT > 0 ? CT : 1); ++c) u.v[c] = 0.f;
        }
        const float w = kept ? (weight ? __ldg(weight + lab) : 1.f) * scale : 0.f, inv = 1.f / sum;
        const int ay0 = (y0 - ry0) * kUpTC, ay1 = (y1 - ry0) * kUpTC, ax0 = x0 - rx0, ax1 = x1 - rx0;
        UP_FOR_CLASSES(c) {
            float gc = 0.f;
            if (kept) gc = w * (expf(u.get(c, lp, g, y0, y1, x0, x1, ly, lx) - mx) * inv - (c == lab ? 1.f : 0.f));
            float a = gc * (1.f - lx), b = gc * lx;           // towards columns x0 and x1
            // segmented sum over the lanes that share x0: after the loop the segment head holds the total
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const float ua = __shfl_down_sync(0xffffffffu, a, 1 << k), ub = __shfl_down_sync(0xffffffffu, b, 1 << k);
                if ((same >> k) & 1u) { a += ua; b += ub; }
            }
            if (head && x < g.W && (a != 0.f || b != 0.f)) {
                float* ac = acc + c * kUpTR * kUpTC;
                atomicAdd(ac + ay0 + ax0, (1.f - ly) * a);
                atomicAdd(ac + ay0 + ax1, (1.f - ly) * b);
                atomicAdd(ac + ay1 + ax0, ly * a);
                atomicAdd(ac + ay1 + ax1, ly * b);
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < g.C * kUpTR * kUpTC; i += kT) {
        const float v = acc[i];
        if (v != 0.f) {
            const int c = i / (kUpTR * kUpTC), r = (i / kUpTC) % kUpTR, q = i % kUpTC;
            const int yy = ry0 + r, xx = rx0 + q;
            if (yy < g.hl && xx < g.wl) atomicAdd(dlow + (((long long)n * g.C + c) * g.hl + yy) * g.wl + xx, v);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Strip form of the three kernels above for a class count known at compile time (2 / 19).  A warp owns 32 adjacent output columns
// and walks kStripRows rows of them; a thread (one column) keeps, per class, the two horizontally interpolated low-resolution rows
// its column sits between (top / bot: the inner two fmaf of up_value, refreshed only when the row pair changes, every ~8 rows at
// ratio 1/8), so a pixel costs one vertical fmaf per class instead of four loads and three interpolations: the values are
// bit-identical to up_value.  The backward accumulates a thread's gradient per class and low-resolution row in registers while the
// row pair stays the same, then reduces over the lanes that share a low-resolution column (segmented shuffle) and adds the four
// corner sums straight into dlow (red.global.add.f32; summation order is not deterministic): no shared memory, ~8x fewer
// reductions and atomics than one per row.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kStripRows = 32, kStripAhead = 4;

template <int CT>
struct UpRows {
    float top[CT], bot[CT];
    int y0, y1;
    __device__ __forceinline__ void fetch(float (&dst)[CT], const float* __restrict__ lp, const UpGeom& g, int yy, int x0, int x1, float lx) {
        const float* r = lp + yy * g.wl;
        const int plane = g.hl * g.wl;
#pragma unroll
        for (int c = 0; c < CT; ++c) dst[c] = fmaf(lx, __ldg(r + c * plane + x1), (1.f - lx) * __ldg(r + c * plane + x0));
    }
    // (ny0, ny1) is warp-uniform
    __device__ __forceinline__ void advance(const float* __restrict__ lp, const UpGeom& g, int ny0, int ny1, int x0, int x1, float lx) {
        if (ny0 == y0 && ny1 == y1) return;
        if (ny0 == y1) {
#pragma unroll
            for (int c = 0; c < CT; ++c) top[c] = bot[c];
        } else {
            fetch(top, lp, g, ny0, x0, x1, lx);
        }
        if (ny1 == ny0) {
#pragma unroll
            for (int c = 0; c < CT; ++c) bot[c] = top[c];
        } else {
            fetch(bot, lp, g, ny1, x0, x1, lx);
        }
        y0 = ny0; y1 = ny1;
    }
    __device__ __forceinline__ float value(int c, float ly) const { return fmaf(ly, bot[c], (1.f - ly) * top[c]); }
};

struct StripIter {          // strips ordered (image, row block, column strip): neighbouring warps share low-resolution rows
    int CS, RB;
    long long total;
    __device__ __forceinline__ StripIter(const UpGeom& g, int n) : CS((g.W + 31) / 32), RB((g.H + kStripRows - 1) / kStripRows) {
        total = (long long)n * RB * CS;
    }
    __device__ __forceinline__ void at(long long s, int& n, int& yb, int& xb) const {
        xb = (int)(s % CS) * 32;
        yb = (int)((s / CS) % RB) * kStripRows;
        n = (int)(s / ((long long)CS * RB));
    }
};

template <int CT>
__global__ void __launch_bounds__(kT)
ohem_strip_prob_kernel(const float* __restrict__ low, const long long* __restrict__ label, float* __restrict__ prob,
                       float* __restrict__ nll_out, UpGeom g, int nimg, long long ignore, unsigned long long* __restrict__ state) {
    const int lane = threadIdx.x & 31;
    const StripIter it(g, nimg);
    unsigned int valid = 0;
    for (long long s = (long long)blockIdx.x * (kT / 32) + (threadIdx.x >> 5); s < it.total; s += (long long)gridDim.x * (kT / 32)) {
        int n, yb, xb;
        it.at(s, n, yb, xb);
        const int x = xb + lane;
        int x0, x1;
        float lx;
        ac_coord(min(x, g.W - 1), g.scx, g.wl, x0, x1, lx);
        const float* lp = low + (long long)n * CT * g.hl * g.wl;
        UpRows<CT> rows;
        rows.y0 = rows.y1 = -1;
        const int yend = min(g.H, yb + kStripRows);
        for (int yq = yb; yq < yend; yq += kStripAhead) {
        long long labs[kStripAhead];          // the labels of kStripAhead rows are requested before the first one is used
#pragma unroll
        for (int j = 0; j < kStripAhead; ++j)
            labs[j] = (x < g.W && yq + j < yend) ? __ldg(label + ((long long)n * g.H + yq + j) * g.W + x) : ignore;
#pragma unroll
        for (int j = 0; j < kStripAhead; ++j) {
            const int y = yq + j;
            if (y >= yend) break;
            int y0, y1;
            float ly;
            ac_coord(y, g.scy, g.hl, y0, y1, ly);
            rows.advance(lp, g, y0, y1, x0, x1, lx);
            if (x >= g.W) continue;
            const long long i = ((long long)n * g.H + y) * g.W + x;
            const long long lab = labs[j];
            float out = __uint_as_float(0x7f800000u), nll = 0.f;
            if (lab != ignore) {
                valid += 1;
                float v[CT], mx = -FLT_MAX;
#pragma unroll
                for (int c = 0; c < CT; ++c) { v[c] = rows.value(c, ly); mx = fmaxf(mx, v[c]); }
                float sum = 0.f, el = 0.f, vl = 0.f;
#pragma unroll
                for (int c = 0; c < CT; ++c) {
                    const float e = expf(v[c] - mx);
                    sum += e;
                    if (c == lab) { el = e; vl = v[c]; }
                }
                out = el / sum;
                nll = -((vl - mx) - logf(sum));      // the loss kernel's own expression on the same values: it only has to read it
            }
            prob[i] = out;
            if (nll_out) nll_out[i] = nll;
        }
        }
    }
    valid = (unsigned int)warp_sumf((float)valid);
    if (lane == 0 && valid) atomicAdd(state, (unsigned long long)valid);
}

// the loss over the kept pixels from the per-pixel negative log-likelihood the probability kernel left behind
__global__ void __launch_bounds__(kT)
ohem_nll_loss_kernel(const long long* __restrict__ label, const float* __restrict__ prob, const float* __restrict__ nll,
                     const float* __restrict__ weight, long long npix, long long ignore, const unsigned long long* __restrict__ state,
                     double* __restrict__ partial) {
    __shared__ double sm[2 * 8];
    const bool keep_all = state[4] != 0ull;
    const float thr = __uint_as_float((unsigned int)state[3]);
    double acc[2] = {0.0, 0.0};
#pragma unroll 4
    for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < npix; i += (long long)gridDim.x * kT) {
        const long long lab = __ldg(label + i);
        if (lab == ignore || !(keep_all || __ldg(prob + i) <= thr)) continue;
        const float w = weight ? __ldg(weight + lab) : 1.f;
        acc[0] += (double)w * (double)__ldg(nll + i);
        acc[1] += (double)w;
    }
    block_sum<2>(acc, sm);
    if (threadIdx.x == 0) { partial[blockIdx.x * 2] = acc[0]; partial[blockIdx.x * 2 + 1] = acc[1]; }
}

template <int CT>
__global__ void __launch_bounds__(kT)
ohem_strip_loss_kernel(const float* __restrict__ low, const long long* __restrict__ label, const float* __restrict__ prob,
                       const float* __restrict__ weight, UpGeom g, int nimg, long long ignore,
                       const unsigned long long* __restrict__ state, double* __restrict__ partial) {
    __shared__ double sm[2 * 8];
    const int lane = threadIdx.x & 31;
    const bool keep_all = state[4] != 0ull;
    const float thr = __uint_as_float((unsigned int)state[3]);
    const StripIter it(g, nimg);
    double acc[2] = {0.0, 0.0};
    for (long long s = (long long)blockIdx.x * (kT / 32) + (threadIdx.x >> 5); s < it.total; s += (long long)gridDim.x * (kT / 32)) {
        int n, yb, xb;
        it.at(s, n, yb, xb);
        const int x = xb + lane;
        int x0, x1;
        float lx;
        ac_coord(min(x, g.W - 1), g.scx, g.wl, x0, x1, lx);
        const float* lp = low + (long long)n * CT * g.hl * g.wl;
        UpRows<CT> rows;
        rows.y0 = rows.y1 = -1;
        const int yend = min(g.H, yb + kStripRows);
        for (int yq = yb; yq < yend; yq += kStripAhead) {
        long long labs[kStripAhead];
        float prs[kStripAhead];
#pragma unroll
        for (int j = 0; j < kStripAhead; ++j) {
            const bool in = x < g.W && yq + j < yend;
            const long long i = ((long long)n * g.H + yq + j) * g.W + x;
            labs[j] = in ? __ldg(label + i) : ignore;
            prs[j] = in ? __ldg(prob + i) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < kStripAhead; ++j) {
            const int y = yq + j;
            if (y >= yend) break;
            const long long lab = labs[j];
            const bool kept = lab != ignore && (keep_all || prs[j] <= thr);
            if (__ballot_sync(0xffffffffu, kept) == 0u) continue;
            int y0, y1;
            float ly;
            ac_coord(y, g.scy, g.hl, y0, y1, ly);
            rows.advance(lp, g, y0, y1, x0, x1, lx);
            if (!kept) continue;
            float v[CT], mx = -FLT_MAX;
#pragma unroll
            for (int c = 0; c < CT; ++c) { v[c] = rows.value(c, ly); mx = fmaxf(mx, v[c]); }
            float sum = 0.f, vl = 0.f;
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                sum += expf(v[c] - mx);
                if (c == lab) vl = v[c];
            }
            const float nll = -((vl - mx) - logf(sum));
            const float w = weight ? __ldg(weight + lab) : 1.f;
            acc[0] += (double)w * (double)nll;
            acc[1] += (double)w;
        }
        }
    }
    block_sum<2>(acc, sm);
    if (threadIdx.x == 0) { partial[blockIdx.x * 2] = acc[0]; partial[blockIdx.x * 2 + 1] = acc[1]; }
}

// (two CTAs per SM: the 128-register cap spills ~200 bytes per thread, and is still 1.5x faster than 210 registers at one CTA per
// SM -- the kernel is occupancy-bound; keeping the row cache in shared memory instead was measured 6 % slower than the spills)
template <int CT>
__global__ void __launch_bounds__(kT, 2)
ohem_strip_grad_kernel(const float* __restrict__ low, const long long* __restrict__ label, const float* __restrict__ prob,
                       const float* __restrict__ weight, UpGeom g, int nimg, long long ignore,
                       const unsigned long long* __restrict__ state, const float* __restrict__ loss_out, const float* __restrict__ gout,
                       float* __restrict__ dlow) {
    const int lane = threadIdx.x & 31;
    const bool keep_all = state[4] != 0ull;
    const float thr = __uint_as_float((unsigned int)state[3]);
    const float scale = __ldg(gout) / __ldg(loss_out + 1);
    const StripIter it(g, nimg);
    const int plane = g.hl * g.wl;
    for (long long s = (long long)blockIdx.x * (kT / 32) + (threadIdx.x >> 5); s < it.total; s += (long long)gridDim.x * (kT / 32)) {
        int n, yb, xb;
        it.at(s, n, yb, xb);
        const int x = xb + lane;
        int x0, x1;
        float lx;
        ac_coord(min(x, g.W - 1), g.scx, g.wl, x0, x1, lx);
        // lanes that share x0 are contiguous: segment heads and the lanes each reduction step may add
        const int prev_x0 = __shfl_up_sync(0xffffffffu, x0, 1);
        const bool head = lane == 0 || prev_x0 != x0;
        unsigned int same = 0u;            // bit k: lane + 2^k is in this lane's segment
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const int ux0 = __shfl_down_sync(0xffffffffu, x0, 1 << k);
            if (lane + (1 << k) < 32 && ux0 == x0) same |= 1u << k;
        }
        const float* lp = low + (long long)n * CT * plane;
        float* dp = dlow + (long long)n * CT * plane;
        UpRows<CT> rows;
        rows.y0 = rows.y1 = -1;
        float a0[CT], a1[CT];              // gradient towards rows y0 / y1 of this thread's column, all classes
#pragma unroll
        for (int c = 0; c < CT; ++c) a0[c] = a1[c] = 0.f;
        bool dirty = false;                // warp-uniform: something was accumulated since the last flush
        auto flush = [&]() {
            const float hx = 1.f - lx;
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                float p = a0[c] * hx, q = a0[c] * lx, r = a1[c] * hx, t = a1[c] * lx;
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    const float up = __shfl_down_sync(0xffffffffu, p, 1 << k), uq = __shfl_down_sync(0xffffffffu, q, 1 << k);
                    const float ur = __shfl_down_sync(0xffffffffu, r, 1 << k), ut = __shfl_down_sync(0xffffffffu, t, 1 << k);
                    if ((same >> k) & 1u) { p += up; q += uq; r += ur; t += ut; }
                }
                if (head) {
                    float* d0 = dp + c * plane + rows.y0 * g.wl;
                    float* d1 = dp + c * plane + rows.y1 * g.wl;
                    if (p != 0.f) atomicAdd(d0 + x0, p);
                    if (q != 0.f) atomicAdd(d0 + x1, q);
                    if (r != 0.f) atomicAdd(d1 + x0, r);
                    if (t != 0.f) atomicAdd(d1 + x1, t);
                }
                a0[c] = a1[c] = 0.f;
            }
        };
        const int yend = min(g.H, yb + kStripRows);
        for (int yq = yb; yq < yend; yq += kStripAhead) {
        long long labs[kStripAhead];
        float prs[kStripAhead];
#pragma unroll
        for (int j = 0; j < kStripAhead; ++j) {
            const bool in = x < g.W && yq + j < yend;
            const long long i = ((long long)n * g.H + yq + j) * g.W + x;
            labs[j] = in ? __ldg(label + i) : ignore;
            prs[j] = in ? __ldg(prob + i) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < kStripAhead; ++j) {
            const int y = yq + j;
            if (y >= yend) break;
            const long long lab = labs[j];
            const bool kept = lab != ignore && (keep_all || prs[j] <= thr);
            if (__ballot_sync(0xffffffffu, kept) == 0u) continue;
            int y0, y1;
            float ly;
            ac_coord(y, g.scy, g.hl, y0, y1, ly);
            if (dirty && (y0 != rows.y0 || y1 != rows.y1)) { flush(); dirty = false; }
            rows.advance(lp, g, y0, y1, x0, x1, lx);
            dirty = true;
            if (!kept) continue;
            float v[CT], mx = -FLT_MAX;
#pragma unroll
            for (int c = 0; c < CT; ++c) { v[c] = rows.value(c, ly); mx = fmaxf(mx, v[c]); }
            float sum = 0.f;
#pragma unroll
            for (int c = 0; c < CT; ++c) { v[c] = __expf(v[c] - mx); sum += v[c]; }      // (ex2.approx: 2e-7 relative here, the gradient tolerates it; the selection and the loss keep expf)
            const float w = (weight ? __ldg(weight + lab) : 1.f) * scale, inv = 1.f / sum, hy = 1.f - ly;
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                const float gc = w * (v[c] * inv - (c == lab ? 1.f : 0.f));
                a0[c] = fmaf(hy, gc, a0[c]);
                a1[c] = fmaf(ly, gc, a1[c]);
            }
        }
        }
        if (dirty) flush();
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// The stem, LearningToDownsample.conv = nn.Conv2d(3, 32, 3, stride 2, padding 0) (models/fast_scnn.py:153, :24-31), directly: through
// im2col its 27-row column matrix (253 MB at config 5, rows of 146 689 pixels that no 16-byte copy can take) is written once and
// read twice.  Same walk as the stride-2 depthwise: a lane owns an output column and carries the 3 x 3 x 3 input window in
// registers down the rows; the 27 x 32 weights sit in shared memory k-major and are read as broadcast float4.
//   forward : 32 accumulators per thread, 864 FMAs per pixel against 216 broadcast LDS.128
//   weight gradient : the 32 output channels in 8 groups of 4 (blockIdx.y), 4 x 27 accumulators per thread over all of a warp's
//                     pixels, then warp shuffle -> shared memory over the CTA's warps -> 108 double atomics per CTA
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kStemCo = 32, kStemK = 27, kStemRows = 16;

struct StemWin { float v[3][3][3]; };      // [channel][row of the window][column]

// rows 2 oy + r0 .. of the window for the three channels; columns 2 ox, 2 ox + 1, 2 ox + 2 (the last one is the next lane's first)
template <bool VEC2>
__device__ __forceinline__ void stem_load_row(const float* __restrict__ xn, int plane, int iy, int H, int W, int ox, int lane, bool want,
                                              float (&e)[3], float (&o)[3], float (&edge)[3]) {
    const bool rok = want && iy < H;
    const int ix = 2 * ox;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float* r = xn + (long long)c * plane + (long long)(rok ? iy : 0) * W;
        if (VEC2) {
            const float2 t = (rok && ix + 1 < W) ? __ldg(reinterpret_cast<const float2*>(r + ix)) : make_float2(0.f, 0.f);
            e[c] = t.x; o[c] = t.y;
        } else {
            e[c] = (rok && ix < W) ? __ldg(r + ix) : 0.f;
            o[c] = (rok && ix + 1 < W) ? __ldg(r + ix + 1) : 0.f;
        }
        edge[c] = (lane == 31 && rok && ix + 2 < W) ? __ldg(r + ix + 2) : 0.f;
    }
}
__device__ __forceinline__ void stem_set_row(StemWin& w, int row, const float (&e)[3], const float (&o)[3], const float (&edge)[3], int lane) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float nx = __shfl_down_sync(kFull, e[c], 1);
        w.v[c][row][0] = e[c]; w.v[c][row][1] = o[c]; w.v[c][row][2] = lane == 31 ? edge[c] : nx;
    }
}

struct StemTasks {      // tasks ordered (image, row block, column block)
    int RB, CB;
    long long total;
    __device__ __forceinline__ void at(long long t, int& n, int& y0, int& x0) const {
        x0 = (int)(t % CB) * 32;
        y0 = (int)((t / CB) % RB) * kStemRows;
        n = (int)(t / ((long long)CB * RB));
    }
};

template <bool VEC2>
__global__ void __launch_bounds__(kT)
stem_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y, int H, int W, int Ho, int Wo, StemTasks tk) {
    __shared__ __align__(16) float ws[kStemK][kStemCo];
    for (int i = threadIdx.x; i < kStemK * kStemCo; i += kT) ws[i % kStemK][i / kStemK] = __ldg(w + i);      // w[co][k] -> ws[k][co]
    __syncthreads();
    const int lane = threadIdx.x & 31, plane = H * W;
    for (long long t = (long long)blockIdx.x * (kT / 32) + (threadIdx.x >> 5); t < tk.total; t += (long long)gridDim.x * (kT / 32)) {
        int n, oy0, ox0;
        tk.at(t, n, oy0, ox0);
        const int ox = ox0 + lane, oy1 = min(Ho, oy0 + kStemRows);
        const float* xn = x + (long long)n * 3 * plane;
        float* yn = y + (long long)n * kStemCo * Ho * Wo;
        StemWin win;
        float e[3], o[3], ed[3];
        stem_load_row<VEC2>(xn, plane, 2 * oy0, H, W, ox, lane, true, e, o, ed);
        stem_set_row(win, 0, e, o, ed, lane);
        for (int oy = oy0; oy < oy1; ++oy) {
            float e1[3], o1[3], d1[3], e2[3], o2[3], d2[3];
            stem_load_row<VEC2>(xn, plane, 2 * oy + 1, H, W, ox, lane, true, e1, o1, d1);
            stem_load_row<VEC2>(xn, plane, 2 * oy + 2, H, W, ox, lane, true, e2, o2, d2);
            stem_set_row(win, 1, e1, o1, d1, lane);
            stem_set_row(win, 2, e2, o2, d2, lane);
            float acc[kStemCo];
#pragma unroll
            for (int c = 0; c < kStemCo; ++c) acc[c] = 0.f;
#pragma unroll
            for (int ci = 0; ci < 3; ++ci)
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const float v = win.v[ci][ky][kx];
                        const float4* wr = reinterpret_cast<const float4*>(ws[ci * 9 + ky * 3 + kx]);
#pragma unroll
                        for (int q = 0; q < kStemCo / 4; ++q) {
                            const float4 w4 = wr[q];
                            acc[4 * q] = fmaf(v, w4.x, acc[4 * q]); acc[4 * q + 1] = fmaf(v, w4.y, acc[4 * q + 1]);
                            acc[4 * q + 2] = fmaf(v, w4.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(v, w4.w, acc[4 * q + 3]);
                        }
                    }
            if (ox < Wo) {
                float* yp = yn + (long long)oy * Wo + ox;
#pragma unroll
                for (int c = 0; c < kStemCo; ++c) yp[(long long)c * Ho * Wo] = acc[c];
            }
#pragma unroll
            for (int ci = 0; ci < 3; ++ci)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) win.v[ci][0][kx] = win.v[ci][2][kx];      // row 2 oy + 2 is row 0 of the next window
        }
    }
}

// grid (pixel CTAs, 8 channel groups); 128 threads
constexpr int kStemWT = 128;
template <bool VEC2>
__global__ void __launch_bounds__(kStemWT)
stem_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, double* __restrict__ dwacc, int H, int W, int Ho, int Wo,
                  StemTasks tk) {
    __shared__ float red[kStemWT / 32][4 * kStemK];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, plane = H * W, g = blockIdx.y;
    float acc[4][kStemK];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < kStemK; ++k) acc[i][k] = 0.f;
    for (long long t = (long long)blockIdx.x * (kStemWT / 32) + warp; t < tk.total; t += (long long)gridDim.x * (kStemWT / 32)) {
        int n, oy0, ox0;
        tk.at(t, n, oy0, ox0);
        const int ox = ox0 + lane, oy1 = min(Ho, oy0 + kStemRows);
        const float* xn = x + (long long)n * 3 * plane;
        const float* gn = dy + ((long long)n * kStemCo + 4 * g) * Ho * Wo;
        StemWin win;
        float e[3], o[3], ed[3];
        stem_load_row<VEC2>(xn, plane, 2 * oy0, H, W, ox, lane, true, e, o, ed);
        stem_set_row(win, 0, e, o, ed, lane);
        for (int oyb = oy0; oyb < oy1; oyb += 2) {      // two output rows per batch of loads (four input rows + eight gradients)
            float re[4][3], ro[4][3], rd[4][3], gv[2][4];
#pragma unroll
            for (int r = 0; r < 4; ++r) stem_load_row<VEC2>(xn, plane, 2 * oyb + 1 + r, H, W, ox, lane, oyb + (r >> 1) < oy1, re[r], ro[r], rd[r]);
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    gv[u][i] = (oyb + u < oy1 && ox < Wo) ? __ldg(gn + (long long)i * Ho * Wo + (long long)(oyb + u) * Wo + ox) : 0.f;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                stem_set_row(win, 1, re[2 * u], ro[2 * u], rd[2 * u], lane);
                stem_set_row(win, 2, re[2 * u + 1], ro[2 * u + 1], rd[2 * u + 1], lane);
#pragma unroll
                for (int ci = 0; ci < 3; ++ci)
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) {
                            const float v = win.v[ci][ky][kx];
#pragma unroll
                            for (int i = 0; i < 4; ++i) acc[i][ci * 9 + ky * 3 + kx] = fmaf(gv[u][i], v, acc[i][ci * 9 + ky * 3 + kx]);
                        }
#pragma unroll
                for (int ci = 0; ci < 3; ++ci)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) win.v[ci][0][kx] = win.v[ci][2][kx];
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < kStemK; ++k) {
            const float s = warp_sumf(acc[i][k]);
            if (lane == 0) red[warp][i * kStemK + k] = s;
        }
    __syncthreads();
    if (threadIdx.x < 4 * kStemK) {
        double tot = 0.0;
#pragma unroll
        for (int wq = 0; wq < kStemWT / 32; ++wq) tot += (double)red[wq][threadIdx.x];
        if (tot != 0.0) atomicAdd(dwacc + (4 * g) * kStemK + threadIdx.x, tot);      // dw[co][k], co = 4 g + i: contiguous over (i, k)
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// the rest of the network's training-mode operators: dense 3x3 convolution (stem, aux head) through im2col + the GEMM
// above, per-channel bias, bilinear resize with align_corners=True, adaptive average pooling (overlapping bins), dropout,
// add + ReLU, and the SGD update
// ---------------------------------------------------------------------------------------------------------------------
// cols[n][(ci*9 + ky*3 + kx)][oy*Wo + ox] = x[n][ci][oy*s + ky - pad][ox*s + kx - pad] (0 outside)
// grid (n * C * 9 column rows, chunks of kEltChunk pixels): the tap of a row is decoded once per CTA, a pixel costs one 32-bit division
__global__ void __launch_bounds__(kT)
im2col3x3_kernel(const float* __restrict__ x, float* __restrict__ cols, int C, int H, int W, int Ho, int Wo, int stride, int pad) {
    const int k = blockIdx.x % (C * 9), n = blockIdx.x / (C * 9);
    const int ci = k / 9, ky = (k % 9) / 3, kx = k % 3;
    const float* xp = x + ((long long)n * C + ci) * H * W;
    float* cp = cols + (long long)blockIdx.x * Ho * Wo;
    const int lo = blockIdx.y * kEltChunk, hi = min(Ho * Wo, lo + kEltChunk);
#pragma unroll 4
    for (int p = lo + threadIdx.x; p < hi; p += kT) {
        const int oy = p / Wo, ox = p - oy * Wo;
        const int iy = oy * stride + ky - pad, ix = ox * stride + kx - pad;
        cp[p] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? __ldg(xp + (long long)iy * W + ix) : 0.f;
    }
}

// dx[n][ci][iy][ix] = sum over the taps that read this pixel of dcols[n][ci*9 + ky*3 + kx][oy*Wo + ox]
// grid (n * C planes, chunks of kEltChunk input pixels)
__global__ void __launch_bounds__(kT)
col2im3x3_kernel(const float* __restrict__ dcols, float* __restrict__ dx, int C, int H, int W, int Ho, int Wo, int stride, int pad) {
    const float* dc = dcols + (long long)blockIdx.x * 9 * Ho * Wo;      // rows (n * C + ci) * 9 + tap
    float* xp = dx + (long long)blockIdx.x * H * W;
    const int lo = blockIdx.y * kEltChunk, hi = min(H * W, lo + kEltChunk);
    for (int p = lo + threadIdx.x; p < hi; p += kT) {
        const int iy = p / W, ix = p - iy * W;
        float acc = 0.f;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int ty = iy + pad - ky;
            if (ty < 0 || ty % stride) continue;
            const int oy = ty / stride;
            if (oy >= Ho) continue;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int tx = ix + pad - kx;
                if (tx < 0 || tx % stride) continue;
                const int ox = tx / stride;
                if (ox < Wo) acc += __ldg(dc + ((long long)(ky * 3 + kx) * Ho + oy) * Wo + ox);
            }
        }
        xp[p] = acc;
    }
}

// grid (n * C planes, chunks of kEltChunk pixels)
__global__ void __launch_bounds__(kT)
bias_add_kernel(float* __restrict__ y, const float* __restrict__ b, int C, int HW) {
    const float bv = __ldg(b + blockIdx.x % C);
    float* yp = y + (long long)blockIdx.x * HW;
    const int lo = blockIdx.y * kEltChunk, hi = min(HW, lo + kEltChunk);
#pragma unroll 4
    for (int p = lo + threadIdx.x; p < hi; p += kT) yp[p] += bv;
}

// grid (C, S): partial[s][c] = sum of dy over this CTA's share of channel c
__global__ void __launch_bounds__(kT)
channel_sum_kernel(const float* __restrict__ dy, double* __restrict__ partial, int N, int C, int HW) {
    __shared__ double sm[8];
    const int c = blockIdx.x, S = gridDim.y;
    const ChanSlice sl = chan_slice(HW, S, blockIdx.y, 1);
    double v[1] = {0.0};
    for (int n = 0; n < N; ++n) {
        const float* dp = dy + ((long long)n * C + c) * HW;
        float part = 0.f;
#pragma unroll 4
        for (int p = sl.lo + threadIdx.x; p < sl.hi; p += kT) part += __ldg(dp + p);
        v[0] += (double)part;
    }
    block_sum<1>(v, sm);
    if (threadIdx.x == 0) partial[(long long)blockIdx.y * C + c] = v[0];
}

__global__ void __launch_bounds__(kT)
bilinear_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int Hi, int Wi, int Ho, int Wo, long long total) {
    const float scy = Ho > 1 ? (float)(Hi - 1) / (float)(Ho - 1) : 0.f, scx = Wo > 1 ? (float)(Wi - 1) / (float)(Wo - 1) : 0.f;
    for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < total; i += (long long)gridDim.x * kT) {
        const int ox = (int)(i % Wo), oy = (int)((i / Wo) % Ho);
        const float* xp = x + (i / ((long long)Wo * Ho)) * Hi * Wi;
        int y0, y1, x0, x1;
        float ly, lx;
        ac_coord(oy, scy, Hi, y0, y1, ly);
        ac_coord(ox, scx, Wi, x0, x1, lx);
        const float top = fmaf(lx, __ldg(xp + y0 * Wi + x1), (1.f - lx) * __ldg(xp + y0 * Wi + x0));
        const float bot = fmaf(lx, __ldg(xp + y1 * Wi + x1), (1.f - lx) * __ldg(xp + y1 * Wi + x0));
        y[i] = fmaf(ly, bot, (1.f - ly) * top);
    }
}

// gather form of the backward: every input pixel collects from the output pixels whose 2 x 2 taps include it (found by running
// the forward's own index computation over the candidate range, so the two passes agree exactly); deterministic, no atomics
__global__ void __launch_bounds__(kT)
bilinear_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int Hi, int Wi, int Ho, int Wo) {      // grid (planes, chunks)
    const float scy = Ho > 1 ? (float)(Hi - 1) / (float)(Ho - 1) : 0.f, scx = Wo > 1 ? (float)(Wi - 1) / (float)(Wo - 1) : 0.f;
    const float* dp = dy + (long long)blockIdx.x * Ho * Wo;
    const long long obase = (long long)blockIdx.x * Hi * Wi;
    const int lo = blockIdx.y * kEltChunk, hi = min(Hi * Wi, lo + kEltChunk);
    for (int pp = lo + threadIdx.x; pp < hi; pp += kT) {
        const int iy = pp / Wi, ix = pp - iy * Wi;
        const long long i = obase + pp;
        // output rows that can touch input row iy: floor(scy * oy) in {iy - 1, iy}
        int oy_lo = 0, oy_hi = Ho - 1, ox_lo = 0, ox_hi = Wo - 1;
        if (scy > 0.f) { oy_lo = max(0, (int)floorf((float)(iy - 1) / scy) - 1); oy_hi = min(Ho - 1, (int)ceilf((float)(iy + 1) / scy) + 1); }
        if (scx > 0.f) { ox_lo = max(0, (int)floorf((float)(ix - 1) / scx) - 1); ox_hi = min(Wo - 1, (int)ceilf((float)(ix + 1) / scx) + 1); }
        float acc = 0.f;
        for (int oy = oy_lo; oy <= oy_hi; ++oy) {
            int y0, y1;
            float ly;
            ac_coord(oy, scy, Hi, y0, y1, ly);
            const float wy = (y0 == iy ? 1.f - ly : 0.f) + (y1 == iy ? ly : 0.f);
            if (wy == 0.f && y0 != iy && y1 != iy) continue;
            for (int ox = ox_lo; ox <= ox_hi; ++ox) {
                int x0, x1;
                float lx;
                ac_coord(ox, scx, Wi, x0, x1, lx);
                const float wx = (x0 == ix ? 1.f - lx : 0.f) + (x1 == ix ? lx : 0.f);
                if (x0 == ix || x1 == ix) acc = fmaf(wy * wx, __ldg(dp + (long long)oy * Wo + ox), acc);
            }
        }
        dx[i] = acc;
    }
}

// nn.AdaptiveAvgPool2d(S): bin i covers [floor(i*h/S), ceil((i+1)*h/S)); bins overlap when h % S != 0
__global__ void __launch_bounds__(kT)
adaptive_pool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int H, int W, int S, long long total) {      // a warp per bin
    const int lane = threadIdx.x & 31;
    for (long long i = (long long)blockIdx.x * (kT / 32) + (threadIdx.x >> 5); i < total; i += (long long)gridDim.x * (kT / 32)) {
        const int bx = (int)(i % S), by = (int)((i / S) % S);
        const float* xp = x + (i / (S * S)) * H * W;
        const int y0 = (by * H) / S, y1 = ((by + 1) * H + S - 1) / S, x0 = (bx * W) / S, x1 = ((bx + 1) * W + S - 1) / S;
        const int bw = x1 - x0, cnt = (y1 - y0) * bw;
        float acc = 0.f;
        for (int e = lane; e < cnt; e += 32) {
            const int r = e / bw;
            acc += __ldg(xp + (y0 + r) * W + x0 + (e - r * bw));
        }
        acc = warp_sumf(acc);
        if (lane == 0) y[i] = acc / (float)cnt;
    }
}

__global__ void __launch_bounds__(kT)
adaptive_pool_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int H, int W, int S, long long total) {
    for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < total; i += (long long)gridDim.x * kT) {
        const int xx = (int)(i % W), yy = (int)((i / W) % H);
        const float* dp = dy + (i / ((long long)W * H)) * S * S;
        float acc = 0.f;
        for (int by = 0; by < S; ++by) {
            const int y0 = (by * H) / S, y1 = ((by + 1) * H + S - 1) / S;
            if (yy < y0 || yy >= y1) continue;
            for (int bx = 0; bx < S; ++bx) {
                const int x0 = (bx * W) / S, x1 = ((bx + 1) * W + S - 1) / S;
                if (xx >= x0 && xx < x1) acc += __ldg(dp + by * S + bx) / (float)((y1 - y0) * (x1 - x0));
            }
        }
        dx[i] = acc;
    }
}

// nn.Dropout(p) in train mode: keep with probability 1 - p, scale by 1 / (1 - p).  The keep decision is a counter-based hash
// of (seed, element index), so forward and backward regenerate it instead of storing a mask.
__device__ __forceinline__ bool dropout_keep(unsigned long long seed, long long i, float p) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1);      // splitmix64
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (float)(z >> 40) * (1.f / 16777216.f) >= p;
}
__global__ void __launch_bounds__(kT)
dropout_kernel(const float* __restrict__ x, float* __restrict__ y, float p, unsigned long long seed,
               const unsigned long long* __restrict__ d_step, long long total) {
    const float scale = 1.f / (1.f - p);
    if (d_step) seed += 0xD1342543DE82EF95ull * (*d_step + 1ull);
    for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < total; i += (long long)gridDim.x * kT)
        y[i] = dropout_keep(seed, i, p) ? x[i] * scale : 0.f;
}

// y = relu(a + b); backward: g = dy where y > 0
__global__ void __launch_bounds__(kT)
add_relu_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y, int relu, long long total) {
    for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < total; i += (long long)gridDim.x * kT) {
        const float v = a[i] + b[i];
        y[i] = relu ? fmaxf(v, 0.f) : v;
    }
}
__global__ void __launch_bounds__(kT)
relu_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy, float* __restrict__ dx, long long total) {
    for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < total; i += (long long)gridDim.x * kT) dx[i] = y[i] > 0.f ? dy[i] : 0.f;
}

// torch.optim.SGD(momentum, weight_decay) on flat buffers (train.py:195-198): g += wd * p; buf = first ? g : m * buf + g; p -= lr * buf
__global__ void __launch_bounds__(kT)
sgd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf, float lr, float momentum, float wd, float gscale,
           int first, long long total) {
    for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < total; i += (long long)gridDim.x * kT) {
        const float gi = fmaf(wd, p[i], g[i] * gscale);
        const float b = first ? gi : fmaf(momentum, buf[i], gi);
        buf[i] = b;
        p[i] -= lr * b;
    }
}

// torch.optim.AdamW (decoupled weight decay; train_bdd100k.py:183-185) on flat buffers: p *= 1 - lr wd; m = b1 m + (1 - b1) g;
// v = b2 v + (1 - b2) g^2; p -= (lr / bc1) m / (sqrt(v) / sqrt(bc2) + eps) with the bias corrections bc = 1 - beta^step from the host
__global__ void __launch_bounds__(kT)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, float lr, float omb1, float b2,
             float omb2, float eps, float wd, float gscale, float step_size, float rsqrt_bc2, long long total) {
    for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < total; i += (long long)gridDim.x * kT) {
        const float gi = g[i] * gscale;
        const float pi = p[i] * (1.f - lr * wd);
        const float mi = m[i] + (gi - m[i]) * omb1;                    // torch's lerp form of b1 m + (1 - b1) g
        const float vi = fmaf(b2, v[i], omb2 * gi * gi);               // 1 - beta comes from the host in double, like torch's scalars
        m[i] = mi;
        v[i] = vi;
        p[i] = pi - step_size * (mi / (sqrtf(vi) * rsqrt_bc2 + eps));
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------------------------------
static int grid_for(long long total) {
    long long b = (total + kT - 1) / kT;
    const long long cap = (long long)num_sms() * 16;
    return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}
static int splits_for(int channels, long long per_channel) {
    long long want = (long long)num_sms() * 4 / (channels > 0 ? channels : 1);
    long long maxs = (per_channel + kT * 4 - 1) / (kT * 4);
    if (want > maxs) want = maxs;
    if (want < 1) want = 1;
    if (want > 64) want = 64;
    return (int)want;
}

// split-K partials of the pointwise weight gradient: as many as fill the GPU a few times over, bounded by 32 MB of scratch
static int pw_parts_max(int cout, int cin) {
    long long p = (8ll << 20) / ((long long)cout * cin);
    return (int)(p < 64 ? 64 : (p > 1024 ? 1024 : p));
}

size_t train_workspace_bytes(int channels_max, int cout, int cin) {
    // reductions: 64 splits x channels x 9 doubles; pointwise weight gradient: pw_parts_max split-K partials of [cout][cin] floats
    return (size_t)64 * channels_max * 9 * sizeof(double) + (size_t)pw_parts_max(cout, cin) * cout * cin * sizeof(float) + 4096;
}

static DwTasks dw_tasks(long long planes, int rows, int cols) {
    DwTasks t{(rows + kDwRows - 1) / kDwRows, (cols + 31) / 32, 0};
    t.total = planes * t.RB * t.CB;
    return t;
}
static int dw_grid(const DwTasks& t) {
    const long long ctas = (t.total + kT / 32 - 1) / (kT / 32), cap = (long long)num_sms() * 16;
    return (int)(ctas < cap ? (ctas > 0 ? ctas : 1) : cap);
}
static bool vec2_ok(int wd, std::initializer_list<const void*> ptrs) {
    if (wd & 1) return false;
    for (const void* p : ptrs)
        if (reinterpret_cast<uintptr_t>(p) & 7) return false;
    return true;
}

static bool dw_v4_ok(int wd, std::initializer_list<const void*> ptrs) {
    if ((wd & 3) || wd > 128 || wd < 4) return false;
    for (const void* p : ptrs)
        if (reinterpret_cast<uintptr_t>(p) & 15) return false;
    return true;
}
static int dw_v4_grid(long long tasks, int wd) {
    const int G = 32 / (wd / 4);
    const long long wtasks = (tasks + G - 1) / G, ctas = (wtasks + kT / 32 - 1) / (kT / 32), cap = (long long)num_sms() * 16;
    return (int)(ctas < cap ? (ctas > 0 ? ctas : 1) : cap);
}

cudaError_t launch_train_dw_fwd(const float* x, const float* w, float* y, int n, int c, int h, int wd, int stride, cudaStream_t s) {
    const int ho = (h - 1) / stride + 1, wo = (wd - 1) / stride + 1;
    if (stride == 1 && dw_v4_ok(wd, {x, y})) {
        const int rb = (h + kDwRows - 1) / kDwRows;
        const long long tasks = (long long)n * c * rb;
        dw_s1v4_kernel<true, false><<<dw_v4_grid(tasks, wd), kT, 0, s>>>(x, nullptr, w, y, nullptr, c, h, wd, rb, tasks, 0);
        return cudaGetLastError();
    }
    const DwTasks tk = dw_tasks((long long)n * c, ho, wo);
    if (stride == 2 && vec2_ok(wd, {x})) dw_fwd_kernel<2, true><<<dw_grid(tk), kT, 0, s>>>(x, w, y, c, h, wd, ho, wo, tk, 0);
    else if (stride == 2) dw_fwd_kernel<2, false><<<dw_grid(tk), kT, 0, s>>>(x, w, y, c, h, wd, ho, wo, tk, 0);
    else dw_fwd_kernel<1, false><<<dw_grid(tk), kT, 0, s>>>(x, w, y, c, h, wd, ho, wo, tk, 0);
    return cudaGetLastError();
}

// ws: c * 9 doubles of weight-gradient accumulators
cudaError_t launch_train_dw_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw, void* ws, int n, int c, int h,
                                int wd, int stride, cudaStream_t s) {
    const int ho = (h - 1) / stride + 1, wo = (wd - 1) / stride + 1;
    double* dwacc = reinterpret_cast<double*>(ws);
    if (dw) {
        cudaError_t e = cudaMemsetAsync(dwacc, 0, (size_t)c * 9 * sizeof(double), s);
        if (e != cudaSuccess) return e;
    }
    const DwTasks tk = dw_tasks((long long)n * c, ho, wo);
    const int grid = dw_grid(tk);
    if (stride == 1 && dw_v4_ok(wd, {x, dy, dx ? dx : dy})) {
        const int rb = (h + kDwRows - 1) / kDwRows;
        const long long tasks = (long long)n * c * rb;
        const int g4 = dw_v4_grid(tasks, wd);
        if (dx && dw) dw_s1v4_kernel<true, true><<<g4, kT, 0, s>>>(dy, x, w, dx, dwacc, c, h, wd, rb, tasks, 1);
        else if (dx) dw_s1v4_kernel<true, false><<<g4, kT, 0, s>>>(dy, nullptr, w, dx, dwacc, c, h, wd, rb, tasks, 1);
        else if (dw) dw_s1v4_kernel<false, true><<<g4, kT, 0, s>>>(dy, x, w, nullptr, dwacc, c, h, wd, rb, tasks, 1);
    } else if (stride == 1) {
        if (dx && dw) dw_bwd1_kernel<true, true><<<grid, kT, 0, s>>>(x, w, dy, dx, dwacc, c, h, wd, tk);
        else if (dx) dw_bwd1_kernel<true, false><<<grid, kT, 0, s>>>(x, w, dy, dx, dwacc, c, h, wd, tk);
        else if (dw) dw_bwd1_kernel<false, true><<<grid, kT, 0, s>>>(x, w, dy, dx, dwacc, c, h, wd, tk);
    } else {
        if (dx && vec2_ok(wd, {dx})) dw_bwd2_data_kernel<true><<<grid, kT, 0, s>>>(dy, w, dx, c, h, wd, ho, wo, tk);
        else if (dx) dw_bwd2_data_kernel<false><<<grid, kT, 0, s>>>(dy, w, dx, c, h, wd, ho, wo, tk);
        if (dw && vec2_ok(wd, {x})) dw_bwd2_weight_kernel<true><<<grid, kT, 0, s>>>(x, dy, dwacc, c, h, wd, ho, wo, tk);
        else if (dw) dw_bwd2_weight_kernel<false><<<grid, kT, 0, s>>>(x, dy, dwacc, c, h, wd, ho, wo, tk);
    }
    if (dw) double_to_float_kernel<<<(c * 9 + kT - 1) / kT, kT, 0, s>>>(dwacc, dw, c * 9);
    return cudaGetLastError();
}

// 0: fp32 FMA kernels (default), 1: TF32 operands on the tensor cores
static int g_train_math = 0;
int train_set_math(int mode) {
    if (mode < 0 || mode > 1) return -1;
    g_train_math = mode;
    return 0;
}
int train_get_math() { return g_train_math; }

static void launch_gemm_pix(const GemmArgs& g, dim3 grid, cudaStream_t s) {
    if (g_train_math == 0) gemm_pix_kernel<<<grid, kT, 0, s>>>(g);
    else gemm_pix_mma_kernel<1><<<grid, kT, 0, s>>>(g);
}

static bool tc_enabled() {
    static const bool off = getenv("FSCNN_TRAIN_NO_TC") != nullptr;      // A/B switch: keep the mma.sync kernels in TF32 mode
    return g_train_math == 1 && !off;
}

cudaError_t launch_train_pw_fwd(const float* x, const float* w, float* y, int n, int cin, int cout, int hw, cudaStream_t s) {
    if (tc_enabled()) {
        const cudaError_t e = launch_gemm_pix_tc(w, x, y, cout, hw, cin, n, cin, true, s);
        if (e != cudaErrorNotSupported) return e;
    }
    GemmArgs g{w, x, y, cout, hw, cin, 0, cin, 1, (long long)cin * hw, hw, 1, (long long)cout * hw, hw, 1};
    launch_gemm_pix(g, dim3((hw + 127) / 128, (cout + 63) / 64, n), s);
    return cudaGetLastError();
}

cudaError_t launch_train_pw_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw, void* ws, int n, int cin,
                                int cout, int hw, cudaStream_t s) {
    if (dx) {      // dx[n] = W^T dy[n]
        cudaError_t e = tc_enabled() ? launch_gemm_pix_tc(w, dy, dx, cin, hw, cout, n, cin, false, s) : cudaErrorNotSupported;
        if (e == cudaErrorNotSupported) {
            GemmArgs g{w, dy, dx, cin, hw, cout, 0, 1, cin, (long long)cout * hw, hw, 1, (long long)cin * hw, hw, 1};
            launch_gemm_pix(g, dim3((hw + 127) / 128, (cin + 63) / 64, n), s);
        } else if (e != cudaSuccess) {
            return e;
        }
    }
    if (dw) {      // dW = sum_n dy[n] x[n]^T, split over the pixels of every image: enough CTAs to fill the GPU ~4 times
        const int tiles = ((cin + 63) / 64) * ((cout + 63) / 64), pmax = pw_parts_max(cout, cin);
        if (n > pmax) return cudaErrorInvalidValue;      // the caller chunks the batch
        int splitk = (4 * num_sms() + n * tiles - 1) / (n * tiles);
        const int maxk = (hw + 255) / 256;
        if (splitk > maxk) splitk = maxk;
        if (splitk > pmax / n) splitk = pmax / n;
        if (splitk < 1) splitk = 1;
        const int parts = n * splitk;
        float* partial = reinterpret_cast<float*>(ws);
        GemmArgs g{dy, x, partial, cout, cin, hw, (long long)cout * hw, hw, 1, (long long)cin * hw, 1, hw, (long long)cout * cin, cin, splitk};
        const dim3 grid((cin + 63) / 64, (cout + 63) / 64, parts);
        if (g_train_math == 0) gemm_wgrad_kernel<<<grid, kT, 0, s>>>(g);
        else gemm_wgrad_mma_kernel<1><<<grid, kT, 0, s>>>(g);
        reduce_partials_f_kernel<<<(cout * cin + kT - 1) / kT, kT, 0, s>>>(partial, dw, cout * cin, parts);
    }
    return cudaGetLastError();
}

static bool vec4_ok(int hw, std::initializer_list<const void*> ptrs) {
    if (hw & 3) return false;
    for (const void* p : ptrs)
        if (reinterpret_cast<uintptr_t>(p) & 15) return false;
    return true;
}

cudaError_t launch_train_bn_fwd(const float* x, const float* gamma, const float* beta, float* running_mean, float* running_var,
                                float* y, float* save_mean, float* save_rstd, void* ws, int n, int c, int hw, float eps,
                                float momentum, int relu, cudaStream_t s) {
    const int S = splits_for(c, (long long)n * hw);
    double* partial = reinterpret_cast<double*>(ws);
    const bool vec = vec4_ok(hw, {x, y});
    if (vec) bn_stats_kernel<true><<<dim3(c, S), kT, 0, s>>>(x, partial, n, c, hw);
    else bn_stats_kernel<false><<<dim3(c, S), kT, 0, s>>>(x, partial, n, c, hw);
    bn_finalize_kernel<<<(c + kT - 1) / kT, kT, 0, s>>>(partial, c, S, (long long)n * hw, eps, momentum, save_mean, save_rstd,
                                                         running_mean, running_var);
    const dim3 grid(n * c, (hw + kEltChunk - 1) / kEltChunk);
    if (vec) bn_apply_kernel<true><<<grid, kT, 0, s>>>(x, save_mean, save_rstd, gamma, beta, y, c, hw, relu);
    else bn_apply_kernel<false><<<grid, kT, 0, s>>>(x, save_mean, save_rstd, gamma, beta, y, c, hw, relu);
    return cudaGetLastError();
}

// `beta` is needed for the ReLU mask (recomputed from x with the forward's own arithmetic: y is never read)
cudaError_t launch_train_bn_bwd(const float* x, const float* dy, const float* gamma, const float* beta, const float* save_mean,
                                const float* save_rstd, float* dx, float* dgamma, float* dbeta, void* ws, int n, int c, int hw,
                                int relu, cudaStream_t s) {
    const int S = splits_for(c, (long long)n * hw);
    double* partial = reinterpret_cast<double*>(ws);
    const bool vec = vec4_ok(hw, {x, dy, dx});
    if (vec) bn_bwd_reduce_kernel<true><<<dim3(c, S), kT, 0, s>>>(x, dy, save_mean, save_rstd, gamma, beta, partial, n, c, hw, relu);
    else bn_bwd_reduce_kernel<false><<<dim3(c, S), kT, 0, s>>>(x, dy, save_mean, save_rstd, gamma, beta, partial, n, c, hw, relu);
    bn_bwd_finalize_kernel<<<(c + kT - 1) / kT, kT, 0, s>>>(partial, c, S, dgamma, dbeta);
    const dim3 grid(n * c, (hw + kEltChunk - 1) / kEltChunk);
    const float inv = 1.f / (float)((long long)n * hw);
    if (vec) bn_bwd_apply_kernel<true><<<grid, kT, 0, s>>>(x, dy, save_mean, save_rstd, gamma, beta, dgamma, dbeta, dx, c, hw, relu, inv);
    else bn_bwd_apply_kernel<false><<<grid, kT, 0, s>>>(x, dy, save_mean, save_rstd, gamma, beta, dgamma, dbeta, dx, c, hw, relu, inv);
    return cudaGetLastError();
}

// ws layout: [0,64) state (8 x u64) | [64, 64+1024) radix histogram | then 2 doubles per CTA of the loss kernel
cudaError_t launch_train_ohem_fwd(const float* logits, const long long* label, const float* weight, float* prob, float* out3, void* ws,
                                  int n, int c, int hw, long long ignore, float thresh, int min_kept, cudaStream_t s) {
    const long long npix = (long long)n * hw;
    unsigned long long* state = reinterpret_cast<unsigned long long*>(ws);
    unsigned int* hist = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(ws) + 64);
    double* partial = reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + 64 + 1024);
    cudaError_t e = cudaMemsetAsync(ws, 0, 64 + 1024, s);
    if (e != cudaSuccess) return e;
    const int grid = grid_for(npix);
    ohem_prob_kernel<<<grid, kT, 0, s>>>(logits, label, prob, c, hw, npix, ignore, state);
    for (int shift = 24; shift >= 0; shift -= 8) {
        ohem_hist_kernel<<<grid, kT, 0, s>>>(prob, npix, shift, state, hist);
        ohem_select_kernel<<<1, 256, 0, s>>>(state, hist, shift, min_kept, thresh);
    }
    ohem_loss_kernel<<<grid, kT, 0, s>>>(logits, label, prob, weight, c, hw, npix, ignore, state, partial);
    ohem_finalize_kernel<<<1, kT, 0, s>>>(partial, grid, state, out3);
    return cudaGetLastError();
}

size_t train_ohem_workspace_bytes() { return 64 + 1024 + (size_t)num_sms() * 16 * 2 * sizeof(double) + 256; }

cudaError_t launch_train_ohem_bwd(const float* logits, const long long* label, const float* weight, const float* prob,
                                  const float* out3, const float* gout, float* dlogits, const void* ws, int n, int c, int hw,
                                  long long ignore, cudaStream_t s) {
    const long long npix = (long long)n * hw;
    ohem_grad_kernel<<<grid_for(npix), kT, 0, s>>>(logits, label, prob, weight, c, hw, npix, ignore,
                                                   reinterpret_cast<const unsigned long long*>(ws), out3, gout, dlogits);
    return cudaGetLastError();
}


// dense 3x3 convolution through im2col: cols [n][cin*9][ho*wo] (caller scratch), then the pointwise GEMMs with cin -> cin*9
cudaError_t launch_train_im2col(const float* x, float* cols, int n, int c, int h, int wd, int stride, int pad, cudaStream_t s) {
    const int ho = (h + 2 * pad - 3) / stride + 1, wo = (wd + 2 * pad - 3) / stride + 1;
    im2col3x3_kernel<<<dim3(n * c * 9, (ho * wo + kEltChunk - 1) / kEltChunk), kT, 0, s>>>(x, cols, c, h, wd, ho, wo, stride, pad);
    return cudaGetLastError();
}
cudaError_t launch_train_col2im(const float* dcols, float* dx, int n, int c, int h, int wd, int stride, int pad, cudaStream_t s) {
    const int ho = (h + 2 * pad - 3) / stride + 1, wo = (wd + 2 * pad - 3) / stride + 1;
    col2im3x3_kernel<<<dim3(n * c, (h * wd + kEltChunk - 1) / kEltChunk), kT, 0, s>>>(dcols, dx, c, h, wd, ho, wo, stride, pad);
    return cudaGetLastError();
}
cudaError_t launch_train_bias_add(float* y, const float* b, int n, int c, int hw, cudaStream_t s) {
    bias_add_kernel<<<dim3(n * c, (hw + kEltChunk - 1) / kEltChunk), kT, 0, s>>>(y, b, c, hw);
    return cudaGetLastError();
}
cudaError_t launch_train_bias_grad(const float* dy, float* db, void* ws, int n, int c, int hw, cudaStream_t s) {
    const int S = splits_for(c, (long long)n * hw);
    channel_sum_kernel<<<dim3(c, S), kT, 0, s>>>(dy, reinterpret_cast<double*>(ws), n, c, hw);
    reduce_partials_kernel<<<(c + kT - 1) / kT, kT, 0, s>>>(reinterpret_cast<const double*>(ws), db, c, S);
    return cudaGetLastError();
}
cudaError_t launch_train_bilinear(const float* in, float* out, int planes, int hi, int wi, int ho, int wo, int backward, cudaStream_t s) {
    if (!backward) {
        const long long total = (long long)planes * ho * wo;
        bilinear_fwd_kernel<<<grid_for(total), kT, 0, s>>>(in, out, hi, wi, ho, wo, total);
    } else {      // in = dy [planes][ho][wo], out = dx [planes][hi][wi]
        bilinear_bwd_kernel<<<dim3(planes, (hi * wi + kEltChunk - 1) / kEltChunk), kT, 0, s>>>(in, out, hi, wi, ho, wo);
    }
    return cudaGetLastError();
}
cudaError_t launch_train_adaptive_pool(const float* in, float* out, int planes, int h, int wd, int bins, int backward, cudaStream_t s) {
    if (!backward) {
        const long long total = (long long)planes * bins * bins;
        adaptive_pool_fwd_kernel<<<grid_for(total * 32), kT, 0, s>>>(in, out, h, wd, bins, total);
    } else {      // in = dy [planes][S][S], out = dx [planes][h][w]
        const long long total = (long long)planes * h * wd;
        adaptive_pool_bwd_kernel<<<grid_for(total), kT, 0, s>>>(in, out, h, wd, bins, total);
    }
    return cudaGetLastError();
}
cudaError_t launch_train_dropout(const float* x, float* y, float p, unsigned long long seed, const unsigned long long* d_step, long long total,
                                 cudaStream_t s) {
    dropout_kernel<<<grid_for(total), kT, 0, s>>>(x, y, p, seed, d_step, total);
    return cudaGetLastError();
}
cudaError_t launch_train_add_relu(const float* a, const float* b, float* y, int relu, long long total, cudaStream_t s) {
    add_relu_kernel<<<grid_for(total), kT, 0, s>>>(a, b, y, relu, total);
    return cudaGetLastError();
}
cudaError_t launch_train_relu_bwd(const float* y, const float* dy, float* dx, long long total, cudaStream_t s) {
    relu_bwd_kernel<<<grid_for(total), kT, 0, s>>>(y, dy, dx, total);
    return cudaGetLastError();
}
cudaError_t launch_train_sgd(float* p, const float* g, float* buf, float lr, float momentum, float wd, float gscale, int first,
                             long long total, cudaStream_t s) {
    sgd_kernel<<<grid_for(total), kT, 0, s>>>(p, g, buf, lr, momentum, wd, gscale, first, total);
    return cudaGetLastError();
}

cudaError_t launch_train_adamw(float* p, const float* g, float* m, float* v, float lr, double b1, double b2, float eps, float wd, float gscale,
                               long long step, long long total, cudaStream_t s) {
    const double bc1 = 1.0 - pow(b1, (double)step), bc2 = 1.0 - pow(b2, (double)step);
    adamw_kernel<<<grid_for(total), kT, 0, s>>>(p, g, m, v, lr, (float)(1.0 - b1), (float)b2, (float)(1.0 - b2), eps, wd, gscale,
                                                (float)((double)lr / bc1), (float)(1.0 / sqrt(bc2)), total);
    return cudaGetLastError();
}


// CTAs of 8 warps over the (image, row block, column strip) strips of the strip kernels, capped like grid_for (the loss kernel writes
// one partial pair per CTA into the num_sms * 16 slots of train_ohem_workspace_bytes)
static int strip_grid(int n, int h, int w) {
    const long long strips = (long long)n * ((h + kStripRows - 1) / kStripRows) * ((w + 31) / 32);
    const long long ctas = (strips + kT / 32 - 1) / (kT / 32), cap = (long long)num_sms() * 16;
    return (int)(ctas < cap ? (ctas > 0 ? ctas : 1) : cap);
}

static StemTasks stem_tasks(int n, int ho, int wo) {
    StemTasks t{(ho + kStemRows - 1) / kStemRows, (wo + 31) / 32, 0};
    t.total = (long long)n * t.RB * t.CB;
    return t;
}
// y [n][32][ho][wo] = conv3x3 stride 2 pad 0 of x [n][3][h][w] with w [32][3][3][3]
cudaError_t launch_train_stem_fwd(const float* x, const float* w, float* y, int n, int h, int wd, cudaStream_t s) {
    const int ho = (h - 3) / 2 + 1, wo = (wd - 3) / 2 + 1;
    const StemTasks tk = stem_tasks(n, ho, wo);
    const long long ctas = (tk.total + kT / 32 - 1) / (kT / 32), cap = (long long)num_sms() * 16;
    const int grid = (int)(ctas < cap ? (ctas > 0 ? ctas : 1) : cap);
    if (vec2_ok(wd, {x})) stem_fwd_kernel<true><<<grid, kT, 0, s>>>(x, w, y, h, wd, ho, wo, tk);
    else stem_fwd_kernel<false><<<grid, kT, 0, s>>>(x, w, y, h, wd, ho, wo, tk);
    return cudaGetLastError();
}
// dw [32][3][3][3]; ws: 864 doubles
cudaError_t launch_train_stem_wgrad(const float* x, const float* dy, float* dw, void* ws, int n, int h, int wd, cudaStream_t s) {
    const int ho = (h - 3) / 2 + 1, wo = (wd - 3) / 2 + 1;
    double* dwacc = reinterpret_cast<double*>(ws);
    cudaError_t e = cudaMemsetAsync(dwacc, 0, (size_t)kStemCo * kStemK * sizeof(double), s);
    if (e != cudaSuccess) return e;
    const StemTasks tk = stem_tasks(n, ho, wo);
    const long long want = (tk.total + kStemWT / 32 - 1) / (kStemWT / 32), cap = (long long)num_sms() * 3 / 4;
    const dim3 grid((unsigned)(want < cap ? (want > 0 ? want : 1) : cap), 8);
    if (vec2_ok(wd, {x})) stem_wgrad_kernel<true><<<grid, kStemWT, 0, s>>>(x, dy, dwacc, h, wd, ho, wo, tk);
    else stem_wgrad_kernel<false><<<grid, kStemWT, 0, s>>>(x, dy, dwacc, h, wd, ho, wo, tk);
    double_to_float_kernel<<<(kStemCo * kStemK + kT - 1) / kT, kT, 0, s>>>(dwacc, dw, kStemCo * kStemK);
    return cudaGetLastError();
}

static UpGeom up_geom(int c, int hl, int wl, int h, int w) {
    UpGeom g{c, hl, wl, h, w, h > 1 ? (float)(hl - 1) / (float)(h - 1) : 0.f, w > 1 ? (float)(wl - 1) / (float)(w - 1) : 0.f};
    return g;
}

cudaError_t launch_train_ohem_up_fwd(const float* low, const long long* label, const float* weight, float* prob, float* nll, float* out3,
                                     void* ws, int n, int c, int hl, int wl, int h, int w, long long ignore, float thresh, int min_kept,
                                     cudaStream_t s) {
    const long long npix = (long long)n * h * w;
    unsigned long long* state = reinterpret_cast<unsigned long long*>(ws);
    unsigned int* hist = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(ws) + 64);
    double* partial = reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + 64 + 1024);
    cudaError_t e = cudaMemsetAsync(ws, 0, 64 + 1024, s);
    if (e != cudaSuccess) return e;
    const UpGeom g = up_geom(c, hl, wl, h, w);
    const int grid = grid_for(npix), sgrid = strip_grid(n, h, w);
    const bool strip = c == 19 || c == 2;
    if (c == 19) ohem_strip_prob_kernel<19><<<sgrid, kT, 0, s>>>(low, label, prob, nll, g, n, ignore, state);
    else if (c == 2) ohem_strip_prob_kernel<2><<<sgrid, kT, 0, s>>>(low, label, prob, nll, g, n, ignore, state);
    else ohem_up_prob_kernel<0><<<grid, kT, 0, s>>>(low, label, prob, g, npix, ignore, state);
    for (int shift = 24; shift >= 0; shift -= 8) {
        ohem_hist_kernel<<<grid, kT, 0, s>>>(prob, npix, shift, state, hist);
        ohem_select_kernel<<<1, 256, 0, s>>>(state, hist, shift, min_kept, thresh);
    }
    if (strip && nll) {
        ohem_nll_loss_kernel<<<grid, kT, 0, s>>>(label, prob, nll, weight, npix, ignore, state, partial);
        ohem_finalize_kernel<<<1, kT, 0, s>>>(partial, grid, state, out3);
        return cudaGetLastError();
    }
    if (c == 19) ohem_strip_loss_kernel<19><<<sgrid, kT, 0, s>>>(low, label, prob, weight, g, n, ignore, state, partial);
    else if (c == 2) ohem_strip_loss_kernel<2><<<sgrid, kT, 0, s>>>(low, label, prob, weight, g, n, ignore, state, partial);
    else ohem_up_loss_kernel<0><<<grid, kT, 0, s>>>(low, label, prob, weight, g, npix, ignore, state, partial);
    ohem_finalize_kernel<<<1, kT, 0, s>>>(partial, strip ? sgrid : grid, state, out3);
    return cudaGetLastError();
}

cudaError_t launch_train_ohem_up_bwd(const float* low, const long long* label, const float* weight, const float* prob, const float* out3,
                                     const float* gout, float* dlow, const void* ws, int n, int c, int hl, int wl, int h, int w,
                                     long long ignore, cudaStream_t s) {
    const bool strip = c == 19 || c == 2;
    const size_t smem = (size_t)c * kUpTR * kUpTC * sizeof(float);
    if (!strip && ((double)(hl - 1) * 7.0 > (double)(h - 1) || (double)(wl - 1) * 7.0 > (double)(w - 1) || smem > 48 * 1024))
        return cudaErrorInvalidValue;
    cudaError_t e = cudaMemsetAsync(dlow, 0, (size_t)n * c * hl * wl * sizeof(float), s);
    if (e != cudaSuccess) return e;
    const dim3 grid((w + 31) / 32, (h + 63) / 64, n);
    const int sgrid = strip_grid(n, h, w);
    const UpGeom g = up_geom(c, hl, wl, h, w);
    const unsigned long long* state = reinterpret_cast<const unsigned long long*>(ws);
    if (c == 19) ohem_strip_grad_kernel<19><<<sgrid, kT, 0, s>>>(low, label, prob, weight, g, n, ignore, state, out3, gout, dlow);
    else if (c == 2) ohem_strip_grad_kernel<2><<<sgrid, kT, 0, s>>>(low, label, prob, weight, g, n, ignore, state, out3, gout, dlow);
    else ohem_up_grad_kernel<0><<<grid, kT, smem, s>>>(low, label, prob, weight, g, ignore, state, out3, gout, dlow);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------------
// The reference's other training criteria (utils/loss.py), per head, with or without the head's final resize fused in:
//   kind 0  nn.CrossEntropyLoss(ignore_index) as MixSoftmaxCrossEntropyLoss applies it to every head (loss.py:103-124): mean of the
//           negative log-likelihood over the pixels whose label is not ignore_label
//   kind 1  DiceLoss (loss.py:12-39, train.py's DEFAULT --loss-type): p = softmax(x)[:, 1] (sigmoid for one channel), t = float(label)
//           of EVERY pixel (no ignore label), 1 - (2 sum(p t) + smooth) / (sum p + sum t + smooth)
//   kind 2  FocalDiceLoss (loss.py:71-100): (1 - dice_weight) * mean(alpha (1 - pt)^gamma ce) + dice_weight * DiceLoss, ce =
//           F.cross_entropy(reduction='none') (labels equal to -100 contribute 0 but count in the mean) or, for one channel,
//           F.binary_cross_entropy(sigmoid(x), t) with its log clamp at -100
// (hl, wl) == (h, w): `logits` are at label resolution and the backward writes d loss / d logits directly (deterministic).
// Otherwise `logits` are a head's low-resolution output and F.interpolate(size=(h, w), 'bilinear', align_corners=True)
// (fast_scnn.py:40, :44) is composed with the criterion exactly as for the OHEM loss above: same interpolation arithmetic
// (up_value), the full-resolution logits and their gradient never exist, the backward scatters through the resize's transpose
// (shared-memory tile + float atomics).  Sums are fp64 per-CTA partials reduced by one CTA in a fixed order.
// Labels outside [0, C) that are not the ignore label (-100 for kind 2) make torch raise; here they are skipped like ignored ones.
// ---------------------------------------------------------------------------------------------------------------------
enum { kCritCE = 0, kCritDice = 1, kCritFocalDice = 2 };
struct CritArgs { int kind; long long ignore; float smooth, alpha, gamma, dice_w; };

template <int CT, bool UP>
struct CritPix {
    float v[CT > 0 ? CT : 1];
    const float* lp;        // UP: the image's low-resolution planes; else: this pixel in plane 0
    long long cs;           // class stride in elements
    int wl, y0, y1, x0, x1;
    float ly, lx;
    __device__ __forceinline__ float fetch(int c) const {
        return UP ? up_value(lp + (long long)c * cs, wl, y0, y1, x0, x1, ly, lx) : __ldg(lp + (long long)c * cs);
    }
    // CT > 0: a compile-time BOUND on the class count (1, 2, 4, 8, 16, 19, 32): the pixel's values are fetched once into registers
    // and every class loop is unrolled CT times with the tail (c >= C) predicated off; CT == 0: any count, values re-fetched per pass
    __device__ __forceinline__ void load(int C) {
        if (CT > 0) {
#pragma unroll
            for (int c = 0; c < (CT > 0 ? CT : 1); ++c) v[c] = c < C ? fetch(c) : 0.f;
        }
    }
    __device__ __forceinline__ float get(int c) const { return CT > 0 ? v[c] : fetch(c); }
};
#define CRIT_FOR_CLASSES(c) _Pragma("unroll") for (int c = 0; c < (CT > 0 ? CT : C); ++c) if (CT == 0 || c < C)

__device__ __forceinline__ float crit_sigmoid(float z) { return 1.f / (1.f + expf(-z)); }
// x^g for the focal term: the reference's default gamma = 2 (and the 1 its derivative needs) without the general powf (uniform branch)
__device__ __forceinline__ float crit_pow(float x, float g) { return g == 2.f ? x * x : (g == 1.f ? x : (g == 3.f ? x * x * x : powf(x, g))); }

// Row cache of the fused-resize kernels whose threads own a COLUMN and walk rows (CT > 0): per class the two horizontally
// interpolated low-resolution rows the current output row sits between -- the inner two fmaf of up_value, refreshed only when the
// row pair changes (every ~8 rows at ratio 1/8) -- so a pixel costs one vertical fmaf per class and the value is bit-identical to
// up_value's.
template <int CT>
struct CritRows {
    float top[CT > 0 ? CT : 1], bot[CT > 0 ? CT : 1];
    int y0 = -1, y1 = -1;
    __device__ __forceinline__ void values(CritPix<CT, true>& u, int C) {
        if (CT == 0) return;
        if (u.y0 != y0 || u.y1 != y1) {          // warp-uniform: a warp's lanes share the output row
            y0 = u.y0; y1 = u.y1;
#pragma unroll
            for (int c = 0; c < (CT > 0 ? CT : 1); ++c) {
                if (c < C) {
                    const float* base = u.lp + (long long)c * u.cs;
                    top[c] = fmaf(u.lx, __ldg(base + y0 * u.wl + u.x1), (1.f - u.lx) * __ldg(base + y0 * u.wl + u.x0));
                    bot[c] = fmaf(u.lx, __ldg(base + y1 * u.wl + u.x1), (1.f - u.lx) * __ldg(base + y1 * u.wl + u.x0));
                }
            }
        }
#pragma unroll
        for (int c = 0; c < (CT > 0 ? CT : 1); ++c) u.v[c] = c < C ? fmaf(u.ly, bot[c], (1.f - u.ly) * top[c]) : 0.f;
    }
};

// acc: kind 0 {sum nll, count, -, -}; kinds 1, 2 {sum p t, sum p, sum t, sum focal}
template <int CT, bool UP>
__device__ __forceinline__ void crit_pixel_fwd(const CritPix<CT, UP>& u, int C, long long lab, const CritArgs& a, double (&acc)[4]) {
    if (C == 1 && a.kind != kCritCE) {
        const float p = crit_sigmoid(u.get(0)), t = (float)lab;
        acc[0] += (double)(p * t); acc[1] += (double)p; acc[2] += (double)t;
        if (a.kind == kCritFocalDice) {
            const float ce = -(t * fmaxf(logf(p), -100.f) + (1.f - t) * fmaxf(logf(1.f - p), -100.f));
            const float pt = t == 1.f ? p : 1.f - p;
            acc[3] += (double)(a.alpha * crit_pow(1.f - pt, a.gamma) * ce);
        }
        return;
    }
    float mx = -FLT_MAX;
    CRIT_FOR_CLASSES(c) mx = fmaxf(mx, u.get(c));
    float sum = 0.f, vl = 0.f, e1 = 0.f;
    CRIT_FOR_CLASSES(c) {
        const float t = u.get(c), e = expf(t - mx);
        sum += e;
        if (c == lab) vl = t;
        if (c == 1) e1 = e;
    }
    const bool valid = lab >= 0 && lab < C && lab != a.ignore;
    if (a.kind == kCritCE) {
        if (valid) { acc[0] += (double)(-((vl - mx) - logf(sum))); acc[1] += 1.0; }
        return;
    }
    const float p1 = e1 / sum, t = (float)lab;
    acc[0] += (double)(p1 * t); acc[1] += (double)p1; acc[2] += (double)t;
    if (a.kind == kCritFocalDice && valid) {
        const float ce = -((vl - mx) - logf(sum)), pt = expf(-ce);
        acc[3] += (double)(a.alpha * crit_pow(1.f - pt, a.gamma) * ce);
    }
}

// scalars of the backward, derived once per thread from the forward's sums: d loss / d p = da * t + db for the dice term
struct CritScal { float sce, da, db, sf; };
__device__ __forceinline__ CritScal crit_scalars(const double* __restrict__ out, float gout, const CritArgs& a) {
    CritScal s{0.f, 0.f, 0.f, 0.f};
    if (a.kind == kCritCE) {
        s.sce = (float)((double)gout / out[2]);
    } else {
        const double I = out[1], D = out[2] + out[3] + (double)a.smooth;
        const double k = (double)gout * (a.kind == kCritFocalDice ? (double)a.dice_w : 1.0);
        s.da = (float)(-2.0 / D * k);
        s.db = (float)((2.0 * I + (double)a.smooth) / (D * D) * k);
        if (a.kind == kCritFocalDice) s.sf = (float)((double)gout * (1.0 - (double)a.dice_w) / out[5]);
    }
    return s;
}

// one channel (sigmoid): d loss / d logit of this pixel
__device__ __forceinline__ float crit_grad_binary(float z, long long lab, const CritArgs& a, const CritScal& s) {
    const float p = crit_sigmoid(z), t = (float)lab;
    float gp = s.da * t + s.db;
    if (a.kind == kCritFocalDice) {
        const float ce = -(t * fmaxf(logf(p), -100.f) + (1.f - t) * fmaxf(logf(1.f - p), -100.f));
        const float dce = (p - t) / fmaxf((1.f - p) * p, 1e-12f);            // binary_cross_entropy's own backward
        const float pt = t == 1.f ? p : 1.f - p, dpt = t == 1.f ? 1.f : -1.f, om = 1.f - pt;
        gp += s.sf * a.alpha * (crit_pow(om, a.gamma) * dce - a.gamma * crit_pow(om, a.gamma - 1.f) * dpt * ce);
    }
    return gp * p * (1.f - p);
}

// softmax statistics of a pixel for the backward: g_c = wce * (p_c - [c == lab]) + wd * ([c == 1] - p_c), p_c = exp(v_c - mx) * inv
// (CT > 0: the exponentials stay in ev[] for the caller's gradient loop)
template <int CT, bool UP>
__device__ __forceinline__ void crit_pixel_bwd_coef(const CritPix<CT, UP>& u, int C, long long lab, const CritArgs& a, const CritScal& s,
                                                    float& mx, float& inv, float& wce, float& wd, float (&ev)[CT > 0 ? CT : 1]) {
    mx = -FLT_MAX;
    CRIT_FOR_CLASSES(c) mx = fmaxf(mx, u.get(c));
    float sum = 0.f, el = 0.f, e1 = 0.f;
    CRIT_FOR_CLASSES(c) {
        const float e = expf(u.get(c) - mx);
        if (CT > 0) ev[CT > 0 ? c : 0] = e;
        sum += e;
        if (c == lab) el = e;
        if (c == 1) e1 = e;
    }
    inv = 1.f / sum;
    const bool valid = lab >= 0 && lab < C && lab != a.ignore;
    wce = 0.f; wd = 0.f;
    if (a.kind == kCritCE) {
        if (valid) wce = s.sce;
        return;
    }
    wd = (s.da * (float)lab + s.db) * (e1 * inv);
    if (a.kind == kCritFocalDice && valid) {
        const float pt = el * inv, ce = -logf(pt), om = 1.f - pt;
        wce = s.sf * a.alpha * (crit_pow(om, a.gamma) + a.gamma * crit_pow(om, a.gamma - 1.f) * pt * ce);
    }
}

template <int CT, bool UP>
__global__ void __launch_bounds__(kT)
crit_fwd_kernel(const float* __restrict__ logits, const long long* __restrict__ label, UpGeom g, long long npix, CritArgs a,
                double* __restrict__ partial) {
    __shared__ double sm[4 * 8];
    const int C = g.C;
    const long long HW = (long long)g.H * g.W;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    const bool small = npix <= 0x7fffffffLL;       // 32-bit index arithmetic (a 64-bit division costs ~100 instructions per pixel)
    for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < npix; i += (long long)gridDim.x * kT) {
        long long n, p;
        if (small) { const unsigned int q = (unsigned int)i / (unsigned int)HW; n = q; p = (unsigned int)i - q * (unsigned int)HW; }
        else { n = i / HW; p = i % HW; }
        CritPix<CT, UP> u;
        if (UP) {
            const int y = small ? (int)((unsigned int)p / (unsigned int)g.W) : (int)(p / g.W), x = (int)(p - (long long)y * g.W);
            ac_coord(y, g.scy, g.hl, u.y0, u.y1, u.ly);
            ac_coord(x, g.scx, g.wl, u.x0, u.x1, u.lx);
            u.lp = logits + n * C * g.hl * g.wl; u.cs = (long long)g.hl * g.wl; u.wl = g.wl;
        } else {
            u.lp = logits + n * C * HW + p; u.cs = HW;
        }
        u.load(C);
        crit_pixel_fwd<CT, UP>(u, C, label[i], a, acc);
    }
    block_sum<4>(acc, sm);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) partial[blockIdx.x * 4 + k] = acc[k];
    }
}

// The fused-resize forward for register-resident class counts: tiles of 64 rows x 32 columns walked by a capped grid (one partial
// per CTA, reduced in a fixed order), lane = column, warp = 8 rows; no per-pixel index divisions, row cache as above.
template <int CT>
__global__ void __launch_bounds__(kT)
crit_strip_fwd_kernel(const float* __restrict__ low, const long long* __restrict__ label, UpGeom g, int tiles_x, int tiles_y, int ntiles,
                      CritArgs a, double* __restrict__ partial) {
    __shared__ double sm[4 * 8];
    const int C = g.C, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int tx = tile % tiles_x, r = tile / tiles_x, ty = r % tiles_y, n = r / tiles_y;
        const int x = tx * 32 + lane;
        if (x >= g.W) continue;
        CritPix<CT, true> u;
        u.lp = low + (long long)n * C * g.hl * g.wl; u.cs = (long long)g.hl * g.wl; u.wl = g.wl;
        ac_coord(x, g.scx, g.wl, u.x0, u.x1, u.lx);
        CritRows<CT> rows;
        const int y_first = ty * 64 + warp * 8;
        const long long* lab = label + ((long long)n * g.H + y_first) * g.W + x;
#pragma unroll 1
        for (int rr = 0; rr < 8; ++rr) {
            if (y_first + rr >= g.H) break;
            ac_coord(y_first + rr, g.scy, g.hl, u.y0, u.y1, u.ly);
            rows.values(u, C);
            crit_pixel_fwd<CT, true>(u, C, lab[(long long)rr * g.W], a, acc);
        }
    }
    block_sum<4>(acc, sm);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) partial[blockIdx.x * 4 + k] = acc[k];
    }
}

// out (doubles): {loss, S0, S1, S2, S3, npix}; kind 0: S0 = sum nll, S1 = count; kinds 1, 2: S0 = sum p t, S1 = sum p, S2 = sum t,
// S3 = sum of the focal terms
__global__ void __launch_bounds__(kT)
crit_finalize_kernel(const double* __restrict__ partial, int nblocks, long long npix, CritArgs a, double* __restrict__ out) {
    __shared__ double sm[4 * 8];
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    for (int i = threadIdx.x; i < nblocks; i += kT) {
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] += partial[4 * i + k];
    }
    block_sum<4>(v, sm);
    if (threadIdx.x != 0) return;
    double loss;
    if (a.kind == kCritCE) {
        loss = v[0] / v[1];            // 0/0 = NaN when every pixel is ignored, like torch
    } else {
        const double dice = (2.0 * v[0] + (double)a.smooth) / (v[1] + v[2] + (double)a.smooth);
        loss = 1.0 - dice;
        if (a.kind == kCritFocalDice) loss = (1.0 - (double)a.dice_w) * (v[3] / (double)npix) + (double)a.dice_w * loss;
    }
    out[0] = loss; out[1] = v[0]; out[2] = v[1]; out[3] = v[2]; out[4] = v[3]; out[5] = (double)npix;
}

// logits at label resolution: dlogits[n][c][p] written directly
template <int CT>
__global__ void __launch_bounds__(kT)
crit_grad_kernel(const float* __restrict__ logits, const long long* __restrict__ label, int C, long long HW, long long npix, CritArgs a,
                 const double* __restrict__ out, const float* __restrict__ gout, float* __restrict__ dlogits) {
    const CritScal s = crit_scalars(out, __ldg(gout), a);
    const bool small = npix <= 0x7fffffffLL;
    for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < npix; i += (long long)gridDim.x * kT) {
        long long n, p;
        if (small) { const unsigned int q = (unsigned int)i / (unsigned int)HW; n = q; p = (unsigned int)i - q * (unsigned int)HW; }
        else { n = i / HW; p = i % HW; }
        const long long lab = label[i];
        CritPix<CT, false> u;
        u.lp = logits + n * C * HW + p; u.cs = HW;
        u.load(C);
        float* dp = dlogits + n * C * HW + p;
        if (C == 1 && a.kind != kCritCE) { dp[0] = crit_grad_binary(u.get(0), lab, a, s); continue; }
        float mx, inv, wce, wd, ev[CT > 0 ? CT : 1];
        crit_pixel_bwd_coef<CT, false>(u, C, lab, a, s, mx, inv, wce, wd, ev);
        CRIT_FOR_CLASSES(c) {
            const float pc = (CT > 0 ? ev[CT > 0 ? c : 0] : expf(u.get(c) - mx)) * inv;
            dp[(long long)c * HW] = wce * (pc - (c == lab ? 1.f : 0.f)) + wd * ((c == 1 ? 1.f : 0.f) - pc);
        }
    }
}

// low-resolution logits: grid (ceil(W/32), ceil(H/64), N), the tile / segmented-shuffle / shared-memory scheme of ohem_up_grad_kernel
template <int CT>
__global__ void __launch_bounds__(kT)
crit_up_grad_kernel(const float* __restrict__ low, const long long* __restrict__ label, UpGeom g, CritArgs a, const double* __restrict__ out,
                    const float* __restrict__ gout, float* __restrict__ dlow) {
    extern __shared__ float acc[];       // [C][kUpTR][kUpTC]
    const int C = g.C;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = blockIdx.z, yb = blockIdx.y * 64, xb = blockIdx.x * 32;
    for (int i = tid; i < C * kUpTR * kUpTC; i += kT) acc[i] = 0.f;
    __syncthreads();
    const CritScal s = crit_scalars(out, __ldg(gout), a);
    const int ry0 = min((int)(g.scy * (float)yb), g.hl - 1), rx0 = min((int)(g.scx * (float)xb), g.wl - 1);
    const int x = xb + lane;
    const bool live_x = x < g.W;
    CritPix<CT, true> u;
    u.lp = low + (long long)n * C * g.hl * g.wl; u.cs = (long long)g.hl * g.wl; u.wl = g.wl;
    u.x0 = 0; u.x1 = 0; u.lx = 0.f;
    if (live_x) ac_coord(x, g.scx, g.wl, u.x0, u.x1, u.lx);
    const int prev_x0 = __shfl_up_sync(0xffffffffu, u.x0, 1);
    const bool head = lane == 0 || prev_x0 != u.x0;
    unsigned int same = 0u;            // bit k: lane + 2^k is in this lane's segment (lanes that share x0 are contiguous)
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const int ux0 = __shfl_down_sync(0xffffffffu, u.x0, 1 << k);
        if (lane + (1 << k) < 32 && ux0 == u.x0) same |= 1u << k;
    }
    CritRows<CT> rows;
    // register-resident counts up to 19: a thread's gradient is accumulated per class for the two low-resolution rows of the current
    // row pair and reduced over the lanes / added to the tile only when the pair changes (every ~8 rows) instead of once per row
    constexpr bool kAccum = CT > 0 && CT <= 19;
    float g0[kAccum ? CT : 1], g1[kAccum ? CT : 1];
    int fy0 = -1, fy1 = -1;
    if (kAccum) {
#pragma unroll
        for (int c = 0; c < (kAccum ? CT : 1); ++c) { g0[c] = 0.f; g1[c] = 0.f; }
    }
    auto flush = [&]() {          // executed by whole warps (the row pair is warp-uniform)
        if (!kAccum || fy0 < 0) return;
        const int ay0 = (fy0 - ry0) * kUpTC, ay1 = (fy1 - ry0) * kUpTC, ax0 = u.x0 - rx0, ax1 = u.x1 - rx0;
#pragma unroll
        for (int c = 0; c < (kAccum ? CT : 1); ++c) {
            if (c < C) {
                float a0 = g0[c] * (1.f - u.lx), b0 = g0[c] * u.lx, a1 = g1[c] * (1.f - u.lx), b1 = g1[c] * u.lx;
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    const float ua0 = __shfl_down_sync(0xffffffffu, a0, 1 << k), ub0 = __shfl_down_sync(0xffffffffu, b0, 1 << k);
                    const float ua1 = __shfl_down_sync(0xffffffffu, a1, 1 << k), ub1 = __shfl_down_sync(0xffffffffu, b1, 1 << k);
                    if ((same >> k) & 1u) { a0 += ua0; b0 += ub0; a1 += ua1; b1 += ub1; }
                }
                if (head && live_x) {
                    float* ac = acc + c * kUpTR * kUpTC;
                    if (a0 != 0.f) atomicAdd(ac + ay0 + ax0, a0);
                    if (b0 != 0.f) atomicAdd(ac + ay0 + ax1, b0);
                    if (a1 != 0.f) atomicAdd(ac + ay1 + ax0, a1);
                    if (b1 != 0.f) atomicAdd(ac + ay1 + ax1, b1);
                }
                g0[c] = 0.f; g1[c] = 0.f;
            }
        }
    };
    for (int rr = 0; rr < 8; ++rr) {
        const int y = yb + warp * 8 + rr;
        if (y >= g.H) break;                                   // warp-uniform
        ac_coord(y, g.scy, g.hl, u.y0, u.y1, u.ly);
        if (kAccum && (u.y0 != fy0 || u.y1 != fy1)) { flush(); fy0 = u.y0; fy1 = u.y1; }
        long long lab = a.ignore;
        float mx = 0.f, inv = 0.f, wce = 0.f, wd = 0.f, gb = 0.f, ev[CT > 0 ? CT : 1];
        const bool binary = C == 1 && a.kind != kCritCE;
        if (live_x) {
            lab = label[((long long)n * g.H + y) * g.W + x];
            rows.values(u, C);
            if (binary) gb = crit_grad_binary(u.get(0), lab, a, s);
            else crit_pixel_bwd_coef<CT, true>(u, C, lab, a, s, mx, inv, wce, wd, ev);
        } else if (CT > 0) {
#pragma unroll
            for (int c = 0; c < (CT > 0 ? CT : 1); ++c) u.v[c] = 0.f;
        }
        const int ay0 = (u.y0 - ry0) * kUpTC, ay1 = (u.y1 - ry0) * kUpTC, ax0 = u.x0 - rx0, ax1 = u.x1 - rx0;
        CRIT_FOR_CLASSES(c) {
            float gc = 0.f;
            if (live_x) {
                if (binary) gc = gb;
                else if (wce != 0.f || wd != 0.f) {
                    const float pc = (CT > 0 ? ev[CT > 0 ? c : 0] : expf(u.get(c) - mx)) * inv;
                    gc = wce * (pc - (c == lab ? 1.f : 0.f)) + wd * ((c == 1 ? 1.f : 0.f) - pc);
                }
            }
            if (kAccum) {
                g0[kAccum ? c : 0] = fmaf(gc, 1.f - u.ly, g0[kAccum ? c : 0]);
                g1[kAccum ? c : 0] = fmaf(gc, u.ly, g1[kAccum ? c : 0]);
                continue;
            }
            float ga = gc * (1.f - u.lx), gbx = gc * u.lx;       // towards columns x0 and x1
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const float ua = __shfl_down_sync(0xffffffffu, ga, 1 << k), ub = __shfl_down_sync(0xffffffffu, gbx, 1 << k);
                if ((same >> k) & 1u) { ga += ua; gbx += ub; }
            }
            if (head && live_x && (ga != 0.f || gbx != 0.f)) {
                float* ac = acc + c * kUpTR * kUpTC;
                atomicAdd(ac + ay0 + ax0, (1.f - u.ly) * ga);
                atomicAdd(ac + ay0 + ax1, (1.f - u.ly) * gbx);
                atomicAdd(ac + ay1 + ax0, u.ly * ga);
                atomicAdd(ac + ay1 + ax1, u.ly * gbx);
            }
        }
    }
    flush();
    __syncthreads();
    for (int i = tid; i < C * kUpTR * kUpTC; i += kT) {
        const float v = acc[i];
        if (v != 0.f) {
            const int c = i / (kUpTR * kUpTC), r = (i / kUpTC) % kUpTR, q = i % kUpTC;
            const int yy = ry0 + r, xx = rx0 + q;
            if (yy < g.hl && xx < g.wl) atomicAdd(dlow + (((long long)n * C + c) * g.hl + yy) * g.wl + xx, v);
        }
    }
}

// register-resident class values up to 32 classes: the smallest compile-time bound that holds c (19 = Cityscapes keeps its own)
static int crit_bucket(int c) { return c <= 2 ? c : c <= 4 ? 4 : c <= 8 ? 8 : c <= 16 ? 16 : c == 19 ? 19 : c <= 32 ? 32 : 0; }

size_t train_criterion_workspace_bytes() { return (size_t)num_sms() * 16 * 4 * sizeof(double) + 256; }

template <bool UP>
static void launch_crit_fwd(int c, int grid, const float* logits, const long long* label, const UpGeom& g, long long npix, const CritArgs& a,
                            double* partial, cudaStream_t s) {
#define CRIT_FWD(CTV) crit_fwd_kernel<CTV, UP><<<grid, kT, 0, s>>>(logits, label, g, npix, a, partial)
    switch (crit_bucket(c)) {
        case 1: CRIT_FWD(1); break;   case 2: CRIT_FWD(2); break;   case 4: CRIT_FWD(4); break;   case 8: CRIT_FWD(8); break;
        case 16: CRIT_FWD(16); break; case 19: CRIT_FWD(19); break; case 32: CRIT_FWD(32); break; default: CRIT_FWD(0); break;
    }
#undef CRIT_FWD
}

cudaError_t launch_train_criterion_fwd(const float* logits, const long long* label, double* out6, void* ws, int kind, int n, int c, int hl,
                                       int wl, int h, int w, long long ignore, float smooth, float alpha, float gamma, float dice_w,
                                       cudaStream_t s) {
    const long long npix = (long long)n * h * w;
    const CritArgs a{kind, ignore, smooth, alpha, gamma, dice_w};
    const UpGeom g = up_geom(c, hl, wl, h, w);
    const int grid = grid_for(npix);
    double* partial = reinterpret_cast<double*>(ws);
    int nparts = grid;
    if (hl == h && wl == w) {
        launch_crit_fwd<false>(c, grid, logits, label, g, npix, a, partial, s);
    } else if (crit_bucket(c) == 0) {
        launch_crit_fwd<true>(c, grid, logits, label, g, npix, a, partial, s);
    } else {
        const int tiles_x = (w + 31) / 32, tiles_y = (h + 63) / 64;
        const long long nt = (long long)tiles_x * tiles_y * n, cap = (long long)num_sms() * 16;
        if (nt > 0x7fffffffLL) return cudaErrorInvalidValue;
        nparts = (int)(nt < cap ? nt : cap);
#define CRIT_SFWD(CTV) crit_strip_fwd_kernel<CTV><<<nparts, kT, 0, s>>>(logits, label, g, tiles_x, tiles_y, (int)nt, a, partial)
        switch (crit_bucket(c)) {
            case 1: CRIT_SFWD(1); break;   case 2: CRIT_SFWD(2); break;   case 4: CRIT_SFWD(4); break;   case 8: CRIT_SFWD(8); break;
            case 16: CRIT_SFWD(16); break; case 19: CRIT_SFWD(19); break; default: CRIT_SFWD(32); break;
        }
#undef CRIT_SFWD
    }
    crit_finalize_kernel<<<1, kT, 0, s>>>(partial, nparts, npix, a, out6);
    return cudaGetLastError();
}

cudaError_t launch_train_criterion_bwd(const float* logits, const long long* label, const double* out6, const float* gout, float* dlogits,
                                       int kind, int n, int c, int hl, int wl, int h, int w, long long ignore, float smooth, float alpha,
                                       float gamma, float dice_w, cudaStream_t s) {
    const long long npix = (long long)n * h * w;
    const CritArgs a{kind, ignore, smooth, alpha, gamma, dice_w};
    if (hl == h && wl == w) {
        const int grid = grid_for(npix);
        const long long hw = (long long)h * w;
#define CRIT_GRAD(CTV) crit_grad_kernel<CTV><<<grid, kT, 0, s>>>(logits, label, c, hw, npix, a, out6, gout, dlogits)
        switch (crit_bucket(c)) {
            case 1: CRIT_GRAD(1); break;   case 2: CRIT_GRAD(2); break;   case 4: CRIT_GRAD(4); break;   case 8: CRIT_GRAD(8); break;
            case 16: CRIT_GRAD(16); break; case 19: CRIT_GRAD(19); break; case 32: CRIT_GRAD(32); break; default: CRIT_GRAD(0); break;
        }
#undef CRIT_GRAD
        return cudaGetLastError();
    }
    const size_t smem = (size_t)c * kUpTR * kUpTC * sizeof(float);
    if ((double)(hl - 1) * 7.0 > (double)(h - 1) || (double)(wl - 1) * 7.0 > (double)(w - 1) || smem > 48 * 1024) return cudaErrorInvalidValue;
    cudaError_t e = cudaMemsetAsync(dlogits, 0, (size_t)n * c * hl * wl * sizeof(float), s);
    if (e != cudaSuccess) return e;
    const dim3 grid((w + 31) / 32, (h + 63) / 64, n);
    const UpGeom g = up_geom(c, hl, wl, h, w);
#define CRIT_UPG(CTV) crit_up_grad_kernel<CTV><<<grid, kT, smem, s>>>(logits, label, g, a, out6, gout, dlogits)
    switch (crit_bucket(c)) {
        case 1: CRIT_UPG(1); break;   case 2: CRIT_UPG(2); break;   case 4: CRIT_UPG(4); break;   case 8: CRIT_UPG(8); break;
        case 16: CRIT_UPG(16); break; case 19: CRIT_UPG(19); break; case 32: CRIT_UPG(32); break; default: CRIT_UPG(0); break;
    }
#undef CRIT_UPG
    return cudaGetLastError();
}

}  // namespace fscnn
