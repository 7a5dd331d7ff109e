// ppm.cu -- PyramidPooling (reference models/fast_scnn.py:118-145) in three launches:
//   1. ppm_rowsum : per (image, row) partial sums of every adaptive-pool column bin (S = 1,2,3,6)
//   2. ppm_branch : per (image, bin) -- finish the mean (bins overlap when h % S != 0, :130-132),
//                   branch conv 128->32 + BN + ReLU (:124-127), and project the 32-vector through its
//                   slice of the out conv (:128): z[bin][128].  Upsampling and the 1x1 conv are both
//                   linear, so conv(upsample(f)) == upsample(conv(f)) and the 256-channel concat
//                   (:143) never exists.
//   3. ppm_out    : out = relu(b + x * Wx + sum_S bilinear_align_corners(z_S)) -- a K = 128
//                   contraction with an interpolated additive term (:134-135, :144).
// No atomics: results are deterministic.
#include "kernels.h"

namespace fscnn {

constexpr int kPpmC = 128;     // channels in / out
constexpr int kPpmBins = 50;   // 1 + 4 + 9 + 36
constexpr int kPpmCols = 12;   // 1 + 2 + 3 + 6 column bins

__device__ __forceinline__ int bin_start(int i, int n, int s) { return (i * n) / s; }
__device__ __forceinline__ int bin_end(int i, int n, int s) { return ((i + 1) * n + s - 1) / s; }

template <typename T>
__global__ void __launch_bounds__(kPpmC) ppm_rowsum_kernel(const T* __restrict__ in, float* __restrict__ rowsum, int h, int wd) {
    const int y = blockIdx.x, n = blockIdx.y, c = threadIdx.x;
    pdl_launch_dependents();
    pdl_wait();
    const int scales[4] = {1, 2, 3, 6};
    const T* row = in + (((size_t)n * h + y) * wd) * kPpmC + c;
    float* o = rowsum + (((size_t)n * h + y) * kPpmCols) * kPpmC + c;
    int col = 0;
#pragma unroll
    for (int si = 0; si < 4; ++si) {
        const int s = scales[si];
        for (int j = 0; j < s; ++j, ++col) {
            float acc = 0.f;
            const int x1 = bin_end(j, wd, s);
            for (int x = bin_start(j, wd, s); x < x1; ++x) acc += Act<T>::ld1(row + (size_t)x * kPpmC);
            o[col * kPpmC] = acc;
        }
    }
}

// z16 (bf16 path): the same vectors as a tcgen05 B-operand image per image, element (co, bin) at bf16 index
// ((bin / 8) * 16 + co / 8) * 64 + (co % 8) * 8 + bin % 8, bins 50..63 zero (blockIdx.x runs to 64 then)
__global__ void __launch_bounds__(kPpmC) ppm_branch_kernel(const float* __restrict__ rowsum, PpmW w, float* __restrict__ z,
                                                            bf16* __restrict__ z16, int h, int wd) {
    __shared__ float mean[kPpmC];
    __shared__ float feat[32];
    const int bin = blockIdx.x, n = blockIdx.y, c = threadIdx.x;
    pdl_launch_dependents();
    pdl_wait();
    if (bin >= kPpmBins) {   // zero padding of the operand image
        z16[(size_t)n * kPpmC * 64 + ((size_t)(bin >> 3) * 16 + (c >> 3)) * 64 + (c & 7) * 8 + (bin & 7)] = __float2bfloat16_rn(0.f);
        return;
    }
    int si, s, local, colbase;
    if (bin < 1) { si = 0; s = 1; local = bin; colbase = 0; }
    else if (bin < 5) { si = 1; s = 2; local = bin - 1; colbase = 1; }
    else if (bin < 14) { si = 2; s = 3; local = bin - 5; colbase = 3; }
    else { si = 3; s = 6; local = bin - 14; colbase = 6; }
    const int by = local / s, bx = local % s;
    const int y0 = bin_start(by, h, s), y1 = bin_end(by, h, s);
    const int cnt = (y1 - y0) * (bin_end(bx, wd, s) - bin_start(bx, wd, s));
    float acc = 0.f;
    for (int y = y0; y < y1; ++y) acc += rowsum[(((size_t)n * h + y) * kPpmCols + colbase + bx) * kPpmC + c];
    mean[c] = acc / (float)cnt;
    __syncthreads();
    if (c < 32) {
        float f = __ldg(w.bc[si] + c);
        for (int k = 0; k < kPpmC; ++k) f = fmaf(mean[k], __ldg(w.wc[si] + k * 32 + c), f);
        feat[c] = relu(f);
    }
    __syncthreads();
    float o = 0.f;
#pragma unroll 8
    for (int j = 0; j < 32; ++j) o = fmaf(feat[j], __ldg(w.wo_s[si] + j * kPpmC + c), o);
    z[((size_t)n * kPpmBins + bin) * kPpmC + c] = o;
    if (z16) z16[(size_t)n * kPpmC * 64 + ((size_t)(bin >> 3) * 16 + (c >> 3)) * 64 + (c & 7) * 8 + (bin & 7)] = __float2bfloat16_rn(o);
}

// bilinear align_corners sample of z_S (an SxS grid of 128-vectors) at output pixel (y, x), 4 channels
__device__ __forceinline__ float4 sample_z(const float* __restrict__ zs, int s, int y, int x, float sy, float sx, int ch) {
    const float fy = sy * (float)y, fx = sx * (float)x;
    const int y0 = min((int)fy, s - 1), x0 = min((int)fx, s - 1);
    const int y1 = min(y0 + 1, s - 1), x1 = min(x0 + 1, s - 1);
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    const float4 v00 = __ldg(reinterpret_cast<const float4*>(zs + (y0 * s + x0) * kPpmC + ch));
    const float4 v01 = __ldg(reinterpret_cast<const float4*>(zs + (y0 * s + x1) * kPpmC + ch));
    const float4 v10 = __ldg(reinterpret_cast<const float4*>(zs + (y1 * s + x0) * kPpmC + ch));
    const float4 v11 = __ldg(reinterpret_cast<const float4*>(zs + (y1 * s + x1) * kPpmC + ch));
    const float hy = 1.f - ly, hx = 1.f - lx;
    float4 r;
    r.x = hy * (hx * v00.x + lx * v01.x) + ly * (hx * v10.x + lx * v11.x);
    r.y = hy * (hx * v00.y + lx * v01.y) + ly * (hx * v10.y + lx * v11.y);
    r.z = hy * (hx * v00.z + lx * v01.z) + ly * (hx * v10.z + lx * v11.z);
    r.w = hy * (hx * v00.w + lx * v01.w) + ly * (hx * v10.w + lx * v11.w);
    return r;
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 2)
ppm_out_kernel(const T* __restrict__ in, PpmW w, const float* __restrict__ z, T* __restrict__ out, int h, int wd) {
    constexpr int KC = 32, TN = 8;
    using CM = ColMap<TN>;
    __shared__ __align__(16) float As[KC * 128];
    __shared__ __align__(16) float Bs[KC * kPpmC];
    const int tid = threadIdx.x, n = blockIdx.y;
    const int npix = h * wd;
    const int p0 = blockIdx.x * 128;
    const int tn = tid & 15, tp = tid >> 4;
    const int cv = tid & 7, pl = tid >> 3;

    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    for (int kc0 = 0; kc0 < kPpmC; kc0 += KC) {
        if (kc0) __syncthreads();
        load_weight_tile<KC, kPpmC>(Bs, w.wo_x + (size_t)kc0 * kPpmC, kPpmC);
        for (int p = pl; p < 128; p += 32) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p0 + p < npix) v = Act<T>::ld4(in + ((size_t)n * npix + p0 + p) * kPpmC + kc0 + 4 * cv);
            const int col = p ^ (cv << 2);
            As[(4 * cv + 0) * 128 + col] = v.x;
            As[(4 * cv + 1) * 128 + col] = v.y;
            As[(4 * cv + 2) * 128 + col] = v.z;
            As[(4 * cv + 3) * 128 + col] = v.w;
        }
        __syncthreads();
        contract_chunk<KC, TN, 128, kPpmC, true>(acc, As, Bs, tp, tn);
    }

    const float* zn = z + (size_t)n * kPpmBins * kPpmC;
    const float sy2 = h > 1 ? 1.f / (float)(h - 1) : 0.f, sx2 = wd > 1 ? 1.f / (float)(wd - 1) : 0.f;
    const float sy3 = h > 1 ? 2.f / (float)(h - 1) : 0.f, sx3 = wd > 1 ? 2.f / (float)(wd - 1) : 0.f;
    const float sy6 = h > 1 ? 5.f / (float)(h - 1) : 0.f, sx6 = wd > 1 ? 5.f / (float)(wd - 1) : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int p = p0 + 8 * tp + i;
        if (p >= npix) continue;
        const int y = p / wd, x = p % wd;
        T* o = out + ((size_t)n * npix + p) * kPpmC;
#pragma unroll
        for (int q = 0; q < CM::NQ; ++q) {
            const int ch = CM::ch(tn, q, 0);
            const float4 b = __ldg(reinterpret_cast<const float4*>(w.bo + ch));
            const float4 z1 = __ldg(reinterpret_cast<const float4*>(zn + ch));   // S = 1: a broadcast
            const float4 z2 = sample_z(zn + 1 * kPpmC, 2, y, x, sy2, sx2, ch);
            const float4 z3 = sample_z(zn + 5 * kPpmC, 3, y, x, sy3, sx3, ch);
            const float4 z6 = sample_z(zn + 14 * kPpmC, 6, y, x, sy6, sx6, ch);
            float4 r;
            r.x = relu(acc[i][q * 4 + 0] + b.x + z1.x + z2.x + z3.x + z6.x);
            r.y = relu(acc[i][q * 4 + 1] + b.y + z1.y + z2.y + z3.y + z6.y);
            r.z = relu(acc[i][q * 4 + 2] + b.z + z1.z + z2.z + z3.z + z6.z);
            r.w = relu(acc[i][q * 4 + 3] + b.w + z1.w + z2.w + z3.w + z6.w);
            Act<T>::st4(o + ch, r);
        }
    }
}

template <typename T>
cudaError_t launch_ppm(const T* in, const PpmW& w, float* rowsum, float* z, T* out, int n, int h, int wd, cudaStream_t s) {
    ppm_rowsum_kernel<T><<<dim3(h, n), kPpmC, 0, s>>>(in, rowsum, h, wd);
    ppm_branch_kernel<<<dim3(kPpmBins, n), kPpmC, 0, s>>>(rowsum, w, z, nullptr, h, wd);
    ppm_out_kernel<T><<<dim3(ceil_div(h * wd, 128), n), kThreads, 0, s>>>(in, w, z, out, h, wd);
    return cudaGetLastError();
}

// bf16 path with the output stage on the tensor core (ppm_tc.cu)
cudaError_t launch_ppm_tc(const bf16* in, const PpmW& w, const bf16* wx_img, float* rowsum, float* z, bf16* z16, bf16* r_img, bf16* out,
                          int n, int h, int wd, cudaStream_t s) {
    cudaError_t e = launch_pdl(ppm_rowsum_kernel<bf16>, dim3(h, n), kPpmC, 0, s, in, rowsum, h, wd);
    if (e != cudaSuccess) return e;
    e = launch_pdl(ppm_branch_kernel, dim3(64, n), kPpmC, 0, s, (const float*)rowsum, w, z, z16, h, wd);
    if (e != cudaSuccess) return e;
    return launch_ppm_out_tc(in, wx_img, z16, w.bo, r_img, out, n, h, wd, s);
}

template cudaError_t launch_ppm<float>(const float*, const PpmW&, float*, float*, float*, int, int, int, cudaStream_t);

}  // namespace fscnn
