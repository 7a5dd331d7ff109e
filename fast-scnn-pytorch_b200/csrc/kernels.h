// kernels.h -- host-side launchers of the Fast-SCNN stage kernels (one .cu per kernel family).
#pragma once
#include "common.cuh"

namespace fscnn {

// Folded (BatchNorm merged) fp32 weights, "k-major": W[k][cout] so that a contraction chunk is a
// dense [KC][COUT] tile.  Depthwise weights are tap-major: W[tap][channel].
struct StemIn { int format; float mean[3]; float inv_std[3]; };   // FSCNN_IN_*; uint8 input: (x/255 - mean) * inv_std
struct StemW { const float *w, *b; };                       // [27][32] (k = ci*9+ky*3+kx), [32]
struct DsW { const float *wd, *bd, *wp, *bp; };             // [9][cin], [cin], [cin][cout], [cout]
struct HeadW { const float *w, *b; int nc, ncp; };          // [cin][ncp], [ncp]; ncp = nc rounded up to 4
struct BneckW { const float *we, *be, *wd, *bd, *wp, *bp; };  // [cin][6cin],[6cin],[9][6cin],[6cin],[6cin][cout],[cout]
struct PpmW {
    const float* wc[4]; const float* bc[4];                 // branch convs [128][32], [32]
    const float* wo_x; const float* wo_s[4]; const float* bo;  // out conv split: [128][128], 4x[32][128], [128]
};
struct FfmW { const float *wd, *bd, *wcat, *bcat; };        // [9][128],[128],[192][128] (64 higher | 128 lower),[128]
struct AuxW { const float *w, *b; HeadW head; };            // [576][32] (k = tap*64+ci), [32]

template <typename T>
cudaError_t launch_stem(const void* x, const StemIn& in, const StemW& w, T* out, int n, int h, int wd, int ho, int wo, cudaStream_t s);

// _DSConv: DW3x3(stride, pad 1)+ReLU -> PW 1x1+ReLU, optionally chained with the classifier's 1x1 head
// (head != nullptr: `out` is not written, low-res logits [n][ho][wo][ncp] fp32 are).
template <typename T>
cudaError_t launch_dsconv(int cin, int cout, int stride, const T* in, const DsW& w, T* out, const HeadW* head,
                          float* logits, int n, int hi, int wi, int ho, int wo, cudaStream_t s);

// LinearBottleneck: PW expand+ReLU -> DW3x3(stride)+ReLU -> PW project (+ residual).
template <typename T>
cudaError_t launch_bottleneck(int cin, int cout, int stride, const T* in, const BneckW& w, T* out,
                              int n, int hi, int wi, int ho, int wo, cudaStream_t s);

// PyramidPooling in three launches: per-row partial bin sums, per-bin reduce + branch conv + projection
// through the out conv, then the K=128 contraction with the interpolated branch term.
template <typename T>
cudaError_t launch_ppm(const T* in, const PpmW& w, float* rowsum, float* z, T* out, int n, int h, int wd, cudaStream_t s);

// FeatureFusionModule: bilinear (align_corners) upsample of `lower` to higher's size, DW3x3+ReLU,
// both 1x1 convs as one K=192 contraction, add, ReLU.
template <typename T>
cudaError_t launch_ffm(const T* higher, const T* lower, const FfmW& w, T* out, int n, int hh, int wh, int hl, int wl,
                       cudaStream_t s);

// Aux head: dense 3x3 (64->32, pad 1)+ReLU -> 1x1 to classes; low-res logits [n][h][w][ncp] fp32.
template <typename T>
cudaError_t launch_aux(const T* higher, const AuxW& w, float* logits, int n, int h, int wd, cudaStream_t s);

// Final bilinear (align_corners) upsample of low-res logits [n][hl][wl][ncp]:
//  - to full-resolution NCHW fp32 logits (API parity with FastSCNN.forward), or
//  - fused with argmax (first max wins, NaN counts as max) and optionally the confusion histogram.
cudaError_t launch_up_logits(const float* low, int nc, int ncp, float* out, int n, int hl, int wl, int h, int w,
                             cudaStream_t s);
cudaError_t launch_up_argmax(const float* low, int nc, int ncp, void* mask, int mask_dtype, const void* labels,
                             int label_dtype, unsigned long long* conf, int n, int hl, int wl, int h, int w,
                             cudaStream_t s, bool prune = true);
cudaError_t launch_confusion(const void* pred, int pred_dtype, const void* label, int label_dtype, long long npix, int nc,
                             unsigned long long* conf, cudaStream_t s);

// camera-frame wrapper (e2e.cu): preprocess frames -> network input; low-res logits -> frame-size probabilities / logits
cudaError_t launch_e2e_preprocess(const void* x, int is_u8, int n, int h, int w, int base, const float* mean3, const float* std3,
                                  float* out, cudaStream_t s);
cudaError_t launch_e2e_postprocess(const float* low, int nc, int ncp, int n, int hl, int wl, int bh, int bw, int oh, int ow,
                                   int apply_softmax, float* out, cudaStream_t s);

// rgb[p] = palette[mask[p] & 255]; palette768 is HOST memory (256 x RGB), passed to the kernel by value
cudaError_t launch_colorize(const void* mask, int dtype, long long npix, const unsigned char* palette768, unsigned char* rgb,
                            cudaStream_t s);

// out = image with the pixels of the drawn classes blended towards their palette colour (demo_tusimple.py:87-104)
cudaError_t launch_overlay(const unsigned char* image, const void* mask, int dtype, long long npix, const unsigned char* palette768,
                           const unsigned int* draw8, double alpha, unsigned char* out, cudaStream_t s);

// BN folding + repack (load time).  out_w[k'][co] = w[co][k] * gamma/sqrt(var+eps); with taps > 1 and
// `tap_major` the k index (ci*taps + tap) is permuted to (tap*cin + ci).  out_b = beta + (cbias - mean)*scale.
cudaError_t launch_fold(const float* w, const float* cbias, const float* gamma, const float* beta, const float* mean,
                        const float* var, int cout, int kdim, int taps, int tap_major, float* out_w, int ld_out,
                        float* out_b, int accumulate_bias, cudaStream_t s);


// bf16 tensor-core (tcgen05 / TMEM) variants ------------------------------------------------------------
cudaError_t launch_fold_umma(const float* w, const float* gamma, const float* var, int nrows, int kdim, int nc, int kc,
                             bf16* out, cudaStream_t s);
// bf16 path: error-diffused rounding of a folded depthwise table [9][c] to bf16-representable values, in place
cudaError_t launch_dw_round_bf16(float* wd, int c, cudaStream_t s);
// stride-1 layers, transposed expand (bottleneck_s1t_tc.cu): the expanded tile goes TMEM -> registers, never through shared
// memory.  we_img / tab from launch_pack_s1t, wp_img = launch_fold_umma(project, nc = cout, kc = 128) into a zeroed buffer.
size_t bottleneck_s1t_we_bytes(int cin);
size_t bottleneck_s1t_wp_bytes(int cin, int cout);
size_t bottleneck_s1t_tab_bytes(int cin, int cout);
cudaError_t launch_pack_s1t(const BneckW& w, int cin, int cout, bf16* we_img, unsigned char* tab, cudaStream_t s);
cudaError_t launch_bottleneck_s1t_tc(int cin, int cout, const bf16* in, const unsigned char* tab, const bf16* we_img,
                                     const bf16* wp_img, bf16* out, int n, int h, int w, cudaStream_t s);
// stride-2 layers, transposed expand over 2x2 sub-tiles of 4x8 output pixels (bottleneck_s2t_tc.cu); same operand images
cudaError_t launch_bottleneck_s2t_tc(int cin, int cout, const bf16* in, const unsigned char* tab, const bf16* we_img,
                                     const bf16* wp_img, bf16* out, int n, int hi, int wi, int ho, int wo, cudaStream_t s);

// wp_img: pointwise weights [cout x cin] as one chunk; wh_img (head != nullptr): head weights [round_up(nc,16) x cout]
cudaError_t launch_dsconv_tc(int cin, int cout, int stride, const bf16* in, const DsW& w, const bf16* wp_img, bf16* out,
                             const HeadW* head, const bf16* wh_img, float* logits, int n, int hi, int wi, int ho, int wo,
                             cudaStream_t s);

// transposed FFM (ffm_t_tc.cu): bilinear resize as an MMA against a per-tile interpolation matrix, depthwise out of TMEM.
// tab = launch_pack_dw_tab(ffm.wd, ffm.bd, ffm.bcat, 128, 128): per-channel {9 bf16 taps, f32 bias} records + f32 fused bias
bool ffm_t_supported(int hh, int wh, int hl, int wl);
cudaError_t launch_pack_dw_tab(const float* wd, const float* bd, const float* bout, int c, int cout, unsigned char* tab, cudaStream_t s);
cudaError_t launch_ffm_t_tc(const bf16* higher, const bf16* lower, const unsigned char* tab, const bf16* wcat_img, bf16* out, int n,
                            int hh, int wh, int hl, int wl, cudaStream_t s);
// bf16 PPM with the output stage on the tensor core (ppm_tc.cu): wx_img = launch_fold_umma(out conv, 128 rows, kdim 256,
// nc 128, kc 128) chunk 0; z16: n x 16 KB, r_img: h x w x 64 bf16 (both workspace)
cudaError_t launch_ppm_out_tc(const bf16* in, const bf16* wx_img, const bf16* z_img, const float* bias, bf16* r_img, bf16* out, int n,
                              int h, int wd, cudaStream_t s);
cudaError_t launch_ppm_tc(const bf16* in, const PpmW& w, const bf16* wx_img, float* rowsum, float* z, bf16* z16, bf16* r_img, bf16* out,
                          int n, int h, int wd, cudaStream_t s);

// w_img: stem weights [32 x 27] padded to K = 32 as one chunk
cudaError_t launch_stem_tc(const void* x, const StemIn& in, const bf16* w_img, const float* bias, bf16* out, int n, int h,
                           int wd, int ho, int wo, cudaStream_t s);

// uint8 input: stem weights / bias with ToTensor + Normalize folded in (run when the input format changes)
cudaError_t launch_stem_refold(const float* w, const float* b, const StemIn& in, bf16* img, float* bias, cudaStream_t s);
// stem images of the fused kernel: 3 x (32 x 16) bf16, k = kx*4 + c over RGBX pixels (norm: uint8 normalisation folded in, bias written)
cudaError_t launch_stem_pack_rgbx(const float* w, const float* b, const StemIn& in, int norm, bf16* img, float* bias, cudaStream_t s);
// stem + dsconv1 fused (the stem's output never reaches HBM); ws_img from launch_stem_pack_rgbx, wp_img as for launch_dsconv_tc
cudaError_t launch_l2d_front_tc(const void* x, const StemIn& in, const bf16* ws_img, const float* bs, const DsW& w,
                                const bf16* wp_img, bf16* out, int n, int h, int wd, int h1, int w1, int h2, int w2, cudaStream_t s);

// the same with the transposed stem (l2d_front_t_tc.cu); cudaErrorNotSupported when the input rows are not 16-byte aligned
cudaError_t launch_l2d_front_t_tc(const void* x, const StemIn& in, const bf16* ws_img, const DsW& w, const bf16* wp_img, bf16* out, int n,
                                  int h, int wd, int h1, int w1, int h2, int w2, cudaStream_t s);

// ---- training step, first slice (train.cu): fp32 NCHW tensors as autograd hands them over ----
size_t train_workspace_bytes(int channels_max, int cout, int cin);
size_t train_ohem_workspace_bytes();
cudaError_t launch_train_dw_fwd(const float* x, const float* w, float* y, int n, int c, int h, int wd, int stride, cudaStream_t s);
cudaError_t launch_train_dw_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw, void* ws, int n, int c, int h,
                                int wd, int stride, cudaStream_t s);
cudaError_t launch_train_pw_fwd(const float* x, const float* w, float* y, int n, int cin, int cout, int hw, cudaStream_t s);
cudaError_t launch_train_pw_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw, void* ws, int n, int cin,
                                int cout, int hw, cudaStream_t s);
// tcgen05 TF32 pointwise contraction (train_tc.cu); cudaErrorNotSupported = shape not 16-byte aligned, use the mma.sync kernel
cudaError_t launch_gemm_pix_tc(const float* A, const float* B, float* C, int M, int N, int K, int nb, long long lda, bool a_kmajor,
                               cudaStream_t s);
cudaError_t launch_train_stem_fwd(const float* x, const float* w, float* y, int n, int h, int wd, cudaStream_t s);
cudaError_t launch_train_stem_wgrad(const float* x, const float* dy, float* dw, void* ws, int n, int h, int wd, cudaStream_t s);
int train_set_math(int mode);
int train_get_math();
cudaError_t launch_train_bn_fwd(const float* x, const float* gamma, const float* beta, float* running_mean, float* running_var,
                                float* y, float* save_mean, float* save_rstd, void* ws, int n, int c, int hw, float eps,
                                float momentum, int relu, cudaStream_t s);
cudaError_t launch_train_bn_bwd(const float* x, const float* dy, const float* gamma, const float* beta, const float* save_mean,
                                const float* save_rstd, float* dx, float* dgamma, float* dbeta, void* ws, int n, int c, int hw,
                                int relu, cudaStream_t s);
cudaError_t launch_train_ohem_up_fwd(const float* low, const long long* label, const float* weight, float* prob, float* nll, float* out3, void* ws,
                                     int n, int c, int hl, int wl, int h, int w, long long ignore, float thresh, int min_kept,
                                     cudaStream_t s);
cudaError_t launch_train_ohem_up_bwd(const float* low, const long long* label, const float* weight, const float* prob, const float* out3,
                                     const float* gout, float* dlow, const void* ws, int n, int c, int hl, int wl, int h, int w,
                                     long long ignore, cudaStream_t s);
cudaError_t launch_train_im2col(const float* x, float* cols, int n, int c, int h, int wd, int stride, int pad, cudaStream_t s);
cudaError_t launch_train_col2im(const float* dcols, float* dx, int n, int c, int h, int wd, int stride, int pad, cudaStream_t s);
cudaError_t launch_train_bias_add(float* y, const float* b, int n, int c, int hw, cudaStream_t s);
cudaError_t launch_train_bias_grad(const float* dy, float* db, void* ws, int n, int c, int hw, cudaStream_t s);
cudaError_t launch_train_bilinear(const float* in, float* out, int planes, int hi, int wi, int ho, int wo, int backward, cudaStream_t s);
cudaError_t launch_train_adaptive_pool(const float* in, float* out, int planes, int h, int wd, int bins, int backward, cudaStream_t s);
cudaError_t launch_train_dropout(const float* x, float* y, float p, unsigned long long seed, const unsigned long long* d_step, long long total,
                                 cudaStream_t s);
cudaError_t launch_train_add_relu(const float* a, const float* b, float* y, int relu, long long total, cudaStream_t s);
cudaError_t launch_train_relu_bwd(const float* y, const float* dy, float* dx, long long total, cudaStream_t s);
cudaError_t launch_train_sgd(float* p, const float* g, float* buf, float lr, float momentum, float wd, float gscale, int first,
                             long long total, cudaStream_t s);
cudaError_t launch_train_adamw(float* p, const float* g, float* m, float* v, float lr, double b1, double b2, float eps, float wd, float gscale,
                               long long step, long long total, cudaStream_t s);
cudaError_t launch_train_ohem_fwd(const float* logits, const long long* label, const float* weight, float* prob, float* out3, void* ws,
                                  int n, int c, int hw, long long ignore, float thresh, int min_kept, cudaStream_t s);
cudaError_t launch_train_ohem_bwd(const float* logits, const long long* label, const float* weight, const float* prob,
                                  const float* out3, const float* gout, float* dlogits, const void* ws, int n, int c, int hw,
                                  long long ignore, cudaStream_t s);
// the reference's other criteria (cross entropy / dice / focal + dice), at label resolution or fused with the head's final resize
size_t train_criterion_workspace_bytes();
cudaError_t launch_train_criterion_fwd(const float* logits, const long long* label, double* out6, void* ws, int kind, int n, int c, int hl,
                                       int wl, int h, int w, long long ignore, float smooth, float alpha, float gamma, float dice_w,
                                       cudaStream_t s);
cudaError_t launch_train_criterion_bwd(const float* logits, const long long* label, const double* out6, const float* gout, float* dlogits,
                                       int kind, int n, int c, int hl, int wl, int h, int w, long long ignore, float smooth, float alpha,
                                       float gamma, float dice_w, cudaStream_t s);

}  // namespace fscnn
