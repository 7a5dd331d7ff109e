/*
 * fscnn_b200.h -- C ABI of libfscnn_b200.so: the B200 (sm_100a) implementation of the
 * Fast-SCNN segmentation forward path (forward -> argmax -> SegmentationMetric counts).
 *
 * This is the drop-in boundary.  The reference (Shinokawa/Fast-SCNN-pytorch) is pure Python
 * and has no FFI of its own; what it binds for this path are torch.nn modules and numpy.  Each
 * entry point below names the reference interface it replaces (file:line into the reference
 * checkout).  The Python host (fast-scnn-pytorch_b200/models/fast_scnn.py, utils/metric.py)
 * binds these with ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *  - plain pointers and sizes only; no torch / ATen / pybind types cross this boundary;
 *  - every pointer named d_* is DEVICE memory owned by the caller (PyTorch allocates it);
 *    pointers named h_* are HOST memory; the library never allocates or frees device memory and
 *    never synchronises the device inside a forward call (CUDA-graph capturable);
 *  - all work is enqueued on the cudaStream_t passed as `stream` (void* here so that C callers
 *    need no CUDA headers);
 *  - every function returns 0 on success or a negative FSCNN_E* code; fscnn_last_error() returns
 *    a thread-local message for the last failure on the calling thread;
 *  - activations inside the workspace are NHWC; public inputs/outputs keep the reference's
 *    layouts: images and logits NCHW fp32, masks / labels [N,H,W].
 */
#ifndef FSCNN_B200_H
#define FSCNN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FSCNN_ABI_VERSION 1

/* error codes */
#define FSCNN_OK 0
#define FSCNN_EINVAL (-22)   /* bad argument (shape, dtype, null pointer, alignment) */
#define FSCNN_ENOENT (-2)    /* a required state_dict tensor / tap name was not found */
#define FSCNN_ENOMEM (-12)   /* caller-provided buffer too small */
#define FSCNN_ECUDA (-5)     /* a CUDA runtime call or kernel launch failed */
#define FSCNN_ESTATE (-1)    /* call order violated (e.g. forward before load_weights) */

/* activation precision of the internal stage tensors */
#define FSCNN_PREC_FP32 0    /* fp32 storage + fp32 FMA: the exactness path (1e-4 vs reference) */
#define FSCNN_PREC_BF16 1    /* bf16 storage + bf16 tensor-core contractions, fp32 accumulate */

/* element types of masks / labels */
#define FSCNN_U8 0
#define FSCNN_I32 1
#define FSCNN_I64 2
#define FSCNN_F32 3   /* float32 frames (fscnn_e2e_preprocess only) */

/* layout of the image batch handed to the forward calls */
#define FSCNN_IN_F32_NCHW 0  /* float32 [n,3,h,w], already normalised: what the reference's transforms produce */
#define FSCNN_IN_U8_NHWC 1   /* uint8 [n,h,w,3] raw RGB: ToTensor + Normalize(mean, std) are applied inside the stem */

typedef struct fscnn_ctx fscnn_ctx;

/* One raw state_dict tensor handed to fscnn_load_weights (reference on-disk layout:
 * train.py:449 torch.save(model.state_dict()); key names as in models/fast_scnn.py). */
typedef struct fscnn_tensor {
    const char* name;       /* state_dict key, without any "module." prefix */
    const float* d_data;    /* fp32, contiguous, device memory */
    int64_t numel;
} fscnn_tensor;

/* Shape of one stage-boundary tensor inside the workspace (for parity tests). */
typedef struct fscnn_tap {
    size_t offset_bytes;    /* from the start of the workspace */
    int n, h, w, c;         /* NHWC extents */
    int c_stride;           /* elements between consecutive pixels (>= c) */
    int elem_bytes;         /* 4 = fp32, 2 = bf16 */
} fscnn_tap;

int fscnn_abi_version(void);
const char* fscnn_last_error(void);

/* Replaces FastSCNN.__init__ (models/fast_scnn.py:16-31): builds the layer plan for
 * `num_classes` outputs, with or without the aux head.  `precision` is FSCNN_PREC_*. */
/* num_classes: 1..240 (FSCNN_PREC_FP32) or 1..128 (FSCNN_PREC_BF16); anything else is refused here with a message. */
int fscnn_create(fscnn_ctx** out, int num_classes, int aux, int precision);
void fscnn_destroy(fscnn_ctx* ctx);

/* Manifest of the state_dict tensors the forward path reads:
 * fscnn_param_count, then name / numel per index.  Lets a host validate a checkpoint. */
int fscnn_param_count(const fscnn_ctx* ctx);
const char* fscnn_param_name(const fscnn_ctx* ctx, int index);
int64_t fscnn_param_numel(const fscnn_ctx* ctx, int index);

/* Bytes of device memory the folded / repacked weights need. */
int fscnn_packed_weight_bytes(const fscnn_ctx* ctx, size_t* out_bytes);

/* Replaces nn.Module.load_state_dict + the eval-mode BatchNorm of every layer
 * (models/fast_scnn.py:55-57, 70-75, 86-88, 107-108, 198-203 and 251-255): folds BN
 * (eps 1e-5) into each convolution and repacks to the kernels' layouts, on the device.
 * `d_packed` (>= fscnn_packed_weight_bytes, 256-byte aligned) stays owned by the caller and
 * must outlive the ctx's forward calls.  May be called again after the weights change. */
int fscnn_load_weights(fscnn_ctx* ctx, const fscnn_tensor* tensors, int n_tensors,
                       void* d_packed, size_t packed_bytes, void* stream);

/* Selects how the forward calls interpret their image pointer.  With FSCNN_IN_U8_NHWC the stem
 * kernel computes (x / 255 - mean[c]) / std[c] on load, replacing transforms.ToTensor() +
 * transforms.Normalize(mean, std) (reference eval.py:22-25, demo.py:37-40); mean / std == NULL means
 * "/255 only" (the custom-dataset convention, data_loader/custom.py:175).  Default: FSCNN_IN_F32_NCHW. */
int fscnn_set_input_format(fscnn_ctx* ctx, int format, const float* mean3, const float* std3);

/* Bytes of workspace a forward over an [n,3,h,w] batch needs (stage tensors, NHWC). */
int fscnn_workspace_bytes(const fscnn_ctx* ctx, int n, int h, int w, size_t* out_bytes);

/* Replaces FastSCNN.forward (models/fast_scnn.py:33-46): x is NCHW fp32 [n,3,h,w] (or uint8 [n,h,w,3] after
 * fscnn_set_input_format(FSCNN_IN_U8_NHWC); the same holds for every forward entry point below);
 * d_logits is NCHW fp32 [n,nc,h,w]; d_aux_logits (same shape) may be NULL (must be non-NULL
 * to get the aux output when the ctx was created with aux=1). */
int fscnn_forward_logits(fscnn_ctx* ctx, const void* d_x, int n, int h, int w,
                         float* d_logits, float* d_aux_logits,
                         void* d_workspace, size_t workspace_bytes, void* stream);

/* Replaces forward + torch.argmax(outputs[0], 1) (eval.py:43-45, demo.py:47-48) without ever
 * materialising the full-resolution logits.  d_mask is [n,h,w] of mask_dtype (FSCNN_U8 needs
 * num_classes <= 256; FSCNN_I64 reproduces torch.argmax's dtype).  Class maps (d_mask here, d_labels / d_mask below) are
 * accessed four elements at a time: their base address must be aligned to 4 elements (4 bytes for FSCNN_U8, 16 bytes for
 * FSCNN_I32 / FSCNN_I64), otherwise the call returns FSCNN_EINVAL. */
int fscnn_forward_mask(fscnn_ctx* ctx, const void* d_x, int n, int h, int w,
                       void* d_mask, int mask_dtype,
                       void* d_workspace, size_t workspace_bytes, void* stream);

/* Replaces forward + argmax + SegmentationMetric.update (eval.py:43-49, utils/metric.py:56-105)
 * for one batch: ADDS this batch's counts into d_conf, an int64[(nc+1)*(nc+1)+2] accumulator
 * (rows = label clipped to nc, cols = pred, then `labeled`, `correct`; see fscnn_conf_len).
 * d_labels is [n,h,w] of label_dtype; d_mask may be NULL. */
int fscnn_forward_confusion(fscnn_ctx* ctx, const void* d_x, const void* d_labels, int label_dtype,
                            int n, int h, int w, long long* d_conf,
                            void* d_mask, int mask_dtype,
                            void* d_workspace, size_t workspace_bytes, void* stream);

/* Number of int64 elements of a confusion accumulator for `num_classes` classes. */
int64_t fscnn_conf_len(int num_classes);

/* Replaces batch_pix_accuracy + batch_intersection_union (utils/metric.py:73-105) for class maps
 * that are already on the device: ADDS into d_conf (layout as above). */
int fscnn_confusion_from_mask(const void* d_pred, int pred_dtype, const void* d_label, int label_dtype,
                              int64_t n_pixels, int num_classes, long long* d_conf, void* stream);

/* Host-side finalisation, utils/metric.py:42-54 and :56-63: turns a confusion accumulator (host
 * copy) into total_inter[nc], total_union[nc], total_correct, total_label. */
int fscnn_conf_to_totals(const long long* h_conf, int num_classes, long long* h_inter, long long* h_union,
                         long long* h_correct, long long* h_label);

/* The step right after the path: replaces get_color_pallete (utils/visualize.py:7-36, a
 * PIL palette on the host) with d_rgb[p] = palette[d_mask[p] & 255] on the device.  h_palette768 is HOST memory,
 * 256 x (R,G,B); d_rgb is [n_pixels][3] uint8. */
int fscnn_colorize(const void* d_mask, int mask_dtype, int64_t n_pixels, const unsigned char* h_palette768,
                   unsigned char* d_rgb, void* stream);

/* Overlay of a class map on the frame it belongs to (the step after the path in demo_tusimple.py:87-104, create_overlay):
 * d_out[p] = (uint8)((1 - alpha) * d_image[p] + alpha * palette[d_mask[p]]) for the pixels whose class is marked in the 256-bit
 * set h_draw_classes8 (bit c of word c / 32), d_image[p] elsewhere.  d_image / d_out uint8 [n_pixels][3]; float64 arithmetic and
 * truncation like the reference's numpy expression; h_palette768 and h_draw_classes8 are HOST memory. */
int fscnn_overlay(const unsigned char* d_image, const void* d_mask, int mask_dtype, int64_t n_pixels, const unsigned char* h_palette768,
                  const unsigned int* h_draw_classes8, double alpha, unsigned char* d_out, void* stream);

/* The camera-frame wrapper around the path: replaces EndToEndPreprocessing.forward and the tail of
 * EndToEndFastSCNN.forward (export_onnx_fixed.py:43-60, :78-98).
 *  preprocess : d_frames [n][3][h][w] uint8 (FSCNN_U8) or float32 (FSCNN_F32), values 0..255 -> bilinear resize
 *               (align_corners=False) to base_size x base_size, / 255, optional (x - mean) / std (HOST float[3] each, both or
 *               neither) -> d_out [n][3][base][base] float32, the input of fscnn_forward_*.
 *  postprocess: the low-resolution logits of the network ([n][hl][wl][padded_classes] float32, the "cls.logits_lowres" tap)
 *               -> x8 align_corners=True upsample to base_h x base_w (models/fast_scnn.py:40) composed with the resize to
 *               out_h x out_w (align_corners=False) and, if apply_softmax, the softmax over the classes
 *               -> d_out [n][num_classes][out_h][out_w] float32.  num_classes <= 32. */
int fscnn_e2e_preprocess(const void* d_frames, int frame_dtype, int n, int h, int w, int base_size, const float* h_mean3,
                         const float* h_std3, float* d_out, void* stream);
int fscnn_e2e_postprocess(const float* d_low_logits, int num_classes, int padded_classes, int n, int hl, int wl, int base_h,
                          int base_w, int out_h, int out_w, int apply_softmax, float* d_out, void* stream);

/* ---- test / profiling hooks ---------------------------------------------------------- */

/* Number of pipeline stages and their names ("stem", "l2d.dsconv1", ... "head"). */
int fscnn_stage_count(const fscnn_ctx* ctx);
const char* fscnn_stage_name(const fscnn_ctx* ctx, int stage);

/* Runs stages [first, last] only, reading / writing the stage tensors in the workspace (the
 * input image is read by stage 0).  Lets a test inject the oracle's tensor before one stage and
 * check that stage in isolation. */
int fscnn_forward_range(fscnn_ctx* ctx, const void* d_x, int n, int h, int w, int first_stage, int last_stage,
                        void* d_workspace, size_t workspace_bytes, void* stream);

/* Where the stage-boundary tensor `tap` ("l2d.conv", "l2d.dsconv1", "l2d.dsconv2",
 * "gfe.bottleneck1.0" ... "gfe.ppm", "ffm", "cls.dsconv1", "cls.logits_lowres",
 * "aux.logits_lowres") lives in a workspace planned for [n,3,h,w]. */
int fscnn_tap_info(const fscnn_ctx* ctx, int n, int h, int w, const char* tap, fscnn_tap* out);

/* The tail of the path as a stage of its own: the x8 bilinear upsample (align_corners=True, models/fast_scnn.py:40)
 * fused with torch.argmax(outputs[0], 1) (eval.py:45) and, when d_labels is given, with SegmentationMetric's counting
 * (utils/metric.py:73-105), from caller-provided low-resolution logits [n][hl][wl][padded_classes] float32 (the layout of
 * the "cls.logits_lowres" tap; padded_classes a multiple of 4, >= num_classes; 16-byte aligned).  d_mask [n][h][w]
 * (mask_dtype; may be NULL when labels are given), d_labels [n][h][w] or NULL, d_conf int64[fscnn_conf_len] accumulated
 * (required with labels).  (hl-1)*7.3 <= h-1 and (wl-1)*7.3 <= w-1 (the resize ratio of the network is 1/8).  Lets a test or a
 * benchmark drive the fused kernel with logits of its own (ties, NaN, class boundaries, adversarial orderings).
 * flags: FSCNN_TAIL_EXHAUSTIVE switches the exact class pruning off (every class is interpolated and compared at every
 * pixel): the result is identical by construction, only the time differs -- the reference point of the pruning tests and
 * of the worst-case timing. */
#define FSCNN_TAIL_EXHAUSTIVE 1
int fscnn_upsample_argmax(const float* d_low_logits, int num_classes, int padded_classes, int n, int hl, int wl, int h, int w,
                          void* d_mask, int mask_dtype, const void* d_labels, int label_dtype, long long* d_conf, int flags,
                          void* stream);

/* Kernel launches issued by this ctx since creation (for bench.py's gpu_launches). */
int64_t fscnn_launch_count(const fscnn_ctx* ctx);

/* Tuning: images pushed through the whole pipeline together (default: about 192 Mpixel of input, nudged to the count
 * whose tiles at the coarsest level fill whole waves of persistent CTAs; at most 128).  0 restores the default. */
int fscnn_set_micro_batch(fscnn_ctx* ctx, int images);

/* Tuning / A-B switches: "fuse_front" (1 = stem + dsconv1 as one kernel in the bf16 path, default), "micro_batch" (as
 * fscnn_set_micro_batch), "front_transposed" (1 = the transposed-stem front kernel whenever the input rows are 16-byte aligned,
 * default; 0 = always the kernel that also serves unaligned widths, l2d_front_tc.cu).  Unknown keys return FSCNN_EINVAL. */
int fscnn_set_option(fscnn_ctx* ctx, const char* key, int value);

/* ---- training step, first slice (reference train.py:253-284) ------------------------------
 * The building blocks of _DSConv / _DWConv / _ConvBNReLU(1x1) / LinearBottleneck in TRAINING mode (models/fast_scnn.py:49-115
 * under model.train(): BatchNorm uses batch statistics and updates its running statistics) and of
 * SoftmaxCrossEntropyOHEMLoss (utils/loss.py:127-182), forward and backward.  Tensors are fp32, contiguous, NCHW -- the layout
 * torch.autograd hands them over in; `hw` = H*W.  Every call takes a caller-owned scratch buffer d_ws of at least
 * fscnn_train_workspace_bytes(channels, cout, cin) bytes (16-byte aligned); reductions are two-stage and deterministic.
 * The Python host wraps these in torch.autograd.Function (fscnn_b200/train_ops.py) so that the covered modules train through
 * ordinary loss.backward().
 *  dwconv3x3 : nn.Conv2d(c, c, 3, stride, 1, groups=c, bias=False)  (fast_scnn.py:70, :86)   d_w [c][3][3]
 *  pwconv    : nn.Conv2d(cin, cout, 1, bias=False)                  (fast_scnn.py:73, :107)  d_w [cout][cin]
 *  batchnorm : nn.BatchNorm2d(c) in train mode (+ the nn.ReLU that follows when relu != 0): y = (x - mean) * rstd * gamma + beta
 *              with the batch's biased variance; d_running_mean / d_running_var (may both be NULL) are updated in place with
 *              `momentum` and the unbiased variance, exactly like PyTorch; d_save_mean / d_save_rstd [c] feed the backward.
 *              backward: d_dy is the gradient of the (ReLU'd) output y; the ReLU mask (y > 0) is recomputed from d_x with the
 *              forward's own arithmetic, so y itself is never read (d_beta is needed for that when relu != 0).
 * Any of d_dx / d_dw may be NULL in the conv backward calls when that gradient is not needed. */
int fscnn_train_workspace_bytes(int max_channels, int max_cout, int max_cin, size_t* out_bytes);
int fscnn_train_dwconv3x3_forward(const float* d_x, const float* d_w, float* d_y, int n, int c, int h, int w, int stride, void* stream);
int fscnn_train_dwconv3x3_backward(const float* d_x, const float* d_w, const float* d_dy, float* d_dx, float* d_dw, void* d_ws,
                                   size_t ws_bytes, int n, int c, int h, int w, int stride, void* stream);
int fscnn_train_pwconv_forward(const float* d_x, const float* d_w, float* d_y, int n, int cin, int cout, int hw, void* stream);
int fscnn_train_pwconv_backward(const float* d_x, const float* d_w, const float* d_dy, float* d_dx, float* d_dw, void* d_ws,
                                size_t ws_bytes, int n, int cin, int cout, int hw, void* stream);
int fscnn_train_batchnorm_forward(const float* d_x, const float* d_gamma, const float* d_beta, float* d_running_mean,
                                  float* d_running_var, float* d_y, float* d_save_mean, float* d_save_rstd, void* d_ws, size_t ws_bytes,
                                  int n, int c, int hw, float eps, float momentum, int relu, void* stream);
int fscnn_train_batchnorm_backward(const float* d_x, const float* d_dy, const float* d_gamma, const float* d_beta, const float* d_save_mean,
                                   const float* d_save_rstd, float* d_dx, float* d_dgamma, float* d_dbeta, void* d_ws, size_t ws_bytes,
                                   int n, int c, int hw, int relu, void* stream);

/* The remaining training-mode operators of the network (fast_scnn.py:24-31, :118-145, :190-237; train.py:195-198):
 *  im2col3x3 / col2im3x3 : the dense 3x3 convolutions (stem: stride 2 pad 0; aux head: stride 1 pad 1) run as im2col + the
 *              pointwise GEMMs above with cin -> cin*9: d_cols [n][c*9][ho*wo], k = ci*9 + ky*3 + kx (PyTorch's weight order);
 *              col2im3x3 folds the column gradient back into d_dx (gather form, deterministic)
 *  bias_add / bias_grad  : the bias of feature_fusion.conv_*_res.0, classifier.conv.1, auxlayer.4 (in place) and its gradient
 *  bilinear  : F.interpolate(mode='bilinear', align_corners=True) of [planes][hi][wi] -> [planes][ho][wo]; backward != 0: d_in is
 *              the gradient [planes][ho][wo], d_out the input gradient [planes][hi][wi] (gather form, deterministic)
 *  adaptive_avg_pool : nn.AdaptiveAvgPool2d(bins) with PyTorch's overlapping bins; backward != 0 as for bilinear
 *  dropout   : nn.Dropout(p) in train mode; the keep mask is a counter-based hash of (seed, element), so the backward is the
 *              same call on the gradient with the same seed; d_step (may be NULL) points to a device counter that is mixed into the
 *              seed when the kernel runs, so a captured CUDA graph of the step draws a fresh mask on every replay
 *  add_relu / relu_backward : y = relu(a + b) (relu = 0: plain add; FFM :217-218, residuals :114) and g = dy * [y > 0]
 *  sgd_step  : torch.optim.SGD(momentum, weight_decay) on flat buffers: g' = grad * grad_scale + wd * p;
 *              buf = first_step ? g' : momentum * buf + g'; p -= lr * buf */
/* The stem, LearningToDownsample.conv = nn.Conv2d(3, 32, 3, stride 2, padding 0, bias=False) (fast_scnn.py:153), without the column
 * matrix: d_x [n][3][h][w], d_w [32][3][3][3], d_y / d_dy [n][32][(h-3)/2+1][(w-3)/2+1]; the weight gradient wants the training
 * workspace (fscnn_train_workspace_bytes(32, 1, 1)).  The input image has no gradient. */
int fscnn_train_stem_forward(const float* d_x, const float* d_w, float* d_y, int n, int h, int w, void* stream);
int fscnn_train_stem_weight_grad(const float* d_x, const float* d_dy, float* d_dw, void* d_ws, size_t ws_bytes, int n, int h, int w,
                                 void* stream);
int fscnn_train_im2col3x3(const float* d_x, float* d_cols, int n, int c, int h, int w, int stride, int pad, void* stream);
int fscnn_train_col2im3x3(const float* d_dcols, float* d_dx, int n, int c, int h, int w, int stride, int pad, void* stream);
int fscnn_train_bias_add(float* d_y, const float* d_bias, int n, int c, int hw, void* stream);
int fscnn_train_bias_grad(const float* d_dy, float* d_dbias, void* d_ws, size_t ws_bytes, int n, int c, int hw, void* stream);
int fscnn_train_bilinear(const float* d_in, float* d_out, int planes, int hi, int wi, int ho, int wo, int backward, void* stream);
int fscnn_train_adaptive_avg_pool(const float* d_in, float* d_out, int planes, int h, int w, int bins, int backward, void* stream);
int fscnn_train_dropout(const float* d_x, float* d_y, float p, unsigned long long seed, const unsigned long long* d_step, int64_t numel,
                        void* stream);
int fscnn_train_add_relu(const float* d_a, const float* d_b, float* d_y, int relu, int64_t numel, void* stream);
int fscnn_train_relu_backward(const float* d_y, const float* d_dy, float* d_dx, int64_t numel, void* stream);
/* Arithmetic of the pointwise / dense-3x3 contractions of the training operators (process-wide; not a per-call argument because
 * torch.backends.cuda.matmul.allow_tf32, the switch it mirrors, is process-wide too):
 *   0  fp32 FMA on the CUDA cores (default; the mode the parity tests pin to 1e-4)
 *   1  TF32 operands (10 mantissa bits) on the tensor cores, fp32 accumulators: what cuDNN does for convolutions when torch's
 *      allow_tf32 is on (its default) and at least what the reference's fp16 autocast (train.py:267-275) keeps.  Forward and data
 *      gradient of 16-byte aligned shapes run tcgen05.mma.kind::tf32 (accumulator in TMEM, the tensor core truncates the fp32
 *      operands); the weight gradient and unaligned shapes run mma.sync.m16n8k8 with operands rounded to nearest
 * Returns FSCNN_EINVAL for any other mode.  fscnn_train_get_math returns the current mode. */
int fscnn_train_set_math(int mode);
int fscnn_train_get_math(void);
int fscnn_train_sgd_step(float* d_param, const float* d_grad, float* d_momentum_buf, float lr, float momentum, float weight_decay,
                         float grad_scale, int first_step, int64_t numel, void* stream);

/* torch.optim.AdamW(lr, betas, eps, weight_decay).step() (train_bdd100k.py:183-185 builds it with torch's defaults betas =
 * (0.9, 0.999), eps = 1e-8) on flat fp32 buffers, one launch: decoupled weight decay, first / second moment buffers d_exp_avg /
 * d_exp_avg_sq (zero before the first step), bias corrections from `step` (1 for the first call).  The gradient is read as
 * d_grad * grad_scale (1 / world size after a summing all-reduce). */
int fscnn_train_adamw_step(float* d_param, const float* d_grad, float* d_exp_avg, float* d_exp_avg_sq, float lr, double beta1, double beta2,
                           float eps, float weight_decay, float grad_scale, int64_t step, int64_t numel, void* stream);

/* SoftmaxCrossEntropyOHEMLoss.forward (utils/loss.py:143-182) without the host round trip: softmax probability of the target
 * class per valid pixel (d_label int64 [n][hw], label != ignore_label), the min(num_valid, min_kept)-th smallest of them by an
 * exact 4-pass radix select (replaces the host numpy argsort, loss.py:167-169), threshold = max(thresh, that value) unless
 * min_kept >= num_valid (keep everything), then nn.CrossEntropyLoss(weight=d_class_weight or NULL, ignore_index) over the
 * kept pixels.  d_prob [n][hw] float32 scratch that the backward re-reads; d_out3 = {loss, sum of kept weights, threshold};
 * d_ws >= fscnn_train_ohem_workspace_bytes().  backward: d_dlogits [n][c][hw] = d_grad_out[0] * d loss / d logits. */
int fscnn_train_ohem_workspace_bytes(size_t* out_bytes);
int fscnn_train_ohem_forward(const float* d_logits, const long long* d_label, const float* d_class_weight, float* d_prob, float* d_out3,
                             void* d_ws, size_t ws_bytes, int n, int c, int hw, long long ignore_label, float thresh, int min_kept,
                             void* stream);
int fscnn_train_ohem_backward(const float* d_logits, const long long* d_label, const float* d_class_weight, const float* d_prob,
                              const float* d_out3, const float* d_grad_out, float* d_dlogits, const void* d_ws, int n, int c, int hw,
                              long long ignore_label, void* stream);

/* The same loss straight from a head's LOW-RESOLUTION logits d_low_logits [n][c][hl][wl]: the final
 * F.interpolate(..., size=(h, w), mode='bilinear', align_corners=True) of models/fast_scnn.py:40 / :44 is composed with the OHEM
 * cross entropy, so the full-resolution logits and their gradient (717 MB each per head at 16 x 19 x 768 x 768) never exist.
 * Selection and loss are bit-identical to fscnn_train_bilinear + fscnn_train_ohem_forward; the backward writes
 * d_dlow [n][c][hl][wl] = d_grad_out[0] * d loss / d low-resolution logits (float atomics: summation order not deterministic).
 * d_nll (may be NULL): [n][h][w] float32 scratch; with it (and 2 or 19 classes) the probability pass leaves each pixel's negative
 * log-likelihood there and the loss pass only reads it instead of interpolating and exponentiating all classes a second time
 * (same expression on the same values: the loss stays bit-identical).
 * The backward needs (hl-1)*7 <= h-1, (wl-1)*7 <= w-1 and c <= 128 unless c is 2 or 19. */
int fscnn_train_ohem_upsampled_forward(const float* d_low_logits, const long long* d_label, const float* d_class_weight, float* d_prob,
                                       float* d_nll, float* d_out3, void* d_ws, size_t ws_bytes, int n, int c, int hl, int wl, int h, int w,
                                       long long ignore_label, float thresh, int min_kept, void* stream);
int fscnn_train_ohem_upsampled_backward(const float* d_low_logits, const long long* d_label, const float* d_class_weight, const float* d_prob,
                                        const float* d_out3, const float* d_grad_out, float* d_dlow, const void* d_ws, int n, int c, int hl,
                                        int wl, int h, int w, long long ignore_label, void* stream);

/* The reference's other training criteria (utils/loss.py), one head per call (the aux-head mix -- loss.py:112-116, :56-68 -- is
 * a weighted sum of calls on the host side):
 *   FSCNN_CRITERION_CE          nn.CrossEntropyLoss(ignore_index = ignore_label), the term MixSoftmaxCrossEntropyLoss applies to every
 *                               head (loss.py:103-124): mean negative log-likelihood over the pixels with label != ignore_label
 *   FSCNN_CRITERION_DICE        DiceLoss (loss.py:12-39; train.py:70's DEFAULT --loss-type, through MixDiceLoss loss.py:42-68):
 *                               p = softmax(x)[:, 1] (sigmoid(x) when c == 1), t = float(label) of EVERY pixel, loss =
 *                               1 - (2 sum(p t) + smooth) / (sum p + sum t + smooth); ignore_label is not used
 *   FSCNN_CRITERION_FOCAL_DICE  FocalDiceLoss (loss.py:71-100): (1 - dice_weight) * mean(alpha (1 - pt)^gamma ce) + dice_weight * DiceLoss
 *                               with ce = F.cross_entropy(reduction='none') (pass ignore_label = -100, torch's default: such pixels add 0
 *                               to the mean over all pixels) or, for c == 1, F.binary_cross_entropy(sigmoid(x), t); gamma >= 1
 * d_logits [n][c][hl][wl] float32, d_label [n][h][w] int64.  (hl, wl) == (h, w): the logits are at label resolution (what
 * FastSCNN.forward returns in training mode) and the backward is deterministic.  Otherwise d_logits is a head's LOW-RESOLUTION output
 * and F.interpolate(size=(h, w), mode='bilinear', align_corners=True) (models/fast_scnn.py:40, :44) is composed with the criterion, as
 * in fscnn_train_ohem_upsampled_*: the full-resolution logits and their gradient never exist; the backward then needs
 * (hl-1)*7 <= h-1, (wl-1)*7 <= w-1, c <= 128 and adds with float atomics.
 * d_out6 (float64): {loss, S0, S1, S2, S3, pixels} -- CE: S0 = sum of nll, S1 = pixels counted; dice: S0 = sum(p t), S1 = sum p,
 * S2 = sum t, S3 = sum of the focal terms; the backward re-reads it.  d_dlogits [n][c][hl][wl] = d_grad_out[0] * d loss / d logits.
 * Labels outside [0, c) other than ignore_label make torch raise; here they count as ignored in the CE / focal terms.
 * d_ws >= fscnn_train_criterion_workspace_bytes(). */
#define FSCNN_CRITERION_CE 0
#define FSCNN_CRITERION_DICE 1
#define FSCNN_CRITERION_FOCAL_DICE 2
int fscnn_train_criterion_workspace_bytes(size_t* out_bytes);
int fscnn_train_criterion_forward(const float* d_logits, const long long* d_label, double* d_out6, void* d_ws, size_t ws_bytes, int kind,
                                  int n, int c, int hl, int wl, int h, int w, long long ignore_label, float smooth, float alpha, float gamma,
                                  float dice_weight, void* stream);
int fscnn_train_criterion_backward(const float* d_logits, const long long* d_label, const double* d_out6, const float* d_grad_out,
                                   float* d_dlogits, int kind, int n, int c, int hl, int wl, int h, int w, long long ignore_label, float smooth,
                                   float alpha, float gamma, float dice_weight, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FSCNN_B200_H */
