"""Training-mode operators of the covered modules (first slice) as ``torch.autograd.Function``s
over the C ABI (``fscnn_train_*`` in include/fscnn_b200.h, kernels in csrc/train.cu).

What is covered: depthwise 3x3 conv, pointwise 1x1 conv, BatchNorm2d with batch statistics (+ the ReLU behind it) -- which is
everything ``_DSConv``, ``_DWConv``, ``_ConvBNReLU(k=1)`` and ``LinearBottleneck`` (reference models/fast_scnn.py:49-115) do
in ``model.train()`` -- and ``SoftmaxCrossEntropyOHEMLoss`` (utils/loss.py:127-182).  Tensors are fp32 NCHW CUDA tensors;
there is no CPU path and no fallback to ATen's convolution / batch-norm kernels.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import native

_ws_cache = {}


MATH_MODES = {'fp32': 0, 'tf32': 1}


def set_matmul_precision(mode: str) -> None:
    """Arithmetic of the pointwise / dense 3x3 contractions of the training operators (process-wide, fscnn_train_set_math):
    'fp32' (default) = FMA on the CUDA cores; 'tf32' = TF32 operands on the tensor cores with fp32 accumulation (what cuDNN does
    under torch's default allow_tf32, and about the precision of the reference's fp16 autocast)."""
    if mode not in MATH_MODES:
        raise ValueError(f'matmul precision must be one of {sorted(MATH_MODES)}, got {mode!r}')
    native.check(native.lib().fscnn_train_set_math(MATH_MODES[mode]), 'fscnn_train_set_math')


def get_matmul_precision() -> str:
    code = native.lib().fscnn_train_get_math()
    return next(k for k, v in MATH_MODES.items() if v == code)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _workspace(device: torch.device, channels: int, cout: int = 1, cin: int = 1) -> torch.Tensor:
    need = C.c_size_t()
    native.check(native.lib().fscnn_train_workspace_bytes(int(channels), int(cout), int(cin), C.byref(need)))
    ws = _ws_cache.get(device)
    if ws is None or ws.numel() < need.value:
        ws = _ws_cache[device] = torch.empty(need.value, dtype=torch.uint8, device=device)
    return ws


def _check(x: torch.Tensor, what: str) -> torch.Tensor:
    if not x.is_cuda:
        raise RuntimeError(f'{what} is on {x.device}: the training operators run on CUDA devices only (there is no CPU fallback)')
    if x.dtype != torch.float32:
        raise ValueError(f'{what} must be float32, got {x.dtype}')
    return x.contiguous()


class DepthwiseConv3x3(torch.autograd.Function):
    """nn.Conv2d(c, c, 3, stride, 1, groups=c, bias=False)"""

    @staticmethod
    def forward(ctx, x, weight, stride):
        x, w = _check(x, 'input'), _check(weight, 'weight')
        n, c, h, wd = x.shape
        if tuple(w.shape) != (c, 1, 3, 3):
            raise ValueError(f'depthwise weight must be [{c},1,3,3], got {tuple(w.shape)}')
        ho, wo = (h - 1) // stride + 1, (wd - 1) // stride + 1
        y = torch.empty((n, c, ho, wo), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            native.check(native.lib().fscnn_train_dwconv3x3_forward(x.data_ptr(), w.data_ptr(), y.data_ptr(), n, c, h, wd, stride, _stream()),
                         'fscnn_train_dwconv3x3_forward')
        ctx.save_for_backward(x, w)
        ctx.stride = stride
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy.contiguous()
        n, c, h, wd = x.shape
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dw = torch.empty_like(w) if ctx.needs_input_grad[1] else None
        if dx is None and dw is None:
            return None, None, None
        ws = _workspace(x.device, c)
        with torch.cuda.device(x.device):
            native.check(native.lib().fscnn_train_dwconv3x3_backward(
                x.data_ptr(), w.data_ptr(), dy.data_ptr(), dx.data_ptr() if dx is not None else None,
                dw.data_ptr() if dw is not None else None, ws.data_ptr(), ws.numel(), n, c, h, wd, ctx.stride, _stream()),
                'fscnn_train_dwconv3x3_backward')
        return dx, dw, None


class PointwiseConv(torch.autograd.Function):
    """nn.Conv2d(cin, cout, 1, bias=False)"""

    @staticmethod
    def forward(ctx, x, weight):
        x, w = _check(x, 'input'), _check(weight, 'weight')
        n, cin, h, wd = x.shape
        cout = w.shape[0]
        if tuple(w.shape) != (cout, cin, 1, 1):
            raise ValueError(f'pointwise weight must be [cout,{cin},1,1], got {tuple(w.shape)}')
        y = torch.empty((n, cout, h, wd), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            native.check(native.lib().fscnn_train_pwconv_forward(x.data_ptr(), w.data_ptr(), y.data_ptr(), n, cin, cout, h * wd, _stream()),
                         'fscnn_train_pwconv_forward')
        ctx.save_for_backward(x, w)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy.contiguous()
        n, cin, h, wd = x.shape
        cout = w.shape[0]
        need_dx, need_dw = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        dx = torch.empty_like(x) if need_dx else None
        dw = (torch.zeros_like(w) if n > 64 else torch.empty_like(w)) if need_dw else None      # chunks of 64 images accumulate
        if dx is None and dw is None:
            return None, None
        ws = _workspace(x.device, 1, cout, cin)
        with torch.cuda.device(x.device):
            for i0 in range(0, n, 64):          # the weight gradient takes at most 64 images per call
                m = min(64, n - i0)
                part = torch.empty_like(w) if (need_dw and n > 64) else dw
                native.check(native.lib().fscnn_train_pwconv_backward(
                    x[i0:i0 + m].data_ptr(), w.data_ptr(), dy[i0:i0 + m].data_ptr(), dx[i0:i0 + m].data_ptr() if need_dx else None,
                    part.data_ptr() if need_dw else None, ws.data_ptr(), ws.numel(), m, cin, cout, h * wd, _stream()),
                    'fscnn_train_pwconv_backward')
                if need_dw and n > 64:
                    dw += part
        return dx, dw


class BatchNormReLU(torch.autograd.Function):
    """nn.BatchNorm2d in training mode (batch statistics, in-place running-statistics update) + optional nn.ReLU"""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, eps, momentum, relu):
        x, gamma, beta = _check(x, 'input'), _check(gamma, 'weight'), _check(beta, 'bias')
        n, c, h, wd = x.shape
        y = torch.empty_like(x)
        mean = torch.empty(c, dtype=torch.float32, device=x.device)
        rstd = torch.empty(c, dtype=torch.float32, device=x.device)
        ws = _workspace(x.device, c)
        with torch.cuda.device(x.device):
            native.check(native.lib().fscnn_train_batchnorm_forward(
                x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), running_mean.data_ptr() if running_mean is not None else None,
                running_var.data_ptr() if running_var is not None else None, y.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                ws.data_ptr(), ws.numel(), n, c, h * wd, float(eps), float(momentum), int(bool(relu)), _stream()),
                'fscnn_train_batchnorm_forward')
        # the running statistics are buffers (no gradient) updated in place by the kernel; autograd does not track them.
        # batchnorm_relu() bumps num_batches_tracked, which is what tells FastSCNN's eval engine to re-fold its weights.
        ctx.save_for_backward(x, gamma, beta, mean, rstd)      # the backward recomputes the ReLU mask from x: y is not kept
        ctx.relu = bool(relu)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, beta, mean, rstd = ctx.saved_tensors
        dy = dy.contiguous()
        n, c, h, wd = x.shape
        dx, dgamma, dbeta = torch.empty_like(x), torch.empty_like(gamma), torch.empty_like(gamma)
        ws = _workspace(x.device, c)
        with torch.cuda.device(x.device):
            native.check(native.lib().fscnn_train_batchnorm_backward(
                x.data_ptr(), dy.data_ptr(), gamma.data_ptr(), beta.data_ptr(), mean.data_ptr(), rstd.data_ptr(), dx.data_ptr(),
                dgamma.data_ptr(), dbeta.data_ptr(), ws.data_ptr(), ws.numel(), n, c, h * wd, int(ctx.relu), _stream()),
                'fscnn_train_batchnorm_backward')
        return dx, dgamma, dbeta, None, None, None, None, None


class OhemCrossEntropy(torch.autograd.Function):
    """SoftmaxCrossEntropyOHEMLoss.forward (reference utils/loss.py:143-182) on the device"""

    @staticmethod
    def forward(ctx, logits, target, class_weight, ignore_label, thresh, min_kept):
        logits = _check(logits, 'logits')
        if target.dtype != torch.int64 or target.dim() != 3 or not target.is_cuda:
            raise ValueError('target must be a CUDA int64 [N,H,W] tensor')
        n, c, h, w = logits.shape
        if tuple(target.shape) != (n, h, w):
            raise ValueError(f'target shape {tuple(target.shape)} does not match logits {tuple(logits.shape)}')
        target = target.contiguous()
        if class_weight is not None:
            class_weight = _check(class_weight, 'class weight')
            if class_weight.numel() != c:
                raise ValueError(f'class weight has {class_weight.numel()} entries for {c} classes')
        need = C.c_size_t()
        native.check(native.lib().fscnn_train_ohem_workspace_bytes(C.byref(need)))
        ws = torch.empty(need.value, dtype=torch.uint8, device=logits.device)
        prob = torch.empty((n, h, w), dtype=torch.float32, device=logits.device)
        out3 = torch.empty(3, dtype=torch.float32, device=logits.device)
        with torch.cuda.device(logits.device):
            native.check(native.lib().fscnn_train_ohem_forward(
                logits.data_ptr(), target.data_ptr(), class_weight.data_ptr() if class_weight is not None else None, prob.data_ptr(),
                out3.data_ptr(), ws.data_ptr(), ws.numel(), n, c, h * w, int(ignore_label), float(thresh), int(min_kept), _stream()),
                'fscnn_train_ohem_forward')
        ctx.save_for_backward(logits, target, prob, out3, ws)
        ctx.class_weight, ctx.ignore_label = class_weight, int(ignore_label)
        ctx.mark_non_differentiable(target)
        return out3[0].clone()

    @staticmethod
    def backward(ctx, gout):
        logits, target, prob, out3, ws = ctx.saved_tensors
        n, c, h, w = logits.shape
        dlogits = torch.empty_like(logits)
        g = gout.to(torch.float32).reshape(1).contiguous()
        cw = ctx.class_weight
        with torch.cuda.device(logits.device):
            native.check(native.lib().fscnn_train_ohem_backward(
                logits.data_ptr(), target.data_ptr(), cw.data_ptr() if cw is not None else None, prob.data_ptr(), out3.data_ptr(),
                g.data_ptr(), dlogits.data_ptr(), ws.data_ptr(), n, c, h * w, ctx.ignore_label, _stream()), 'fscnn_train_ohem_backward')
        return dlogits, None, None, None, None, None


class OhemCrossEntropyUpsampled(torch.autograd.Function):
    """ohem_cross_entropy(bilinear_resize(low, size), target) without materialising the full-resolution logits or their gradient"""

    @staticmethod
    def forward(ctx, low, target, class_weight, ignore_label, thresh, min_kept):
        low = _check(low, 'low-resolution logits')
        if target.dtype != torch.int64 or target.dim() != 3 or not target.is_cuda:
            raise ValueError('target must be a CUDA int64 [N,H,W] tensor')
        n, c, hl, wl = low.shape
        if target.shape[0] != n:
            raise ValueError(f'target batch {target.shape[0]} does not match logits batch {n}')
        h, w = int(target.shape[1]), int(target.shape[2])
        target = target.contiguous()
        if class_weight is not None:
            class_weight = _check(class_weight, 'class weight')
            if class_weight.numel() != c:
                raise ValueError(f'class weight has {class_weight.numel()} entries for {c} classes')
        need = C.c_size_t()
        native.check(native.lib().fscnn_train_ohem_workspace_bytes(C.byref(need)))
        ws = torch.empty(need.value, dtype=torch.uint8, device=low.device)
        prob = torch.empty((n, h, w), dtype=torch.float32, device=low.device)
        nll = torch.empty((n, h, w), dtype=torch.float32, device=low.device)      # scratch of the forward only (not saved)
        out3 = torch.empty(3, dtype=torch.float32, device=low.device)
        with torch.cuda.device(low.device):
            native.check(native.lib().fscnn_train_ohem_upsampled_forward(
                low.data_ptr(), target.data_ptr(), class_weight.data_ptr() if class_weight is not None else None, prob.data_ptr(),
                nll.data_ptr(), out3.data_ptr(), ws.data_ptr(), ws.numel(), n, c, hl, wl, h, w, int(ignore_label), float(thresh), int(min_kept), _stream()),
                'fscnn_train_ohem_upsampled_forward')
        ctx.save_for_backward(low, target, prob, out3, ws)
        ctx.class_weight, ctx.ignore_label = class_weight, int(ignore_label)
        ctx.mark_non_differentiable(target)
        return out3[0].clone()

    @staticmethod
    def backward(ctx, gout):
        low, target, prob, out3, ws = ctx.saved_tensors
        n, c, hl, wl = low.shape
        h, w = int(target.shape[1]), int(target.shape[2])
        dlow = torch.empty_like(low)
        g = gout.to(torch.float32).reshape(1).contiguous()
        cw = ctx.class_weight
        with torch.cuda.device(low.device):
            native.check(native.lib().fscnn_train_ohem_upsampled_backward(
                low.data_ptr(), target.data_ptr(), cw.data_ptr() if cw is not None else None, prob.data_ptr(), out3.data_ptr(), g.data_ptr(),
                dlow.data_ptr(), ws.data_ptr(), n, c, hl, wl, h, w, ctx.ignore_label, _stream()), 'fscnn_train_ohem_upsampled_backward')
        return dlow, None, None, None, None, None


def ohem_cross_entropy_upsampled(low_logits, target, class_weight: Optional[torch.Tensor] = None, ignore_label=-1, thresh=0.7, min_kept=256):
    """``ohem_cross_entropy(F.interpolate(low_logits, target.shape[1:], 'bilinear', align_corners=True), target, ...)`` fused: what the
    reference computes with models/fast_scnn.py:40 followed by utils/loss.py:143-182.  Falls back to the two-step form when the
    upsampling ratio is below 7 or there are more than 128 classes."""
    n, c, hl, wl = low_logits.shape
    h, w = target.shape[1], target.shape[2]
    if (hl - 1) * 7 > h - 1 or (wl - 1) * 7 > w - 1 or c > 128:
        return ohem_cross_entropy(bilinear_resize(low_logits, (h, w)), target, class_weight, ignore_label, thresh, min_kept)
    return OhemCrossEntropyUpsampled.apply(low_logits, target, class_weight, ignore_label, thresh, min_kept)


CRITERIA = {'ce': 0, 'dice': 1, 'focal_dice': 2}


class Criterion(torch.autograd.Function):
    """One head's cross entropy / dice / focal + dice term (reference utils/loss.py:103-124, :12-39, :71-100) on the device.
    ``logits`` are either at the labels' resolution or a head's low-resolution output, in which case the final
    ``F.interpolate(..., 'bilinear', align_corners=True)`` (models/fast_scnn.py:40, :44) is composed with the criterion and the
    full-resolution logits never exist (fscnn_train_criterion_*)."""

    @staticmethod
    def forward(ctx, logits, target, kind, ignore_label, smooth, alpha, gamma, dice_weight):
        logits = _check(logits, 'logits')
        if target.dtype != torch.int64 or target.dim() != 3 or not target.is_cuda:
            raise ValueError('target must be a CUDA int64 [N,H,W] tensor')
        n, c, hl, wl = logits.shape
        if target.shape[0] != n or target.device != logits.device:
            raise ValueError(f'target [{target.shape[0]}, ...] on {target.device} does not match logits [{n}, ...] on {logits.device}')
        h, w = int(target.shape[1]), int(target.shape[2])
        if hl > h or wl > w:
            raise ValueError(f'logits {hl}x{wl} are larger than the target {h}x{w}')
        target = target.contiguous()
        need = C.c_size_t()
        native.check(native.lib().fscnn_train_criterion_workspace_bytes(C.byref(need)))
        ws = torch.empty(need.value, dtype=torch.uint8, device=logits.device)
        out6 = torch.empty(6, dtype=torch.float64, device=logits.device)
        ctx.args = (int(kind), n, c, hl, wl, h, w, int(ignore_label), float(smooth), float(alpha), float(gamma), float(dice_weight))
        with torch.cuda.device(logits.device):
            native.check(native.lib().fscnn_train_criterion_forward(logits.data_ptr(), target.data_ptr(), out6.data_ptr(), ws.data_ptr(),
                                                                     ws.numel(), *ctx.args, _stream()), 'fscnn_train_criterion_forward')
        ctx.save_for_backward(logits, target, out6)
        ctx.mark_non_differentiable(target)
        return out6[0].to(torch.float32)

    @staticmethod
    def backward(ctx, gout):
        logits, target, out6 = ctx.saved_tensors
        dlogits = torch.empty_like(logits)
        g = gout.to(torch.float32).reshape(1).contiguous()
        with torch.cuda.device(logits.device):
            native.check(native.lib().fscnn_train_criterion_backward(logits.data_ptr(), target.data_ptr(), out6.data_ptr(), g.data_ptr(),
                                                                      dlogits.data_ptr(), *ctx.args, _stream()), 'fscnn_train_criterion_backward')
        return dlogits, None, None, None, None, None, None, None


def criterion(logits, target, kind: str, ignore_label=-1, smooth=1e-6, alpha=0.5, gamma=2.0, dice_weight=0.5):
    """``kind``: 'ce' = nn.CrossEntropyLoss(ignore_index=ignore_label) (the per-head term of MixSoftmaxCrossEntropyLoss), 'dice' =
    DiceLoss(smooth), 'focal_dice' = FocalDiceLoss(alpha, gamma, dice_weight, smooth) -- reference utils/loss.py.  When ``logits`` are
    smaller than ``target`` they are a head's low-resolution output and the x8 bilinear resize is fused (ratios below 7 or more than 128
    classes take the resize kernel first)."""
    if kind not in CRITERIA:
        raise ValueError(f'criterion must be one of {sorted(CRITERIA)}, got {kind!r}')
    n, c, hl, wl = logits.shape
    h, w = int(target.shape[1]), int(target.shape[2])
    if (hl, wl) != (h, w) and ((hl - 1) * 7 > h - 1 or (wl - 1) * 7 > w - 1 or c > 128):
        logits = bilinear_resize(logits, (h, w))
    if kind == 'ce' and (hl, wl) != (h, w) and tuple(logits.shape[2:]) == (hl, wl) and c in (2, 19):
        # plain cross entropy = the OHEM loss that keeps every valid pixel (min_kept >= the number of valid pixels, loss.py:160) without
        # class weights: for 2 / 19 classes that path has the strip kernels (class values in registers, gradients accumulated per
        # low-resolution row pair) -- 1.7 ms per config-5 step faster than the generic shared-memory scatter
        return ohem_cross_entropy_upsampled(logits, target, None, ignore_label, 0.7, 2 ** 31 - 1)
    if kind == 'focal_dice' and ignore_label == -1:
        ignore_label = -100       # F.cross_entropy's default ignore_index: what FocalDiceLoss.focal_loss runs with
    return Criterion.apply(logits, target, CRITERIA[kind], ignore_label, smooth, alpha, gamma, dice_weight)


def depthwise_conv3x3(x, weight, stride=1):
    return DepthwiseConv3x3.apply(x, weight, int(stride))


def pointwise_conv(x, weight):
    return PointwiseConv.apply(x, weight)


def batchnorm_relu(x, bn: torch.nn.BatchNorm2d, relu: bool):
    """``bn`` in training mode followed by ReLU when ``relu``; updates ``bn``'s running statistics like nn.BatchNorm2d does
    (momentum None = cumulative average is not used by the reference and is not supported)."""
    if bn.momentum is None:
        raise NotImplementedError('BatchNorm2d(momentum=None) is not used by the reference and not supported')
    track = bn.track_running_stats and bn.running_mean is not None
    y = BatchNormReLU.apply(x, bn.weight, bn.bias, bn.running_mean if track else None, bn.running_var if track else None,
                            bn.eps, bn.momentum, relu)
    if track and bn.num_batches_tracked is not None:
        if _deferred_batch_counters is not None:
            _deferred_batch_counters.append(bn.num_batches_tracked)
        else:
            bn.num_batches_tracked += 1
    return y


# Inside `deferred_batch_counters()` the num_batches_tracked increments of all BatchNorm layers of a forward are collected and
# applied by ONE multi-tensor add on exit (45 one-element kernels per step otherwise).
_deferred_batch_counters = None


class deferred_batch_counters:
    def __enter__(self):
        global _deferred_batch_counters
        self._outer, _deferred_batch_counters = _deferred_batch_counters, []
        return self

    def __exit__(self, *exc):
        global _deferred_batch_counters
        counters, _deferred_batch_counters = _deferred_batch_counters, self._outer
        if counters and exc[0] is None:
            torch._foreach_add_(counters, 1)
        return False


def ohem_cross_entropy(logits, target, class_weight: Optional[torch.Tensor] = None, ignore_label=-1, thresh=0.7, min_kept=256):
    return OhemCrossEntropy.apply(logits, target, class_weight, ignore_label, thresh, min_kept)


# ---- the remaining operators of the network (dense 3x3 conv, bias, bilinear resize, adaptive pooling, dropout, add + ReLU) ----
def _lib():
    return native.lib()


class Conv3x3Dense(torch.autograd.Function):
    """nn.Conv2d(cin, cout, 3, stride, pad, bias=False): im2col + the pointwise GEMM with cin -> cin * 9 (stem, aux head)"""

    @staticmethod
    def forward(ctx, x, weight, stride, pad):
        x, w = _check(x, 'input'), _check(weight, 'weight')
        n, c, h, wd = x.shape
        cout = w.shape[0]
        if tuple(w.shape) != (cout, c, 3, 3):
            raise ValueError(f'dense weight must be [cout,{c},3,3], got {tuple(w.shape)}')
        ho, wo = (h + 2 * pad - 3) // stride + 1, (wd + 2 * pad - 3) // stride + 1
        cols = torch.empty((n, c * 9, ho * wo), dtype=torch.float32, device=x.device)
        y = torch.empty((n, cout, ho, wo), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            native.check(_lib().fscnn_train_im2col3x3(x.data_ptr(), cols.data_ptr(), n, c, h, wd, stride, pad, _stream()), 'fscnn_train_im2col3x3')
            native.check(_lib().fscnn_train_pwconv_forward(cols.data_ptr(), w.data_ptr(), y.data_ptr(), n, c * 9, cout, ho * wo, _stream()),
                         'fscnn_train_pwconv_forward')
        ctx.save_for_backward(cols, w)
        ctx.geom = (n, c, h, wd, stride, pad, ho, wo)
        return y

    @staticmethod
    def backward(ctx, dy):
        cols, w = ctx.saved_tensors
        n, c, h, wd, stride, pad, ho, wo = ctx.geom
        cout = w.shape[0]
        dy = dy.contiguous()
        need_dx, need_dw = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        dcols = torch.empty_like(cols) if need_dx else None
        dw = (torch.zeros_like(w) if n > 64 else torch.empty_like(w)) if need_dw else None      # chunks of 64 images accumulate
        ws = _workspace(dy.device, 1, cout, c * 9)
        with torch.cuda.device(dy.device):
            for i0 in range(0, n, 64):
                m = min(64, n - i0)
                part = torch.empty_like(w) if (need_dw and n > 64) else dw
                native.check(_lib().fscnn_train_pwconv_backward(
                    cols[i0:i0 + m].data_ptr(), w.data_ptr(), dy[i0:i0 + m].data_ptr(), dcols[i0:i0 + m].data_ptr() if need_dx else None,
                    part.data_ptr() if need_dw else None, ws.data_ptr(), ws.numel(), m, c * 9, cout, ho * wo, _stream()),
                    'fscnn_train_pwconv_backward')
                if need_dw and n > 64:
                    dw += part
            dx = None
            if need_dx:
                dx = torch.empty((n, c, h, wd), dtype=torch.float32, device=dy.device)
                native.check(_lib().fscnn_train_col2im3x3(dcols.data_ptr(), dx.data_ptr(), n, c, h, wd, stride, pad, _stream()),
                             'fscnn_train_col2im3x3')
        return dx, dw, None, None


class BiasAdd(torch.autograd.Function):
    """y = x + bias[c] (the bias of a 1x1 convolution)"""

    @staticmethod
    def forward(ctx, x, bias):
        x, b = _check(x, 'input'), _check(bias, 'bias')
        n, c, h, w = x.shape
        y = x.clone()
        with torch.cuda.device(x.device):
            native.check(_lib().fscnn_train_bias_add(y.data_ptr(), b.data_ptr(), n, c, h * w, _stream()), 'fscnn_train_bias_add')
        ctx.shape = (n, c, h * w)
        return y

    @staticmethod
    def backward(ctx, dy):
        n, c, hw = ctx.shape
        dy = dy.contiguous()
        db = None
        if ctx.needs_input_grad[1]:
            db = torch.empty(c, dtype=torch.float32, device=dy.device)
            ws = _workspace(dy.device, c)
            with torch.cuda.device(dy.device):
                native.check(_lib().fscnn_train_bias_grad(dy.data_ptr(), db.data_ptr(), ws.data_ptr(), ws.numel(), n, c, hw, _stream()),
                             'fscnn_train_bias_grad')
        return dy, db


class BilinearResize(torch.autograd.Function):
    """F.interpolate(x, size, mode='bilinear', align_corners=True)"""

    @staticmethod
    def forward(ctx, x, ho, wo):
        x = _check(x, 'input')
        n, c, hi, wi = x.shape
        y = torch.empty((n, c, ho, wo), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            native.check(_lib().fscnn_train_bilinear(x.data_ptr(), y.data_ptr(), n * c, hi, wi, ho, wo, 0, _stream()), 'fscnn_train_bilinear')
        ctx.geom = (n, c, hi, wi, ho, wo)
        return y

    @staticmethod
    def backward(ctx, dy):
        n, c, hi, wi, ho, wo = ctx.geom
        dy = dy.contiguous()
        dx = torch.empty((n, c, hi, wi), dtype=torch.float32, device=dy.device)
        with torch.cuda.device(dy.device):
            native.check(_lib().fscnn_train_bilinear(dy.data_ptr(), dx.data_ptr(), n * c, hi, wi, ho, wo, 1, _stream()), 'fscnn_train_bilinear')
        return dx, None, None


class AdaptiveAvgPool(torch.autograd.Function):
    """nn.AdaptiveAvgPool2d(bins)"""

    @staticmethod
    def forward(ctx, x, bins):
        x = _check(x, 'input')
        n, c, h, w = x.shape
        y = torch.empty((n, c, bins, bins), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            native.check(_lib().fscnn_train_adaptive_avg_pool(x.data_ptr(), y.data_ptr(), n * c, h, w, bins, 0, _stream()),
                         'fscnn_train_adaptive_avg_pool')
        ctx.geom = (n, c, h, w, bins)
        return y

    @staticmethod
    def backward(ctx, dy):
        n, c, h, w, bins = ctx.geom
        dy = dy.contiguous()
        dx = torch.empty((n, c, h, w), dtype=torch.float32, device=dy.device)
        with torch.cuda.device(dy.device):
            native.check(_lib().fscnn_train_adaptive_avg_pool(dy.data_ptr(), dx.data_ptr(), n * c, h, w, bins, 1, _stream()),
                         'fscnn_train_adaptive_avg_pool')
        return dx, None


# A device int64 counter mixed into every dropout seed when set (Trainer's CUDA-graph mode bumps it before each replay, so a
# captured step draws fresh masks); None: the seed alone decides.
_dropout_step: Optional[torch.Tensor] = None


def set_dropout_step_counter(counter: Optional[torch.Tensor]) -> None:
    global _dropout_step
    if counter is not None and (not counter.is_cuda or counter.dtype != torch.int64 or counter.numel() != 1):
        raise ValueError('the dropout step counter must be a CUDA int64 tensor with one element')
    _dropout_step = counter


class Dropout(torch.autograd.Function):
    """nn.Dropout(p) in train mode; the mask is regenerated from the seed in the backward"""

    @staticmethod
    def forward(ctx, x, p, seed):
        x = _check(x, 'input')
        y = torch.empty_like(x)
        step = _dropout_step.data_ptr() if _dropout_step is not None else None
        with torch.cuda.device(x.device):
            native.check(_lib().fscnn_train_dropout(x.data_ptr(), y.data_ptr(), float(p), int(seed), step, x.numel(), _stream()),
                         'fscnn_train_dropout')
        ctx.p, ctx.seed, ctx.step = float(p), int(seed), step
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = dy.contiguous()
        dx = torch.empty_like(dy)
        with torch.cuda.device(dy.device):
            native.check(_lib().fscnn_train_dropout(dy.data_ptr(), dx.data_ptr(), ctx.p, ctx.seed, ctx.step, dy.numel(), _stream()),
                         'fscnn_train_dropout')
        return dx, None, None


class AddReLU(torch.autograd.Function):
    """relu(a + b) (relu=False: a + b)"""

    @staticmethod
    def forward(ctx, a, b, relu):
        a, b = _check(a, 'a'), _check(b, 'b')
        if a.shape != b.shape:
            raise ValueError(f'shapes differ: {tuple(a.shape)} vs {tuple(b.shape)}')
        y = torch.empty_like(a)
        with torch.cuda.device(a.device):
            native.check(_lib().fscnn_train_add_relu(a.data_ptr(), b.data_ptr(), y.data_ptr(), int(bool(relu)), a.numel(), _stream()),
                         'fscnn_train_add_relu')
        ctx.relu = bool(relu)
        if relu:
            ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = dy.contiguous()
        if not ctx.relu:
            return dy, dy, None
        (y,) = ctx.saved_tensors
        g = torch.empty_like(dy)
        with torch.cuda.device(dy.device):
            native.check(_lib().fscnn_train_relu_backward(y.data_ptr(), dy.data_ptr(), g.data_ptr(), dy.numel(), _stream()),
                         'fscnn_train_relu_backward')
        return g, g, None


class StemConv(torch.autograd.Function):
    """nn.Conv2d(3, 32, 3, stride=2, padding=0, bias=False), the network's first layer (models/fast_scnn.py:153), without the column
    matrix: the input window walks down the rows in registers (csrc/train.cu, stem_fwd_kernel / stem_wgrad_kernel).  The input
    image takes no gradient."""

    @staticmethod
    def forward(ctx, x, weight):
        x, w = _check(x, 'input'), _check(weight, 'weight')
        n, c, h, wd = x.shape
        ho, wo = (h - 3) // 2 + 1, (wd - 3) // 2 + 1
        y = torch.empty((n, 32, ho, wo), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            native.check(_lib().fscnn_train_stem_forward(x.data_ptr(), w.data_ptr(), y.data_ptr(), n, h, wd, _stream()), 'fscnn_train_stem_forward')
        ctx.save_for_backward(x, w)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        if not ctx.needs_input_grad[1]:
            return None, None
        dy = dy.contiguous()
        n, c, h, wd = x.shape
        dw = torch.empty_like(w)
        ws = _workspace(x.device, 32)
        with torch.cuda.device(x.device):
            native.check(_lib().fscnn_train_stem_weight_grad(x.data_ptr(), dy.data_ptr(), dw.data_ptr(), ws.data_ptr(), ws.numel(), n, h, wd,
                                                             _stream()), 'fscnn_train_stem_weight_grad')
        return None, dw


def conv3x3_dense(x, weight, stride, pad):
    if (tuple(weight.shape) == (32, 3, 3, 3) and int(stride) == 2 and int(pad) == 0 and x.dim() == 4 and x.shape[1] == 3
            and x.shape[2] >= 3 and x.shape[3] >= 3 and not x.requires_grad):
        return StemConv.apply(x, weight)
    return Conv3x3Dense.apply(x, weight, int(stride), int(pad))


def bias_add(x, bias):
    return BiasAdd.apply(x, bias)


def bilinear_resize(x, size):
    return BilinearResize.apply(x, int(size[0]), int(size[1]))


def adaptive_avg_pool(x, bins):
    return AdaptiveAvgPool.apply(x, int(bins))


_dropout_calls = [0]


def dropout(x, p, training=True, seed=None):
    """nn.Dropout(p).  ``seed`` None: drawn from torch's CPU generator, so ``torch.manual_seed`` makes runs reproducible."""
    if not training or p == 0.0:
        return x
    if seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    return Dropout.apply(x, float(p), seed)


def add_relu(a, b, relu=True):
    return AddReLU.apply(a, b, bool(relu))


def sgd_step(param_flat, grad_flat, momentum_buf, lr, momentum=0.9, weight_decay=1e-4, grad_scale=1.0, first_step=False):
    """torch.optim.SGD(momentum, weight_decay) (reference train.py:195-198) on flat fp32 buffers, one launch."""
    for t in (param_flat, grad_flat, momentum_buf):
        if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != param_flat.numel():
            raise ValueError('sgd_step takes three contiguous CUDA float32 buffers of equal length')
    with torch.cuda.device(param_flat.device):
        native.check(_lib().fscnn_train_sgd_step(param_flat.data_ptr(), grad_flat.data_ptr(), momentum_buf.data_ptr(), float(lr), float(momentum),
                                                 float(weight_decay), float(grad_scale), int(bool(first_step)), param_flat.numel(), _stream()),
                     'fscnn_train_sgd_step')


def adamw_step(param_flat, grad_flat, exp_avg, exp_avg_sq, lr, step, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, grad_scale=1.0):
    """torch.optim.AdamW (reference train_bdd100k.py:183-185) on flat fp32 buffers, one launch; ``step`` counts from 1."""
    for t in (param_flat, grad_flat, exp_avg, exp_avg_sq):
        if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != param_flat.numel():
            raise ValueError('adamw_step takes four contiguous CUDA float32 buffers of equal length')
    with torch.cuda.device(param_flat.device):
        native.check(_lib().fscnn_train_adamw_step(param_flat.data_ptr(), grad_flat.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(),
                                                   float(lr), float(betas[0]), float(betas[1]), float(eps), float(weight_decay),
                                                   float(grad_scale), int(step), param_flat.numel(), _stream()), 'fscnn_train_adamw_step')
