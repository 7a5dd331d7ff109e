"""Drop-in for the reference ``utils/loss.py``: the criteria ``train.py`` builds (train.py:182-192) under their own names and
call signatures, computed by the training kernels of ``libfscnn_b200.so`` (no host round trip, no ATen loss kernels, no CPU path).

    MixSoftmaxCrossEntropyLoss       loss.py:103-124   nn.CrossEntropyLoss(ignore_index) on every head, aux heads x aux_weight
    SoftmaxCrossEntropyOHEMLoss      loss.py:127-182   the host numpy argsort becomes a device radix select
    MixSoftmaxCrossEntropyOHEMLoss   loss.py:185-206   BASELINE config 5's criterion
    DiceLoss / MixDiceLoss           loss.py:12-68     train.py's default --loss-type
    FocalDiceLoss                    loss.py:71-100

Every class takes what the model returns in training mode (a tuple of full-resolution logits).  It also accepts the heads'
LOW-RESOLUTION logits (``FastSCNN._train_forward_lowres``): any prediction smaller than the target is resized inside the loss
kernels (``F.interpolate(..., 'bilinear', align_corners=True)`` of models/fast_scnn.py:40, :44 composed with the criterion), which
is how ``fscnn_b200.Trainer`` avoids the full-resolution logits and their gradient.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from fscnn_b200 import train_ops
from fscnn_b200.trainer import OHEM_CLASS_WEIGHTS

__all__ = ['MixSoftmaxCrossEntropyLoss', 'MixSoftmaxCrossEntropyOHEMLoss', 'DiceLoss', 'MixDiceLoss']


def _pred4(pred, who):
    if not torch.is_tensor(pred):      # the reference fails the same way on a tuple: pred.dim() does not exist
        raise AttributeError(f"'{type(pred).__name__}' object has no attribute 'dim'")
    if pred.dim() != 4:
        raise NotImplementedError(f'{who}: only [N,C,H,W] logits are supported (the reference also takes ready probabilities), '
                                  f'got {tuple(pred.shape)}')
    return pred


class DiceLoss(nn.Module):
    """loss.py:12-39: 1 - (2 sum(p t) + smooth) / (sum p + sum t + smooth), p = softmax(pred)[:, 1] (sigmoid for one channel),
    t = target.float() of every pixel."""

    def __init__(self, smooth=1e-6, **kwargs):
        super().__init__()
        self.smooth = smooth

    def forward(self, pred, target):
        return train_ops.criterion(_pred4(pred, 'DiceLoss'), target, 'dice', smooth=self.smooth)


class MixDiceLoss(nn.Module):
    """loss.py:42-68: DiceLoss on the main head + aux_weight x DiceLoss on the second prediction (only the second, like the
    reference) when ``aux``."""

    def __init__(self, aux=True, aux_weight=0.4, smooth=1e-6, **kwargs):
        super().__init__()
        self.aux = aux
        self.aux_weight = aux_weight
        self.dice_loss = DiceLoss(smooth=smooth)

    def forward(self, preds, target):
        if isinstance(preds, tuple):
            loss = self.dice_loss(preds[0], target)
            if self.aux and len(preds) > 1:
                loss = loss + self.aux_weight * self.dice_loss(preds[1], target)
            return loss
        return self.dice_loss(preds, target)


class FocalDiceLoss(nn.Module):
    """loss.py:71-100: (1 - dice_weight) * mean(alpha (1 - pt)^gamma ce) + dice_weight * DiceLoss on ONE prediction tensor (the
    reference's forward calls pred.dim(), so a tuple raises AttributeError there and here)."""

    def __init__(self, alpha=0.5, gamma=2.0, dice_weight=0.5, smooth=1e-6, **kwargs):
        super().__init__()
        self.alpha = alpha
        self.gamma = gamma
        self.dice_weight = dice_weight
        self.dice_loss = DiceLoss(smooth=smooth)

    def forward(self, pred, target):
        return train_ops.criterion(_pred4(pred, 'FocalDiceLoss'), target, 'focal_dice', ignore_label=-100, smooth=self.dice_loss.smooth,
                                   alpha=self.alpha, gamma=self.gamma, dice_weight=self.dice_weight)


class MixSoftmaxCrossEntropyLoss(nn.CrossEntropyLoss):
    """loss.py:103-124.  forward(preds, target): ``preds`` is the tuple the model returns."""

    def __init__(self, aux=True, aux_weight=0.2, ignore_label=-1, **kwargs):
        super().__init__(ignore_index=ignore_label)
        self.aux = aux
        self.aux_weight = aux_weight

    def _term(self, pred, target):
        return train_ops.criterion(_pred4(pred, 'MixSoftmaxCrossEntropyLoss'), target, 'ce', ignore_label=self.ignore_index)

    def forward(self, *inputs, **kwargs):
        preds, target = tuple(inputs)
        preds = list(preds)
        if self.aux:
            loss = self._term(preds[0], target)
            for p in preds[1:]:
                loss = loss + self.aux_weight * self._term(p, target)
            return loss
        if len(preds) != 1:      # the reference passes every prediction to nn.CrossEntropyLoss.forward(input, target)
            raise TypeError(f'forward() takes 3 positional arguments but {len(preds) + 2} were given')
        return self._term(preds[0], target)


class SoftmaxCrossEntropyOHEMLoss(nn.Module):
    """loss.py:127-182.  ``use_weight`` installs the reference's 19 Cityscapes class weights (valid for 19 classes only, as
    there)."""

    def __init__(self, ignore_label=-1, thresh=0.7, min_kept=256, use_weight=True, **kwargs):
        super().__init__()
        self.ignore_label = ignore_label
        self.thresh = float(thresh)
        self.min_kept = int(min_kept)
        if use_weight:
            self.register_buffer('weight', torch.tensor(OHEM_CLASS_WEIGHTS, dtype=torch.float32), persistent=False)
        else:
            self.weight = None

    def forward(self, predict, target, weight=None):
        # the reference's own argument checks (loss.py:144-147) are assertions; keep the exception type
        if target.requires_grad or predict.dim() != 4 or target.dim() != 3 or predict.shape[0] != target.shape[0]:
            raise AssertionError(f'expected logits [N,C,h,w] and a label map [N,H,W] without gradient, got {tuple(predict.shape)} and '
                                 f'{tuple(target.shape)}')
        w = self.weight
        if w is not None and w.device != predict.device:
            w = self.weight = w.to(predict.device)
        if tuple(predict.shape[2:]) != tuple(target.shape[1:]):
            return train_ops.ohem_cross_entropy_upsampled(predict, target, w, self.ignore_label, self.thresh, self.min_kept)
        return train_ops.ohem_cross_entropy(predict, target, w, self.ignore_label, self.thresh, self.min_kept)


class MixSoftmaxCrossEntropyOHEMLoss(SoftmaxCrossEntropyOHEMLoss):
    """loss.py:185-206"""

    def __init__(self, aux=False, aux_weight=0.2, ignore_index=-1, **kwargs):
        super().__init__(ignore_label=ignore_index, **kwargs)
        self.aux = aux
        self.aux_weight = aux_weight

    def forward(self, *inputs, **kwargs):
        preds, target = tuple(inputs)
        preds = list(preds)
        one = super().forward
        if self.aux:
            loss = one(preds[0], target)
            for p in preds[1:]:
                loss = loss + self.aux_weight * one(p, target)
            return loss
        if len(preds) > 2:
            raise TypeError(f'forward() takes from 3 to 4 positional arguments but {len(preds) + 2} were given')
        return one(preds[0], target)     # (a second prediction would land in the unused `weight` argument of loss.py:143)
