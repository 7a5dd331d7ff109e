"""The reference's camera-frame wrapper as a drop-in (export_onnx_fixed.py:34-98): ``EndToEndFastSCNN(backbone, input_size,
base_size, mean, std, apply_softmax)`` takes raw frames ``[B,3,H,W]`` (uint8 or float32, 0..255), resizes them to
``base_size x base_size`` (bilinear, align_corners=False), scales by 1/255 (+ optional mean/std), runs the backbone, resizes
the logits back to ``input_size`` (given as (W, H), like the reference) and applies the softmax.

Here the preprocessing is one CUDA kernel and the tail is one CUDA kernel that composes the backbone's final x8 upsample
(align_corners=True), the resize back and the softmax straight from the low-resolution logits: the ``nc x base x base``
full-resolution logits of the reference never exist.  The ONNX-only PyramidPooling rewrite of the same reference file
(:100-137) is an export workaround with different arithmetic and is not reproduced: the backbone keeps the PyTorch semantics.
"""
from __future__ import annotations

import torch
import torch.nn as nn

__all__ = ['EndToEndFastSCNN', 'EndToEndPreprocessing']


class EndToEndPreprocessing(nn.Module):
    """export_onnx_fixed.py:62-98."""

    def __init__(self, input_size=(640, 360), base_size=1024, mean=None, std=None):
        super().__init__()
        self.input_size = input_size
        self.base_size = base_size
        if mean is not None and std is not None:
            self.register_buffer('mean', torch.tensor(mean, dtype=torch.float32).view(1, 3, 1, 1))
            self.register_buffer('std', torch.tensor(std, dtype=torch.float32).view(1, 3, 1, 1))
        else:
            self.mean = None
            self.std = None
        self._backbone = None      # set by EndToEndFastSCNN: the native engine lives with the backbone

    def forward(self, x):
        if self._backbone is None:
            raise RuntimeError('EndToEndPreprocessing runs through its EndToEndFastSCNN (the CUDA engine belongs to the backbone)')
        if x.device.type != 'cuda':
            raise RuntimeError('the B200 build runs on CUDA tensors only (no CPU fallback)')
        eng = self._backbone[0]._engine(x.device)
        return eng.e2e_preprocess(x.detach(), self.base_size, *self._host_norm())

    def _host_norm(self):
        """mean / std as host lists for the kernel's arguments, read back from the buffers only when they change (a .tolist() per
        call would synchronise the device on every frame batch)."""
        if self.mean is None:
            return None, None
        key = (self.mean.data_ptr(), self.mean._version, self.std.data_ptr(), self.std._version)
        if getattr(self, '_norm_cache', (None,))[0] != key:
            self._norm_cache = (key, self.mean.flatten().tolist(), self.std.flatten().tolist())
        return self._norm_cache[1], self._norm_cache[2]


class EndToEndFastSCNN(nn.Module):
    """export_onnx_fixed.py:34-60; ``backbone_model`` is a ``models.fast_scnn.FastSCNN`` of this package."""

    def __init__(self, backbone_model, input_size=(640, 360), base_size=1024, mean=None, std=None, apply_softmax=True):
        super().__init__()
        self.backbone = backbone_model
        self.preprocessor = EndToEndPreprocessing(input_size, base_size, mean, std)
        self.preprocessor._backbone = (backbone_model,)        # a tuple: not registered as a second copy of the submodule
        self.apply_softmax = apply_softmax
        self.input_size = input_size

    @torch.no_grad()
    def forward(self, x):
        pre = self.preprocessor(x)
        eng = self.backbone._engine(pre.device)
        return eng.e2e_forward(pre, self.input_size[1], self.input_size[0], self.apply_softmax)
