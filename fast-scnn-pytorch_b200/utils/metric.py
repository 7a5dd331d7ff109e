"""``SegmentationMetric`` (pixAcc / mIoU) with the reference's interface (utils/metric.py:12-70)
and a device-side accumulator.

Reference behaviour kept: ``SegmentationMetric(nclass)``; ``update(preds, labels)`` with a numpy
array or a list/tuple of them (one image pair per element); ``get() -> (pixAcc, mIoU)`` as float64
with ``np.spacing(1)`` in the denominators and the mean over ALL ``nclass`` classes; ``reset()``; the
public int64 accumulators ``total_inter / total_union / total_correct / total_label``; updates are
serialised by ``self.lock``.

What is different: the counting itself (reference metric.py:73-105, three ``np.histogram`` passes on
the host, ~205 ms per 1024x2048 image) runs as a CUDA kernel on a confusion accumulator
``int64[(nclass+1)^2 + 2]`` that lives on the GPU; numpy inputs are copied to the device first.
CUDA tensors are accepted as well (the reference silently ignores them).  ``FastSCNN.evaluate``
feeds the same accumulator straight from the fused upsample+argmax kernel.  Counts are integers, so
results are bit-identical to the reference's for the same predictions.  Multi-GPU evaluation calls
``all_reduce()`` once at the end (one NCCL sum over the accumulator).
"""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np
import torch

__all__ = ['SegmentationMetric']

_DTYPE_CODE = {torch.uint8: 0, torch.int32: 1, torch.int64: 2}


class SegmentationMetric(object):
    """Computes pixAcc and mIoU metric scores."""

    def __init__(self, nclass, device=None):
        super(SegmentationMetric, self).__init__()
        self.nclass = int(nclass)
        self.lock = threading.Lock()
        self._device = torch.device(device) if device is not None else None
        self._conf = None
        self.reset()

    # ---- device-side state -----------------------------------------------------------------------
    def conf_len(self):
        return (self.nclass + 1) * (self.nclass + 1) + 2

    def device_confusion(self, device=None):
        """The int64 accumulator on ``device`` (created on first use)."""
        if self._conf is None:
            dev = torch.device(device) if device is not None else (self._device or torch.device('cuda', torch.cuda.current_device()))
            if dev.type != 'cuda':
                raise RuntimeError('SegmentationMetric (B200 build) counts on a CUDA device; there is no CPU fallback')
            self._device = dev
            self._conf = torch.zeros(self.conf_len(), dtype=torch.int64, device=dev)
        elif device is not None and torch.device(device) != self._conf.device:
            raise RuntimeError(f'metric state lives on {self._conf.device}, got data on {device}')
        return self._conf

    def _to_device_map(self, a):
        if isinstance(a, np.ndarray):
            if a.dtype not in (np.uint8, np.int32, np.int64):
                a = a.astype(np.int64)
            a = torch.from_numpy(np.ascontiguousarray(a))
        if a.dtype not in _DTYPE_CODE:
            a = a.to(torch.int64)
        conf = self.device_confusion(a.device if a.is_cuda else None)
        return a.to(conf.device, non_blocking=True).contiguous()

    def _count(self, pred, label):
        from fscnn_b200 import native
        if tuple(pred.shape) != tuple(label.shape):
            raise AssertionError('predict and target shapes differ')   # reference asserts, metric.py:76/:89
        pred, label = self._to_device_map(pred), self._to_device_map(label)
        conf = self.device_confusion(pred.device)
        with torch.cuda.device(conf.device):
            native.check(native.lib().fscnn_confusion_from_mask(
                pred.data_ptr(), _DTYPE_CODE[pred.dtype], label.data_ptr(), _DTYPE_CODE[label.dtype], pred.numel(),
                self.nclass, conf.data_ptr(), torch.cuda.current_stream().cuda_stream), 'fscnn_confusion_from_mask')
            # pred/label staging tensors must outlive the kernel
            pred.record_stream(torch.cuda.current_stream())
            label.record_stream(torch.cuda.current_stream())

    # ---- reference interface ----------------------------------------------------------------------
    def update(self, preds, labels):
        """Adds predictions (class maps) and labels: arrays / tensors, or lists of them."""
        if isinstance(preds, (np.ndarray, torch.Tensor)):
            pairs = [(preds, labels)]
        elif isinstance(preds, (list, tuple)):
            pairs = list(zip(preds, labels))
        else:
            return
        with self.lock:
            for pred, label in pairs:
                self._count(pred, label)

    def add_confusion(self, conf):
        """Adds a pre-computed confusion accumulator (e.g. from ``FastSCNN.evaluate``)."""
        with self.lock:
            mine = self.device_confusion(conf.device)
            if conf.data_ptr() != mine.data_ptr():
                mine += conf.to(mine.device)

    def all_reduce(self, group=None):
        """Sums the accumulator over the ranks of ``group`` (one NCCL all-reduce of (nc+1)^2+2 int64)."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            with self.lock:
                dist.all_reduce(self.device_confusion(), op=dist.ReduceOp.SUM, group=group)

    def _totals(self):
        from fscnn_b200 import native
        if self._conf is None:
            return (np.zeros(self.nclass, np.int64), np.zeros(self.nclass, np.int64), 0, 0)
        host = self._conf.cpu().numpy().astype(np.int64)   # device -> host read (synchronises this stream)
        inter = np.zeros(self.nclass, dtype=np.int64)
        union = np.zeros(self.nclass, dtype=np.int64)
        correct, labeled = C.c_longlong(), C.c_longlong()
        ll = C.POINTER(C.c_longlong)
        native.check(native.lib().fscnn_conf_to_totals(host.ctypes.data_as(ll), self.nclass, inter.ctypes.data_as(ll),
                                                       union.ctypes.data_as(ll), C.byref(correct), C.byref(labeled)))
        return inter, union, correct.value, labeled.value

    @property
    def total_inter(self):
        return self._totals()[0]

    @property
    def total_union(self):
        return self._totals()[1]

    @property
    def total_correct(self):
        return self._totals()[2]

    @property
    def total_label(self):
        return self._totals()[3]

    def get(self):
        """Returns (pixAcc, mIoU), float64, exactly as reference metric.py:42-54."""
        inter, union, correct, labeled = self._totals()
        pixAcc = 1.0 * correct / (np.spacing(1) + labeled)
        IoU = 1.0 * inter / (np.spacing(1) + union)
        mIoU = IoU.mean()
        return pixAcc, mIoU

    def reset(self):
        """Resets the internal evaluation result to initial state."""
        if self._conf is not None:
            self._conf.zero_()
