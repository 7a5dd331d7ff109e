"""CPU oracle for ``SegmentationMetric`` (pixAcc / mIoU).  TEST INFRASTRUCTURE ONLY.

Restates reference utils/metric.py:12-105 for integer class maps.  Two equivalent
formulations are kept so that they check each other:

* ``pixel_counts`` / ``inter_union`` follow the reference's own route (shift by one,
  zero the ignored pixels, count values 1..nclass) -- metric.py:73-105.  The reference
  counts with ``np.histogram(v, bins=nclass, range=(1, nclass))``; for integer ``v`` that
  maps value k (1 <= k <= nclass) to bin k-1 and drops everything else, which is what
  ``_count_1_to_n`` does with ``np.bincount``.
* ``confusion_counts`` is the layout the CUDA kernel accumulates:
  ``int64[(nclass+1)*(nclass+1) + 2]`` = confusion rows (label 0..nclass-1, plus one overflow
  row for label >= nclass) x columns (pred 0..nclass-1, plus one overflow column for preds
  outside the class range), then ``labeled`` and ``correct``.
  ``totals_from_confusion`` turns that into the reference's four accumulators.

Parity status: PINNED against the unmodified reference class run in the build container
(``oracle/gen_golden.py`` -> ``tests/golden/metric_*.npz``), including labels < -1,
labels >= nclass and list inputs.  Only tests/, smoke() and bench.py's CPU-baseline leg
may import this file.
"""
from __future__ import annotations

import numpy as np

SPACING1 = float(np.spacing(1))  # 2.220446049250313e-16, metric.py:50-51


def _count_1_to_n(values, n):
    v = values.reshape(-1)
    v = v[(v >= 1) & (v <= n)]
    return np.bincount(v, minlength=n + 1)[1:n + 1].astype(np.int64)


def pixel_counts(pred, label):
    """(correct, labeled) as metric.py:73-83: a pixel is labeled iff label+1 > 0."""
    pred = np.asarray(pred).astype(np.int64)
    label = np.asarray(label).astype(np.int64)
    if pred.shape != label.shape:
        raise AssertionError('shape mismatch')
    valid = label >= 0
    return int(np.count_nonzero((pred == label) & valid)), int(np.count_nonzero(valid))


def inter_union(pred, label, nclass):
    """(area_inter[nclass], area_union[nclass]) as metric.py:86-105."""
    pred = np.asarray(pred).astype(np.int64) + 1
    label = np.asarray(label).astype(np.int64) + 1
    if pred.shape != label.shape:
        raise AssertionError('shape mismatch')
    pred = np.where(label > 0, pred, 0)
    inter = np.where(pred == label, pred, 0)
    a_inter = _count_1_to_n(inter, nclass)
    a_pred = _count_1_to_n(pred, nclass)
    a_lab = _count_1_to_n(label, nclass)
    return a_inter, a_pred + a_lab - a_inter


def conf_len(nclass):
    """Length of the device-side accumulator for ``nclass`` classes."""
    return (nclass + 1) * (nclass + 1) + 2


def confusion_counts(pred, label, nclass):
    """int64[(nclass+1)*(nclass+1) + 2]: the device-side accumulator layout.

    Only labeled pixels (label >= 0, i.e. label+1 > 0, metric.py:79/:96) are counted.
    conf[r*(nclass+1) + c] with r = min(label, nclass) (last row: label >= nclass) and
    c = pred if 0 <= pred < nclass else nclass (last column: pred outside the class range);
    conf[-2] = labeled pixels, conf[-1] = correct pixels (pred == label, metric.py:81)."""
    pred = np.asarray(pred).astype(np.int64).reshape(-1)
    label = np.asarray(label).astype(np.int64).reshape(-1)
    valid = label >= 0
    w = nclass + 1
    rows = np.minimum(label[valid], nclass)
    pv = pred[valid]
    cols = np.where((pv >= 0) & (pv < nclass), pv, nclass)
    out = np.zeros(conf_len(nclass), dtype=np.int64)
    out[:w * w] = np.bincount(rows * w + cols, minlength=w * w)
    out[-2] = np.count_nonzero(valid)
    out[-1] = np.count_nonzero(valid & (pred == label))
    return out


def totals_from_confusion(conf, nclass):
    """(total_inter[nclass], total_union[nclass], total_correct, total_label) from the
    accumulator of ``confusion_counts``: inter = diagonal; area_pred = column sums over every
    labeled row (metric.py:96,101: pred is zeroed only where the pixel is unlabeled, so pixels
    labeled >= nclass still count towards their predicted class); area_lab = row sums over every
    column (metric.py:102) for rows < nclass."""
    conf = np.asarray(conf, dtype=np.int64)
    w = nclass + 1
    m = conf[:w * w].reshape(w, w)
    inter = np.diagonal(m)[:nclass].copy()
    area_pred = m[:, :nclass].sum(axis=0)
    area_lab = m[:nclass, :].sum(axis=1)
    return inter, area_pred + area_lab - inter, int(conf[-1]), int(conf[-2])


class SegmentationMetricOracle:
    """State machine of metric.py:12-70 (update / get / reset) on top of the counts above."""

    def __init__(self, nclass):
        self.nclass = nclass
        self.reset()

    def reset(self):
        self.total_inter = np.zeros(self.nclass, dtype=np.int64)
        self.total_union = np.zeros(self.nclass, dtype=np.int64)
        self.total_correct = 0
        self.total_label = 0

    def update(self, preds, labels):
        pairs = [(preds, labels)] if isinstance(preds, np.ndarray) else list(zip(preds, labels))
        for p, l in pairs:
            c, n = pixel_counts(p, l)
            i, u = inter_union(p, l, self.nclass)
            self.total_correct += c
            self.total_label += n
            self.total_inter += i
            self.total_union += u

    def get(self):
        pix_acc = 1.0 * self.total_correct / (SPACING1 + self.total_label)
        iou = 1.0 * self.total_inter / (SPACING1 + self.total_union)
        return pix_acc, iou.mean()
