"""numpy (float64) restatement of the reference's cross-entropy / dice / focal + dice criteria (utils/loss.py:12-124) and of the
gradients torch.autograd derives for them, plus the composition with the heads' final bilinear resize (models/fast_scnn.py:40,
:44).  TEST INFRASTRUCTURE: pinned against fixtures produced by the unmodified reference classes (oracle/gen_golden_loss.py ->
tests/golden/train_loss_*.npz, tests/test_oracle_golden.py); only tests/ import it."""
import numpy as np

import fastscnn_oracle as fo


def _softmax(x):
    x = x - x.max(axis=1, keepdims=True)
    e = np.exp(x)
    return e / e.sum(axis=1, keepdims=True)


def cross_entropy(logits, target, ignore_label=-1):
    """nn.CrossEntropyLoss(ignore_index=ignore_label) as MixSoftmaxCrossEntropyLoss applies it to one head (loss.py:103-111):
    mean negative log-likelihood over the pixels whose label is not ignored.  Returns (loss, d loss / d logits)."""
    x = logits.astype(np.float64)
    n, c, h, w = x.shape
    p = _softmax(x)
    valid = target != ignore_label
    lab = np.where(valid, target, 0)
    onehot = np.eye(c)[lab].transpose(0, 3, 1, 2)
    pl = np.take_along_axis(p, lab[:, None], axis=1)[:, 0]
    cnt = valid.sum()
    loss = -(np.log(pl) * valid).sum() / cnt
    grad = (p - onehot) * (valid / cnt)[:, None]
    return loss, grad


def _dice_terms(logits, target):
    """p (probability of the positive class), dp/dlogits as a function applying the chain rule, t = target.float() (loss.py:26-33)."""
    x = logits.astype(np.float64)
    t = target.astype(np.float64)
    if x.shape[1] > 1:
        sm = _softmax(x)
        p = sm[:, 1]

        def back(gp):      # d softmax_1 / d x_c = p1 ([c == 1] - p_c)
            delta = np.zeros(x.shape[1])
            delta[1] = 1.0
            return (gp * p)[:, None] * (delta[None, :, None, None] - sm)
    else:
        p = 1.0 / (1.0 + np.exp(-x[:, 0]))

        def back(gp):
            return (gp * p * (1.0 - p))[:, None]
    return p, t, back


def dice(logits, target, smooth=1e-6):
    """DiceLoss.forward (loss.py:19-39).  Returns (loss, d loss / d logits)."""
    p, t, back = _dice_terms(logits, target)
    inter, den = (p * t).sum(), p.sum() + t.sum() + smooth
    loss = 1.0 - (2.0 * inter + smooth) / den
    gp = -(2.0 * t * den - (2.0 * inter + smooth)) / den ** 2
    return loss, back(gp)


def focal_dice(logits, target, alpha=0.5, gamma=2.0, dice_weight=0.5, smooth=1e-6):
    """FocalDiceLoss.forward (loss.py:81-100): labels equal to -100 (F.cross_entropy's default ignore_index) contribute 0 to the
    focal mean, which still divides by every pixel."""
    x = logits.astype(np.float64)
    n, c, h, w = x.shape
    npix = n * h * w
    if c > 1:
        p = _softmax(x)
        valid = target != -100
        lab = np.where(valid, target, 0)
        onehot = np.eye(c)[lab].transpose(0, 3, 1, 2)
        pt = np.take_along_axis(p, lab[:, None], axis=1)[:, 0]
        ce = -np.log(pt) * valid
        pt = np.where(valid, pt, 1.0)
        focal = (alpha * (1.0 - pt) ** gamma * ce).sum() / npix
        dfdce = alpha * ((1.0 - pt) ** gamma + gamma * (1.0 - pt) ** (gamma - 1.0) * pt * ce) * valid
        gfocal = (p - onehot) * (dfdce / npix)[:, None]
    else:
        pr = 1.0 / (1.0 + np.exp(-x[:, 0]))
        t = target.astype(np.float64)
        ce = -(t * np.maximum(np.log(pr), -100.0) + (1.0 - t) * np.maximum(np.log(1.0 - pr), -100.0))     # F.binary_cross_entropy
        pt = np.where(t == 1.0, pr, 1.0 - pr)
        dpt = np.where(t == 1.0, 1.0, -1.0)
        focal = (alpha * (1.0 - pt) ** gamma * ce).sum() / npix
        dce = (pr - t) / np.maximum((1.0 - pr) * pr, 1e-12)
        dfdp = alpha * ((1.0 - pt) ** gamma * dce - gamma * (1.0 - pt) ** (gamma - 1.0) * dpt * ce)
        gfocal = (dfdp * pr * (1.0 - pr) / npix)[:, None]
    dl, dg = dice(logits, target, smooth)
    return (1.0 - dice_weight) * focal + dice_weight * dl, (1.0 - dice_weight) * gfocal + dice_weight * dg


CRITERIA = {'ce': cross_entropy, 'dice': dice, 'focal_dice': focal_dice}


def resize_transpose(g, hl, wl):
    """Gradient of ``bilinear_ac(low, H, W)`` with respect to ``low`` given d loss / d output ``g`` [N,C,H,W] (float64)."""
    n, c, H, W = g.shape
    y0, y1, ly = fo._ac_coords(hl, H, np.float32)
    x0, x1, lx = fo._ac_coords(wl, W, np.float32)
    ly, lx = ly.astype(np.float64), lx.astype(np.float64)
    rows = np.zeros((n, c, hl, W))
    np.add.at(rows, (slice(None), slice(None), y0), g * (1.0 - ly)[None, None, :, None])
    np.add.at(rows, (slice(None), slice(None), y1), g * ly[None, None, :, None])
    out = np.zeros((n, c, hl, wl))
    np.add.at(out, (slice(None), slice(None), slice(None), x0), rows * (1.0 - lx)[None, None, None, :])
    np.add.at(out, (slice(None), slice(None), slice(None), x1), rows * lx[None, None, None, :])
    return out


def criterion_upsampled(kind, low, target, **kw):
    """criterion(F.interpolate(low, target.shape[1:], 'bilinear', align_corners=True), target) and its gradient w.r.t. ``low``."""
    H, W = target.shape[1:]
    full = fo.bilinear_ac(low.astype(np.float32), H, W)
    loss, g = CRITERIA[kind](full, target, **kw)
    return loss, resize_transpose(g, low.shape[2], low.shape[3])
