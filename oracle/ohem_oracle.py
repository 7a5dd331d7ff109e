"""numpy restatement of SoftmaxCrossEntropyOHEMLoss.forward (reference utils/loss.py:143-182) and of the gradient
torch.autograd derives for it.  TEST INFRASTRUCTURE: pinned against fixtures produced by the unmodified reference
(oracle/gen_golden_train.py -> tests/golden/train_ohem_*.npz, tests/test_oracle_golden.py); only tests/ import it."""
import numpy as np


def ohem_select(logits, target, ignore_label=-1, thresh=0.7, min_kept=256):
    """Returns (kept flag [N,H,W], threshold actually used or None when every valid pixel is kept).  loss.py:150-175."""
    n, c, h, w = logits.shape
    label = target.ravel().astype(np.int32)
    x = np.rollaxis(logits, 1).reshape((c, -1))                      # loss.py:152
    prob = np.exp(x - x.max(axis=0).reshape((1, -1)))                # loss.py:153
    prob /= prob.sum(axis=0).reshape((1, -1))                        # loss.py:154
    valid = label != ignore_label
    num_valid = int(valid.sum())
    kept = valid.copy()
    threshold = None
    if min_kept < num_valid and num_valid > 0:                       # loss.py:160-175
        pred = prob[label[valid], np.arange(label.size)[valid]]
        threshold = np.float32(thresh)
        if min_kept > 0:
            kth = np.sort(pred, kind='stable')[min(pred.size, min_kept) - 1]
            if kth > thresh:
                threshold = kth
        kept[valid] = pred <= threshold
    return kept.reshape(n, h, w), threshold


def ohem_loss_and_grad(logits, target, weight=None, ignore_label=-1, thresh=0.7, min_kept=256):
    """(loss, d loss / d logits): nn.CrossEntropyLoss(weight, ignore_index) over the kept pixels (loss.py:135-138, :182),
    weighted mean reduction."""
    n, c, h, w = logits.shape
    kept, _ = ohem_select(logits, target, ignore_label, thresh, min_kept)
    x = logits.astype(np.float64)
    x = x - x.max(axis=1, keepdims=True)
    logsm = x - np.log(np.exp(x).sum(axis=1, keepdims=True))
    lab = np.where(kept, target, 0)
    wpix = (np.ones(c) if weight is None or len(weight) == 0 else np.asarray(weight, np.float64))[lab] * kept
    nll = -np.take_along_axis(logsm, lab[:, None], axis=1)[:, 0]
    wsum = wpix.sum()
    loss = (wpix * nll).sum() / wsum
    onehot = np.eye(c)[lab].transpose(0, 3, 1, 2)
    grad = (np.exp(logsm) - onehot) * (wpix / wsum)[:, None]
    return np.float32(loss), grad.astype(np.float32)
