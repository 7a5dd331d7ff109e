"""Golden vectors for the reference's cross-entropy / dice / focal + dice criteria (reference
utils/loss.py:12-124), produced by the UNMODIFIED reference classes under torch.autograd on the CPU in the build container
(TEST INFRASTRUCTURE; /root/reference does not exist on the GPU box, so the vectors are committed together with this script):

    python oracle/gen_golden_loss.py        ->  tests/golden/train_loss_cases.npz

Every case holds seeded logits (one or two heads) and labels, the reference's loss and its gradient with respect to every head.
'*_low' cases feed the criterion with F.interpolate(low, size, 'bilinear', align_corners=True) of LOW-RESOLUTION logits -- what
models/fast_scnn.py:40 / :44 do in front of it -- and record the gradient with respect to the low-resolution tensors."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

REF = '/root/reference'
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
from utils.loss import DiceLoss, FocalDiceLoss, MixDiceLoss, MixSoftmaxCrossEntropyLoss  # noqa: E402


def case(out, name, crit, heads_shape, target, seed, size=None, single=False):
    g = torch.Generator().manual_seed(seed)
    heads = [(torch.randn(s, generator=g) * 2.0).requires_grad_(True) for s in heads_shape]
    preds = [F.interpolate(t, size, mode='bilinear', align_corners=True) if size else t for t in heads]
    loss = crit(preds[0], target) if single else crit(tuple(preds), target)
    loss.backward()
    out[name + '/target'] = target.numpy()
    out[name + '/loss'] = np.float64(loss.item())
    for i, t in enumerate(heads):
        out[f'{name}/logits{i}'] = t.detach().numpy()
        out[f'{name}/grad{i}'] = t.grad.numpy()
    print(f'{name}: loss {loss.item():.6f}')


def main():
    out = {}
    g = torch.Generator().manual_seed(5)
    n, h, w = 2, 32, 48
    lane = (torch.rand((n, h, w), generator=g) < 0.2).long()                       # binary lane labels (train.py's dice setting)
    lab19 = torch.randint(-1, 19, (n, h, w), generator=g)                          # -1 = ignore (train.py:190-191)
    lab3 = torch.randint(0, 3, (n, h, w), generator=g)
    lab3_ign = lab3.clone()
    lab3_ign[torch.rand((n, h, w), generator=g) < 0.1] = -100                       # F.cross_entropy's default ignore_index
    full2, low2 = (n, 2, h, w), (n, 2, h // 8, w // 8)
    # DiceLoss / MixDiceLoss (loss.py:12-68)
    case(out, 'dice_c2', DiceLoss(), [full2], lane, 11, single=True)
    case(out, 'dice_c1_sigmoid', DiceLoss(), [(n, 1, h, w)], lane, 12, single=True)
    case(out, 'dice_smooth1', DiceLoss(smooth=1.0), [full2], lane, 13, single=True)
    case(out, 'mixdice_aux', MixDiceLoss(aux=True, aux_weight=0.4), [full2, full2], lane, 14)
    case(out, 'mixdice_aux_low', MixDiceLoss(aux=True, aux_weight=0.4), [low2, low2], lane, 15, size=(h, w))
    case(out, 'dice_c19_labels', DiceLoss(), [(n, 19, h, w)], lab19, 16, single=True)      # t = float(label), -1 included
    # MixSoftmaxCrossEntropyLoss (loss.py:103-124)
    case(out, 'ce_c19_aux', MixSoftmaxCrossEntropyLoss(aux=True, aux_weight=0.4), [(n, 19, h, w)] * 2, lab19, 21)
    case(out, 'ce_c19_aux_low', MixSoftmaxCrossEntropyLoss(aux=True, aux_weight=0.4), [(n, 19, h // 8, w // 8)] * 2, lab19, 22, size=(h, w))
    case(out, 'ce_c2_noaux', MixSoftmaxCrossEntropyLoss(aux=False), [full2], lane, 23)
    case(out, 'ce_c5_low_odd', MixSoftmaxCrossEntropyLoss(aux=False), [(n, 5, 7, 9)], torch.randint(-1, 5, (n, 51, 67), generator=g), 24,
         size=(51, 67))
    # FocalDiceLoss (loss.py:71-100)
    case(out, 'focal_c2', FocalDiceLoss(), [full2], lane, 31, single=True)
    case(out, 'focal_c2_low', FocalDiceLoss(), [low2], lane, 32, size=(h, w), single=True)
    case(out, 'focal_c3_ignore100', FocalDiceLoss(alpha=0.25, gamma=3.0, dice_weight=0.3), [(n, 3, h, w)], lab3_ign, 33, single=True)
    case(out, 'focal_c1_bce', FocalDiceLoss(), [(n, 1, h, w)], lane, 34, single=True)
    np.savez_compressed(os.path.join(OUT, 'train_loss_cases.npz'), **out)


if __name__ == '__main__':
    main()
