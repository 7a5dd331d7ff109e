"""Shared helpers for the parity tests."""
import os

import numpy as np

import fastscnn_oracle as fo

from conftest import GOLDEN


def load_case(name):
    g = np.load(os.path.join(GOLDEN, name + '.npz'))
    nc, aux, n, h, w, wseed, xseed = (int(v) for v in g['meta'])
    sd = fo.make_state_dict(nc, bool(aux), wseed)
    sd['classifier.conv.1.bias'] = g['cls_bias']
    x = fo.make_input(n, h, w, xseed)
    return g, sd, x, nc, bool(aux)


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def torch_state_dict(sd):
    import torch
    return {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}


def build_model(sd, nc, aux, device, precision='fp32'):
    from models.fast_scnn import FastSCNN
    model = FastSCNN(nc, aux=aux, precision=precision).eval()
    model.load_state_dict(torch_state_dict(sd))
    return model.to(device)


# ---- the bf16 path's stated tolerance, anchored on the reference's own bf16 behaviour ------------------------------------
# Measured on a B200 (profiles/r02_bf16_error_budget.md): the path's logit error is the sum of ~45 independent bf16
# roundings (stage tensors, expanded / depthwise tiles, weights) spread evenly over the stages, and its rms equals that
# of the reference run under torch.autocast(bfloat16) on the same input (ratio 0.74-0.99 over six configurations, CPU and
# cuDNN autocast alike), with FEWER differing mask pixels in all six.  The max over 1e5-1e7 logits of that heavy-tailed
# error is a noisy statistic: ours / reference ranged 0.59-1.73 at equal rms.  The sharp criterion is therefore the rms,
# the max gets the slack its variance needs, and fixed backstops guard against both being bad together.
BF16_RMS_VS_REF, BF16_MAX_VS_REF, BF16_MASK_SLACK = 1.25, 2.0, 0.005
BF16_MAX_BACKSTOP, BF16_MASK_BACKSTOP = 6e-2, 5e-2


def bf16_errors(got_logits, ref_logits):
    d = np.abs(np.asarray(got_logits, np.float64) - np.asarray(ref_logits, np.float64))
    scale = float(np.abs(ref_logits).max())
    return {'max': float(d.max()) / scale, 'rms': float(np.sqrt((d * d).mean())) / scale,
            'mask': float((np.argmax(got_logits, 1) != np.argmax(ref_logits, 1)).mean())}


def check_bf16_against_yardstick(ours, yard):
    """ours / yard: dicts from bf16_errors / fastscnn_torch_port.bf16_yardstick.  Returns a list of violated criteria."""
    bad = []
    if not ours['rms'] <= BF16_RMS_VS_REF * yard['rms']:
        bad.append(f"rms {ours['rms']:.3e} > {BF16_RMS_VS_REF} x reference bf16 rms {yard['rms']:.3e}")
    if not ours['max'] <= min(BF16_MAX_VS_REF * yard['max'], BF16_MAX_BACKSTOP):
        bad.append(f"max {ours['max']:.3e} > min({BF16_MAX_VS_REF} x reference bf16 max {yard['max']:.3e}, {BF16_MAX_BACKSTOP})")
    if not ours['mask'] <= min(yard['mask'] + BF16_MASK_SLACK, BF16_MASK_BACKSTOP):
        bad.append(f"mask disagreement {ours['mask']:.3%} > reference bf16 {yard['mask']:.3%} + {BF16_MASK_SLACK:.1%}")
    return bad


# tests/golden/train_loss_cases.npz (oracle/gen_golden_loss.py): criterion, its keyword arguments, weight of the second head
LOSS_CASES = {
    'dice_c2': ('dice', {}, None),
    'dice_c1_sigmoid': ('dice', {}, None),
    'dice_smooth1': ('dice', {'smooth': 1.0}, None),
    'mixdice_aux': ('dice', {}, 0.4),
    'mixdice_aux_low': ('dice', {}, 0.4),
    'dice_c19_labels': ('dice', {}, None),
    'ce_c19_aux': ('ce', {'ignore_label': -1}, 0.4),
    'ce_c19_aux_low': ('ce', {'ignore_label': -1}, 0.4),
    'ce_c2_noaux': ('ce', {'ignore_label': -1}, None),
    'ce_c5_low_odd': ('ce', {'ignore_label': -1}, None),
    'focal_c2': ('focal_dice', {}, None),
    'focal_c2_low': ('focal_dice', {}, None),
    'focal_c3_ignore100': ('focal_dice', {'alpha': 0.25, 'gamma': 3.0, 'dice_weight': 0.3}, None),
    'focal_c1_bce': ('focal_dice', {}, None),
}
