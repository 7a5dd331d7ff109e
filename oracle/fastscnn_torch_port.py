"""CPU baseline port of the reference path, on the SAME vendor kernels the reference dispatches.
TEST / BENCH INFRASTRUCTURE ONLY.

The reference's arithmetic is ATen/oneDNN behind ``nn.Conv2d`` / ``nn.BatchNorm2d`` /
``F.interpolate`` / ``nn.AdaptiveAvgPool2d`` / ``torch.argmax``, followed by
numpy histograms for the metric.  /root/reference does not exist on the GPU box, so this file
re-expresses ``FastSCNN.forward`` (models/fast_scnn.py:33-237) functionally over a plain
state_dict with ``torch.nn.functional`` calls -- the same ATen ops, in the same order, with
unfolded BatchNorm -- which makes it the fairest stand-in for "the reference's CPU path" when the
GPU numbers are reported (bench.py: ``cpu_baseline`` with kind "port", and ``--impl reference``).
It is pinned against the reference-generated golden vectors by tests/test_oracle_golden.py.
Only tests/ and bench.py may import it; the product never does.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from fastscnn_oracle import BOTTLENECK_PLAN


def _bn(sd, p, x):
    return F.batch_norm(x, sd[p + '.running_mean'], sd[p + '.running_var'], sd[p + '.weight'], sd[p + '.bias'], False, 0.1, 1e-5)


def _cbr(sd, p, x, stride=1, padding=0, groups=1):
    return F.relu(_bn(sd, p + '.1', F.conv2d(x, sd[p + '.0.weight'], None, stride, padding, 1, groups)))


def _dsconv(sd, p, x, stride):
    c = x.shape[1]
    x = F.relu(_bn(sd, p + '.conv.1', F.conv2d(x, sd[p + '.conv.0.weight'], None, stride, 1, 1, c)))
    return F.relu(_bn(sd, p + '.conv.4', F.conv2d(x, sd[p + '.conv.3.weight'])))


def _bottleneck(sd, p, x, stride, cout):
    y = _cbr(sd, p + '.block.0.conv', x)
    y = _cbr(sd, p + '.block.1.conv', y, stride, 1, y.shape[1])
    y = _bn(sd, p + '.block.3', F.conv2d(y, sd[p + '.block.2.weight']))
    return x + y if (stride == 1 and x.shape[1] == cout) else y


def _up(x, size):
    return F.interpolate(x, size, mode='bilinear', align_corners=True)


@torch.no_grad()
def forward(sd, x, aux=False):
    """sd: name -> torch tensor (reference key names); x: [N,3,H,W] fp32.  Returns the reference's tuple."""
    size = x.shape[2:]
    p = 'learning_to_downsample'
    t = _cbr(sd, p + '.conv.conv', x, 2, 0)
    t = _dsconv(sd, p + '.dsconv1', t, 2)
    higher = _dsconv(sd, p + '.dsconv2', t, 2)
    p = 'global_feature_extractor'
    t = higher
    for name, cout, stride in BOTTLENECK_PLAN:
        t = _bottleneck(sd, f'{p}.{name}', t, stride, cout)
    hw = t.shape[2:]
    feats = [t] + [_up(_cbr(sd, f'{p}.ppm.conv{i}.conv', F.adaptive_avg_pool2d(t, s)), hw)
                   for i, s in enumerate((1, 2, 3, 6), start=1)]
    t = _cbr(sd, p + '.ppm.out.conv', torch.cat(feats, 1))
    p = 'feature_fusion'
    low = _cbr(sd, p + '.dwconv.conv', _up(t, higher.shape[2:]), 1, 1, 128)
    low = _bn(sd, p + '.conv_lower_res.1', F.conv2d(low, sd[p + '.conv_lower_res.0.weight'], sd[p + '.conv_lower_res.0.bias']))
    hi = _bn(sd, p + '.conv_higher_res.1', F.conv2d(higher, sd[p + '.conv_higher_res.0.weight'], sd[p + '.conv_higher_res.0.bias']))
    t = F.relu(hi + low)
    t = _dsconv(sd, 'classifier.dsconv1', t, 1)
    t = _dsconv(sd, 'classifier.dsconv2', t, 1)
    outs = [_up(F.conv2d(t, sd['classifier.conv.1.weight'], sd['classifier.conv.1.bias']), size)]
    if aux:
        a = F.relu(_bn(sd, 'auxlayer.1', F.conv2d(higher, sd['auxlayer.0.weight'], None, 1, 1)))
        outs.append(_up(F.conv2d(a, sd['auxlayer.4.weight'], sd['auxlayer.4.bias']), size))
    return tuple(outs)


@torch.no_grad()
def eval_step(sd, x, labels, nclass, metric):
    """One pass of the reference's eval loop body (eval.py:43-49): forward, argmax, host metric update."""
    pred = torch.argmax(forward(sd, x)[0], 1)
    metric.update(pred.numpy(), labels)
    return pred


@torch.no_grad()
def forward_bf16_autocast(sd, x):
    """The reference's OWN reduced-precision behaviour on this input: the same op sequence under
    ``torch.autocast(bfloat16)`` on the CPU (ATen casts the convolutions to bf16, keeps BatchNorm / interpolation
    accumulation in fp32).  Returns float32 logits.  This is the yardstick of the bf16 path's tolerance
    (tests/test_gpu_bf16.py, __graft_entry__.smoke): how far does the reference itself move when it runs in bf16?"""
    with torch.autocast('cpu', dtype=torch.bfloat16):
        out = forward(sd, x)
    return tuple(o.float() for o in out)


def bf16_yardstick(sd_np, x_np, ref_logits=None):
    """(max, rms) error / absmax and mask disagreement of the reference's bf16 autocast against its fp32 run, both
    through this port, for a numpy state_dict and input.  ``ref_logits`` (numpy, fp32) may be passed to save the
    fp32 run."""
    sd = to_torch_state_dict(sd_np)
    x = torch.from_numpy(np.asarray(x_np))
    ref = torch.from_numpy(np.asarray(ref_logits)) if ref_logits is not None else forward(sd, x)[0]
    low = forward_bf16_autocast(sd, x)[0]
    d = (low - ref).abs()
    scale = float(ref.abs().max())
    return {'max': float(d.max()) / scale, 'rms': float(d.pow(2).mean().sqrt()) / scale,
            'mask': float((low.argmax(1) != ref.argmax(1)).float().mean())}


def to_torch_state_dict(sd_np):
    return {k: torch.from_numpy(np.asarray(v)) for k, v in sd_np.items()}
