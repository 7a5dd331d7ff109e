"""GPU box utility: where does the bf16 path's logit error come from?

For one seeded configuration (the smoke() input by default) it runs the fp32 engine (the exactness path, within 1e-5 of the
oracle) and the bf16 engine and prints, per stage:
  isolated : the bf16 stage fed with the fp32 stage input  -> error of that stage alone (max and rms, / absmax of the tensor)
  chained  : the bf16 pipeline's tensor at that point      -> accumulated error
  tail-from: low-res logit error when only the stages from this one on run in bf16 (everything before it exact)
and, for scale, the error of torch's own bf16 autocast (cuDNN eager on the same GPU) on the same input.

    python tools/bf16_chain_report.py [nc n h w wseed xseed]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ('fast-scnn-pytorch_b200', 'oracle', 'tests'):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np
import torch

import fastscnn_oracle as fo
import fastscnn_torch_port as port
from helpers import build_model
from test_gpu_parity import STAGE_IO

args = [int(v) for v in sys.argv[1:7]]
nc, n, h, w, wseed, xseed = args + [19, 2, 96, 160, 7, 21][len(args):]
dev = torch.device('cuda', 0)
sd = fo.make_state_dict(nc, False, seed=wseed)
x = fo.make_input(n, h, w, seed=xseed)
xd = torch.from_numpy(x).to(dev)
# recipe D2's calibration, done with the fp32 engine (the oracle is slow at large sizes)
m32 = build_model(sd, nc, False, dev, precision='fp32')
e32 = m32._engine(dev)
names = e32.stage_names()
last = names.index('cls.dsconv2+head')
e32.forward_range(xd, 0, last)
low = e32.tap_view('cls.logits_lowres', n, h, w).float()
sd['classifier.conv.1.bias'] = (sd['classifier.conv.1.bias'] - low.mean(dim=(0, 1, 2)).cpu().numpy()).astype(np.float32)
m32 = build_model(sd, nc, False, dev, precision='fp32')
e32 = m32._engine(dev)
m16 = build_model(sd, nc, False, dev, precision='bf16')
e16 = m16._engine(dev)
for kv in os.environ.get('OPTS', '').split(','):
    if '=' in kv:
        k, v = kv.split('=')
        e16.set_option(k, int(v))
stages = [s for s in STAGE_IO if s[0] in names and s[0] != 'aux']
e32.forward_range(xd, 0, last)
ref = {out: e32.tap_view(out, n, h, w).float().clone() for _, _, out in stages}
e16.forward_range(xd, 0, last)
chained = {out: e16.tap_view(out, n, h, w).float().clone() for _, _, out in stages}


def errs(a, b):
    d = (a - b).abs()
    s = b.abs().max().item()
    return d.max().item() / s, d.pow(2).mean().sqrt().item() / s


print(f'nc {nc}  {n}x{h}x{w}  weights seed {wseed}  input seed {xseed}   OPTS={os.environ.get("OPTS", "")}')
print(f'{"stage":22s} {"isolated max":>12s} {"rms":>9s} | {"chained max":>11s} {"rms":>9s} | {"tail-from max":>13s} {"rms":>9s}')
for stage, ins, out in stages:
    idx = names.index(stage)
    # isolated
    for tap in ins or []:
        v = e16.tap_view(tap, n, h, w)
        v.copy_(ref[tap].to(v.dtype))
    e16.forward_range(xd, idx, idx)
    iso = errs(e16.tap_view(out, n, h, w).float(), ref[out])
    # tail-from: all fp32 taps before this stage, bf16 from here on
    for s2, _, o2 in stages:
        if names.index(s2) < idx:
            v = e16.tap_view(o2, n, h, w)
            v.copy_(ref[o2].to(v.dtype))
    e16.forward_range(xd, idx, last)
    tail = errs(e16.tap_view('cls.logits_lowres', n, h, w).float(), ref['cls.logits_lowres'])
    ch = errs(chained[out], ref[out])
    print(f'{stage:22s} {iso[0]:12.3e} {iso[1]:9.2e} | {ch[0]:11.3e} {ch[1]:9.2e} | {tail[0]:13.3e} {tail[1]:9.2e}')

full32 = m32(xd)[0]
full16 = m16(xd)[0]
e = errs(full16, full32)
mask32, mask16 = full32.argmax(1), full16.argmax(1)
print(f'full-res logits: bf16 path vs fp32 path  max {e[0]:.3e} rms {e[1]:.2e}   mask diff {(mask32 != mask16).float().mean().item():.3%}')
tsd = {k: v.to(dev) for k, v in port.to_torch_state_dict(sd).items()}
eager32 = port.forward(tsd, xd)[0]
e = errs(full32, eager32)
print(f'fp32 path vs cuDNN eager fp32            max {e[0]:.3e} rms {e[1]:.2e}')
with torch.autocast('cuda', dtype=torch.bfloat16):
    eager16 = port.forward(tsd, xd)[0].float()
e = errs(eager16, eager32)
print(f'cuDNN eager bf16 autocast vs eager fp32  max {e[0]:.3e} rms {e[1]:.2e}   mask diff {(eager16.argmax(1) != eager32.argmax(1)).float().mean().item():.3%}')
x16 = xd.cpu()
csd = port.to_torch_state_dict(sd)
with torch.autocast('cpu', dtype=torch.bfloat16):
    cpu16 = port.forward(csd, x16)[0].float()
e = errs(cpu16.to(dev), eager32)
print(f'CPU bf16 autocast vs eager fp32          max {e[0]:.3e} rms {e[1]:.2e}   mask diff {(cpu16.to(dev).argmax(1) != eager32.argmax(1)).float().mean().item():.3%}')
