"""Prints per-stage isolation errors (stage fed with the reference's tensor) and end-to-end errors of the
CUDA path against the golden fixtures, for a given precision.  GPU box utility:
    python tools/stage_report.py [fp32|bf16] [case]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ('fast-scnn-pytorch_b200', 'oracle', 'tests'):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np
import torch

from helpers import build_model, load_case, rel_err
from test_gpu_parity import STAGE_IO, nhwc

prec = sys.argv[1] if len(sys.argv) > 1 else 'bf16'
case = sys.argv[2] if len(sys.argv) > 2 else 'fwd_nc19_aux_n2_65x97'
dev = torch.device('cuda', 0)
g, sd, x, nc, aux = load_case(case)
model = build_model(sd, nc, aux, dev, precision=prec)
xd = torch.from_numpy(x).to(dev)
eng = model._engine(dev)
n, _, h, w = x.shape
names = eng.stage_names()
print(f'precision {prec}, case {case}')
for stage, ins, out in STAGE_IO:
    if stage not in names or ('tap/' + out) not in g.files:
        continue
    idx = names.index(stage)
    for tap in ins or []:
        v = eng.tap_view(tap, n, h, w)
        v.copy_(nhwc(g['tap/' + tap]).to(dev).to(v.dtype))
    eng.forward_range(xd, idx, idx)
    torch.cuda.synchronize()
    got = eng.tap_view(out, n, h, w).permute(0, 3, 1, 2).float().cpu().numpy()
    print(f'  isolated {stage:20s} rel err {rel_err(got, g["tap/" + out]):.3e}  nan={np.isnan(got).any()}')
outs = model(xd)
logits = outs[0].cpu().numpy()
if 'logits' in g.files:
    print(f'end-to-end logits rel err {rel_err(logits, g["logits"]):.3e}')
for key in g.files:
    if key.startswith('tap/'):
        got = eng.tap_view(key[4:], n, h, w).permute(0, 3, 1, 2).float().cpu().numpy() if key[4:] != 'cls.dsconv2' else None
        if got is not None:
            print(f'  chained  {key[4:]:20s} rel err {rel_err(got, g[key]):.3e}')
mask = model.predict(xd).cpu().numpy()
print(f'mask disagreement vs reference: {(mask != g["mask"]).mean():.4%}')
