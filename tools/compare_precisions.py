"""GPU box utility: run the fp32 and the bf16 engine on the same input and print, per stage tensor, the relative
error of the bf16 one (max |diff| / max |fp32|) -- localises a broken bf16 kernel on shapes that have no golden taps.
    python tools/compare_precisions.py [case | HxW] [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ('fast-scnn-pytorch_b200', 'oracle', 'tests'):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np
import torch

from helpers import build_model, load_case
from test_gpu_parity import STAGE_IO

arg = sys.argv[1] if len(sys.argv) > 1 else 'fwd_nc2_n1_360x640'
dev = torch.device('cuda', 0)
if 'x' in arg and arg[0].isdigit():
    import fastscnn_oracle as fo
    h, w = (int(v) for v in arg.split('x'))
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    nc, aux = 19, False
    sd = fo.make_state_dict(nc, aux, seed=3)
    x = fo.make_input(n, h, w, seed=5)
else:
    g, sd, x, nc, aux = load_case(arg)
n, _, h, w = x.shape
xd = torch.from_numpy(x).to(dev)
taps = {}
for prec in ('fp32', 'bf16'):
    model = build_model(sd, nc, aux, dev, precision=prec)
    eng = model._engine(dev)
    names = eng.stage_names()
    eng.forward_range(xd, 0, len(names) - 1)
    torch.cuda.synchronize()
    for stage, ins, out in STAGE_IO:
        if stage in names:
            try:
                taps[(prec, out)] = eng.tap_view(out, n, h, w).float().cpu().numpy().copy()
            except Exception as e:  # noqa: BLE001
                pass
for stage, ins, out in STAGE_IO:
    a, b = taps.get(('fp32', out)), taps.get(('bf16', out))
    if a is None or b is None or a.shape != b.shape:
        continue
    d = np.abs(a - b)
    idx = np.unravel_index(np.argmax(d), d.shape)
    print(f'{stage:28s} {out:24s} rel err {d.max() / max(np.abs(a).max(), 1e-30):9.3e}  nan {int(np.isnan(b).sum()):6d}  worst at {idx} of {a.shape}')
if len(sys.argv) > 3 or os.environ.get('MAP'):
    out = os.environ.get('MAP', 'gfe.bottleneck1.1')
    a, b = taps[('fp32', out)], taps[('bf16', out)]
    e = np.abs(a - b).max(axis=-1)[0] / np.abs(a).max()
    print('error map of', out, '(. < 2%, + < 10%, # >= 10%)')
    for row in e:
        print(''.join('.' if v < 0.02 else ('+' if v < 0.1 else '#') for v in row))
    ec = np.abs(a - b)[0].reshape(-1, a.shape[-1]).max(axis=0) / np.abs(a).max()
    print('per-channel max error:', ' '.join(f'{v:.2f}' for v in ec))
