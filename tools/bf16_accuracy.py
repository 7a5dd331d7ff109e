"""GPU box utility: end-to-end accuracy of the bf16 path against the golden fixtures (logit error / absmax, mask mismatch)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ('fast-scnn-pytorch_b200', 'oracle', 'tests'):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np, torch
from helpers import build_model, load_case
dev = torch.device('cuda', 0)
for case in ['fwd_nc19_aux_n2_65x97', 'fwd_nc2_n1_360x640', 'fwd_nc19_n1_256x512', 'fwd_nc3_aux_n3_64x40']:
    g, sd, x, nc, aux = load_case(case)
    for prec in ('fp32', 'bf16'):
        model = build_model(sd, nc, aux, dev, precision=prec)
        xd = torch.from_numpy(x).to(dev)
        logits = model(xd)[0].cpu().numpy()
        scale = float(g['logits_absmax'])
        err = (np.abs(logits - g['logits']).max() if 'logits' in g.files else np.abs(logits[:, :, ::7, ::11] - g['logits_sample']).max()) / scale
        mask = model.predict(xd).cpu().numpy()
        print(f'{case:26s} {prec}: logits err/absmax {err:.3e}   mask mismatch {(mask != g["mask"]).mean():.4%}')
