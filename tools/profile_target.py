"""One warm-up + one profiled pass of the hot path, for ncu:  python tools/profile_target.py [fp32|bf16] [batch]
Recipe-D2 weights and inputs (class regions and boundaries: what the data-dependent tail kernel has to be profiled on)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'fast-scnn-pytorch_b200'))
import torch

import bench
from models.fast_scnn import FastSCNN

prec = sys.argv[1] if len(sys.argv) > 1 else 'bf16'
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 8
nc, h, w = 19, 1024, 2048
dev = torch.device('cuda', 0)
model = FastSCNN(nc, precision=prec).eval()
bench.init_recipe_d2(model, 7)
model.to(dev)
x = bench.smooth_images(batch, h, w, dev, 1234)
eng = model._engine(dev)
with torch.no_grad():
    low = bench.lowres_logits(eng, x[:min(batch, 8)].contiguous(), h, w)
    model.classifier.conv[1].bias -= low[..., :nc].mean(dim=(0, 1, 2))
labels = torch.randint(-1, nc, (batch, h, w), device=dev)
conf = torch.zeros(model._engine(dev).conf_len(), dtype=torch.int64, device=dev)
for _ in range(2):
    model.evaluate(x, labels, conf=conf)
torch.cuda.synchronize()
print('done', conf[-2:].tolist())
