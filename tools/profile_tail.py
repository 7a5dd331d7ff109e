"""ncu target for the tail kernel alone: recipe-D2 network logits of `batch` 1024x2048 images, then two launches of
fscnn_upsample_argmax in metric mode (int64 labels).    python tools/profile_tail.py [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'fast-scnn-pytorch_b200'))
import torch

import bench
from models.fast_scnn import FastSCNN

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 37
nc, h, w = 19, 1024, 2048
dev = torch.device('cuda', 0)
model = FastSCNN(nc, precision='bf16').eval()
bench.init_recipe_d2(model, 7)
model.to(dev)
x = bench.smooth_images(batch, h, w, dev, 1234)
eng = model._engine(dev)
with torch.no_grad():
    low = bench.lowres_logits(eng, x[:8].contiguous(), h, w)
    model.classifier.conv[1].bias -= low[..., :nc].mean(dim=(0, 1, 2))
eng = model._engine(dev)
low = bench.lowres_logits(eng, x, h, w)
if len(sys.argv) > 2 and sys.argv[2] == 'tied':
    low.zero_()
labels = torch.randint(-1, nc, (batch, h, w), device=dev)
conf = torch.zeros(eng.conf_len(), dtype=torch.int64, device=dev)
for _ in range(2):
    eng.upsample_argmax(low, h, w, labels=labels, conf=conf, want_mask=False)
torch.cuda.synchronize()
print('done', conf[-2:].tolist())
