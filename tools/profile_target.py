"""One warm-up + one profiled pass of the hot path, for ncu:  python tools/profile_target.py [fp32|bf16] [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'fast-scnn-pytorch_b200'))
import torch
from models.fast_scnn import FastSCNN

prec = sys.argv[1] if len(sys.argv) > 1 else 'bf16'
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device('cuda', 0)
torch.manual_seed(0)
model = FastSCNN(19, precision=prec).eval().to(dev)
x = torch.randn(batch, 3, 1024, 2048, device=dev)
labels = torch.randint(-1, 19, (batch, 1024, 2048), device=dev)
conf = torch.zeros(402, dtype=torch.int64, device=dev)
for _ in range(2):
    model.evaluate(x, labels, conf=conf)
torch.cuda.synchronize()
print('done', conf[-2:].tolist())
