"""ncu target: one warm-up + one profiled training step (BASELINE config 5 shapes by default).
    python tools/profile_train.py [batch=16] [crop=768]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'fast-scnn-pytorch_b200'))
import torch

import bench
from fscnn_b200 import Trainer
from models.fast_scnn import FastSCNN

tb = int(sys.argv[1]) if len(sys.argv) > 1 else 16
crop = int(sys.argv[2]) if len(sys.argv) > 2 else 768
dev = torch.device('cuda', 0)
m = FastSCNN(19, aux=True).train()
bench.init_recipe_d2(m, 3)
m.to(dev)
tr = Trainer(m)
x = bench.smooth_images(tb, crop, crop, dev, 1, chunk=16)
t = torch.randint(-1, 19, (tb, crop, crop), device=dev)
for _ in range(2):
    loss = tr.step(x, t)
torch.cuda.synchronize()
print('loss', float(loss))
