"""Times the training step of BASELINE config 5 (16 x 3 x 768 x 768, aux head, OHEM mix loss, SGD) with CUDA events.
    python tools/train_bench.py [batch=16] [crop=768] [reps=10] [fp32|tf32] [graph]
Prints: ms per step (device), images / s, the host time to enqueue one step, last loss."""
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
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
math = sys.argv[4] if len(sys.argv) > 4 else 'fp32'
graph = len(sys.argv) > 5 and sys.argv[5] == 'graph'
dev = torch.device('cuda', 0)
m = FastSCNN(19, aux=True).train()
bench.init_recipe_d2(m, 3)
m.to(dev)
tr = Trainer(m, cuda_graph=graph, matmul_precision=math)
x = bench.smooth_images(tb, crop, crop, dev, 1, chunk=16)
t = torch.randint(-1, 19, (tb, crop, crop), device=dev)
for _ in range(5):
    loss = tr.step(x, t)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
import time
t0 = time.perf_counter()
e0.record()
for _ in range(reps):
    loss = tr.step(x, t)
e1.record()
host_ms = (time.perf_counter() - t0) * 1e3 / reps      # time the host needs to enqueue one step
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f'{math}{" graph" if graph else ""}: {ms:.3f} ms/step  {tb / ms * 1e3:.1f} images/s  host enqueue {host_ms:.3f} ms/step  loss {float(loss):.4f}')
