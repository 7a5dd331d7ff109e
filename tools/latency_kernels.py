"""Where the batch-1 latency goes: per-kernel device time and the gaps between kernels of one CUDA-graph replay of
FastSCNN.predict on a 1 x 3 x 1024 x 2048 image (CUPTI activity trace through torch.profiler).
    python tools/latency_kernels.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'fast-scnn-pytorch_b200'))
import torch
from torch.profiler import ProfilerActivity, profile

import bench
from models.fast_scnn import FastSCNN

dev = torch.device('cuda', 0)
m = FastSCNN(19, precision='bf16').eval()
bench.init_recipe_d2(m, 3)
m.to(dev)
x = bench.smooth_images(1, 1024, 2048, dev, 1, chunk=1)
mask = torch.empty((1, 1024, 2048), dtype=torch.uint8, device=dev)
for _ in range(3):
    m.predict(x, out=mask)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    m.predict(x, out=mask)
for _ in range(50):
    g.replay()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
ev = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
per = len(ev) // 5
last = ev[-per:]
t0 = last[0].time_range.start
busy = 0.0
print('| kernel | start us | duration us | gap before us |\n|---|---|---|---|')
prev_end = None
for e in last:
    s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
    gap = (e.time_range.start - prev_end) if prev_end is not None else 0.0
    prev_end = e.time_range.end
    busy += d
    print(f"| `{e.name.split('(')[0].replace('void ', '').replace('fscnn::', '')[:60]}` | {s:.1f} | {d:.1f} | {gap:.1f} |")
print(f'{per} kernels, span {prev_end - t0:.1f} us, busy {busy:.1f} us')
