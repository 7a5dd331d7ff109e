"""Per-kernel device time of one training step (BASELINE config 5) from the CUPTI activity trace (torch.profiler: kernels run
back to back as in a real step, unlike the serialised cold-cache ncu launch list).
    python tools/train_kernel_times.py [batch=16] [crop=768] [steps=3] [fp32|tf32] [ohem|ce|dice|focal_dice] [classes=19] [sgd|adamw]"""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'fast-scnn-pytorch_b200'))
import torch
from torch.profiler import ProfilerActivity, profile

import bench
from fscnn_b200 import Trainer
from models.fast_scnn import FastSCNN

tb = int(sys.argv[1]) if len(sys.argv) > 1 else 16
crop = int(sys.argv[2]) if len(sys.argv) > 2 else 768
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
if len(sys.argv) > 4:
    from fscnn_b200 import train_ops
    train_ops.set_matmul_precision(sys.argv[4])
loss_type = sys.argv[5] if len(sys.argv) > 5 else 'ohem'
nc = int(sys.argv[6]) if len(sys.argv) > 6 else 19
optimizer = sys.argv[7] if len(sys.argv) > 7 else 'sgd'
dev = torch.device('cuda', 0)
m = FastSCNN(nc, aux=True).train()
bench.init_recipe_d2(m, 3)
m.to(dev)
tr = Trainer(m, loss_type=loss_type, optimizer=optimizer)
x = bench.smooth_images(tb, crop, crop, dev, 1, chunk=16)
if loss_type in ('dice', 'focal_dice'):      # lane-like binary labels (train.py's default setting)
    t = (torch.rand((tb, crop, crop), device=dev) < 0.1).long()
else:
    t = torch.randint(-1, nc, (tb, crop, crop), device=dev)
for _ in range(3):
    tr.step(x, t)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(steps):
        tr.step(x, t)
    torch.cuda.synchronize()
agg = collections.OrderedDict()
for ev in prof.events():
    if ev.device_type != torch.autograd.DeviceType.CUDA:
        continue
    name = ev.name.split('(')[0].replace('void ', '').replace('fscnn::', '')
    if name.startswith('Memset') or name.startswith('Memcpy'):
        name = name.split(' ')[0]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ev.device_time_total if hasattr(ev, 'device_time_total') else ev.cuda_time_total
tot = sum(a[1] for a in agg.values())
print(f'{steps} steps, batch {tb}, crop {crop}, {nc} classes, loss {loss_type}, {optimizer}: {tot / steps / 1e3:.2f} ms of kernel time per step')
print('| kernel | launches / step | us / step | us / launch | share |\n|---|---|---|---|---|')
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if a[1] / tot < 0.002:
        continue
    print(f'| `{k[:70]}` | {a[0] / steps:.0f} | {a[1] / steps:.1f} | {a[1] / a[0]:.1f} | {a[1] / tot:.1%} |')
