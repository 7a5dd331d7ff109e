"""Per-shape device time of the pointwise / dense-3x3 contractions inside one training step (BASELINE config 5): wraps the two C-ABI
entry points with CUDA events.  Columns: the GEMM shape, calls per step, ms, achieved TFLOP/s and GB/s on the algorithmic bytes.
    python tools/train_gemm_shapes.py [fp32|tf32] [batch=16] [crop=768]"""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'fast-scnn-pytorch_b200'))
import torch

import bench
from fscnn_b200 import Trainer, native
from models.fast_scnn import FastSCNN

math = sys.argv[1] if len(sys.argv) > 1 else 'fp32'
tb = int(sys.argv[2]) if len(sys.argv) > 2 else 16
crop = int(sys.argv[3]) if len(sys.argv) > 3 else 768
dev = torch.device('cuda', 0)
m = FastSCNN(19, aux=True).train()
bench.init_recipe_d2(m, 3)
m.to(dev)
tr = Trainer(m, matmul_precision=math)
x = bench.smooth_images(tb, crop, crop, dev, 1, chunk=16)
t = torch.randint(-1, 19, (tb, crop, crop), device=dev)
for _ in range(2):
    tr.step(x, t)
torch.cuda.synchronize()
lib = native.lib()
records = []
real_f, real_b = lib.fscnn_train_pwconv_forward, lib.fscnn_train_pwconv_backward


def fwd(xp, wp, yp, n, cin, cout, hw, s):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = real_f(xp, wp, yp, n, cin, cout, hw, s)
    e1.record()
    records.append((('fwd', n, cin, cout, hw), e0, e1))
    return rc


def bwd(xp, wp, dyp, dxp, dwp, ws, wsb, n, cin, cout, hw, s):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = real_b(xp, wp, dyp, dxp, dwp, ws, wsb, n, cin, cout, hw, s)
    e1.record()
    records.append((('bwd' + ('' if dxp else '_nodx'), n, cin, cout, hw), e0, e1))
    return rc


lib.fscnn_train_pwconv_forward, lib.fscnn_train_pwconv_backward = fwd, bwd
tr.step(x, t)
torch.cuda.synchronize()
lib.fscnn_train_pwconv_forward, lib.fscnn_train_pwconv_backward = real_f, real_b
agg = collections.OrderedDict()
for key, e0, e1 in records:
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += e0.elapsed_time(e1)
print(f'{math}: pointwise / dense contractions of one step, batch {tb}, crop {crop}')
print('| pass | n | cin | cout | hw | calls | ms | TFLOP/s | GB/s (x + y [+ dy]) |\n|---|---|---|---|---|---|---|---|---|')
tot = 0.0
for (kind, n, cin, cout, hw), (calls, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    passes = 1 if kind == 'fwd' else (2 if kind == 'bwd' else 1)
    flops = 2.0 * n * cin * cout * hw * passes * calls
    byts = 4.0 * n * hw * calls * ((cin + cout) if kind == 'fwd' else (2 * cin + 2 * cout if kind == 'bwd' else cin + cout))
    tot += ms
    print(f'| {kind} | {n} | {cin} | {cout} | {hw} | {calls} | {ms:.3f} | {flops / ms / 1e9:.1f} | {byts / ms / 1e6:.0f} |')
print(f'total {tot:.2f} ms')
