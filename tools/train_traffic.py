"""Algorithmic HBM traffic of one training step (BASELINE config 5) at operator granularity, the plan-P convention of SURVEY 8(d):
every operator reads each of its input tensors once and writes each of its outputs once (forward: inputs + outputs; backward:
incoming gradients + saved tensors + produced gradients).  Tallied by wrapping every torch.autograd.Function of fscnn_b200.train_ops.
    python tools/train_traffic.py [batch=16] [crop=768]"""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'fast-scnn-pytorch_b200'))
import torch

import bench
from fscnn_b200 import Trainer, train_ops
from models.fast_scnn import FastSCNN

tb = int(sys.argv[1]) if len(sys.argv) > 1 else 16
crop = int(sys.argv[2]) if len(sys.argv) > 2 else 768
dev = torch.device('cuda', 0)
tally = collections.OrderedDict()


def nbytes(objs):
    return sum(o.numel() * o.element_size() for o in objs if isinstance(o, torch.Tensor))


def wrap(cls):
    fwd, bwd = cls.forward, cls.backward

    def forward(ctx, *args):
        out = fwd(ctx, *args)
        outs = out if isinstance(out, tuple) else (out,)
        t = tally.setdefault(cls.__name__, [0, 0, 0])
        t[0] += 1
        t[1] += nbytes(args) + nbytes(outs)
        return out

    def backward(ctx, *grads):
        out = bwd(ctx, *grads)
        outs = out if isinstance(out, tuple) else (out,)
        tally.setdefault(cls.__name__, [0, 0, 0])[2] += nbytes(grads) + nbytes(ctx.saved_tensors) + nbytes(outs)
        return out

    cls.forward, cls.backward = staticmethod(forward), staticmethod(backward)


for name in dir(train_ops):
    obj = getattr(train_ops, name)
    if isinstance(obj, type) and issubclass(obj, torch.autograd.Function) and obj is not torch.autograd.Function:
        wrap(obj)
m = FastSCNN(19, aux=True).train()
bench.init_recipe_d2(m, 3)
m.to(dev)
tr = Trainer(m)
x = bench.smooth_images(tb, crop, crop, dev, 1, chunk=16)
t = torch.randint(-1, 19, (tb, crop, crop), device=dev)
tr.step(x, t)
torch.cuda.synchronize()
print(f'algorithmic bytes of one training step, batch {tb}, crop {crop} (operator granularity)')
print('| operator | calls | forward GB | backward GB |\n|---|---|---|---|')
f = b = 0
for k, (calls, fb, bb) in sorted(tally.items(), key=lambda kv: -(kv[1][1] + kv[1][2])):
    print(f'| {k} | {calls} | {fb / 1e9:.3f} | {bb / 1e9:.3f} |')
    f += fb
    b += bb
params = sum(p.numel() for p in m.parameters()) * 4
print(f'total forward {f / 1e9:.2f} GB, backward {b / 1e9:.2f} GB, optimizer {5 * params / 1e9:.3f} GB -> {(f + b + 5 * params) / 1e9:.2f} GB per step')
