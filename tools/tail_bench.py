"""GPU box utility: time the tail kernel (x8 upsample + argmax + metric counting, fscnn_upsample_argmax) alone, with its own
CUDA-event pair, on logits of different character (its time depends on how many classes stay candidates per block):
    python tools/tail_bench.py [images=37] [reps=10]
Prints microseconds per 1024x2048 image for: D2 network logits (recipe D2 weights + multi-scale input, calibrated: all 19
classes present), D1 logits (plain random-init network, randn input), iid noise, all-tied, a checkerboard class-order flip
(no class dominates another anywhere: nothing can be pruned) and every one of them with the pruning switched off."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ('fast-scnn-pytorch_b200', 'oracle', 'tests'):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np
import torch

import fastscnn_oracle as fo
from helpers import build_model
from models.fast_scnn import FastSCNN

nimg = int(sys.argv[1]) if len(sys.argv) > 1 else 37
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
nc, h, w = 19, 1024, 2048
dev = torch.device('cuda', 0)


def network_logits(model, x):
    eng = model._engine(dev)
    names = eng.stage_names()
    eng.forward_range(x, 0, names.index('cls.dsconv2+head'))
    v = eng.tap_view('cls.logits_lowres', x.shape[0], h, w)
    base = v._base if v._base is not None else v
    return v, eng


cases = {}
# D2: variance-preserving weights, multi-scale smooth input, calibrated classifier bias
sd = fo.make_state_dict(nc, False, 7)
xs = torch.from_numpy(fo.make_input(4, h, w, 31)).to(dev)
m = build_model(sd, nc, False, dev, precision='bf16')
v, eng = network_logits(m, xs)
sd['classifier.conv.1.bias'] = (sd['classifier.conv.1.bias'] - v.float().mean(dim=(0, 1, 2)).cpu().numpy()).astype(np.float32)
m = build_model(sd, nc, False, dev, precision='bf16')
v, eng = network_logits(m, xs)
ncp = v.shape[-1] if v.stride(2) == v.shape[-1] else v.stride(2)
hl, wl = v.shape[1], v.shape[2]


def padded(t):      # [n,hl,wl,nc] view -> contiguous [nimg,hl,wl,ncp]
    out = torch.zeros((nimg, hl, wl, ncp), dtype=torch.float32, device=dev)
    reps_ = -(-nimg // t.shape[0])
    out[..., :nc] = t.float().repeat(reps_, 1, 1, 1)[:nimg]
    return out


cases['D2 network logits'] = padded(v)
torch.manual_seed(0)
m1 = FastSCNN(nc, precision='bf16').eval().to(dev)
v1, _ = network_logits(m1, torch.randn(4, 3, h, w, device=dev))
cases['D1 plain-init logits'] = padded(v1)
g = torch.Generator(device=dev).manual_seed(1)
cases['iid noise'] = padded(torch.randn((4, hl, wl, nc), device=dev, generator=g))
cases['all tied'] = padded(torch.zeros((1, hl, wl, nc), device=dev))
checker = ((torch.arange(hl, device=dev)[:, None] + torch.arange(wl, device=dev)[None, :]) % 2 * 2 - 1).float()
cases['checkerboard order flip'] = padded((checker[None, :, :, None] * torch.arange(nc, device=dev).float()).contiguous())
# piecewise-constant "real segmentation like" logits: large regions, one clear winner
big = torch.randn((4, 8, 16, nc), device=dev, generator=g).permute(0, 3, 1, 2)
cases['large regions'] = padded(torch.nn.functional.interpolate(big, size=(hl, wl), mode='bilinear', align_corners=True).permute(0, 2, 3, 1) * 4)

labels64 = torch.randint(-1, nc, (nimg, h, w), device=dev, dtype=torch.int64)
labels8 = torch.randint(0, nc + 1, (nimg, h, w), device=dev, dtype=torch.uint8)
conf = torch.zeros(eng.conf_len(), dtype=torch.int64, device=dev)


def timeit(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps / nimg * 1e3


print(f'{nimg} images of {h}x{w}, {nc} classes; microseconds per image')
print(f'{"logits":28s} {"metric,i64":>11s} {"metric,u8":>10s} {"mask u8":>9s} {"mask+metric":>12s} | exhaustive: {"metric,i64":>10s} {"mask u8":>9s} | mean cand/px')
for name, low in cases.items():
    t = [timeit(lambda: eng.upsample_argmax(low, h, w, labels=labels64, conf=conf, want_mask=False)),
         timeit(lambda: eng.upsample_argmax(low, h, w, labels=labels8, conf=conf, want_mask=False)),
         timeit(lambda: eng.upsample_argmax(low, h, w)),
         timeit(lambda: eng.upsample_argmax(low, h, w, labels=labels8, conf=conf)),
         timeit(lambda: eng.upsample_argmax(low, h, w, labels=labels64, conf=conf, want_mask=False, exhaustive=True)),
         timeit(lambda: eng.upsample_argmax(low, h, w, exhaustive=True))]
    assert torch.equal(eng.upsample_argmax(low, h, w), eng.upsample_argmax(low, h, w, exhaustive=True)), name
    frac = [float((eng.upsample_argmax(low[:2].contiguous(), h, w) == c).float().mean()) for c in range(nc)]
    print(f'{name:28s} {t[0]:11.2f} {t[1]:10.2f} {t[2]:9.2f} {t[3]:12.2f} |             {t[4]:10.2f} {t[5]:9.2f} | classes present {sum(f > 1e-4 for f in frac)}, largest {max(frac):.2f}')
