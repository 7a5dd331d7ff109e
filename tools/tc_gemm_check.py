"""Debug aid for train_tc.cu: pointwise forward / data gradient in TF32 mode against torch on small aligned shapes, with error maps.
    python tools/tc_gemm_check.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'fast-scnn-pytorch_b200'))
import torch

from fscnn_b200 import train_ops

dev = torch.device('cuda', 0)
torch.manual_seed(0)
train_ops.set_matmul_precision('tf32')
for (n, cin, cout, h, w) in [(1, 32, 128, 8, 16), (1, 8, 128, 8, 16), (2, 64, 48, 24, 28), (1, 128, 384, 16, 24)]:
    x = torch.randn(n, cin, h, w, device=dev, requires_grad=True)
    wt = (torch.randn(cout, cin, 1, 1, device=dev) / cin ** 0.5).requires_grad_(True)
    dy = torch.randn(n, cout, h, w, device=dev)
    y = train_ops.pointwise_conv(x, wt)
    y.backward(dy)
    torch.cuda.synchronize()
    yr = torch.einsum('oc,nchw->nohw', wt.detach()[:, :, 0, 0].double(), x.detach().double()).float()
    dxr = torch.einsum('oc,nohw->nchw', wt.detach()[:, :, 0, 0].double(), dy.double()).float()
    ey = (y.detach() - yr).abs().max().item() / yr.abs().max().item()
    ex = (x.grad - dxr).abs().max().item() / dxr.abs().max().item()
    print(f'n{n} cin{cin} cout{cout} hw{h * w}: fwd rel err {ey:.3e}  dgrad rel err {ex:.3e}  |y| max {y.abs().max().item():.3f} (ref {yr.abs().max().item():.3f})')
    if ey > 1e-2:
        yy, rr = y.detach()[0].reshape(cout, -1), yr[0].reshape(cout, -1)
        print('  y[0:4, 0:8]  ', yy[0:4, 0:8].cpu().numpy().round(3).tolist())
        print('  ref[0:4, 0:8]', rr[0:4, 0:8].cpu().numpy().round(3).tolist())
        # does y match the reference under a permutation of rows / columns?  correlate a few rows
        c = (yy[:8, :64] @ rr[:16, :64].T)
        print('  row match (argmax of correlation with ref rows 0..15):', c.argmax(1).tolist())
        c2 = (yy[:4, :32].T @ rr[:4, :64])
        print('  col match:', c2.argmax(1).tolist())
    if ex > 1e-2:
        a, b = x.grad[0].reshape(cin, -1), dxr[0].reshape(cin, -1)
        print('  dx[0:4, 0:8] ', a[0:4, 0:8].cpu().numpy().round(3).tolist())
        print('  ref[0:4, 0:8]', b[0:4, 0:8].cpu().numpy().round(3).tolist())
train_ops.set_matmul_precision('fp32')
