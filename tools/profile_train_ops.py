"""ncu target: every training kernel family once, at the largest BASELINE config-5 shape it sees (batch 16, crop 768): the 96 x 96
x 128-channel classifier / FFM layers, the stride-2 depthwise of bottleneck 1.0, the fused resize + OHEM loss.
    python tools/profile_train_ops.py [fp32|tf32]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'fast-scnn-pytorch_b200'))
import torch

from fscnn_b200 import train_ops

dev = torch.device('cuda', 0)
train_ops.set_matmul_precision(sys.argv[1] if len(sys.argv) > 1 else 'tf32')
torch.manual_seed(0)
n = 16
for rep in range(2):      # the first pass warms up (workspaces, function attributes); ncu skips it with -s
    x = torch.randn(n, 128, 96, 96, device=dev, requires_grad=True)
    w = (torch.randn(128, 128, 1, 1, device=dev) / 11).requires_grad_(True)
    y = train_ops.pointwise_conv(x, w)                                  # forward GEMM
    bn = torch.nn.BatchNorm2d(128).to(dev).train()
    z = train_ops.batchnorm_relu(y, bn, True)                           # stats + apply
    wd = torch.randn(128, 1, 3, 3, device=dev, requires_grad=True)
    d = train_ops.depthwise_conv3x3(z, wd, 1)                           # stride-1 depthwise (float4 groups)
    d.sum().backward()                                                  # dgrad + wgrad (one pass), BN backward, GEMM dgrad + wgrad
    x2 = torch.randn(n, 384, 96, 96, device=dev, requires_grad=True)
    wd2 = torch.randn(384, 1, 3, 3, device=dev, requires_grad=True)
    train_ops.depthwise_conv3x3(x2, wd2, 2).sum().backward()            # stride-2 depthwise forward + both gradients
    low = torch.randn(n, 19, 96, 96, device=dev, requires_grad=True)
    t = torch.randint(-1, 19, (n, 768, 768), device=dev)
    cw = torch.ones(19, device=dev)
    train_ops.ohem_cross_entropy_upsampled(low, t, cw, -1, 0.7, 256).backward()
    torch.cuda.synchronize()
print('ok')
