"""ncu target: the criterion kernels (fscnn_train_criterion_*: cross entropy / dice / focal + dice) at the BASELINE config-5 loss shape --
16 images, low-resolution logits 96 x 96 -> labels 768 x 768 -- fused with the resize, and the dice kernels at label resolution.
    python tools/profile_train_criteria.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'fast-scnn-pytorch_b200'))
import torch

from fscnn_b200 import train_ops

dev = torch.device('cuda', 0)
torch.manual_seed(0)
n = 16
for rep in range(2):      # the first pass warms up; ncu skips it with -s
    lane = (torch.rand((n, 768, 768), device=dev) < 0.1).long()
    low2 = torch.randn(n, 2, 96, 96, device=dev, requires_grad=True)
    train_ops.criterion(low2, lane, 'dice').backward()                  # crit_fwd<2, up>, crit_up_grad<2>
    train_ops.criterion(low2, lane, 'focal_dice').backward()
    low7 = torch.randn(n, 7, 96, 96, device=dev, requires_grad=True)
    t7 = torch.randint(-1, 7, (n, 768, 768), device=dev)
    train_ops.criterion(low7, t7, 'ce').backward()                      # register bucket of 8
    full2 = torch.randn(n, 2, 768, 768, device=dev, requires_grad=True)
    train_ops.criterion(full2, lane, 'dice').backward()                 # label resolution: crit_fwd<2, direct>, crit_grad<2>
    torch.cuda.synchronize()
print('ok')
