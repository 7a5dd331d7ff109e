import os, sys
ROOT = '/root/repo'
sys.path.insert(0, os.path.join(ROOT, 'fast-scnn-pytorch_b200'))
import torch
from fscnn_b200 import train_ops, native
dev = torch.device('cuda', 0)
torch.manual_seed(0)
train_ops.set_matmul_precision('tf32')
lib = native.lib()
n, cin, cout, hw = 1, 32, 128, 128
x = torch.randn(n, cin, hw, device=dev)
w = torch.randn(cout, cin, device=dev) / cin ** 0.5
y = torch.full((n, cout, hw), 7.0, device=dev)
rc = lib.fscnn_train_pwconv_forward(x.data_ptr(), w.data_ptr(), y.data_ptr(), n, cin, cout, hw, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print('rc', rc, 'y unique-ish', y.flatten()[:8].tolist(), 'count of 7.0:', int((y == 7.0).sum()), 'zeros:', int((y == 0).sum()), 'nan:', int(torch.isnan(y).sum()))
ref = w @ x[0]
print('ref', ref.flatten()[:8].tolist())
# ones test: x = 1, w = 1 -> y = cin
x1 = torch.ones(n, cin, hw, device=dev); w1 = torch.ones(cout, cin, device=dev); y1 = torch.full((n, cout, hw), 7.0, device=dev)
lib.fscnn_train_pwconv_forward(x1.data_ptr(), w1.data_ptr(), y1.data_ptr(), n, cin, cout, hw, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print('ones:', y1.flatten()[:8].tolist(), 'min', y1.min().item(), 'max', y1.max().item())
