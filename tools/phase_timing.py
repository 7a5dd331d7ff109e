"""Debug utility (library built with -DFSCNN_PHASE_TIMING): prints the clock64 deltas between the phases of one CTA of
the fused stem+dsconv1 kernel while the GPU is busy with a full batch."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'fast-scnn-pytorch_b200'))
import torch
from models.fast_scnn import FastSCNN
from fscnn_b200 import native

dev = torch.device('cuda', 0)
model = FastSCNN(19, precision='bf16').eval().to(dev)
x = torch.randn(8, 3, 1024, 2048, device=dev)
for _ in range(3):
    model.predict(x)
torch.cuda.synchronize()
lib = C.CDLL(native.lib_path())
buf = (C.c_longlong * 16)()
assert lib.fscnn_debug_front_phases(buf) == 0
t = list(buf)
names = ['alloc/barriers', 'stage patch (+sync)', 'im2col gather (+sync)', 'stem MMAs (+wait)', 'stem epilogue (+sync)', 'depthwise (+sync)',
         'pointwise MMAs (+wait)', 'output epilogue (+sync)', 'dealloc']
for i, nme in enumerate(names):
    print(f'{nme:28s} {t[i + 1] - t[i]:8d} cycles')
print(f'{"total":28s} {t[9] - t[0]:8d} cycles')
