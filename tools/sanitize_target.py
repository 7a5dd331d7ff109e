"""Small end-to-end run of every kernel (both precisions, aux head, odd sizes, uint8 input, metric) for
compute-sanitizer:   compute-sanitizer --tool memcheck python tools/sanitize_target.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'fast-scnn-pytorch_b200'))
import torch
from models.fast_scnn import FastSCNN
from utils.metric import SegmentationMetric

dev = torch.device('cuda', 0)
torch.manual_seed(0)
for prec in ('fp32', 'bf16'):
    for nc, aux, shape in ((19, True, (2, 3, 97, 161)), (2, False, (1, 3, 200, 264))):
        model = FastSCNN(nc, aux=aux, precision=prec).eval().to(dev)
        x = torch.randn(*shape, device=dev)
        labels = torch.randint(-1, nc, (shape[0], shape[2], shape[3]), device=dev)
        outs = model(x)
        mask = model.predict(x)
        metric = SegmentationMetric(nc, device=dev)
        model.evaluate(x, labels, metric)
        xu = torch.randint(0, 256, (shape[0], shape[2], shape[3], 3), dtype=torch.uint8, device=dev)
        model.evaluate(xu, labels.clamp(min=0).to(torch.uint8), metric)
        metric.update(mask.cpu().numpy(), labels.cpu().numpy())
        torch.cuda.synchronize()
        print(prec, nc, aux, shape, [tuple(o.shape) for o in outs], metric.get())
print('sanitize target done')
