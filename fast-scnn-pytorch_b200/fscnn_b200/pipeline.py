"""StreamingEvaluator: the batched evaluation loop of the reference (eval.py:38-55) as a double-buffered
host -> device pipeline.  While batch i runs through the fused forward + argmax + metric kernels on the
compute stream, batch i+1 is copied from pinned host memory on a copy stream; the metric state stays
on the device and is read back asynchronously.

    ev = StreamingEvaluator(model, metric, batch_shape=(16, 1024, 2048, 3))
    for images_u8, labels_u8 in loader:          # pinned host tensors (uint8 HWC images, uint8/int labels)
        ev.submit(images_u8, labels_u8)
    pixAcc, mIoU = ev.result()
"""
from __future__ import annotations

import torch


class StreamingEvaluator:
    def __init__(self, model, metric, images_like: torch.Tensor, labels_like: torch.Tensor, device=None, normalize='default',
                 depth: int = 2):
        self.model, self.metric = model, metric
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        self.depth = depth
        self.kw = {} if normalize == 'default' else {'normalize': normalize}
        with torch.cuda.device(self.device):
            self.copy_stream = torch.cuda.Stream()
            self.img = [torch.empty_like(images_like, device=self.device) for _ in range(depth)]
            self.lab = [torch.empty_like(labels_like, device=self.device) for _ in range(depth)]
            self.copied = [torch.cuda.Event() for _ in range(depth)]
            self.consumed = [torch.cuda.Event() for _ in range(depth)]
            self.conf_host = torch.zeros(metric.conf_len(), dtype=torch.int64).pin_memory()
        self.k = 0
        metric.device_confusion(self.device)

    def submit(self, images: torch.Tensor, labels: torch.Tensor) -> None:
        """Enqueue one host batch (ideally pinned); returns immediately."""
        slot = self.k % self.depth
        compute = torch.cuda.current_stream(self.device)
        if self.k >= self.depth:
            self.copy_stream.wait_event(self.consumed[slot])      # the kernels that read this slot have finished
        with torch.cuda.stream(self.copy_stream):
            self.img[slot].copy_(images, non_blocking=True)
            self.lab[slot].copy_(labels, non_blocking=True)
            self.copied[slot].record(self.copy_stream)
        compute.wait_event(self.copied[slot])
        self.model.evaluate(self.img[slot], self.lab[slot], self.metric, **self.kw)
        self.consumed[slot].record(compute)
        self.conf_host.copy_(self.metric.device_confusion(self.device), non_blocking=True)   # device -> host read of the step's result
        self.k += 1

    def result(self):
        """(pixAcc, mIoU) over everything submitted so far (synchronises)."""
        torch.cuda.current_stream(self.device).synchronize()
        return self.metric.get()
