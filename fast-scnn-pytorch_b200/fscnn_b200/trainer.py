"""One training step of the reference (train.py:241-306) on the B200 training operators.

``Trainer`` owns what ``train.py`` sets up around the model: the loss (``MixSoftmaxCrossEntropyOHEMLoss``: OHEM cross entropy
on the main head + ``aux_weight`` x the aux head, utils/loss.py:185-206), SGD with momentum and weight decay on ALL parameters
(one param group, train.py:195-198), the poly learning-rate schedule (utils/lr_scheduler.py:66-91) and, under
``torch.distributed``, DDP semantics: per-rank BatchNorm statistics, gradients averaged over the ranks with ONE NCCL all-reduce of
a flat fp32 buffer (1.16 M parameters, 4.6 MB) followed by the same fused SGD update on every rank.

Parameters and gradients live in two flat buffers (the module's parameters and their .grad are views into them), so the
collective and the optimizer are one launch each.  The reference trains under fp16 autocast + GradScaler (train.py:73-74,
:267-275: ``--use-fp16`` defaults to True); this path keeps fp32 tensors throughout and fp32 arithmetic by default
(``matmul_precision='tf32'`` moves the contractions to the tensor cores with TF32 operands and fp32 accumulation, still at least the
precision of fp16 autocast).  ``cuda_graph=True`` replays zero_grad + forward + loss + backward from a captured CUDA graph.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.distributed as dist

from . import train_ops

# the 19 Cityscapes class weights hard-coded by the reference (utils/loss.py:135-137)
OHEM_CLASS_WEIGHTS = (0.8373, 0.918, 0.866, 1.0345, 1.0166, 0.9969, 0.9754, 1.0489, 0.8786, 1.0023, 0.9539, 0.9843, 1.1116, 0.9037,
                      1.0865, 1.0955, 1.0865, 1.1529, 1.0507)


def poly_lr(base_lr: float, cur_iter: int, nepochs: int, iters_per_epoch: int, power: float = 0.9) -> float:
    """LRScheduler(mode='poly') (utils/lr_scheduler.py:67-76, :91): base * (1 - T / N) ** power with N = nepochs * iters - 1."""
    n = nepochs * iters_per_epoch - 1
    return base_lr * (1.0 - cur_iter / n) ** power if n > 0 else base_lr


class Trainer:
    loss_type = 'ohem'      # class-level default: MixSoftmaxCrossEntropyOHEMLoss (BASELINE config 5)

    def __init__(self, model, base_lr=1e-2, momentum=0.9, weight_decay=1e-4, aux_weight=0.4, ignore_label=-1, ohem_thresh=0.7,
                 ohem_min_kept=256, use_class_weights: Optional[bool] = None, nepochs=160, iters_per_epoch=1000, process_group=None,
                 fused_loss: bool = True, cuda_graph: bool = False, graph_warmup: int = 3, matmul_precision: Optional[str] = None,
                 loss_type: str = 'ohem', dice_smooth: float = 1e-6, focal_alpha: float = 0.5, focal_gamma: float = 2.0,
                 focal_dice_weight: float = 0.5, optimizer: str = 'sgd', betas=(0.9, 0.999), adam_eps: float = 1e-8):
        """``loss_type``: the criterion train.py builds (train.py:182-192) -- 'ohem' = MixSoftmaxCrossEntropyOHEMLoss (its
        --loss-type ce, BASELINE config 5), 'dice' = MixDiceLoss (train.py's default), 'ce' = MixSoftmaxCrossEntropyLoss,
        'focal_dice' = FocalDiceLoss on the main head (the reference's own call hands it the output tuple and fails in
        loss.py:82, pred.dim()).  ``optimizer``: 'sgd' = torch.optim.SGD(momentum, weight_decay) (train.py:195-198,
        train_custom_finetune.py:102) or 'adamw' = torch.optim.AdamW(lr, weight_decay) with ``betas`` / ``adam_eps`` (train_bdd100k.py:183;
        pass the weight decay that script uses).  ``cuda_graph``: after ``graph_warmup`` eager steps, zero_grad + forward + loss + backward are captured once into a CUDA
        graph per input shape and replayed (the ~780 kernel launches of a step cost more host time than the kernels take on a
        B200); the gradient all-reduce and the SGD update stay outside the graph, so the learning rate remains a host value.
        ``matmul_precision``: 'fp32' or 'tf32' for the pointwise / dense 3x3 contractions (train_ops.set_matmul_precision;
        process-wide); None leaves the current setting."""
        self.model = model
        self.fused_loss = bool(fused_loss)
        if loss_type not in ('ohem', 'ce', 'dice', 'focal_dice'):
            raise ValueError(f"loss_type must be 'ohem', 'ce', 'dice' or 'focal_dice', got {loss_type!r}")
        self.loss_type = loss_type
        self.dice_smooth, self.focal = float(dice_smooth), (float(focal_alpha), float(focal_gamma), float(focal_dice_weight))
        if matmul_precision is not None:
            train_ops.set_matmul_precision(matmul_precision)
        self.cuda_graph, self.graph_warmup = bool(cuda_graph), int(graph_warmup)
        self._graph = self._graph_key = self._static_x = self._static_t = self._static_loss = None
        params = [p for p in model.parameters() if p.requires_grad]
        if not params or not params[0].is_cuda:
            raise RuntimeError('move the model to a CUDA device before building the Trainer (there is no CPU path)')
        dev = params[0].device
        total = sum(p.numel() for p in params)
        self.flat_param = torch.empty(total, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.momentum_buf = torch.zeros(total, dtype=torch.float32, device=dev)      # SGD momentum / AdamW first moment
        if optimizer not in ('sgd', 'adamw'):
            raise ValueError(f"optimizer must be 'sgd' or 'adamw', got {optimizer!r}")
        self.optimizer, self.betas, self.adam_eps = optimizer, (float(betas[0]), float(betas[1])), float(adam_eps)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev) if optimizer == 'adamw' else None
        off = 0
        self._grad_views = []
        with torch.no_grad():
            for p in params:      # parameters and their gradients become views into the flat buffers
                n = p.numel()
                self.flat_param[off:off + n].copy_(p.reshape(-1))
                p.data = self.flat_param[off:off + n].view_as(p)
                self._grad_views.append(self.flat_grad[off:off + n].view_as(p))
                p.grad = self._grad_views[-1]
                off += n
        self.params = params
        self.base_lr, self.momentum, self.weight_decay = float(base_lr), float(momentum), float(weight_decay)
        self.aux_weight, self.ignore_label = float(aux_weight), int(ignore_label)
        self.ohem_thresh, self.ohem_min_kept = float(ohem_thresh), int(ohem_min_kept)
        self.nepochs, self.iters_per_epoch = int(nepochs), int(iters_per_epoch)
        if use_class_weights is None:      # the reference always uses its 19 weights (only valid for 19 classes)
            use_class_weights = model.num_classes == len(OHEM_CLASS_WEIGHTS)
        self.class_weight = torch.tensor(OHEM_CLASS_WEIGHTS, dtype=torch.float32, device=dev) if use_class_weights else None
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.iteration = 0
        self._dropout_step = torch.zeros(1, dtype=torch.int64, device=dev) if self.cuda_graph else None
        if self.world > 1:      # every rank starts from rank 0's weights
            dist.broadcast(self.flat_param, src=0, group=process_group)

    def _other_criterion(self, outputs, target):
        """MixSoftmaxCrossEntropyLoss (loss.py:103-124), MixDiceLoss (:42-68) or FocalDiceLoss (:71-100) over the heads; logits
        smaller than the target are resized inside the loss kernels (train_ops.criterion)."""
        if self.loss_type == 'focal_dice':
            a, g, dw = self.focal
            return train_ops.criterion(outputs[0], target, 'focal_dice', ignore_label=-100, smooth=self.dice_smooth, alpha=a, gamma=g,
                                       dice_weight=dw)
        heads = outputs if self.loss_type == 'ce' else outputs[:2]      # MixDiceLoss looks at the first aux prediction only
        total = train_ops.criterion(heads[0], target, self.loss_type, ignore_label=self.ignore_label, smooth=self.dice_smooth)
        for aux_out in heads[1:]:
            total = total + self.aux_weight * train_ops.criterion(aux_out, target, self.loss_type, ignore_label=self.ignore_label,
                                                                   smooth=self.dice_smooth)
        return total

    def loss_from_lowres(self, lowres_outputs, target):
        """The same loss from the heads' LOW-RESOLUTION logits: the final x8 bilinear resize (models/fast_scnn.py:40, :44) is fused
        into the OHEM kernels (train_ops.ohem_cross_entropy_upsampled)."""
        if self.loss_type != 'ohem':
            return self._other_criterion(lowres_outputs, target)
        total = train_ops.ohem_cross_entropy_upsampled(lowres_outputs[0], target, self.class_weight, self.ignore_label, self.ohem_thresh,
                                                       self.ohem_min_kept)
        for aux_out in lowres_outputs[1:]:
            total = total + self.aux_weight * train_ops.ohem_cross_entropy_upsampled(aux_out, target, self.class_weight, self.ignore_label,
                                                                                      self.ohem_thresh, self.ohem_min_kept)
        return total

    def loss(self, outputs, target):
        """MixSoftmaxCrossEntropyOHEMLoss.forward (utils/loss.py:191-206), or the criterion ``loss_type`` names."""
        if self.loss_type != 'ohem':
            return self._other_criterion(outputs, target)
        total = train_ops.ohem_cross_entropy(outputs[0], target, self.class_weight, self.ignore_label, self.ohem_thresh, self.ohem_min_kept)
        for aux_out in outputs[1:]:
            total = total + self.aux_weight * train_ops.ohem_cross_entropy(aux_out, target, self.class_weight, self.ignore_label,
                                                                            self.ohem_thresh, self.ohem_min_kept)
        return total

    def step(self, images, target, lr: Optional[float] = None):
        """zero_grad -> forward -> loss -> backward -> gradient all-reduce (mean) -> SGD (train.py:253-284).  Returns the loss
        as a 0-d device tensor (no host synchronisation; the reference's loss.item() is the caller's choice)."""
        self.model.train()
        if lr is None:
            lr = poly_lr(self.base_lr, self.iteration, self.nepochs, self.iters_per_epoch)
        if self.cuda_graph and self.iteration >= self.graph_warmup:
            loss = self._graph_forward_backward(images, target)
        else:
            loss = self._forward_backward(images, target)
        if self.world > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.group)
        if self.optimizer == 'adamw':
            train_ops.adamw_step(self.flat_param, self.flat_grad, self.momentum_buf, self.exp_avg_sq, lr, self.iteration + 1, self.betas,
                                 self.adam_eps, self.weight_decay, grad_scale=1.0 / self.world)
        else:
            train_ops.sgd_step(self.flat_param, self.flat_grad, self.momentum_buf, lr, self.momentum, self.weight_decay,
                               grad_scale=1.0 / self.world, first_step=self.iteration == 0)
        self.iteration += 1
        return loss

    def _forward_backward(self, images, target):
        """zero_grad + forward + loss + backward, leaving the gradients in the flat buffer.  With .grad pointing at the flat views
        autograd would ADD every produced gradient to it (148 small kernels per step on top of zeroing the buffer); instead .grad is
        cleared, autograd stores the gradients it produces, and one multi-tensor copy moves them into the flat views, which .grad
        then points at again (what zero_grad(set_to_none=True) + backward leaves, packed)."""
        for p in self.params:
            p.grad = None
        with train_ops.deferred_batch_counters():
            if self.fused_loss and hasattr(self.model, '_train_forward_lowres'):
                loss = self.loss_from_lowres(self.model._train_forward_lowres(images), target)
            else:
                loss = self.loss(self.model(images), target)
        loss.backward()
        with torch.no_grad():
            missing = [v for p, v in zip(self.params, self._grad_views) if p.grad is None]
            if missing:                      # a parameter that took no part in this forward: its gradient is zero
                torch._foreach_zero_(missing)
            have = [(v, p.grad) for p, v in zip(self.params, self._grad_views) if p.grad is not None]
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        for p, v in zip(self.params, self._grad_views):
            p.grad = v
        return loss.detach()

    def _graph_forward_backward(self, images, target):
        """Replays the captured zero_grad + forward + loss + backward on static copies of the inputs (captured on first use per input
        shape).  Dropout masks change from replay to replay through a device step counter mixed into every dropout seed.  The
        returned loss is the graph's static output tensor: it is overwritten by the next step."""
        key = (tuple(images.shape), images.dtype, tuple(target.shape), target.dtype)
        if self._graph is None or key != self._graph_key:
            self._static_x, self._static_t = images.detach().clone(), target.detach().clone()
            torch.cuda.synchronize(images.device)
            graph = torch.cuda.CUDAGraph()
            train_ops.set_dropout_step_counter(self._dropout_step)
            try:
                with torch.cuda.graph(graph):
                    self._static_loss = self._forward_backward(self._static_x, self._static_t)
            finally:
                train_ops.set_dropout_step_counter(None)
            self._graph, self._graph_key = graph, key
        if images.data_ptr() != self._static_x.data_ptr():
            self._static_x.copy_(images, non_blocking=True)
        if target.data_ptr() != self._static_t.data_ptr():
            self._static_t.copy_(target, non_blocking=True)
        self._dropout_step += 1
        self._graph.replay()
        return self._static_loss
