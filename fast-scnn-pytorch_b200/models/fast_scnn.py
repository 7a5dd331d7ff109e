"""Fast-SCNN as a drop-in ``torch.nn.Module`` whose eval-mode forward runs on hand-written
sm_100a kernels (csrc/, through the C ABI in include/fscnn_b200.h).

Mirrors the public surface of the reference ``models/fast_scnn.py``:

* ``FastSCNN(num_classes, aux=False, **kwargs)`` (reference :16-46) -- same module tree, so
  ``state_dict()`` / ``load_state_dict()`` use the reference's ``.pth`` layout (268 tensors, or 276
  with ``aux=True``), including ``module.``-prefixed DataParallel checkpoints;
  ``forward(x)`` returns the tuple ``(logits,)`` or ``(logits, aux_logits)`` of NCHW fp32 tensors;
* ``get_fast_scnn(dataset, pretrained, root, map_cpu, **kwargs)`` (reference :240-256).

The nn.Modules below only *hold* the parameters in the reference layout.  The arithmetic -- BN
folded at load time, fused depthwise-separable / bottleneck / PPM / FFM / classifier kernels, the
final upsample (+ argmax, + metric histogram) -- lives in ``libfscnn_b200.so``.  There is no CPU or
eager-PyTorch fallback: a CPU tensor or training mode raises.

Additions over the reference (optional, all keyword-only or new methods):
``precision='fp32'|'bf16'`` constructor keyword, ``predict(x)`` (fused upsample+argmax mask),
``evaluate(x, labels, metric)`` (fused SegmentationMetric counting), and raw ``uint8 [N,H,W,3]`` image
batches: ``transforms.ToTensor()`` + ``Normalize(mean, std)`` (eval.py:22-25) then run inside the stem
kernel (``normalize=(mean, std)``, default ImageNet as in the reference; ``normalize=None`` = /255 only).
"""
from __future__ import annotations

import os
from collections import OrderedDict
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

__all__ = ['FastSCNN', 'get_fast_scnn']

# dataset name -> number of classes; the reference reads ``datasets[name].NUM_CLASS``
# (data_loader/cityscapes.py:41, tusimple.py:41, bdd100k.py:52, custom.py:18).
NUM_CLASS = {'citys': 19, 'tusimple': 2, 'bdd100k': 2, 'custom': 2}
# checkpoint file acronyms, reference models/fast_scnn.py:241-248
_ACRONYMS = {'pascal_voc': 'voc', 'pascal_aug': 'voc', 'ade20k': 'ade', 'coco': 'coco', 'citys': 'citys',
             'tusimple': 'tusimple'}


# transforms.Normalize constants every reference driver uses (eval.py:22-25, demo.py:37-40)
IMAGENET_MEAN, IMAGENET_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


class _ParamHolder(nn.Module):
    """Base of the parameter-holding submodules.  In eval mode they are never called on their own (the fused CUDA path has
    no per-module entry point); in TRAINING mode the covered ones (ConvBNReLU 1x1, DSConv, DWConv, LinearBottleneck: the
    first slice of SURVEY.md section 8 row f3) run through the training operators of fscnn_b200/train_ops.py and are
    differentiable with ordinary ``loss.backward()``."""

    def forward(self, *args, **kwargs):
        if self.training:
            return self._train_forward(*args, **kwargs)
        raise RuntimeError(f'{type(self).__name__} only holds parameters in eval mode; call FastSCNN.forward / predict / evaluate '
                           '(the fused CUDA path has no per-module entry point and no CPU fallback)')

    def _train_forward(self, *args, **kwargs):
        raise NotImplementedError(f'{type(self).__name__} has no training-mode forward yet: the training step covers ConvBNReLU '
                                  '(1x1), DSConv, DWConv and LinearBottleneck so far (there is no eager-PyTorch fallback)')


def _train_seq(seq, x):
    """Runs an nn.Sequential of the covered layer kinds in training mode through the CUDA training operators: depthwise 3x3
    / pointwise 1x1 convolutions without bias, BatchNorm2d with batch statistics fused with the ReLU that follows it."""
    from fscnn_b200 import train_ops
    layers = list(seq)
    i = 0
    while i < len(layers):
        m = layers[i]
        if isinstance(m, nn.Conv2d):
            if m.dilation != (1, 1):
                raise NotImplementedError('training operators cover undilated convolutions')
            if m.kernel_size == (1, 1) and m.groups == 1 and m.stride == (1, 1) and m.padding == (0, 0):
                x = train_ops.pointwise_conv(x, m.weight)
            elif (m.kernel_size == (3, 3) and m.groups == m.in_channels == m.out_channels and m.groups > 1 and m.padding == (1, 1)
                  and m.stride[0] == m.stride[1]):
                x = train_ops.depthwise_conv3x3(x, m.weight, m.stride[0])
            elif m.kernel_size == (3, 3) and m.groups == 1 and m.stride[0] == m.stride[1] and m.padding in ((0, 0), (1, 1)):
                x = train_ops.conv3x3_dense(x, m.weight, m.stride[0], m.padding[0])
            else:
                raise NotImplementedError(f'no training operator for {m} (covered: dense 3x3, depthwise 3x3 pad 1, pointwise 1x1)')
            if m.bias is not None:
                x = train_ops.bias_add(x, m.bias)
            i += 1
        elif isinstance(m, nn.BatchNorm2d):
            relu = i + 1 < len(layers) and isinstance(layers[i + 1], nn.ReLU)
            x = train_ops.batchnorm_relu(x, m, relu)
            i += 2 if relu else 1
        elif isinstance(m, nn.ReLU):          # a ReLU that does not follow a BatchNorm (none in this network)
            x = train_ops.add_relu(x, torch.zeros_like(x), True)
            i += 1
        elif isinstance(m, nn.Dropout):
            x = train_ops.dropout(x, m.p, m.training)
            i += 1
        elif isinstance(m, _ParamHolder):
            x = m(x)
            i += 1
        else:
            raise NotImplementedError(f'no training operator for {type(m).__name__} yet')
    return x


def _conv_bn_relu(cin, cout, k=3, stride=1, padding=0):
    return nn.Sequential(nn.Conv2d(cin, cout, k, stride, padding, bias=False), nn.BatchNorm2d(cout), nn.ReLU(True))


class ConvBNReLU(_ParamHolder):
    """Conv + BN + ReLU; parameters under ``conv.0`` / ``conv.1`` (reference _ConvBNReLU, :49-61)."""

    def __init__(self, cin, cout, k=3, stride=1, padding=0, **_):
        super().__init__()
        self.conv = _conv_bn_relu(cin, cout, k, stride, padding)

    def _train_forward(self, x):
        return _train_seq(self.conv, x)


class DSConv(_ParamHolder):
    """Depthwise 3x3 + BN + ReLU, pointwise 1x1 + BN + ReLU: ``conv.0,1,3,4`` (reference _DSConv, :64-79)."""

    def __init__(self, channels, cout, stride=1, **_):
        super().__init__()
        self.conv = nn.Sequential(
            nn.Conv2d(channels, channels, 3, stride, 1, groups=channels, bias=False), nn.BatchNorm2d(channels),
            nn.ReLU(True), nn.Conv2d(channels, cout, 1, bias=False), nn.BatchNorm2d(cout), nn.ReLU(True))

    def _train_forward(self, x):
        return _train_seq(self.conv, x)


class DWConv(_ParamHolder):
    """Depthwise 3x3 + BN + ReLU: ``conv.0,1`` (reference _DWConv, :82-92)."""

    def __init__(self, channels, cout, stride=1, **_):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv2d(channels, cout, 3, stride, 1, groups=channels, bias=False),
                                  nn.BatchNorm2d(cout), nn.ReLU(True))

    def _train_forward(self, x):
        return _train_seq(self.conv, x)


class LinearBottleneck(_ParamHolder):
    """MobileNetV2 inverted residual: ``block.0`` expand, ``block.1`` depthwise, ``block.2/3`` linear
    projection (reference :95-115)."""

    def __init__(self, cin, cout, t=6, stride=2, **_):
        super().__init__()
        self.use_shortcut = stride == 1 and cin == cout
        self.block = nn.Sequential(ConvBNReLU(cin, cin * t, 1), DWConv(cin * t, cin * t, stride),
                                   nn.Conv2d(cin * t, cout, 1, bias=False), nn.BatchNorm2d(cout))

    def _train_forward(self, x):
        from fscnn_b200 import train_ops
        out = _train_seq(self.block, x)
        return train_ops.add_relu(x, out, False) if self.use_shortcut else out      # reference :111-115


class PyramidPooling(_ParamHolder):
    """``conv1..conv4`` (128->32 each) and ``out`` (256->128) (reference :118-145)."""

    def __init__(self, cin, cout, **_):
        super().__init__()
        inter = cin // 4
        for i in range(1, 5):
            setattr(self, f'conv{i}', ConvBNReLU(cin, inter, 1))
        self.out = ConvBNReLU(cin * 2, cout, 1)

    def pool(self, x, size):   # kept: export scripts introspect ``ppm.pool`` (export_onnx_fixed.py:143-146)
        return nn.AdaptiveAvgPool2d(size)(x)

    def _train_forward(self, x):   # reference :137-145
        from fscnn_b200 import train_ops
        size = x.shape[2:]
        feats = [x]
        for i, bins in enumerate((1, 2, 3, 6), start=1):
            f = getattr(self, f'conv{i}')(train_ops.adaptive_avg_pool(x, bins))
            feats.append(train_ops.bilinear_resize(f, size))
        return self.out(torch.cat(feats, dim=1))


class LearningToDownsample(_ParamHolder):
    """``conv`` (3->32 s2), ``dsconv1`` (32->48 s2), ``dsconv2`` (48->64 s2) (reference :148-161)."""

    def __init__(self, c1=32, c2=48, cout=64, **_):
        super().__init__()
        self.conv = ConvBNReLU(3, c1, 3, 2)
        self.dsconv1 = DSConv(c1, c2, 2)
        self.dsconv2 = DSConv(c2, cout, 2)

    def _train_forward(self, x):   # reference :157-161
        return self.dsconv2(self.dsconv1(self.conv(x)))


class GlobalFeatureExtractor(_ParamHolder):
    """``bottleneck1..3`` (three LinearBottlenecks each) and ``ppm`` (reference :164-187)."""

    def __init__(self, cin=64, block_channels=(64, 96, 128), cout=128, t=6, num_blocks=(3, 3, 3), **_):
        super().__init__()
        strides = (2, 2, 1)
        for i, (planes, blocks, stride) in enumerate(zip(block_channels, num_blocks, strides), start=1):
            layers = [LinearBottleneck(cin if j == 0 else planes, planes, t, stride if j == 0 else 1) for j in range(blocks)]
            setattr(self, f'bottleneck{i}', nn.Sequential(*layers))
            cin = planes
        self.ppm = PyramidPooling(block_channels[2], cout)

    def _train_forward(self, x):   # reference :182-187
        for i in (1, 2, 3):
            for block in getattr(self, f'bottleneck{i}'):
                x = block(x)
        return self.ppm(x)


class FeatureFusionModule(_ParamHolder):
    """``dwconv``, ``conv_lower_res`` and ``conv_higher_res`` (both 1x1 with bias + BN) (reference :190-218)."""

    def __init__(self, highter_in_channels, lower_in_channels, out_channels, scale_factor=4, **_):
        super().__init__()
        self.scale_factor = scale_factor
        self.dwconv = DWConv(lower_in_channels, out_channels, 1)
        self.conv_lower_res = nn.Sequential(nn.Conv2d(out_channels, out_channels, 1), nn.BatchNorm2d(out_channels))
        self.conv_higher_res = nn.Sequential(nn.Conv2d(highter_in_channels, out_channels, 1), nn.BatchNorm2d(out_channels))
        self.relu = nn.ReLU(True)

    def _train_forward(self, higher_res_feature, lower_res_feature):   # reference :207-218
        from fscnn_b200 import train_ops
        lower = train_ops.bilinear_resize(lower_res_feature, higher_res_feature.shape[2:])
        lower = _train_seq(self.conv_lower_res, self.dwconv(lower))
        higher = _train_seq(self.conv_higher_res, higher_res_feature)
        return train_ops.add_relu(higher, lower, True)


class Classifer(_ParamHolder):
    """``dsconv1``, ``dsconv2`` and ``conv`` = [Dropout, Conv2d(128, nc, 1)] (reference :221-237; the
    class name keeps the reference's spelling)."""

    def __init__(self, channels, num_classes, stride=1, **_):
        super().__init__()
        self.dsconv1 = DSConv(channels, channels, stride)
        self.dsconv2 = DSConv(channels, channels, stride)
        self.conv = nn.Sequential(nn.Dropout(0.1), nn.Conv2d(channels, num_classes, 1))

    def _train_forward(self, x):   # reference :233-237
        return _train_seq(self.conv, self.dsconv2(self.dsconv1(x)))


class _Runtime:
    """Per-model native state shared by DataParallel replicas: one Engine per device, plus the
    fingerprint of the parameters each engine was packed from."""

    def __init__(self):
        self.engines: Dict[torch.device, object] = {}
        self.fingerprints: Dict[torch.device, Tuple] = {}


class FastSCNN(nn.Module):
    def __init__(self, num_classes, aux=False, **kwargs):
        super().__init__()
        self.aux = aux
        self.num_classes = int(num_classes)
        self.precision = kwargs.get('precision') or os.environ.get('FSCNN_PRECISION', 'fp32')
        self.learning_to_downsample = LearningToDownsample(32, 48, 64)
        self.global_feature_extractor = GlobalFeatureExtractor(64, [64, 96, 128], 128, 6, [3, 3, 3])
        self.feature_fusion = FeatureFusionModule(64, 128, 128)
        self.classifier = Classifer(128, num_classes)
        if self.aux:
            self.auxlayer = nn.Sequential(nn.Conv2d(64, 32, 3, padding=1, bias=False), nn.BatchNorm2d(32), nn.ReLU(True),
                                          nn.Dropout(0.1), nn.Conv2d(32, num_classes, 1))
        self._rt = _Runtime()
        self._fp_tensors = None

    def __getstate__(self):   # the native runtime is rebuilt lazily; it never travels with copies / pickles
        state = self.__dict__.copy()
        state['_rt'] = None
        state['_fp_tensors'] = None
        return state

    def __setstate__(self, state):
        super().__setstate__(state)
        self.__dict__['_rt'] = _Runtime()

    # ---- checkpoint compatibility ------------------------------------------------------------
    def load_state_dict(self, state_dict, strict=True, **kwargs):
        """Accepts the reference's bare state_dict, a DataParallel one (``module.`` prefix) or a wrapper
        dict holding it under 'state_dict' / 'model' / 'model_state_dict' (SURVEY.md section 5)."""
        for key in ('state_dict', 'model_state_dict', 'model'):
            if isinstance(state_dict, dict) and key in state_dict and isinstance(state_dict[key], dict):
                state_dict = state_dict[key]
                break
        if state_dict and all(k.startswith('module.') for k in state_dict):
            state_dict = OrderedDict((k[len('module.'):], v) for k, v in state_dict.items())
        self._fp_tensors = None
        return super().load_state_dict(state_dict, strict=strict, **kwargs)

    # ---- native engine management ------------------------------------------------------------
    def _apply(self, fn, recurse=True):
        # .to() / .cuda() / .float() replace buffer objects: drop the cached tensor list the fingerprint walks
        self._fp_tensors = None
        return super()._apply(fn, recurse)

    def _fingerprint(self):
        """(storage address, in-place version) of every parameter and buffer: changes whenever the weights the engine was
        packed from may have changed.  Called on every forward (eval.py:43 runs batch 1), so the tensor list is cached
        instead of rebuilding ``state_dict()``; ``_apply`` and ``load_state_dict`` invalidate it."""
        ts = self.__dict__.get('_fp_tensors')      # per module object: a DataParallel replica walks its own tensors
        if ts is None:
            ts = self._fp_tensors = list(self.state_dict(keep_vars=True).values())
        return tuple([(t.data_ptr(), t._version) for t in ts])

    def _engine(self, device: torch.device):
        from fscnn_b200 import Engine
        if self.training:
            raise RuntimeError('the fused inference engine serves eval mode; in training mode FastSCNN.forward runs the training '
                               'operators (fscnn_b200/train_ops.py) and predict / evaluate are not available: call model.eval()')
        if device.type != 'cuda':
            raise RuntimeError(f'FastSCNN (B200 build) runs on CUDA devices only, got a tensor on {device}; move the '
                               'model and the input to the GPU (there is no CPU fallback)')
        rt = self._rt
        eng = rt.engines.get(device)
        if eng is None or eng.precision != self.precision:
            eng = rt.engines[device] = Engine(self.num_classes, self.aux, self.precision)
            rt.fingerprints.pop(device, None)
        fp = self._fingerprint()
        if rt.fingerprints.get(device) != fp:
            eng.load_state_dict(self.state_dict(keep_vars=True), device)   # BN folding + repack on the device
            rt.fingerprints[device] = fp
        return eng

    @staticmethod
    def _as_input(x: torch.Tensor) -> torch.Tensor:
        if x.dim() != 4:
            raise ValueError(f'expected an [N,3,H,W] float batch or an [N,H,W,3] uint8 batch, got {tuple(x.shape)}')
        if x.dtype == torch.uint8:
            if x.size(3) != 3:
                raise ValueError(f'uint8 images must be [N,H,W,3], got {tuple(x.shape)}')
            return x.detach().contiguous()
        if x.size(1) != 3:
            raise ValueError(f'expected an [N,3,H,W] batch, got {tuple(x.shape)}')
        return x.detach().to(torch.float32).contiguous()

    # ---- reference API --------------------------------------------------------------------------
    def forward(self, x, normalize=(IMAGENET_MEAN, IMAGENET_STD)):
        """Returns ``(logits,)`` or ``(logits, aux_logits)``: NCHW fp32 at the input resolution
        (reference models/fast_scnn.py:33-46).  ``normalize`` only applies to raw uint8 input."""
        if self.training:
            return self._train_forward(x)
        x = self._as_input(x)
        logits, aux = self._engine(x.device).forward_logits(x, want_aux=self.aux, norm=normalize)
        return (logits, aux) if self.aux else (logits,)

    def _train_forward(self, x):
        """The reference forward (models/fast_scnn.py:33-46) in TRAINING mode: BatchNorm with batch statistics (running statistics
        updated), Dropout active, every operator a CUDA kernel of csrc/train.cu behind torch.autograd.Function wrappers
        (fscnn_b200/train_ops.py), so ``loss.backward()`` works as usual.  fp32 NCHW tensors."""
        from fscnn_b200 import train_ops
        if x.dim() != 4 or x.size(1) != 3 or x.dtype != torch.float32:
            raise ValueError(f'training forward expects a float32 [N,3,H,W] batch, got {x.dtype} {tuple(x.shape)}')
        if not x.is_cuda:
            raise RuntimeError(f'FastSCNN (B200 build) runs on CUDA devices only, got a tensor on {x.device} (there is no CPU fallback)')
        size = x.shape[2:]
        return tuple(train_ops.bilinear_resize(t, size) for t in self._train_forward_lowres(x))

    def _train_forward_lowres(self, x):
        """The training forward up to the heads' low-resolution logits (1/8 scale), before the final x8 resize of :40 / :44.
        ``fscnn_b200.Trainer`` feeds these to the fused upsample + OHEM loss, so the full-resolution logits are never stored."""
        if x.dim() != 4 or x.size(1) != 3 or x.dtype != torch.float32:
            raise ValueError(f'training forward expects a float32 [N,3,H,W] batch, got {x.dtype} {tuple(x.shape)}')
        if not x.is_cuda:
            raise RuntimeError(f'FastSCNN (B200 build) runs on CUDA devices only, got a tensor on {x.device} (there is no CPU fallback)')
        higher = self.learning_to_downsample(x)
        t = self.global_feature_extractor(higher)
        t = self.feature_fusion(higher, t)
        outputs = [self.classifier(t)]
        if self.aux:
            outputs.append(_train_seq(self.auxlayer, higher))
        return tuple(outputs)

    # ---- fused fast paths (additions) ------------------------------------------------------------
    @torch.no_grad()
    def predict(self, x, out_dtype=torch.uint8, out: Optional[torch.Tensor] = None, normalize=(IMAGENET_MEAN, IMAGENET_STD)):
        """``torch.argmax(model(x)[0], 1)`` (eval.py:43-45) without materialising full-resolution
        logits.  ``out_dtype=torch.int64`` reproduces torch.argmax's dtype."""
        x = self._as_input(x)
        return self._engine(x.device).forward_mask(x, out_dtype, out, norm=normalize)

    @torch.no_grad()
    def evaluate(self, x, labels, metric=None, conf: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None,
                 normalize=(IMAGENET_MEAN, IMAGENET_STD)):
        """forward + argmax + ``SegmentationMetric.update`` (eval.py:43-49) in one pass: accumulates this
        batch into an int64 confusion tensor on the device and returns it.  Pass ``metric`` (a
        ``utils.metric.SegmentationMetric``) to accumulate into its device-side state instead."""
        x = self._as_input(x)
        eng = self._engine(x.device)
        if metric is not None:
            conf = metric.device_confusion(x.device)
        elif conf is None:
            conf = torch.zeros(eng.conf_len(), dtype=torch.int64, device=x.device)
        return eng.forward_confusion(x, labels.to(x.device).contiguous(), conf, mask, norm=normalize)


def get_fast_scnn(dataset='citys', pretrained=False, root='./weights', map_cpu=False, **kwargs):
    """Reference models/fast_scnn.py:240-256: class count from the dataset table, optional
    ``root/fast_scnn_<acronym>.pth`` (only 'citys' and 'tusimple' have acronyms, like the reference)."""
    model = FastSCNN(NUM_CLASS[dataset], **kwargs)
    if pretrained:
        path = os.path.join(root, 'fast_scnn_%s.pth' % _ACRONYMS[dataset])
        model.load_state_dict(torch.load(path, map_location='cpu') if map_cpu else torch.load(path))
    return model
