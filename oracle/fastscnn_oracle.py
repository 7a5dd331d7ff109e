"""CPU oracle for the Fast-SCNN segmentation forward path.  TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the arithmetic the reference runs for
``FastSCNN.forward`` -> ``torch.argmax`` (reference: models/fast_scnn.py:16-237,
eval.py:43-45).  It exists so that the CUDA path can be checked on a box that has
no copy of the reference.  Nothing in the shipped package may import it: only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do.

Parity status: PINNED.  The reference is pure Python and imports in the build
container, so ``oracle/gen_golden.py`` runs the *unmodified* reference
(``/root/reference/models/fast_scnn.py``) on seeded weights/inputs and stores
its stage-boundary tensors, logits and masks under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this restatement against every one of
them (fp32 tolerance 2e-5 of the tensor's absmax, masks equal outside
near-ties).  The reference itself ships no golden vectors.

All tensors are NCHW numpy arrays, exactly like the reference.  Each function
cites the reference lines whose arithmetic it restates; the third-party
arithmetic (ATen conv / batch_norm / adaptive_avg_pool2d / upsample_bilinear2d,
torch 2.11) is restated from its published definition.
"""
from __future__ import annotations

import numpy as np

BN_EPS = 1e-5  # nn.BatchNorm2d default, used by every BN in models/fast_scnn.py


# --------------------------------------------------------------------------------------
# Primitive ops (ATen semantics restated)
# --------------------------------------------------------------------------------------
def conv_out_size(n: int, k: int, s: int, p: int) -> int:
    """floor((n + 2p - k)/s) + 1  -- nn.Conv2d output size rule."""
    return (n + 2 * p - k) // s + 1


def conv3x3_dense(x, w, stride, pad, bias=None):
    """Dense 3x3 convolution, ``nn.Conv2d(cin, cout, 3, stride, pad)``.

    Used by LearningToDownsample.conv (fast_scnn.py:153: 3->32, stride 2, pad 0) and
    the aux head (fast_scnn.py:26: 64->32, stride 1, pad 1).
    """
    n, c, h, wd = x.shape
    co = w.shape[0]
    if pad:
        x = np.pad(x, ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    ho = conv_out_size(h, 3, stride, pad)
    wo = conv_out_size(wd, 3, stride, pad)
    out = np.zeros((n, co, ho, wo), dtype=x.dtype)
    wmat = w.reshape(co, c, 9)
    for ky in range(3):
        for kx in range(3):
            patch = x[:, :, ky:ky + stride * (ho - 1) + 1:stride, kx:kx + stride * (wo - 1) + 1:stride]
            # [n,c,ho,wo] x [co,c] -> [n,co,ho,wo]
            out += np.einsum('nchw,oc->nohw', patch, wmat[:, :, ky * 3 + kx], optimize=True)
    if bias is not None:
        out += bias.reshape(1, -1, 1, 1)
    return out


def dwconv3x3(x, w, stride):
    """Depthwise 3x3, pad 1: ``nn.Conv2d(c, c, 3, stride, 1, groups=c, bias=False)``
    (fast_scnn.py:70 in _DSConv, :86 in _DWConv)."""
    n, c, h, wd = x.shape
    xp = np.pad(x, ((0, 0), (0, 0), (1, 1), (1, 1)))
    ho = conv_out_size(h, 3, stride, 1)
    wo = conv_out_size(wd, 3, stride, 1)
    out = np.zeros((n, c, ho, wo), dtype=x.dtype)
    wk = w.reshape(c, 9)
    for ky in range(3):
        for kx in range(3):
            patch = xp[:, :, ky:ky + stride * (ho - 1) + 1:stride, kx:kx + stride * (wo - 1) + 1:stride]
            out += patch * wk[:, ky * 3 + kx].reshape(1, c, 1, 1)
    return out


def pwconv(x, w, bias=None):
    """Pointwise 1x1 convolution ``nn.Conv2d(cin, cout, 1)``: a per-pixel matrix product."""
    n, c, h, wd = x.shape
    co = w.shape[0]
    xm = x.transpose(0, 2, 3, 1).reshape(-1, c)
    om = xm @ w.reshape(co, c).T
    if bias is not None:
        om = om + bias.reshape(1, co)
    return np.ascontiguousarray(om.reshape(n, h, wd, co).transpose(0, 3, 1, 2))


def batchnorm(x, p):
    """Eval-mode ``nn.BatchNorm2d``: (x - mean) / sqrt(var + 1e-5) * gamma + beta."""
    g, b, m, v = (p[k].astype(x.dtype).reshape(1, -1, 1, 1) for k in ('weight', 'bias', 'running_mean', 'running_var'))
    return (x - m) / np.sqrt(v + x.dtype.type(BN_EPS)) * g + b


def relu(x):
    return np.maximum(x, 0)


def adaptive_avg_pool(x, s):
    """``nn.AdaptiveAvgPool2d(s)`` (fast_scnn.py:130-132): bin i spans
    [floor(i*n/s), ceil((i+1)*n/s)); bins overlap when n % s != 0."""
    n, c, h, w = x.shape
    out = np.empty((n, c, s, s), dtype=x.dtype)
    for i in range(s):
        y0, y1 = (i * h) // s, -((-(i + 1) * h) // s)
        for j in range(s):
            x0, x1 = (j * w) // s, -((-(j + 1) * w) // s)
            out[:, :, i, j] = x[:, :, y0:y1, x0:x1].mean(axis=(2, 3), dtype=x.dtype)
    return out


def _ac_coords(n_in, n_out, dtype):
    """Source index / weight for bilinear ``align_corners=True`` (ATen
    area_pixel_compute_scale + compute_source_index): scale=(in-1)/(out-1) in the
    tensor dtype, src=scale*dst, i0=floor(src), i1=min(i0+1,in-1), lam=src-i0."""
    ft = np.dtype(dtype).type
    scale = ft(n_in - 1) / ft(n_out - 1) if n_out > 1 else ft(0)
    src = (scale * np.arange(n_out).astype(dtype)).astype(dtype)
    i0 = np.minimum(src.astype(np.int64), n_in - 1)
    i1 = np.minimum(i0 + 1, n_in - 1)
    lam = (src - i0.astype(dtype)).astype(dtype)
    return i0, i1, lam


def bilinear_ac(x, out_h, out_w):
    """``F.interpolate(x, (out_h,out_w), mode='bilinear', align_corners=True)``
    (fast_scnn.py:40, :44, :135, :212)."""
    dt = x.dtype
    y0, y1, ly = _ac_coords(x.shape[2], out_h, dt)
    x0, x1, lx = _ac_coords(x.shape[3], out_w, dt)
    ly = ly.reshape(1, 1, -1, 1)
    lx = lx.reshape(1, 1, 1, -1)
    one = dt.type(1)
    top = x[:, :, y0][:, :, :, x0] * (one - lx) + x[:, :, y0][:, :, :, x1] * lx
    bot = x[:, :, y1][:, :, :, x0] * (one - lx) + x[:, :, y1][:, :, :, x1] * lx
    return (one - ly) * top + ly * bot


# --------------------------------------------------------------------------------------
# Module graph (models/fast_scnn.py)
# --------------------------------------------------------------------------------------
class _SD:
    """Read-only view on a state_dict of numpy arrays, cast to the compute dtype."""

    def __init__(self, sd, dtype):
        self.sd, self.dtype = sd, dtype

    def w(self, key):
        return np.asarray(self.sd[key]).astype(self.dtype)

    def opt(self, key):
        return np.asarray(self.sd[key]).astype(self.dtype) if key in self.sd else None

    def bn(self, prefix):
        return {k: np.asarray(self.sd[f'{prefix}.{k}']) for k in ('weight', 'bias', 'running_mean', 'running_var')}


def _conv_bn_relu_1x1(sd, prefix, x):
    """_ConvBNReLU with kernel 1 (fast_scnn.py:49-61)."""
    return relu(batchnorm(pwconv(x, sd.w(prefix + '.conv.0.weight')), sd.bn(prefix + '.conv.1')))


def _dsconv(sd, prefix, x, stride):
    """_DSConv (fast_scnn.py:64-79): DW3x3+BN+ReLU then PW1x1+BN+ReLU."""
    x = relu(batchnorm(dwconv3x3(x, sd.w(prefix + '.conv.0.weight'), stride), sd.bn(prefix + '.conv.1')))
    return relu(batchnorm(pwconv(x, sd.w(prefix + '.conv.3.weight')), sd.bn(prefix + '.conv.4')))


def _bottleneck(sd, prefix, x, stride, cout):
    """LinearBottleneck (fast_scnn.py:95-115)."""
    cin = x.shape[1]
    y = _conv_bn_relu_1x1(sd, prefix + '.block.0', x)
    y = relu(batchnorm(dwconv3x3(y, sd.w(prefix + '.block.1.conv.0.weight'), stride), sd.bn(prefix + '.block.1.conv.1')))
    y = batchnorm(pwconv(y, sd.w(prefix + '.block.2.weight')), sd.bn(prefix + '.block.3'))
    if stride == 1 and cin == cout:
        y = x + y
    return y


def _ppm(sd, prefix, x):
    """PyramidPooling (fast_scnn.py:118-145)."""
    h, w = x.shape[2:]
    feats = [x]
    for i, s in enumerate((1, 2, 3, 6), start=1):
        f = _conv_bn_relu_1x1(sd, f'{prefix}.conv{i}', adaptive_avg_pool(x, s))
        feats.append(bilinear_ac(f, h, w))
    return _conv_bn_relu_1x1(sd, prefix + '.out', np.concatenate(feats, axis=1))


def _ffm(sd, prefix, higher, lower):
    """FeatureFusionModule (fast_scnn.py:190-218)."""
    lower = bilinear_ac(lower, higher.shape[2], higher.shape[3])
    lower = relu(batchnorm(dwconv3x3(lower, sd.w(prefix + '.dwconv.conv.0.weight'), 1), sd.bn(prefix + '.dwconv.conv.1')))
    lower = batchnorm(pwconv(lower, sd.w(prefix + '.conv_lower_res.0.weight'), sd.w(prefix + '.conv_lower_res.0.bias')),
                      sd.bn(prefix + '.conv_lower_res.1'))
    hi = batchnorm(pwconv(higher, sd.w(prefix + '.conv_higher_res.0.weight'), sd.w(prefix + '.conv_higher_res.0.bias')),
                   sd.bn(prefix + '.conv_higher_res.1'))
    return relu(hi + lower)


BOTTLENECK_PLAN = (  # (name, cout, stride) from GlobalFeatureExtractor._make_layer, fast_scnn.py:170-180
    ('bottleneck1.0', 64, 2), ('bottleneck1.1', 64, 1), ('bottleneck1.2', 64, 1),
    ('bottleneck2.0', 96, 2), ('bottleneck2.1', 96, 1), ('bottleneck2.2', 96, 1),
    ('bottleneck3.0', 128, 1), ('bottleneck3.1', 128, 1), ('bottleneck3.2', 128, 1),
)


def forward(state_dict, x, aux=False, dtype=np.float32, taps=None, full_res=True):
    """FastSCNN.forward (fast_scnn.py:33-46) in eval mode.

    state_dict: mapping name -> numpy array with the reference's key names (an optional
    ``module.`` prefix must already be stripped).  Returns a tuple like the reference
    (logits[, aux_logits]).  ``taps``: optional dict filled with the stage-boundary
    tensors.  ``full_res=False`` skips the final upsample (returns low-res logits).
    """
    sd = _SD(state_dict, dtype)
    x = np.asarray(x).astype(dtype)
    size = x.shape[2:]

    def tap(name, t):
        if taps is not None:
            taps[name] = t
        return t

    # LearningToDownsample (fast_scnn.py:148-161)
    p = 'learning_to_downsample'
    t = relu(batchnorm(conv3x3_dense(x, sd.w(p + '.conv.conv.0.weight'), 2, 0), sd.bn(p + '.conv.conv.1')))
    tap('l2d.conv', t)
    t = tap('l2d.dsconv1', _dsconv(sd, p + '.dsconv1', t, 2))
    higher = tap('l2d.dsconv2', _dsconv(sd, p + '.dsconv2', t, 2))

    # GlobalFeatureExtractor (fast_scnn.py:164-187)
    p = 'global_feature_extractor'
    t = higher
    for name, cout, stride in BOTTLENECK_PLAN:
        t = tap('gfe.' + name, _bottleneck(sd, f'{p}.{name}', t, stride, cout))
    t = tap('gfe.ppm', _ppm(sd, p + '.ppm', t))

    t = tap('ffm', _ffm(sd, 'feature_fusion', higher, t))

    # Classifer (fast_scnn.py:221-237); Dropout is the identity in eval mode.
    t = tap('cls.dsconv1', _dsconv(sd, 'classifier.dsconv1', t, 1))
    t = tap('cls.dsconv2', _dsconv(sd, 'classifier.dsconv2', t, 1))
    low = tap('cls.logits_lowres', pwconv(t, sd.w('classifier.conv.1.weight'), sd.w('classifier.conv.1.bias')))
    outs = [bilinear_ac(low, *size) if full_res else low]
    if aux:
        a = conv3x3_dense(higher, sd.w('auxlayer.0.weight'), 1, 1)
        a = relu(batchnorm(a, sd.bn('auxlayer.1')))
        a = tap('aux.logits_lowres', pwconv(a, sd.w('auxlayer.4.weight'), sd.w('auxlayer.4.bias')))
        outs.append(bilinear_ac(a, *size) if full_res else a)
    return tuple(outs)


def argmax_classes(logits):
    """``torch.argmax(outputs[0], 1)`` (eval.py:45): first maximal index wins; NaN is maximal."""
    nan = np.isnan(logits)
    if nan.any():
        logits = np.where(nan, np.inf, logits)
    return np.argmax(logits, axis=1).astype(np.int64)


def upsample_argmax(low_logits, out_h, out_w, block_rows=64):
    """argmax over classes of the bilinear (align_corners) upsample of low-res logits,
    computed in row blocks so the full-resolution logits are never materialised."""
    n = low_logits.shape[0]
    dt = low_logits.dtype
    y0, y1, ly = _ac_coords(low_logits.shape[2], out_h, dt)
    x0, x1, lx = _ac_coords(low_logits.shape[3], out_w, dt)
    lx = lx.reshape(1, 1, 1, -1)
    one = dt.type(1)
    mask = np.empty((n, out_h, out_w), dtype=np.int64)
    for r0 in range(0, out_h, block_rows):
        r1 = min(out_h, r0 + block_rows)
        a, b = low_logits[:, :, y0[r0:r1]], low_logits[:, :, y1[r0:r1]]
        top = a[:, :, :, x0] * (one - lx) + a[:, :, :, x1] * lx
        bot = b[:, :, :, x0] * (one - lx) + b[:, :, :, x1] * lx
        l = ly[r0:r1].reshape(1, 1, -1, 1)
        mask[:, r0:r1] = argmax_classes((one - l) * top + l * bot)
    return mask


def top2_margin(logits):
    """Per-pixel gap between the best and second-best class (for near-tie exclusion)."""
    part = np.partition(logits, -2, axis=1)
    return part[:, -1] - part[:, -2]


# --------------------------------------------------------------------------------------
# Seeded synthetic weights / inputs (SURVEY.md Appendix D recipe D2, numpy RNG so that the
# same tensors can be rebuilt on any box without torch's generator)
# --------------------------------------------------------------------------------------
def state_dict_spec(num_classes, aux=False):
    """Ordered (name, shape, kind) list for the reference state_dict layout
    (fast_scnn.py:16-31, 148-237). kind in
    {'conv','cbias','bn_w','bn_b','bn_m','bn_v','bn_n'}."""
    spec = []

    def conv(name, co, ci, k, bias=False):
        spec.append((name + '.weight', (co, ci, k, k), 'conv'))
        if bias:
            spec.append((name + '.bias', (co,), 'cbias'))

    def bn(name, c):
        spec.extend([(name + '.weight', (c,), 'bn_w'), (name + '.bias', (c,), 'bn_b'),
                     (name + '.running_mean', (c,), 'bn_m'), (name + '.running_var', (c,), 'bn_v'),
                     (name + '.num_batches_tracked', (), 'bn_n')])

    p = 'learning_to_downsample'
    conv(p + '.conv.conv.0', 32, 3, 3); bn(p + '.conv.conv.1', 32)
    for name, ci, co in (('dsconv1', 32, 48), ('dsconv2', 48, 64)):
        conv(f'{p}.{name}.conv.0', ci, 1, 3); bn(f'{p}.{name}.conv.1', ci)
        conv(f'{p}.{name}.conv.3', co, ci, 1); bn(f'{p}.{name}.conv.4', co)
    p = 'global_feature_extractor'
    cin = 64
    for name, cout, _ in BOTTLENECK_PLAN:
        q = f'{p}.{name}.block'
        conv(q + '.0.conv.0', cin * 6, cin, 1); bn(q + '.0.conv.1', cin * 6)
        conv(q + '.1.conv.0', cin * 6, 1, 3); bn(q + '.1.conv.1', cin * 6)
        conv(q + '.2', cout, cin * 6, 1); bn(q + '.3', cout)
        cin = cout
    for i in range(1, 5):
        conv(f'{p}.ppm.conv{i}.conv.0', 32, 128, 1); bn(f'{p}.ppm.conv{i}.conv.1', 32)
    conv(p + '.ppm.out.conv.0', 128, 256, 1); bn(p + '.ppm.out.conv.1', 128)
    p = 'feature_fusion'
    conv(p + '.dwconv.conv.0', 128, 1, 3); bn(p + '.dwconv.conv.1', 128)
    conv(p + '.conv_lower_res.0', 128, 128, 1, bias=True); bn(p + '.conv_lower_res.1', 128)
    conv(p + '.conv_higher_res.0', 128, 64, 1, bias=True); bn(p + '.conv_higher_res.1', 128)
    for name in ('dsconv1', 'dsconv2'):
        conv(f'classifier.{name}.conv.0', 128, 1, 3); bn(f'classifier.{name}.conv.1', 128)
        conv(f'classifier.{name}.conv.3', 128, 128, 1); bn(f'classifier.{name}.conv.4', 128)
    conv('classifier.conv.1', num_classes, 128, 1, bias=True)
    if aux:
        conv('auxlayer.0', 32, 64, 3); bn('auxlayer.1', 32)
        conv('auxlayer.4', num_classes, 32, 1, bias=True)
    return spec


def make_state_dict(num_classes, aux=False, seed=7):
    """Variance-preserving random weights with non-trivial BN statistics (recipe D2):
    conv ~ N(0, sqrt(2/fan_in)), conv bias ~ N(0, .05), BN gamma ~ U(.8,1.2), beta ~ N(0,.1),
    mean ~ N(0,.1), var ~ U(.8,1.2).  ``np.random.RandomState`` streams are frozen by numpy,
    so every box rebuilds identical tensors."""
    rng = np.random.RandomState(seed)
    sd = {}
    for name, shape, kind in state_dict_spec(num_classes, aux):
        if kind == 'conv':
            fan_in = shape[1] * shape[2] * shape[3]
            v = rng.standard_normal(shape) * np.sqrt(2.0 / fan_in)
        elif kind == 'cbias':
            v = rng.standard_normal(shape) * 0.05
        elif kind == 'bn_w' or kind == 'bn_v':
            v = rng.uniform(0.8, 1.2, shape)
        elif kind in ('bn_b', 'bn_m'):
            v = rng.standard_normal(shape) * 0.1
        else:
            sd[name] = np.array(0, dtype=np.int64)
            continue
        sd[name] = v.astype(np.float32)
    return sd


def make_input(n, h, w, seed=11):
    """Multi-scale smooth noise image batch [n,3,h,w] fp32 (recipe D2): sum of bilinear
    upsampled Gaussian noise at strides 64/16/4/1 with amplitudes 1/.6/.3/.15."""
    rng = np.random.RandomState(seed)
    x = np.zeros((n, 3, h, w), dtype=np.float32)
    for stride, amp in ((64, 1.0), (16, 0.6), (4, 0.3), (1, 0.15)):
        hs, ws = max(2, -(-h // stride) + 1), max(2, -(-w // stride) + 1)
        z = rng.standard_normal((n, 3, hs, ws)).astype(np.float32)
        x += np.float32(amp) * (z if stride == 1 and (hs, ws) == (h, w) else bilinear_ac(z, h, w)).astype(np.float32)
    return x


def make_labels(n, h, w, num_classes, seed=13, adversarial=False):
    """int64 labels uniform in [-1, nc-1] (-1 = ignore).  ``adversarial`` adds values
    >= nc and < -1, which metric.py:73-105 must handle."""
    rng = np.random.RandomState(seed)
    lo, hi = (-3, num_classes + 2) if adversarial else (-1, num_classes)
    return rng.randint(lo, hi, size=(n, h, w)).astype(np.int64)


def calibrate_classifier_bias(state_dict, x, dtype=np.float32):
    """Recipe D2's last step: subtract the per-class mean logit from classifier.conv.1.bias
    so that every class wins somewhere (otherwise mask parity is vacuous)."""
    low = forward(state_dict, x, aux=False, dtype=dtype, full_res=False)[0]
    sd = dict(state_dict)
    sd['classifier.conv.1.bias'] = (sd['classifier.conv.1.bias'] - low.mean(axis=(0, 2, 3))).astype(np.float32)
    return sd
