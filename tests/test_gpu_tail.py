"""The tail stage on its own (fscnn_upsample_argmax): x8 align_corners bilinear upsample (models/fast_scnn.py:40) fused with
torch.argmax(outputs[0], 1) (eval.py:45) and SegmentationMetric's counting (utils/metric.py:73-105), driven with crafted
low-resolution logits.

The kernel prunes classes by pairwise dominance per low-resolution cell (head.cu); the pruning must be EXACT: the class
map is compared bit for bit with the same kernel run exhaustively (FSCNN_TAIL_EXHAUSTIVE), with torch.argmax of torch's
own interpolation outside near-ties, and with the numpy oracle on small cases; metric counts are bit-exact."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import fastscnn_oracle as fo
import metric_oracle as mo

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda', 0)


def _engine(nc):
    from fscnn_b200.engine import Engine
    return Engine(nc, False, 'fp32')


def _nhwc_padded(low_nchw):
    """[N,nc,hl,wl] -> contiguous NHWC float32 with the class dim padded to a multiple of 4 (zeros, like the head kernel)."""
    n, nc, hl, wl = low_nchw.shape
    ncp = (nc + 3) // 4 * 4
    out = torch.zeros((n, hl, wl, ncp), dtype=torch.float32, device=low_nchw.device)
    out[..., :nc] = low_nchw.permute(0, 2, 3, 1)
    return out.contiguous()


def _smooth(rng, n, nc, hl, wl, strides=(16, 4, 1), amps=(1.0, 0.5, 0.1)):
    x = np.zeros((n, nc, hl, wl), np.float32)
    for s, a in zip(strides, amps):
        hs, ws = max(2, -(-hl // s) + 1), max(2, -(-wl // s) + 1)
        z = rng.standard_normal((n, nc, hs, ws)).astype(np.float32)
        x += np.float32(a) * fo.bilinear_ac(z, hl, wl).astype(np.float32)
    return x


def _cases(nc, hl, wl, rng):
    n = 2
    idx = np.arange(nc, dtype=np.float32)[None, :, None, None]
    checker = ((np.arange(hl)[:, None] + np.arange(wl)[None, :]) % 2 * 2 - 1).astype(np.float32)[None, None]
    smooth = _smooth(rng, n, nc, hl, wl)
    yield 'smooth regions', smooth
    yield 'iid noise', rng.standard_normal((n, nc, hl, wl)).astype(np.float32)
    yield 'all tied', np.zeros((n, nc, hl, wl), np.float32)
    yield 'tied constant 3.25', np.full((n, nc, hl, wl), 3.25, np.float32)
    yield 'checkerboard order flip (nothing dominates)', np.broadcast_to(idx * checker, (n, nc, hl, wl)).copy()
    yield 'last class wins by one ulp', np.broadcast_to(np.where(idx == nc - 1, np.nextafter(np.float32(1000.0), np.float32(2000.0)),
                                                               np.float32(1000.0)), (n, nc, hl, wl)).astype(np.float32).copy()
    yield 'huge offset, tiny gaps', (1e4 + 1e-3 * smooth).astype(np.float32)
    yield 'tiny values', (1e-30 * smooth).astype(np.float32)
    two = np.full((n, nc, hl, wl), -5.0, np.float32)
    ramp = np.linspace(-1, 1, wl, dtype=np.float32)[None, None, :]
    two[:, 0] = ramp
    two[:, nc - 1] = -ramp
    yield 'two-class boundary', two
    q = np.round(smooth * 2) / 2
    yield 'quantised (many exact ties)', q.astype(np.float32)


@pytest.mark.parametrize('nc,hl,wl,h,w', [(19, 23, 40, 180, 317), (2, 45, 80, 360, 640), (32, 12, 20, 96, 160), (5, 9, 13, 65, 97)])
def test_pruned_argmax_equals_exhaustive_and_torch(nc, hl, wl, h, w):
    eng = _engine(nc)
    rng = np.random.RandomState(nc * 1000 + hl)
    for name, low in _cases(nc, hl, wl, rng):
        lowd = torch.from_numpy(low).to(DEV)
        packed = _nhwc_padded(lowd)
        pruned = eng.upsample_argmax(packed, h, w)
        exhaustive = eng.upsample_argmax(packed, h, w, exhaustive=True)
        assert torch.equal(pruned, exhaustive), name
        up = F.interpolate(lowd, size=(h, w), mode='bilinear', align_corners=True)
        ref = up.argmax(1)
        top2 = up.topk(2, dim=1).values
        near_tie = (top2[:, 0] - top2[:, 1]) <= 1e-5 * up.abs().amax().clamp_min(1e-37)
        bad = (pruned.long() != ref) & ~near_tie
        assert int(bad.sum()) == 0, (name, int(bad.sum()))
        # where the interpolated values tie exactly (constant inputs), the first maximal class must win
        if name.startswith(('all tied', 'tied constant')):
            assert int(pruned.max()) == 0, name
    # int64 / int32 class maps carry the same values
    assert torch.equal(eng.upsample_argmax(packed, h, w, out_dtype=torch.int64), pruned.long())
    assert torch.equal(eng.upsample_argmax(packed, h, w, out_dtype=torch.int32), pruned.int())


def test_tail_matches_numpy_oracle():
    nc, hl, wl, h, w = 7, 8, 12, 57, 89
    rng = np.random.RandomState(3)
    low = _smooth(rng, 1, nc, hl, wl, strides=(4, 1), amps=(1.0, 0.3))
    want = fo.upsample_argmax(low, h, w)
    up = fo.bilinear_ac(low, h, w)
    near_tie = fo.top2_margin(up) < 1e-5 * np.abs(up).max()
    got = _engine(nc).upsample_argmax(_nhwc_padded(torch.from_numpy(low).to(DEV)), h, w).cpu().numpy()
    assert int(((got != want) & ~near_tie).sum()) == 0


def test_more_than_32_classes_and_nonfinite_take_the_exact_loop():
    for nc in (40, 150):
        rng = np.random.RandomState(nc)
        low = torch.from_numpy(_smooth(rng, 1, nc, 10, 14)).to(DEV)
        got = _engine(nc).upsample_argmax(_nhwc_padded(low), 73, 105)
        up = F.interpolate(low, size=(73, 105), mode='bilinear', align_corners=True)
        top2 = up.topk(2, dim=1).values
        near_tie = (top2[:, 0] - top2[:, 1]) <= 1e-5 * up.abs().amax()
        assert int(((got.long() != up.argmax(1)) & ~near_tie).sum()) == 0
    nc = 6
    low = torch.from_numpy(_smooth(np.random.RandomState(1), 1, nc, 10, 14)).to(DEV)
    low[0, 4, 3, 5] = float('nan')
    low[0, 2, 7, 9] = float('inf')
    low[0, 1, 0, 0] = float('-inf')
    eng = _engine(nc)
    got = eng.upsample_argmax(_nhwc_padded(low), 73, 105)
    up = F.interpolate(low, size=(73, 105), mode='bilinear', align_corners=True)
    ref = up.argmax(1)            # torch: NaN is maximal, first one wins
    finite = torch.isfinite(up).all(1)
    top2 = torch.nan_to_num(up, nan=0.0, posinf=0.0, neginf=0.0).topk(2, dim=1).values
    near_tie = (top2[:, 0] - top2[:, 1]) <= 1e-5
    nanpix = torch.isnan(up).any(1)
    assert torch.equal(got.long()[nanpix], ref[nanpix])
    assert int(((got.long() != ref) & finite & ~near_tie).sum()) == 0
    assert torch.equal(got, eng.upsample_argmax(_nhwc_padded(low), 73, 105, exhaustive=True))


@pytest.mark.parametrize('label_dtype', [torch.int64, torch.int32, torch.uint8])
def test_tail_counts_are_bit_exact(label_dtype):
    nc, hl, wl, h, w = 19, 23, 40, 180, 316
    rng = np.random.RandomState(5)
    low = torch.from_numpy(_smooth(rng, 3, nc, hl, wl)).to(DEV)
    labels = fo.make_labels(3, h, w, nc, seed=9, adversarial=(label_dtype != torch.uint8))
    if label_dtype == torch.uint8:
        labels = np.where(labels < 0, 255, labels)
    eng = _engine(nc)
    conf = torch.zeros(eng.conf_len(), dtype=torch.int64, device=DEV)
    ld = torch.from_numpy(labels).to(DEV).to(label_dtype)
    mask = eng.upsample_argmax(_nhwc_padded(low), h, w, labels=ld, conf=conf)
    eng.upsample_argmax(_nhwc_padded(low), h, w, labels=ld, conf=conf, want_mask=False)      # accumulates; no mask written
    assert np.array_equal(conf.cpu().numpy(), 2 * mo.confusion_counts(mask.cpu().numpy(), labels, nc))


def test_unaligned_class_maps_are_refused():
    """Class maps are accessed four elements at a time: a view that starts at an unaligned storage offset must be refused
    with an error instead of faulting on the device (include/fscnn_b200.h)."""
    from fscnn_b200 import NativeError
    from helpers import build_model
    from utils.metric import SegmentationMetric
    nc, n, h, w = 3, 1, 64, 96
    model = build_model(fo.make_state_dict(nc, False, 1), nc, False, DEV)
    x = torch.from_numpy(fo.make_input(n, h, w, 2)).to(DEV)
    store = torch.zeros(n * h * w + 8, dtype=torch.int64, device=DEV)
    lab = store[1:1 + n * h * w].view(n, h, w)                   # 8-byte offset: not 16-byte aligned
    with pytest.raises((ValueError, NativeError)):
        model.evaluate(x, lab, SegmentationMetric(nc))
    bad_out = torch.zeros(n * h * w + 8, dtype=torch.uint8, device=DEV)[1:1 + n * h * w].view(n, h, w)
    with pytest.raises((ValueError, NativeError)):
        model.predict(x, out=bad_out)
    with pytest.raises(ValueError):
        model.predict(x, out=torch.zeros((n, h, w + 1), dtype=torch.uint8, device=DEV))      # wrong shape
    with pytest.raises(ValueError):
        model.predict(x, out=torch.zeros((n, h, w), dtype=torch.float32, device=DEV))        # wrong dtype
    ok = store[2:2 + n * h * w].view(n, h, w)                    # 16-byte offset is fine
    model.evaluate(x, ok, SegmentationMetric(nc))
