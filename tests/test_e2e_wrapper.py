"""the camera-frame wrapper (reference export_onnx_fixed.EndToEndFastSCNN).  CPU: the numpy restatement
against the fixtures produced by the unmodified reference; GPU: the CUDA path (preprocess kernel -> network -> fused upsample +
resize-back + softmax kernel, through the C ABI) against the same fixtures."""
import os

import numpy as np
import pytest

import e2e_oracle as eo
import fastscnn_oracle as fo
from conftest import GOLDEN

IMAGENET = ([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])
CASES = ['e2e_nc2_n2_90x160_b256', 'e2e_nc19_n1_75x131_b224']


def load(name):
    g = np.load(os.path.join(GOLDEN, name + '.npz'))
    nc, n, h, w, base, wseed, xseed, norm, sm = (int(v) for v in g['meta'])
    sd = fo.make_state_dict(nc, False, wseed)
    sd['classifier.conv.1.bias'] = g['cls_bias']
    frames = eo.make_frames(n, h, w, xseed)
    mean, std = IMAGENET if norm else (None, None)
    return g, sd, frames, nc, (w, h), base, mean, std, bool(sm)


@pytest.mark.parametrize('case', CASES)
def test_oracle_matches_reference_fixture(case):
    g, sd, frames, nc, size, base, mean, std, sm = load(case)
    pre = eo.preprocess(frames, base, mean, std)
    assert np.abs(pre[:, :, ::5, ::7] - g['pre_sample']).max() < 1e-5
    out = eo.end_to_end(sd, frames, size, base, mean, std, sm)
    ref = g['out']
    assert out.shape == ref.shape
    assert np.abs(out - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max())


def test_half_pixel_resize_identity_and_edges():
    x = np.arange(2 * 3 * 5 * 7, dtype=np.float32).reshape(2, 3, 5, 7)
    assert np.array_equal(eo.bilinear_hp(x, 5, 7), x)                      # same size: every source index is exact
    up = eo.bilinear_hp(x, 10, 14)
    assert np.allclose(up[:, :, 0, 0], x[:, :, 0, 0]) and np.allclose(up[:, :, -1, -1], x[:, :, -1, -1])   # clamped borders
    assert eo.softmax(np.zeros((1, 4, 2, 2), np.float32)).sum(1).max() == pytest.approx(1.0)


@pytest.mark.gpu
@pytest.mark.parametrize('precision,tol', [('fp32', 1e-4), ('bf16', 6e-2)])
@pytest.mark.parametrize('case', CASES)
def test_cuda_wrapper_matches_reference_fixture(case, precision, tol):
    import torch
    from helpers import build_model
    from models.end_to_end import EndToEndFastSCNN
    dev = torch.device('cuda', 0)
    g, sd, frames, nc, size, base, mean, std, sm = load(case)
    backbone = build_model(sd, nc, False, dev, precision=precision)
    model = EndToEndFastSCNN(backbone, input_size=size, base_size=base, mean=mean, std=std, apply_softmax=sm).eval().to(dev)
    x = torch.from_numpy(frames).to(dev)
    pre = model.preprocessor(x).cpu().numpy()
    assert np.abs(pre[:, :, ::5, ::7] - g['pre_sample']).max() < 1e-5
    out = model(x).cpu().numpy()
    ref = g['out']
    assert out.shape == ref.shape and np.isfinite(out).all()
    scale = max(1.0, float(np.abs(ref).max()))
    if precision == 'fp32':
        assert np.abs(out - ref).max() <= tol * scale
    else:   # bf16 backbone: logits within the bf16 tolerance; probabilities agree except near class boundaries
        if sm:
            assert (np.abs(out - ref) > 0.25).mean() < 0.05
        else:
            assert np.abs(out - ref).max() <= tol * scale
    assert np.array_equal(out, model(x.float()).cpu().numpy())             # float32 frames take the same path
    if sm:
        assert np.abs(out.sum(1) - 1.0).max() < 1e-5
