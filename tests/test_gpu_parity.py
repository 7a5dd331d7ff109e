"""Parity of the CUDA path (through the C ABI) against the reference-generated golden vectors and
the numpy oracle.  Run on the B200 box: ``pytest -m gpu``.

Tolerances (fp32 path): every stage tensor and the logits within 1e-4 of the tensor's absmax
(BASELINE.json north_star: "fp32 logits within 1e-4 relative"); masks identical outside the pixels
whose top-2 logit margin is below 1e-4 * absmax; metric counts bit-exact."""
import numpy as np
import pytest
import torch

import fastscnn_oracle as fo
import metric_oracle as mo
from helpers import build_model, load_case, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4
DEV = torch.device('cuda', 0)

STAGE_IO = [  # (stage name, input taps, output tap)
    ('stem', None, 'l2d.conv'),
    ('l2d.dsconv1', ['l2d.conv'], 'l2d.dsconv1'),
    ('l2d.dsconv2', ['l2d.dsconv1'], 'l2d.dsconv2'),
    ('gfe.bottleneck1.0', ['l2d.dsconv2'], 'gfe.bottleneck1.0'),
    ('gfe.bottleneck1.1', ['gfe.bottleneck1.0'], 'gfe.bottleneck1.1'),
    ('gfe.bottleneck1.2', ['gfe.bottleneck1.1'], 'gfe.bottleneck1.2'),
    ('gfe.bottleneck2.0', ['gfe.bottleneck1.2'], 'gfe.bottleneck2.0'),
    ('gfe.bottleneck2.1', ['gfe.bottleneck2.0'], 'gfe.bottleneck2.1'),
    ('gfe.bottleneck2.2', ['gfe.bottleneck2.1'], 'gfe.bottleneck2.2'),
    ('gfe.bottleneck3.0', ['gfe.bottleneck2.2'], 'gfe.bottleneck3.0'),
    ('gfe.bottleneck3.1', ['gfe.bottleneck3.0'], 'gfe.bottleneck3.1'),
    ('gfe.bottleneck3.2', ['gfe.bottleneck3.1'], 'gfe.bottleneck3.2'),
    ('gfe.ppm', ['gfe.bottleneck3.2'], 'gfe.ppm'),
    ('ffm', ['l2d.dsconv2', 'gfe.ppm'], 'ffm'),
    ('cls.dsconv1', ['ffm'], 'cls.dsconv1'),
    ('cls.dsconv2+head', ['cls.dsconv1'], 'cls.logits_lowres'),
    ('aux', ['l2d.dsconv2'], 'aux.logits_lowres'),
]


def nhwc(t):
    return torch.from_numpy(np.ascontiguousarray(t.transpose(0, 2, 3, 1)))


@pytest.mark.parametrize('case', ['fwd_nc19_aux_n2_65x97', 'fwd_nc3_aux_n3_64x40'])
def test_each_stage_in_isolation(case):
    """Feed the REFERENCE's tensor into each stage and compare that stage's output alone."""
    g, sd, x, nc, aux = load_case(case)
    model = build_model(sd, nc, aux, DEV)
    xd = torch.from_numpy(x).to(DEV)
    eng = model._engine(DEV)
    n, _, h, w = x.shape
    names = eng.stage_names()
    failures = []
    for stage, ins, out in STAGE_IO:
        idx = names.index(stage)
        for tap in ins or []:
            eng.tap_view(tap, n, h, w).copy_(nhwc(g['tap/' + tap]).to(DEV))
        eng.forward_range(xd, idx, idx)
        got = eng.tap_view(out, n, h, w).permute(0, 3, 1, 2).float().cpu().numpy()
        err = rel_err(got, g['tap/' + out])
        if not err < TOL:
            failures.append((stage, err))
    assert not failures, failures


@pytest.mark.parametrize('case', ['fwd_nc19_aux_n2_65x97', 'fwd_nc3_aux_n3_64x40', 'fwd_nc2_n1_360x640', 'fwd_nc19_n1_256x512'])
def test_forward_matches_reference(case):
    g, sd, x, nc, aux = load_case(case)
    model = build_model(sd, nc, aux, DEV)
    xd = torch.from_numpy(x).to(DEV)
    outs = model(xd)
    assert isinstance(outs, tuple) and len(outs) == (2 if aux else 1)
    logits = outs[0].cpu().numpy()
    n, _, h, w = x.shape
    assert logits.shape == (n, nc, h, w)
    if 'logits' in g.files:
        assert rel_err(logits, g['logits']) < TOL
        near_tie = g['margin'].astype(np.float32) < 1e-4 * float(g['logits_absmax'])
    else:
        scale = float(g['logits_absmax'])
        assert np.abs(logits[:, :, ::7, ::11] - g['logits_sample']).max() / scale < TOL
        assert np.abs(logits[:, :, 40:72, 96:160] - g['logits_window']).max() / scale < TOL
        near_tie = np.unpackbits(g['margin_small'])[:n * h * w].reshape(n, h, w).astype(bool)
    if aux:
        assert np.abs(outs[1].cpu().numpy()[:, :, ::3, ::5] - g['aux_logits_sample']).max() / np.abs(g['aux_logits_sample']).max() < TOL
    # torch.argmax on our logits, and the fused upsample+argmax kernel, against the reference mask
    for mask in (torch.argmax(outs[0], 1).cpu().numpy(), model.predict(xd).cpu().numpy(),
                 model.predict(xd, out_dtype=torch.int64).cpu().numpy()):
        differ = (mask != g['mask']) & ~near_tie
        assert not differ.any(), f'{int(differ.sum())} mask pixels differ outside near-ties'


def test_fused_mask_equals_argmax_of_own_logits():
    """Self-consistency at an odd size: the fused kernel and up_logits+argmax interpolate identically."""
    nc = 19
    sd = fo.make_state_dict(nc, False, 3)
    x = fo.make_input(2, 203, 333, 4)
    model = build_model(sd, nc, False, DEV)
    xd = torch.from_numpy(x).to(DEV)
    a = torch.argmax(model(xd)[0], 1)
    for dt in (torch.uint8, torch.int32, torch.int64):
        assert torch.equal(a, model.predict(xd, out_dtype=dt).long())


def test_full_size_against_oracle():
    """BASELINE config 1 (19 classes, 1x3x1024x2048): low-res logits and mask against the numpy oracle."""
    nc, h, w = 19, 1024, 2048
    sd = fo.make_state_dict(nc, False, 7)
    x = fo.make_input(1, h, w, 31)
    sd = fo.calibrate_classifier_bias(sd, x)
    taps = {}
    low = fo.forward(sd, x, full_res=False, taps=taps)[0]
    ref_mask = fo.upsample_argmax(low, h, w)
    model = build_model(sd, nc, False, DEV)
    xd = torch.from_numpy(x).to(DEV)
    mask = model.predict(xd).cpu().numpy()
    eng = model._engine(DEV)
    got_low = eng.tap_view('cls.logits_lowres', 1, h, w).permute(0, 3, 1, 2).cpu().numpy()
    assert rel_err(got_low, low) < TOL
    for tap in ('l2d.conv', 'l2d.dsconv2', 'gfe.bottleneck3.2', 'gfe.ppm', 'ffm'):
        got = eng.tap_view(tap, 1, h, w).permute(0, 3, 1, 2).cpu().numpy()
        assert rel_err(got, taps[tap]) < TOL, tap
    # near ties: recompute the margin from the oracle at full resolution in row blocks
    differ = mask != ref_mask
    if differ.any():
        ys, xs = np.nonzero(differ[0])
        full = fo.bilinear_ac(low, h, w)
        margin = fo.top2_margin(full)[0]
        assert (margin[ys, xs] < 1e-4 * np.abs(full).max()).all(), 'mask differs outside near-ties'
    assert differ.mean() < 1e-3
    counts = np.bincount(mask.reshape(-1), minlength=nc)
    assert (counts > 0).sum() >= 15   # the comparison is not vacuous


def test_batch_invariance_and_micro_batches():
    """Image i of a batch gives bit-identical results alone, and micro-batching does not change them."""
    nc = 2
    sd = fo.make_state_dict(nc, False, 5)
    x = fo.make_input(5, 120, 168, 6)
    model = build_model(sd, nc, False, DEV)
    xd = torch.from_numpy(x).to(DEV)
    full = model(xd)[0]
    eng = model._engine(DEV)
    eng.set_micro_batch(2)   # 5 images -> micro-batches of 2,2,1
    assert torch.equal(model(xd)[0], full)
    assert torch.equal(model.predict(xd), torch.argmax(full, 1).to(torch.uint8))
    eng.set_micro_batch(0)
    for i in (0, 4):
        assert torch.equal(model(xd[i:i + 1].contiguous())[0], full[i:i + 1])


def test_metric_kernel_matches_reference_cases(golden_dir):
    import os
    from utils.metric import SegmentationMetric
    g = np.load(os.path.join(golden_dir, 'metric_cases.npz'))
    for i in range(int(g['ncases'])):
        nc = int(g[f'c{i}/nclass'])
        m = SegmentationMetric(nc, device=DEV)
        m.update(g[f'c{i}/pred'], g[f'c{i}/label'])
        assert np.array_equal(m.total_inter, g[f'c{i}/inter']), i
        assert np.array_equal(m.total_union, g[f'c{i}/union']), i
        assert m.total_correct == int(g[f'c{i}/correct']) and m.total_label == int(g[f'c{i}/labeled'])
        pix, miou = m.get()
        assert pix == float(g[f'c{i}/pixacc']) and miou == float(g[f'c{i}/miou'])
    m = SegmentationMetric(19, device=DEV)
    preds, labels = g['list/pred'], g['list/label']
    m.update([preds[0], preds[1]], [labels[0], labels[1]])
    m.update(torch.from_numpy(preds[2]).to(DEV), torch.from_numpy(labels[2]).to(DEV))   # CUDA tensors work too
    assert np.array_equal(m.total_inter, g['list/inter']) and np.array_equal(m.total_union, g['list/union'])
    assert m.get() == (float(g['list/pixacc']), float(g['list/miou']))
    m.reset()
    assert m.total_label == 0 and m.get() == (0.0, 0.0)


@pytest.mark.parametrize('nc,label_dtype', [(19, torch.int64), (2, torch.uint8), (19, torch.int32)])
def test_fused_evaluate_counts(nc, label_dtype):
    """forward+argmax+histogram in one pass == oracle counting on the mask the GPU produced."""
    from utils.metric import SegmentationMetric
    n, h, w = 3, 200, 264
    sd = fo.make_state_dict(nc, False, 11)
    x = fo.make_input(n, h, w, 12)
    sd = fo.calibrate_classifier_bias(sd, x)
    labels = fo.make_labels(n, h, w, nc, seed=13, adversarial=(label_dtype != torch.uint8))
    if label_dtype == torch.uint8:
        labels = np.where(labels < 0, 255, labels)   # uint8 label maps mark "ignore" as 255 (>= nclass)
    model = build_model(sd, nc, False, DEV)
    xd = torch.from_numpy(x).to(DEV)
    ld = torch.from_numpy(labels).to(DEV).to(label_dtype)
    metric = SegmentationMetric(nc)
    mask = torch.empty((n, h, w), dtype=torch.uint8, device=DEV)
    for _ in range(2):   # accumulates across calls
        model.evaluate(xd, ld, metric, mask=mask)
    assert torch.equal(mask, model.predict(xd))
    ref = 2 * mo.confusion_counts(mask.cpu().numpy(), labels, nc)
    assert np.array_equal(metric.device_confusion().cpu().numpy(), ref)
    inter, union, correct, labeled = mo.totals_from_confusion(ref, nc)
    o = mo.SegmentationMetricOracle(nc)
    o.update(mask.cpu().numpy().astype(np.int64), labels.astype(np.int64))
    o.update(mask.cpu().numpy().astype(np.int64), labels.astype(np.int64))
    assert np.array_equal(metric.total_inter, o.total_inter) and np.array_equal(metric.total_union, o.total_union)
    assert metric.get() == o.get()


def test_nan_and_ties_follow_torch_argmax():
    """argmax semantics: first maximal class wins; NaN counts as maximal."""
    nc = 4
    sd = fo.make_state_dict(nc, False, 2)
    for k in list(sd):
        if k.startswith('classifier.conv.1'):
            sd[k] = np.zeros_like(sd[k])           # all logits equal -> class 0 everywhere
    model = build_model(sd, nc, False, DEV)
    xd = torch.from_numpy(fo.make_input(1, 64, 64, 1)).to(DEV)
    assert int(model.predict(xd).max()) == 0
    sd['classifier.conv.1.bias'] = np.array([0, np.nan, 1, np.nan], dtype=np.float32)
    model = build_model(sd, nc, False, DEV)
    assert torch.equal(model.predict(xd).long(), torch.argmax(model(xd)[0], 1))
    assert int(model.predict(xd).min()) == 1 and int(model.predict(xd).max()) == 1


def test_errors_are_loud():
    from fscnn_b200 import NativeError
    model = build_model(fo.make_state_dict(2, False, 1), 2, False, DEV)
    with pytest.raises(RuntimeError):
        model(torch.zeros(1, 3, 64, 64))                      # CPU tensor: no fallback
    with pytest.raises(NativeError):
        model(torch.zeros(1, 3, 2, 2, device=DEV))            # too small
    model.train()
    with pytest.raises(RuntimeError):
        model.predict(torch.zeros(1, 3, 64, 64, device=DEV))   # training mode runs the training operators; the fused engine is eval only


@pytest.mark.parametrize('shape', [(2, 97, 131, 3), (1, 100, 132, 3)])   # row pitch not / is a multiple of 4 bytes
@pytest.mark.parametrize('precision,tol', [('fp32', 1e-4), ('bf16', 4e-2)])
def test_uint8_input_fuses_totensor_normalize(precision, tol, shape):
    """Raw uint8 HWC images: the stem applies ToTensor + Normalize(mean, std) (eval.py:22-25) on load."""
    from models.fast_scnn import IMAGENET_MEAN, IMAGENET_STD
    nc = 19
    sd = fo.make_state_dict(nc, False, 7)
    rng = np.random.RandomState(4)
    img = rng.randint(0, 256, size=shape).astype(np.uint8)
    mean, std = np.array(IMAGENET_MEAN, np.float32), np.array(IMAGENET_STD, np.float32)
    x = ((img.astype(np.float32) / np.float32(255) - mean) / std).transpose(0, 3, 1, 2).copy()   # what the reference's transforms produce
    ref = fo.forward(sd, x)[0]
    model = build_model(sd, nc, False, DEV, precision=precision)
    got = model(torch.from_numpy(img).to(DEV))[0].cpu().numpy()
    assert rel_err(got, ref) < tol
    # and the /255-only convention of the custom dataset path (data_loader/custom.py:175)
    ref2 = fo.forward(sd, (img.astype(np.float32) / np.float32(255)).transpose(0, 3, 1, 2).copy())[0]
    got2 = model(torch.from_numpy(img).to(DEV), normalize=None)[0].cpu().numpy()
    assert rel_err(got2, ref2) < tol
    # same mask through the fused path, float and uint8 inputs agree
    m1 = model.predict(torch.from_numpy(img).to(DEV))
    m2 = model.predict(torch.from_numpy(x).to(DEV))
    assert (m1 != m2).float().mean().item() < (1e-3 if precision == 'fp32' else 2e-2)


def test_streaming_evaluator_matches_batchwise():
    from fscnn_b200 import StreamingEvaluator
    from utils.metric import SegmentationMetric
    nc = 19
    sd = fo.make_state_dict(nc, False, 7)
    model = build_model(sd, nc, False, DEV)
    rng = np.random.RandomState(9)
    batches = [(torch.from_numpy(rng.randint(0, 256, size=(2, 96, 128, 3)).astype(np.uint8)).pin_memory(),
                torch.from_numpy(rng.randint(0, nc + 1, size=(2, 96, 128)).astype(np.uint8)).pin_memory()) for _ in range(5)]
    m_stream, m_ref = SegmentationMetric(nc, device=DEV), SegmentationMetric(nc, device=DEV)
    ev = StreamingEvaluator(model, m_stream, batches[0][0], batches[0][1], device=DEV)
    for img, lab in batches:
        ev.submit(img, lab)
        model.evaluate(img.to(DEV), lab.to(DEV), m_ref)
    assert ev.result() == m_ref.get()
    assert np.array_equal(m_stream.total_inter, m_ref.total_inter) and m_stream.total_label == 5 * 2 * 96 * 128


def test_argmax_pruning_is_exact_on_adversarial_logits():
    """The fused argmax skips classes that provably cannot win inside a thread's pixel block.  Stress it with logits
    whose classes are nearly tied / cross inside blocks: the mask must still equal argmax of the path's own logits."""
    nc = 19
    rng = np.random.RandomState(17)
    for scale_w, bias_scale in ((1e-3, 0.0), (1.0, 0.0), (0.05, 5.0)):
        sd = fo.make_state_dict(nc, False, 23)
        sd['classifier.conv.1.weight'] = (sd['classifier.conv.1.weight'] * scale_w).astype(np.float32)
        sd['classifier.conv.1.bias'] = (rng.standard_normal(nc) * bias_scale).astype(np.float32)
        model = build_model(sd, nc, False, DEV)
        xd = torch.from_numpy(fo.make_input(2, 264, 392, 5)).to(DEV)
        assert torch.equal(model.predict(xd).long(), torch.argmax(model(xd)[0], 1))
    # all logits identical everywhere: nothing can be pruned, class 0 wins every pixel
    sd['classifier.conv.1.weight'] = np.zeros_like(sd['classifier.conv.1.weight'])
    sd['classifier.conv.1.bias'] = np.zeros(nc, np.float32)
    model = build_model(sd, nc, False, DEV)
    assert int(model.predict(xd).max()) == 0


def test_colorize_kernel():
    from utils.visualize import colorize, palette_for
    rng = np.random.RandomState(2)
    for dt in (torch.uint8, torch.int64):
        m = torch.from_numpy(rng.randint(0, 19, size=(2, 37, 53)).astype(np.int64)).to(dt).to(DEV)
        rgb = colorize(m, 'citys').cpu().numpy()
        assert np.array_equal(rgb, palette_for('citys')[m.cpu().numpy().astype(np.int64)])


def test_colorize_and_overlay_against_reference_vectors():
    """Device palette rendering against the reference's complete tables, and the overlay kernel against create_overlay()
    (demo_tusimple.py:87-104) run from the unmodified reference source (oracle/gen_golden_visual.py), bit for bit."""
    import os
    from conftest import GOLDEN
    from utils.visualize import colorize, overlay
    g = np.load(os.path.join(GOLDEN, 'palettes.npz'))
    m = torch.from_numpy(g['cls_map']).to(DEV)
    assert np.array_equal(colorize(m, 'citys').cpu().numpy(), g['rgb_citys'])
    assert np.array_equal(colorize(m.to(torch.uint8), 'tusimple').cpu().numpy(), g['rgb_voc'])
    allc = torch.arange(256, device=DEV, dtype=torch.int32).view(1, 16, 16)
    assert np.array_equal(colorize(allc, 'tusimple').cpu().numpy()[0].reshape(256, 3), g['voc'])
    o = np.load(os.path.join(GOLDEN, 'overlay.npz'))
    frame = torch.from_numpy(o['frame']).to(DEV)
    lane = torch.from_numpy(o['lane'].astype(np.uint8)).to(DEV)          # class 1 = lane
    for alpha in (0.5, 0.3, 0.85):
        for dt in (torch.uint8, torch.int64):
            got = overlay(frame, lane.to(dt), classes=(1,), colors={1: (0, 255, 0)}, alpha=alpha).cpu().numpy()
            assert np.array_equal(got, o[f'overlay_{alpha}']), (alpha, dt)
    assert torch.equal(overlay(frame, lane, classes=(), alpha=0.5), frame)      # nothing drawn: the frame comes back
