"""bf16 fast path (bf16 stage tensors, tcgen05 tensor-core contractions with fp32 accumulation)
against the reference fixtures and the oracle.

Stated tolerance.  (a) Fixtures: every stage fed with the reference's input within 1.5e-2 of the output's absmax;
end-to-end logits within 4e-2 of absmax; at most 5 % of mask pixels differ.  (b) Seeded configurations against the
oracle (test_bf16_against_the_references_own_bf16, also what __graft_entry__.smoke() checks): anchored on the reference's
OWN bf16 behaviour on the same input (the reference op sequence under torch.autocast(bfloat16), oracle/
fastscnn_torch_port.bf16_yardstick): logit rms error <= 1.25 x the reference's, max error <= 2 x the reference's (and
<= 6e-2), mask disagreement <= the reference's + 0.5 points (and <= 5 %).  Why these: helpers.py / profiles/
r02_bf16_error_budget.md -- the error is ~45 independent bf16 roundings spread evenly over the stages, rms equal to
cuDNN's / ATen's bf16 autocast (0.74-0.99 x), the max over 1e5-1e7 logits fluctuates 0.59-1.73 x at equal rms."""
import numpy as np
import pytest
import torch

import fastscnn_oracle as fo
import metric_oracle as mo
import fastscnn_torch_port as port
from helpers import bf16_errors, build_model, check_bf16_against_yardstick, load_case, rel_err
from test_gpu_parity import STAGE_IO, nhwc

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda', 0)
STAGE_TOL, LOGIT_TOL, MASK_TOL = 1.5e-2, 4e-2, 5e-2


@pytest.mark.parametrize('case', ['fwd_nc19_aux_n2_65x97', 'fwd_nc3_aux_n3_64x40'])
def test_bf16_each_stage_in_isolation(case):
    g, sd, x, nc, aux = load_case(case)
    model = build_model(sd, nc, aux, DEV, precision='bf16')
    xd = torch.from_numpy(x).to(DEV)
    eng = model._engine(DEV)
    n, _, h, w = x.shape
    names = eng.stage_names()
    failures = []
    for stage, ins, out in STAGE_IO:
        idx = names.index(stage)
        for tap in ins or []:
            v = eng.tap_view(tap, n, h, w)
            v.copy_(nhwc(g['tap/' + tap]).to(DEV).to(v.dtype))
        eng.forward_range(xd, idx, idx)
        got = eng.tap_view(out, n, h, w).permute(0, 3, 1, 2).float().cpu().numpy()
        err = rel_err(got, g['tap/' + out])
        if not err < STAGE_TOL:
            failures.append((stage, err))
    assert not failures, failures


@pytest.mark.parametrize('case', ['fwd_nc19_aux_n2_65x97', 'fwd_nc2_n1_360x640', 'fwd_nc19_n1_256x512'])
def test_bf16_forward_and_mask(case):
    g, sd, x, nc, aux = load_case(case)
    model = build_model(sd, nc, aux, DEV, precision='bf16')
    xd = torch.from_numpy(x).to(DEV)
    logits = model(xd)[0].cpu().numpy()
    scale = float(g['logits_absmax'])
    if 'logits' in g.files:
        assert np.abs(logits - g['logits']).max() / scale < LOGIT_TOL
    else:
        assert np.abs(logits[:, :, ::7, ::11] - g['logits_sample']).max() / scale < LOGIT_TOL
    mask = model.predict(xd).cpu().numpy()
    assert (mask != g['mask']).mean() < MASK_TOL
    assert np.array_equal(mask, np.argmax(logits, 1))      # fused argmax == argmax of the path's own logits


@pytest.mark.parametrize('nc,n,h,w,wseed,xseed', [
    (19, 2, 96, 160, 7, 21),        # the configuration __graft_entry__.smoke() runs
    (19, 2, 96, 160, 101, 102),     # unseen seeds
    (19, 1, 360, 640, 103, 104),
    (2, 1, 360, 640, 105, 106),
    (19, 3, 96, 160, 107, 108),
])
def test_bf16_against_the_references_own_bf16(nc, n, h, w, wseed, xseed):
    """Seeded D2 configurations against the numpy ORACLE, with the bound anchored on what the reference itself does in bf16
    on the same input (module docstring, criterion b)."""
    sd = fo.make_state_dict(nc, False, seed=wseed)
    x = fo.make_input(n, h, w, seed=xseed)
    sd = fo.calibrate_classifier_bias(sd, x)
    ref = fo.forward(sd, x)[0]
    yard = port.bf16_yardstick(sd, x, ref)
    model = build_model(sd, nc, False, DEV, precision='bf16')
    xd = torch.from_numpy(x).to(DEV)
    got = model(xd)[0].cpu().numpy()
    ours = bf16_errors(got, ref)
    assert not check_bf16_against_yardstick(ours, yard), (ours, yard)
    mask = model.predict(xd).cpu().numpy()
    assert np.array_equal(mask, np.argmax(got, 1))          # the fused argmax is the argmax of the path's own logits
    assert (mask != fo.argmax_classes(ref)).mean() <= min(yard['mask'] + 0.005, MASK_TOL)


def test_bf16_full_size_and_metric():
    """1x3x1024x2048, 19 classes: mask agreement with the fp32 oracle and exact metric counting."""
    from utils.metric import SegmentationMetric
    nc, h, w = 19, 1024, 2048
    sd = fo.make_state_dict(nc, False, 7)
    x = fo.make_input(1, h, w, 31)
    sd = fo.calibrate_classifier_bias(sd, x)
    low = fo.forward(sd, x, full_res=False)[0]
    ref_mask = fo.upsample_argmax(low, h, w)
    labels = fo.make_labels(1, h, w, nc, seed=3)
    model = build_model(sd, nc, False, DEV, precision='bf16')
    xd = torch.from_numpy(x).to(DEV)
    metric = SegmentationMetric(nc)
    mask = torch.empty((1, h, w), dtype=torch.uint8, device=DEV)
    model.evaluate(xd, torch.from_numpy(labels).to(DEV), metric, mask=mask)
    mask = mask.cpu().numpy()
    assert (mask != ref_mask).mean() < MASK_TOL
    got_low = model._engine(DEV).tap_view('cls.logits_lowres', 1, h, w).permute(0, 3, 1, 2).cpu().numpy()
    assert rel_err(got_low, low) < LOGIT_TOL
    assert np.array_equal(metric.device_confusion().cpu().numpy(), mo.confusion_counts(mask, labels, nc))
    o = mo.SegmentationMetricOracle(nc)
    o.update(mask.astype(np.int64), labels)
    assert metric.get() == o.get()


def test_bf16_batch_invariance():
    nc = 2
    sd = fo.make_state_dict(nc, False, 5)
    x = fo.make_input(5, 120, 168, 6)
    model = build_model(sd, nc, False, DEV, precision='bf16')
    xd = torch.from_numpy(x).to(DEV)
    full = model.predict(xd)
    eng = model._engine(DEV)
    eng.set_micro_batch(2)
    assert torch.equal(model.predict(xd), full)
    eng.set_micro_batch(0)
    assert torch.equal(model.predict(xd[3:4].contiguous()), full[3:4])


@pytest.mark.parametrize('shape,batch,nc', [((520, 776), 16, 19), ((328, 520), 24, 3)])
def test_bf16_persistent_pipelines_many_tiles(shape, batch, nc):
    """The bf16 kernels are persistent, software-pipelined CTAs: a CTA's result for a tile must not depend on which tiles
    it processed before or which CTA got it.  Odd sizes (partial tiles in both directions at every scale) and a batch that
    gives every CTA several tiles at each stage; per-image runs, micro-batched runs and the full batch must agree bit for
    bit (logits and mask), with float32 and uint8 input."""
    h, w = shape
    sd = fo.make_state_dict(nc, False, 11)
    x = fo.make_input(batch, h, w, 12)
    model = build_model(sd, nc, False, DEV, precision='bf16')
    eng = model._engine(DEV)
    xd = torch.from_numpy(x).to(DEV)
    full_logits = model(xd)[0]
    full_mask = model.predict(xd)
    assert torch.isfinite(full_logits).all()
    eng.set_micro_batch(5)                      # different tile -> CTA assignment, ragged last micro-batch
    assert torch.equal(model.predict(xd), full_mask)
    assert torch.equal(model(xd)[0], full_logits)
    eng.set_micro_batch(0)
    for i in (0, batch // 2, batch - 1):        # a single image: every CTA gets at most a few tiles
        assert torch.equal(model.predict(xd[i:i + 1].contiguous()), full_mask[i:i + 1])
    # the same through the uint8 HWC input format (normalisation folded into the stem weights)
    xu = torch.randint(0, 256, (batch, h, w, 3), dtype=torch.uint8, device=DEV)
    mu = model.predict(xu)
    eng.set_micro_batch(3)
    assert torch.equal(model.predict(xu), mu)
    eng.set_micro_batch(0)
    assert torch.equal(model.predict(xu[batch - 1:].contiguous()), mu[batch - 1:])
    # and against the fp32 path: the masks agree within the bf16 tolerance on every image
    m32 = build_model(sd, nc, False, DEV, precision='fp32').predict(xd)
    assert (m32 != full_mask).float().mean().item() < MASK_TOL


def test_bf16_front_kernel_isolated_and_fused():
    """Stem + dsconv1 run as ONE kernel whenever both stages are requested (the transposed front kernel for inputs with 16-byte
    aligned rows, l2d_front_tc.cu otherwise); on its own the stem runs stem_tc.cu.  All of them against the reference fixtures."""
    for case in ('fwd_nc19_aux_n2_65x97', 'fwd_nc3_aux_n3_64x40'):
        g, sd, x, nc, aux = load_case(case)
        model = build_model(sd, nc, aux, DEV, precision='bf16')
        xd = torch.from_numpy(x).to(DEV)
        eng = model._engine(DEV)
        n, _, h, w = x.shape
        names = eng.stage_names()
        got = {}
        try:
            for gen in (1, 0):
                eng.set_option('front_transposed', gen)
                eng.forward_range(xd, names.index('stem'), names.index('l2d.dsconv1'))
                got[gen] = eng.tap_view('l2d.dsconv1', n, h, w).permute(0, 3, 1, 2).float().cpu().numpy()
                assert rel_err(got[gen], g['tap/l2d.dsconv1']) < STAGE_TOL, (case, gen)
        finally:
            eng.set_option('front_transposed', 1)
        assert rel_err(got[1], got[0]) < STAGE_TOL


def test_bf16_front_kernel_input_formats():
    """The transposed front kernel needs 16-byte aligned input rows (TMA); other widths fall back to the previous kernel.
    Aligned and unaligned widths, float32 NCHW and uint8 HWC input: both kernel generations agree on the mask."""
    nc = 3
    sd = fo.make_state_dict(nc, False, 21)
    for h, w in ((200, 336), (203, 333)):        # 336 * 3 bytes and 336 * 4 bytes are multiples of 16; 333 is not
        model = build_model(sd, nc, False, DEV, precision='bf16')
        eng = model._engine(DEV)
        xf = torch.from_numpy(fo.make_input(3, h, w, 22)).to(DEV)
        xu = torch.randint(0, 256, (3, h, w, 3), dtype=torch.uint8, device=DEV)
        for x in (xf, xu):
            new = model.predict(x)
            eng.set_option('front_transposed', 0)
            old = model.predict(x)
            eng.set_option('front_transposed', 1)
            assert (new != old).float().mean().item() < MASK_TOL


def test_bf16_many_tiles_against_the_fp32_path():
    """Every persistent CTA of the bf16 kernels walks many tiles at 328x520 x 12 images (partial tiles in both directions at
    every scale, border tiles on all four sides).  The exactness path (fp32, 1e-4 of the reference) is the truth here: logits
    within the bf16 tolerance (rms and max), masks within the mask tolerance, everything finite."""
    nc, h, w, n = 19, 328, 520, 12
    sd = fo.make_state_dict(nc, False, 13)
    x = fo.make_input(n, h, w, 14)
    xd = torch.from_numpy(x).to(DEV)
    fast = build_model(sd, nc, False, DEV, precision='bf16')
    exact = build_model(sd, nc, False, DEV, precision='fp32')
    lf, le = fast(xd)[0], exact(xd)[0]
    assert torch.isfinite(lf).all()
    scale = le.abs().max().item()
    d = (lf - le).abs()
    assert d.max().item() / scale < 6e-2 and d.pow(2).mean().sqrt().item() / scale < 1.2e-2
    assert (fast.predict(xd) != exact.predict(xd)).float().mean().item() < MASK_TOL
