"""The numpy oracle against the golden vectors produced by the unmodified reference
(oracle/gen_golden.py).  CPU only.  Tolerance: 2e-5 of each tensor's absmax (the
reference's own fp32-vs-fp64 noise floor is ~3e-6)."""
import os

import numpy as np
import pytest

import fastscnn_oracle as fo
import metric_oracle as mo

from conftest import GOLDEN

FWD_CASES = ['fwd_nc19_aux_n2_65x97', 'fwd_nc3_aux_n3_64x40', 'fwd_nc2_n1_360x640', 'fwd_nc19_n1_256x512']
TOL = 2e-5


def load_case(name):
    g = np.load(os.path.join(GOLDEN, name + '.npz'))
    nc, aux, n, h, w, wseed, xseed = (int(v) for v in g['meta'])
    sd = fo.make_state_dict(nc, bool(aux), wseed)
    sd['classifier.conv.1.bias'] = g['cls_bias']
    x = fo.make_input(n, h, w, xseed)
    return g, sd, x, nc, bool(aux)


def rel_err(a, b):
    return float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize('name', FWD_CASES)
def test_forward_matches_reference(name):
    g, sd, x, nc, aux = load_case(name)
    taps = {}
    outs = fo.forward(sd, x, aux=aux, taps=taps)
    for key in g.files:
        if key.startswith('tap/'):
            assert taps[key[4:]].shape == g[key].shape, key
            assert rel_err(taps[key[4:]], g[key]) < TOL, key
    logits = outs[0]
    if 'logits' in g.files:
        assert rel_err(logits, g['logits']) < TOL
        margin = g['margin'].astype(np.float32)
        near_tie = margin < 1e-4 * float(g['logits_absmax'])
    else:
        assert rel_err(logits[:, :, ::7, ::11], g['logits_sample']) < TOL
        assert rel_err(logits[:, :, 40:72, 96:160], g['logits_window']) < TOL
        near_tie = np.unpackbits(g['margin_small'])[:logits.shape[0] * logits.shape[2] * logits.shape[3]]
        near_tie = near_tie.reshape(logits.shape[0], logits.shape[2], logits.shape[3]).astype(bool)
    if aux:
        assert rel_err(outs[1][:, :, ::3, ::5], g['aux_logits_sample']) < TOL
    mask = fo.argmax_classes(logits)
    differ = (mask != g['mask']) & ~near_tie
    assert not differ.any(), f'{int(differ.sum())} mask pixels differ outside near-ties'
    # fused upsample+argmax restatement gives the same mask as upsample-then-argmax
    low = taps['cls.logits_lowres']
    assert np.array_equal(fo.upsample_argmax(low, x.shape[2], x.shape[3]), mask)


def test_fp64_oracle_agrees():
    g, sd, x, nc, aux = load_case('fwd_nc3_aux_n3_64x40')
    o32 = fo.forward(sd, x, aux=aux)[0]
    o64 = fo.forward(sd, x, aux=aux, dtype=np.float64)[0]
    assert rel_err(o32, o64) < TOL


def test_metric_matches_reference():
    g = np.load(os.path.join(GOLDEN, 'metric_cases.npz'))
    for i in range(int(g['ncases'])):
        nc = int(g[f'c{i}/nclass'])
        pred, label = g[f'c{i}/pred'], g[f'c{i}/label']
        m = mo.SegmentationMetricOracle(nc)
        m.update(pred, label)
        assert np.array_equal(m.total_inter, g[f'c{i}/inter']), i
        assert np.array_equal(m.total_union, g[f'c{i}/union']), i
        assert m.total_correct == int(g[f'c{i}/correct']) and m.total_label == int(g[f'c{i}/labeled'])
        pix, miou = m.get()
        assert pix == float(g[f'c{i}/pixacc']) and miou == float(g[f'c{i}/miou'])  # bit-exact float64
        # confusion-matrix formulation (what the GPU accumulates) gives the same totals
        conf = mo.confusion_counts(pred, label, nc)
        inter, union, correct, labeled = mo.totals_from_confusion(conf, nc)
        assert np.array_equal(inter, g[f'c{i}/inter']) and np.array_equal(union, g[f'c{i}/union']), i
        assert correct == int(g[f'c{i}/correct']) and labeled == int(g[f'c{i}/labeled'])


def test_metric_list_input():
    g = np.load(os.path.join(GOLDEN, 'metric_cases.npz'))
    m = mo.SegmentationMetricOracle(19)
    preds, labels = g['list/pred'], g['list/label']
    m.update([preds[0], preds[1]], [labels[0], labels[1]])
    m.update(preds[2], labels[2])
    assert np.array_equal(m.total_inter, g['list/inter']) and np.array_equal(m.total_union, g['list/union'])
    pix, miou = m.get()
    assert pix == float(g['list/pixacc']) and miou == float(g['list/miou'])


def test_adaptive_pool_bins_overlap():
    x = np.arange(32 * 5, dtype=np.float32).reshape(1, 1, 32, 5)
    out = fo.adaptive_avg_pool(x, 3)
    # rows [0,11) [10,22) [21,32)
    assert np.isclose(out[0, 0, 1, 0], x[0, 0, 10:22, 0:2].mean())


@pytest.mark.parametrize('name', ['fwd_nc19_aux_n2_65x97', 'fwd_nc2_n1_360x640'])
def test_torch_port_matches_reference(name):
    """The ATen-functional port used as the CPU baseline reproduces the reference's outputs."""
    import torch
    import fastscnn_torch_port as tp
    g, sd, x, nc, aux = load_case(name)
    outs = tp.forward(tp.to_torch_state_dict(sd), torch.from_numpy(x), aux=aux)
    logits = outs[0].numpy()
    if 'logits' in g.files:
        assert rel_err(logits, g['logits']) < 2e-6
        assert rel_err(outs[1].numpy()[:, :, ::3, ::5], g['aux_logits_sample']) < 2e-6
    else:
        assert rel_err(logits[:, :, ::7, ::11], g['logits_sample']) < 2e-6
    assert (np.argmax(logits, 1) != g['mask']).mean() < 1e-4


@pytest.mark.parametrize('name', ['train_ohem_kth', 'train_ohem_thresh', 'train_ohem_keepall', 'train_ohem_nc2'])
def test_ohem_oracle_matches_reference_fixtures(name):
    """The numpy restatement of SoftmaxCrossEntropyOHEMLoss (oracle/ohem_oracle.py) against loss / gradient vectors produced by
    the unmodified reference under torch.autograd (oracle/gen_golden_train.py)."""
    import os
    import ohem_oracle as oo
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, name + '.npz'))
    w = g['weight'] if g['weight'].size else None
    loss, grad = oo.ohem_loss_and_grad(g['logits'], g['target'], w, -1, float(g['thresh']), int(g['min_kept']))
    assert abs(loss - g['loss']) <= 1e-6 * abs(g['loss'])
    assert np.abs(grad - g['dlogits']).max() <= 1e-6 * np.abs(g['dlogits']).max()
    kept, thr = oo.ohem_select(g['logits'], g['target'], -1, float(g['thresh']), int(g['min_kept']))
    assert np.array_equal(kept, np.abs(g['dlogits']).sum(1) > 0)


def test_loss_oracle_matches_reference_fixtures():
    """The numpy restatement of the cross-entropy / dice / focal + dice criteria (oracle/loss_oracle.py), alone and composed with
    the heads' final bilinear resize, against loss / gradient vectors produced by the unmodified reference classes under
    torch.autograd (oracle/gen_golden_loss.py)."""
    import os
    import loss_oracle as lo
    from conftest import GOLDEN
    from helpers import LOSS_CASES
    g = np.load(os.path.join(GOLDEN, 'train_loss_cases.npz'))
    for name, (kind, kw, aux_weight) in LOSS_CASES.items():
        target = g[name + '/target']
        total, grads = 0.0, []
        for i in range(2 if aux_weight is not None else 1):
            low = g[f'{name}/logits{i}']
            scale = 1.0 if i == 0 else aux_weight
            if low.shape[2:] != target.shape[1:]:
                loss, grad = lo.criterion_upsampled(kind, low, target, **kw)
            else:
                loss, grad = lo.CRITERIA[kind](low, target, **kw)
            total += scale * loss
            grads.append(scale * grad)
        assert abs(total - float(g[name + '/loss'])) <= 2e-6 * abs(float(g[name + '/loss'])), name
        for i, grad in enumerate(grads):
            ref = g[f'{name}/grad{i}']
            assert np.abs(grad - ref).max() <= 2e-5 * np.abs(ref).max(), (name, i)
