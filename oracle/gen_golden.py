"""Generate tests/golden/*.npz by running the UNMODIFIED reference in the build container.

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden.py [--reference /root/reference]

The reference (pure Python: models/fast_scnn.py, utils/metric.py) is imported from
``/root/reference`` -- never copied.  Weights and inputs come from the numpy-seeded
recipes in ``oracle/fastscnn_oracle.py`` so that tests can rebuild them anywhere; the
fixtures hold only what the reference *computed* (stage taps, logits, masks, metric
totals) plus the calibrated classifier bias.  /root/reference does not exist on the GPU
box, so nothing at test time calls this script; the fixtures are committed.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import fastscnn_oracle as fo  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), 'tests', 'golden')

# (case name, num_classes, aux, N, H, W, weight seed, input seed, what to keep)
FORWARD_CASES = (
    ('fwd_nc19_aux_n2_65x97', 19, True, 2, 65, 97, 7, 11, 'all'),        # odd sizes, every stage tap
    ('fwd_nc2_n1_360x640', 2, False, 1, 360, 640, 7, 12, 'sample'),      # BASELINE config 0
    ('fwd_nc19_n1_256x512', 19, False, 1, 256, 512, 8, 13, 'sample'),    # reference's own smoke size
    ('fwd_nc3_aux_n3_64x40', 3, True, 3, 64, 40, 9, 14, 'all'),          # tiny: 2x1 after /32
)

TAP_MODULES = {
    'l2d.conv': 'learning_to_downsample.conv',
    'l2d.dsconv1': 'learning_to_downsample.dsconv1',
    'l2d.dsconv2': 'learning_to_downsample.dsconv2',
    'gfe.ppm': 'global_feature_extractor.ppm',
    'ffm': 'feature_fusion',
    'cls.dsconv1': 'classifier.dsconv1',
    'cls.dsconv2': 'classifier.dsconv2',
    'cls.logits_lowres': 'classifier',
}
for _name, _, _ in fo.BOTTLENECK_PLAN:
    TAP_MODULES['gfe.' + _name] = 'global_feature_extractor.' + _name


def run_forward_case(torch, FastSCNN, case):
    name, nc, aux, n, h, w, wseed, xseed, keep = case
    sd_np = fo.make_state_dict(nc, aux, wseed)
    x_np = fo.make_input(n, h, w, xseed)
    model = FastSCNN(nc, aux=aux).eval()
    missing = model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd_np.items()}, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    x = torch.from_numpy(x_np)
    with torch.no_grad():
        # recipe D2 calibration on the reference itself: centre the per-class mean logit
        mean = model(x)[0].mean((0, 2, 3))
        model.classifier.conv[1].bias -= mean
        taps = {}
        mods = dict(model.named_modules())
        hooks = [mods[path].register_forward_hook(lambda m, i, o, key=key: taps.__setitem__(key, o.detach().numpy().copy()))
                 for key, path in TAP_MODULES.items()]
        if aux:
            hooks.append(model.auxlayer.register_forward_hook(
                lambda m, i, o: taps.__setitem__('aux.logits_lowres', o.detach().numpy().copy())))
        outs = model(x)
        for hk in hooks:
            hk.remove()
        logits = outs[0].numpy()
        mask = torch.argmax(outs[0], 1).numpy()
    out = {
        'meta': np.array([nc, int(aux), n, h, w, wseed, xseed], dtype=np.int64),
        'cls_bias': model.classifier.conv[1].bias.detach().numpy().copy(),
        'mask': mask.astype(np.uint8),
        'margin': fo.top2_margin(logits).astype(np.float32).astype(np.float16) if keep == 'all' else np.zeros(0, np.float16),
        'logits_absmax': np.array(np.abs(logits).max(), dtype=np.float32),
    }
    if keep == 'all':
        for k, v in taps.items():
            out['tap/' + k] = v
        out['logits'] = logits
        if aux:
            out['aux_logits_sample'] = outs[1].numpy()[:, :, ::3, ::5].copy()
    else:
        for k in ('l2d.dsconv2', 'gfe.bottleneck3.2', 'gfe.ppm', 'cls.logits_lowres'):
            out['tap/' + k] = taps[k]
        out['logits_sample'] = logits[:, :, ::7, ::11].copy()   # strided probe of the full-res logits
        out['logits_window'] = logits[:, :, 40:72, 96:160].copy()
        out['margin_small'] = (fo.top2_margin(logits) < 1e-4 * np.abs(logits).max())  # near-tie pixels
        out['margin_small'] = np.packbits(out['margin_small'])
    counts = np.bincount(mask.reshape(-1), minlength=nc)
    print(f'{name}: logits absmax {np.abs(logits).max():.3f}, class share max {counts.max() / counts.sum():.3f}, '
          f'classes present {(counts > 0).sum()}/{nc}')
    np.savez_compressed(os.path.join(GOLDEN, name + '.npz'), **out)


def run_metric_cases(SegmentationMetric):
    rng = np.random.RandomState(5)
    out = {}
    cases = []
    for i, (nc, shape, lo, hi) in enumerate((
            (19, (2, 33, 47), -1, 19),      # normal labels in [-1, nc)
            (19, (1, 40, 40), -3, 23),      # adversarial: labels < -1 and >= nc
            (2, (3, 17, 29), -1, 2),
            (3, (1, 1, 1), 0, 3),           # a single pixel
            (150, (1, 64, 64), -2, 160),    # many classes
            (5, (1, 0, 7), -1, 5),          # empty image
    )):
        pred = rng.randint(0, nc, size=shape).astype(np.int64)
        label = rng.randint(lo, hi, size=shape).astype(np.int64)
        if i == 1:
            pred = rng.randint(-2, nc + 3, size=shape).astype(np.int64)  # preds outside the class range too
        m = SegmentationMetric(nc)
        m.update(pred, label)
        pix, miou = m.get()
        out[f'c{i}/pred'], out[f'c{i}/label'] = pred, label
        out[f'c{i}/nclass'] = np.array(nc)
        out[f'c{i}/inter'] = np.asarray(m.total_inter, dtype=np.int64).reshape(-1) if shape[1] else np.zeros(nc, np.int64)
        out[f'c{i}/union'] = np.asarray(m.total_union, dtype=np.int64).reshape(-1) if shape[1] else np.zeros(nc, np.int64)
        out[f'c{i}/correct'], out[f'c{i}/labeled'] = np.array(int(m.total_correct)), np.array(int(m.total_label))
        out[f'c{i}/pixacc'], out[f'c{i}/miou'] = np.array(pix, np.float64), np.array(miou, np.float64)
        cases.append(i)
    # list input + accumulation over two updates (metric.py:34-40, 56-63)
    nc = 19
    preds = [rng.randint(0, nc, size=(1, 20, 30)).astype(np.int64) for _ in range(3)]
    labels = [rng.randint(-1, nc, size=(1, 20, 30)).astype(np.int64) for _ in range(3)]
    m = SegmentationMetric(nc)
    m.update(preds[:2], labels[:2])
    m.update(preds[2], labels[2])
    pix, miou = m.get()
    out['list/pred'], out['list/label'] = np.stack(preds), np.stack(labels)
    out['list/inter'], out['list/union'] = np.asarray(m.total_inter, np.int64), np.asarray(m.total_union, np.int64)
    out['list/correct'], out['list/labeled'] = np.array(int(m.total_correct)), np.array(int(m.total_label))
    out['list/pixacc'], out['list/miou'] = np.array(pix, np.float64), np.array(miou, np.float64)
    out['ncases'] = np.array(len(cases))
    np.savez_compressed(os.path.join(GOLDEN, 'metric_cases.npz'), **out)
    print('metric_cases: %d cases + list case' % len(cases))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reference', default='/root/reference')
    args = ap.parse_args()
    sys.dont_write_bytecode = True
    sys.path.insert(0, args.reference)
    import torch
    from models.fast_scnn import FastSCNN          # the reference, unmodified
    from utils.metric import SegmentationMetric    # the reference, unmodified
    torch.manual_seed(0)
    os.makedirs(GOLDEN, exist_ok=True)
    for case in FORWARD_CASES:
        run_forward_case(torch, FastSCNN, case)
    run_metric_cases(SegmentationMetric)


if __name__ == '__main__':
    main()
