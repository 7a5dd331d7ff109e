"""Generate tests/golden/e2e_*.npz by running the UNMODIFIED reference wrapper (export_onnx_fixed.EndToEndFastSCNN) in the build
container:   PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden_e2e.py [--reference /root/reference]
Weights / frames come from the seeded numpy recipes (fastscnn_oracle.make_state_dict, e2e_oracle.make_frames); the fixtures hold
what the reference computed plus the calibrated classifier bias."""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import e2e_oracle as eo  # noqa: E402
import fastscnn_oracle as fo  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), 'tests', 'golden')
IMAGENET = ([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])
# (name, nc, n, frame h, frame w, base size, weight seed, frame seed, normalise, softmax)
CASES = (
    ('e2e_nc2_n2_90x160_b256', 2, 2, 90, 160, 256, 31, 32, False, True),      # the deployed shape scaled by 1/4 (640x360 -> 1024)
    ('e2e_nc19_n1_75x131_b224', 19, 1, 75, 131, 224, 33, 34, True, False),    # odd sizes, ImageNet normalisation, raw logits
)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reference', default='/root/reference')
    args = ap.parse_args()
    sys.path.insert(0, args.reference)
    import torch
    from export_onnx_fixed import EndToEndFastSCNN
    from models.fast_scnn import FastSCNN
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    for name, nc, n, h, w, base, wseed, xseed, norm, sm in CASES:
        sd = fo.make_state_dict(nc, False, wseed)
        frames = eo.make_frames(n, h, w, xseed)
        mean, std = IMAGENET if norm else (None, None)
        backbone = FastSCNN(nc).eval()
        backbone.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
        model = EndToEndFastSCNN(backbone, input_size=(w, h), base_size=base, mean=mean, std=std, apply_softmax=sm).eval()
        x = torch.from_numpy(frames)
        with torch.no_grad():
            pre = model.preprocessor(x)
            backbone.classifier.conv[1].bias -= backbone(pre)[0].mean((0, 2, 3))     # recipe D2: centre the per-class mean logit
            out = model(x).numpy()
            out_f32_in = model(x.float()).numpy()                                     # float32 frames take the same path
        assert np.array_equal(out, out_f32_in)
        np.savez_compressed(os.path.join(GOLDEN, name + '.npz'),
                            meta=np.array([nc, n, h, w, base, wseed, xseed, int(norm), int(sm)], dtype=np.int64),
                            cls_bias=backbone.classifier.conv[1].bias.detach().numpy().copy(), out=out.astype(np.float32),
                            pre_sample=pre.numpy()[:, :, ::5, ::7].copy())
        print(name, out.shape, 'range', float(out.min()), float(out.max()))


if __name__ == '__main__':
    main()
