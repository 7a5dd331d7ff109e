"""Shared helpers for the parity tests."""
import os

import numpy as np

import fastscnn_oracle as fo

from conftest import GOLDEN


def load_case(name):
    g = np.load(os.path.join(GOLDEN, name + '.npz'))
    nc, aux, n, h, w, wseed, xseed = (int(v) for v in g['meta'])
    sd = fo.make_state_dict(nc, bool(aux), wseed)
    sd['classifier.conv.1.bias'] = g['cls_bias']
    x = fo.make_input(n, h, w, xseed)
    return g, sd, x, nc, bool(aux)


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def torch_state_dict(sd):
    import torch
    return {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}


def build_model(sd, nc, aux, device, precision='fp32'):
    from models.fast_scnn import FastSCNN
    model = FastSCNN(nc, aux=aux, precision=precision).eval()
    model.load_state_dict(torch_state_dict(sd))
    return model.to(device)
