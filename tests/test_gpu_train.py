"""Training-step slice: the covered modules in train mode and the OHEM loss, through the C ABI
training operators (fscnn_train_*), against golden vectors produced by torch.autograd on the UNMODIFIED reference modules
(oracle/gen_golden_train.py).  Tolerance: fp32 path, 1e-4 of each tensor's absmax (outputs, input gradients, parameter
gradients, BatchNorm running statistics); OHEM loss 1e-5 relative, its gradient 1e-5 of absmax."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from helpers import rel_err

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda', 0)
TOL = 1e-4


def _load(name):
    return np.load(os.path.join(GOLDEN, name + '.npz'))


def _run_module(module, g):
    sd = {k[len('before/'):]: torch.from_numpy(g[k]) for k in g.files if k.startswith('before/')}
    module.load_state_dict(sd)
    module.to(DEV).train()
    x = torch.from_numpy(g['x']).to(DEV).requires_grad_(True)
    y = module(x)
    y.backward(torch.from_numpy(g['gy']).to(DEV))
    assert rel_err(y.detach().cpu().numpy(), g['y']) < TOL
    assert rel_err(x.grad.cpu().numpy(), g['dx']) < TOL
    for k, p in module.named_parameters():
        assert p.grad is not None, k
        assert rel_err(p.grad.cpu().numpy(), g['grad/' + k]) < TOL, k
    after = module.state_dict()
    for k in g.files:
        if k.startswith('after/'):
            got = after[k[len('after/'):]].cpu().numpy()
            if 'num_batches' in k:
                assert int(got) == int(g[k]), k
            else:
                assert rel_err(got, g[k]) < TOL, k


@pytest.mark.parametrize('name,cin,cout,stride', [('train_dsconv_32_48_s2', 32, 48, 2), ('train_dsconv_16_16_s1', 16, 16, 1)])
def test_dsconv_train_step_matches_reference_autograd(name, cin, cout, stride):
    from models.fast_scnn import DSConv
    _run_module(DSConv(cin, cout, stride), _load(name))


@pytest.mark.parametrize('name,cin,cout,stride', [('train_bottleneck_16_16_s1', 16, 16, 1), ('train_bottleneck_16_24_s2', 16, 24, 2)])
def test_linear_bottleneck_train_step_matches_reference_autograd(name, cin, cout, stride):
    from models.fast_scnn import LinearBottleneck
    _run_module(LinearBottleneck(cin, cout, 6, stride), _load(name))


@pytest.mark.parametrize('name', ['train_ohem_kth', 'train_ohem_thresh', 'train_ohem_keepall', 'train_ohem_nc2'])
def test_ohem_loss_matches_reference(name):
    from fscnn_b200 import train_ops
    g = _load(name)
    logits = torch.from_numpy(g['logits']).to(DEV).requires_grad_(True)
    target = torch.from_numpy(g['target']).to(DEV)
    weight = torch.from_numpy(g['weight']).to(DEV) if g['weight'].size else None
    loss = train_ops.ohem_cross_entropy(logits, target, weight, -1, float(g['thresh']), int(g['min_kept']))
    (2.0 * loss).backward()
    assert abs(float(loss.detach()) - float(g['loss'])) <= 1e-5 * abs(float(g['loss']))
    assert rel_err(logits.grad.cpu().numpy(), 2.0 * g['dlogits']) < 1e-5
    # the same pixels are kept (their gradient rows are the non-zero ones)
    assert np.array_equal(np.abs(logits.grad.cpu().numpy()).sum(1) > 0, np.abs(g['dlogits']).sum(1) > 0)


def test_train_ops_match_torch_at_baseline_like_shapes():
    """Shapes of BASELINE config 5 scaled down (crop 96, batch 4): stride-2 bottleneck 64 -> 96 with t = 6 and a DSConv 48 -> 64.
    Truth = torch.autograd in float64 on ATen's kernels (the checker, not the product).  Outputs and input gradients within
    1e-4 of absmax; parameter gradients are sums of N*H*W largely cancelling products (BN gamma / beta especially), so their
    yardstick is what ATen's own float32 kernels lose against the same float64 truth: ours must not be worse than 2x that."""
    import torch.nn.functional as F
    from models.fast_scnn import DSConv, LinearBottleneck
    torch.manual_seed(3)
    torch.backends.cudnn.allow_tf32 = False

    def ref_seq(seq, t, dt):
        for m in seq:
            if isinstance(m, torch.nn.Conv2d):
                t = F.conv2d(t, m.weight.to(dt), None, m.stride, m.padding, 1, m.groups)
            elif isinstance(m, torch.nn.BatchNorm2d):
                t = F.batch_norm(t, None, None, m.weight.to(dt), m.bias.to(dt), True, 0.1, m.eps)
            elif isinstance(m, torch.nn.ReLU):
                t = F.relu(t)
            else:
                t = ref_seq(m.conv, t, dt)
        return t

    for mod, cin, hw in ((LinearBottleneck(64, 96, 6, 2), 64, (24, 24)), (DSConv(48, 64, 2), 48, (47, 49))):
        mod.to(DEV).train()
        seq = mod.block if hasattr(mod, 'block') else mod.conv
        x = torch.randn(4, cin, *hw, device=DEV, requires_grad=True)
        y = mod(x)
        gy = torch.randn_like(y)
        y.backward(gy)
        ours = {k: p.grad.double().cpu().numpy() for k, p in mod.named_parameters()}
        ours_dx = x.grad.double().cpu().numpy()
        refs = {}
        for dt in (torch.float64, torch.float32):
            mod.zero_grad()
            xr = x.detach().to(dt).requires_grad_(True)
            yr = ref_seq(seq, xr, dt)
            yr.backward(gy.to(dt))
            refs[dt] = (yr.detach().double().cpu().numpy(), xr.grad.double().cpu().numpy(),
                        {k: p.grad.double().cpu().numpy() for k, p in mod.named_parameters()})
        y64, dx64, g64 = refs[torch.float64]
        _, _, g32 = refs[torch.float32]
        assert rel_err(y.detach().cpu().numpy(), y64) < TOL
        assert rel_err(ours_dx, dx64) < TOL
        for k in ours:
            mine, aten = np.abs(ours[k] - g64[k]).max(), np.abs(g32[k] - g64[k]).max()
            assert mine <= max(2.0 * aten, TOL * np.abs(g64[k]).max()), (k, mine, aten)


def test_uncovered_training_paths_are_loud():
    from fscnn_b200 import train_ops
    from models.fast_scnn import DSConv, FastSCNN
    model = FastSCNN(3).to(DEV).train()
    with pytest.raises(NotImplementedError):
        model(torch.zeros(1, 3, 64, 64, device=DEV))                          # whole-network training forward: not yet
    with pytest.raises(NotImplementedError):
        model.global_feature_extractor.ppm(torch.zeros(1, 128, 8, 8, device=DEV))   # PPM has no training operator yet
    with pytest.raises(RuntimeError):
        DSConv(16, 16, 1).train()(torch.zeros(1, 16, 8, 8))                   # CPU tensor: no fallback
    with pytest.raises(ValueError):
        train_ops.ohem_cross_entropy(torch.zeros(1, 3, 4, 4, device=DEV), torch.zeros(1, 4, 5, dtype=torch.int64, device=DEV))
