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


def _no_dropout(model):
    """Loss-goes-down checks on a 2-image batch must not depend on the dropout masks (dropout has its own tests)."""
    for mod in model.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    return model


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


def test_training_paths_are_loud_where_unsupported():
    from fscnn_b200 import train_ops
    from models.fast_scnn import DSConv, FastSCNN
    model = FastSCNN(3).to(DEV).train()
    with pytest.raises(RuntimeError):
        model.predict(torch.zeros(1, 3, 64, 64, device=DEV))                  # the fused inference engine serves eval mode only
    with pytest.raises(ValueError):
        model(torch.zeros(1, 3, 64, 64, device=DEV, dtype=torch.float16))     # training forward is fp32
    with pytest.raises(RuntimeError):
        DSConv(16, 16, 1).train()(torch.zeros(1, 16, 8, 8))                   # CPU tensor: no fallback
    with pytest.raises(ValueError):
        train_ops.ohem_cross_entropy(torch.zeros(1, 3, 4, 4, device=DEV), torch.zeros(1, 4, 5, dtype=torch.int64, device=DEV))


def test_whole_network_train_step_matches_reference_fixture():
    """One training step of the whole network (aux head on, Dropout p = 0): loss, outputs, BatchNorm buffers and every parameter
    gradient against the vectors produced by the unmodified reference FastSCNN + MixSoftmaxCrossEntropyOHEMLoss under
    torch.autograd (oracle/gen_golden_train.py network_case).  Gradients: relative L2 distance on the stored sample < 1e-3
    (they pass through ~45 layers of fp32 reductions in a different order), norms within 1e-3."""
    import fastscnn_oracle as fo
    from fscnn_b200 import Trainer
    from models.fast_scnn import FastSCNN
    g = _load('train_net_nc19_aux')
    nc, aux, seed, n, h, w = (int(v) for v in g['meta'])
    model = FastSCNN(nc, aux=bool(aux))
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in fo.make_state_dict(nc, bool(aux), seed).items()})
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    model.to(DEV).train()
    x = torch.from_numpy(fo.make_input(n, h, w, seed + 1)).to(DEV)
    target = torch.from_numpy(fo.make_labels(n, h, w, nc, seed + 2)).to(DEV)
    trainer = Trainer(model, aux_weight=0.4)
    outputs = model(x)
    loss = trainer.loss(outputs, target)
    loss.backward()
    assert abs(float(loss.detach()) - float(g['loss'])) <= 1e-4 * abs(float(g['loss']))
    for i, o in enumerate(outputs):
        got = o.detach().cpu().numpy()[:, :, ::5, ::7]
        assert np.abs(got - g[f'out{i}_sample']).max() <= 2e-4 * float(g[f'out{i}_absmax']), i
    after = model.state_dict()
    for k in g.files:
        if k.startswith('after/'):
            got = after[k[len('after/'):]].cpu().numpy()
            if 'num_batches' in k:
                assert int(got) == int(g[k]), k
            else:
                assert rel_err(got, g[k]) < 2e-4, k
    # Parameter gradients are long fp32 reduction chains (the stem's passes back through ~45 layers): the yardstick is a
    # float64 run of the same operator sequence (ATen, the checker) with the SAME pixels kept by the OHEM selection.  Ours must
    # be as close to that truth as the reference's own float32 run is (factor 2).
    for m in model.modules():      # a second forward only to read the OHEM selection: it must not move the BatchNorm buffers
        if isinstance(m, torch.nn.BatchNorm2d):
            m.momentum = 0.0
    outputs2 = model(x)
    kept = [(o_grad.abs().sum(1) > 0) for o_grad in torch.autograd.grad(trainer.loss(outputs2, target), outputs2)]
    truth = _float64_truth(model, x, target, kept, trainer)
    worst, all_ours, all_ref = 0.0, [], []
    for k, p in model.named_parameters():
        gr = p.grad.detach().cpu().numpy().ravel().astype(np.float64)
        ref_s = g['gsample/' + k].astype(np.float64)
        t = truth[k].ravel()
        # (a conv bias in front of a BatchNorm has an exactly zero gradient: float32 runs return 1e-7 noise there)
        scale = max(np.linalg.norm(t[::13]), 1e-4 * np.sqrt(t[::13].size))
        d_ours, d_ref = np.linalg.norm(gr[::13] - t[::13]) / scale, np.linalg.norm(ref_s - t[::13]) / scale
        worst = max(worst, d_ours)
        assert d_ours <= max(1e-3, 2.0 * d_ref), (k, d_ours, d_ref)
        all_ours.append(gr[::13]); all_ref.append(ref_s)
    # and against the reference's own float32 vector, over all parameters at once (single BatchNorm gammas are cancelling sums
    # that the reference's float32 run itself only gets to a few per cent)
    a, b = np.concatenate(all_ours), np.concatenate(all_ref)
    glob = np.linalg.norm(a - b) / np.linalg.norm(b)
    print('worst relative gradient distance to the float64 truth', worst, '; all gradients vs the reference fixture', glob)
    assert glob < 2e-2


def _float64_truth(model, x, target, kept, trainer, dtype=torch.float64):
    """Parameter gradients of the same step in float64 on ATen (the checker): training-mode BatchNorm, no dropout, weighted cross
    entropy over exactly the pixels `kept` per head, aux weight as in the trainer.  dtype=torch.float32 gives the same operator
    sequence through cuDNN in float32 (with whatever torch.backends.cudnn.allow_tf32 says): the yardstick of the TF32 mode."""
    import torch.nn.functional as F
    import fastscnn_torch_port as port
    sd = {k: (v.detach().to(dtype) if v.dtype.is_floating_point else v.detach()).requires_grad_(v.dtype.is_floating_point and 'running' not in k)
          for k, v in model.state_dict().items()}
    real_bn = port._bn
    port._bn = lambda s, p, t: F.batch_norm(t, None, None, s[p + '.weight'], s[p + '.bias'], True, 0.1, 1e-5)
    try:
        with torch.enable_grad():
            fwd = getattr(port.forward, '__wrapped__', port.forward)
            outs = fwd(sd, x.to(dtype), aux=len(kept) > 1)
    finally:
        port._bn = real_bn
    total = 0.0
    for i, (o, keep) in enumerate(zip(outs, kept)):
        tgt = torch.where(keep, target, torch.full_like(target, -1))
        w = trainer.class_weight.to(dtype) if trainer.class_weight is not None else None
        li = F.cross_entropy(o, tgt, weight=w, ignore_index=-1)
        total = total + (li if i == 0 else trainer.aux_weight * li)
    names = [k for k, v in sd.items() if v.requires_grad]
    grads = torch.autograd.grad(total, [sd[k] for k in names])
    return {k: g_.cpu().numpy() for k, g_ in zip(names, grads)}


def test_trainer_steps_match_torch_sgd():
    """Three Trainer steps (flat-buffer SGD with momentum and weight decay, poly learning rate) against torch.optim.SGD driving
    the same operators' gradients: parameters must agree to fp32 rounding, and the loss must go down on a fixed batch."""
    import fastscnn_oracle as fo
    from fscnn_b200 import Trainer, poly_lr
    from models.fast_scnn import FastSCNN
    nc = 19
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in fo.make_state_dict(nc, True, 5).items()}
    x = torch.from_numpy(fo.make_input(2, 96, 96, 6)).to(DEV)
    target = torch.from_numpy(fo.make_labels(2, 96, 96, nc, 7)).to(DEV)
    models = []
    for _ in range(2):
        m = FastSCNN(nc, aux=True)
        m.load_state_dict(sd)
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
        models.append(m.to(DEV).train())
    ours, ref = models
    # fused_loss=False: both sides then run the same deterministic kernels in the same order, so the comparison isolates the optimizer
    # (the fused loss sums its backward with float atomics, and OHEM's pixel selection amplifies last-bit differences over steps)
    trainer = Trainer(ours, base_lr=0.01, nepochs=1, iters_per_epoch=10, fused_loss=False)
    helper = Trainer.__new__(Trainer)          # the reference side uses the same loss definition, torch's optimizer
    helper.__dict__.update(class_weight=trainer.class_weight, ignore_label=-1, ohem_thresh=0.7, ohem_min_kept=256, aux_weight=0.4)
    opt = torch.optim.SGD(ref.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    losses = []
    for it in range(3):
        losses.append(float(trainer.step(x, target)))
        for gp in opt.param_groups:
            gp['lr'] = poly_lr(0.01, it, 1, 10)
        opt.zero_grad()
        Trainer.loss(helper, ref(x), target).backward()
        opt.step()
    assert losses[2] < losses[0]
    for (k, a), (_, b) in zip(ours.named_parameters(), ref.named_parameters()):
        # (1e-4, not fp32 epsilon: the two optimizers round differently by an ulp, and two more steps through batch-statistics
        # BatchNorm and the OHEM pixel selection amplify that -- measured 1e-5 ... 4e-5 on the small BatchNorm biases; a wrong
        # momentum / weight-decay / learning-rate rule shows up at 1e-2)
        assert rel_err(a.detach().cpu().numpy(), b.detach().cpu().numpy()) < 1e-4, k
    ours.eval()                                   # and the trained weights go straight back into the fused inference engine
    with torch.no_grad():
        mask = ours.predict(x)
    assert mask.shape == (2, 96, 96)


@pytest.mark.parametrize('n,c,hl,wl,h,w,scale,min_kept', [(2, 19, 12, 16, 96, 128, 3.0, 256), (1, 19, 9, 13, 65, 97, 9.0, 64),
                                                          (2, 2, 8, 8, 60, 57, 2.0, 100000), (1, 5, 10, 11, 75, 80, 4.0, 32)])
def test_fused_upsample_ohem_equals_the_two_step_form(n, c, hl, wl, h, w, scale, min_kept):
    """ohem_cross_entropy_upsampled(low) against ohem_cross_entropy(bilinear_resize(low)): the loss (same arithmetic, same summation
    order: bit-identical), the kept pixels, and the low-resolution gradient (the fused backward sums with float atomics: 1e-5)."""
    from fscnn_b200 import train_ops
    g = torch.Generator(device=DEV).manual_seed(c * 100 + hl)
    low = (torch.randn((n, c, hl, wl), device=DEV, generator=g) * scale)
    target = torch.randint(-1, c, (n, h, w), device=DEV, generator=g)
    weight = torch.rand(c, device=DEV, generator=g) + 0.5 if c == 19 else None
    a = low.clone().requires_grad_(True)
    la = train_ops.ohem_cross_entropy_upsampled(a, target, weight, -1, 0.7, min_kept)
    (3.0 * la).backward()
    b = low.clone().requires_grad_(True)
    lb = train_ops.ohem_cross_entropy(train_ops.bilinear_resize(b, (h, w)), target, weight, -1, 0.7, min_kept)
    (3.0 * lb).backward()
    assert float(la.detach()) == float(lb.detach())
    assert rel_err(a.grad.cpu().numpy(), b.grad.cpu().numpy()) < 1e-5


def test_tf32_matmul_mode_is_close_to_fp32_and_really_different():
    """fscnn_train_set_math(1): the pointwise forward / data gradient (tcgen05.mma.kind::tf32, accumulator in TMEM; mma.sync for rows
    that are not 16-byte aligned) and weight gradient (mma.sync) on the tensor cores with TF32 operands.  Against the fp32 FMA kernels
    the error must sit at the TF32 level (10 mantissa bits per operand, averaged over the contraction), well above fp32 noise (so the
    tensor-core kernels really ran) and below 3e-3 of the tensor's absmax; the mode is process-wide and is restored."""
    from fscnn_b200 import train_ops
    g = torch.Generator(device='cpu').manual_seed(11)
    results = {}
    try:
        for mode in ('fp32', 'tf32'):
            train_ops.set_matmul_precision(mode)
            assert train_ops.get_matmul_precision() == mode
            outs = []
            # aligned (tcgen05 kernels: 128-pixel tiles, 256-pixel tiles from 2048 pixels up, several channel blocks, partial tiles)
            # and ragged (mma.sync kernels)
            for n, cin, cout, h, w in [(2, 64, 384, 24, 28), (3, 130, 50, 17, 19), (1, 32, 19, 40, 36), (2, 48, 160, 44, 60), (1, 20, 64, 52, 48)]:
                gg = torch.Generator(device='cpu').manual_seed(100 + cin)
                x = torch.randn(n, cin, h, w, generator=gg).to(DEV).requires_grad_(True)
                wt = (torch.randn(cout, cin, 1, 1, generator=gg) / cin ** 0.5).to(DEV).requires_grad_(True)
                dy = torch.randn(n, cout, h, w, generator=gg).to(DEV)
                y = train_ops.pointwise_conv(x, wt)
                y.backward(dy)
                outs += [y.detach().cpu().numpy(), x.grad.cpu().numpy(), wt.grad.cpu().numpy()]
            results[mode] = outs
    finally:
        train_ops.set_matmul_precision('fp32')
    del g
    for a, b in zip(results['fp32'], results['tf32']):
        e = rel_err(b, a)
        assert 1e-6 < e < 3e-3, e
    with pytest.raises(ValueError):
        train_ops.set_matmul_precision('fp16')


def test_trainer_cuda_graph_replays_match_eager_steps():
    """Trainer(cuda_graph=True): two eager warm-up steps, then the captured zero_grad + forward + loss + backward replayed; with
    Dropout switched off the parameters after five steps must agree with an eager Trainer's to float-atomics noise, on inputs that
    CHANGE from step to step (the graph copies them into its static buffers)."""
    import fastscnn_oracle as fo
    from fscnn_b200 import Trainer
    from models.fast_scnn import FastSCNN
    nc = 19
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in fo.make_state_dict(nc, True, 9).items()}
    batches = [(torch.from_numpy(fo.make_input(2, 96, 128, 20 + i)).to(DEV), torch.from_numpy(fo.make_labels(2, 96, 128, nc, 40 + i)).to(DEV))
               for i in range(5)]
    trainers = []
    for use_graph in (False, True):
        m = FastSCNN(nc, aux=True)
        m.load_state_dict(sd)
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
        # fused_loss=False: deterministic loss kernels on both sides (the fused loss sums its backward with float atomics, and OHEM's pixel
        # selection amplifies last-bit differences over steps)
        trainers.append(Trainer(m.to(DEV).train(), base_lr=0.01, nepochs=1, iters_per_epoch=10, cuda_graph=use_graph, graph_warmup=2,
                                fused_loss=False))
    losses = [[], []]
    for x, t in batches:
        for k, tr in enumerate(trainers):
            losses[k].append(float(tr.step(x, t)))
    assert trainers[1]._graph is not None
    np.testing.assert_allclose(losses[1], losses[0], rtol=2e-4)
    for (k, a), (_, b) in zip(trainers[0].model.named_parameters(), trainers[1].model.named_parameters()):
        assert rel_err(b.detach().cpu().numpy(), a.detach().cpu().numpy()) < 2e-4, k
    for (k, a), (_, b) in zip(trainers[0].model.named_buffers(), trainers[1].model.named_buffers()):
        assert rel_err(b.detach().float().cpu().numpy(), a.detach().float().cpu().numpy()) < 2e-4, k


def test_dropout_step_counter_changes_the_mask():
    """The device step counter behind CUDA-graph replays: same seed + same counter = same mask, a bumped counter = a new mask with the
    same keep rate, and the backward regenerates the forward's mask."""
    from fscnn_b200 import train_ops
    x = torch.ones(4, 8, 32, 32, device=DEV, requires_grad=True)
    counter = torch.zeros(1, dtype=torch.int64, device=DEV)
    train_ops.set_dropout_step_counter(counter)
    try:
        a = train_ops.dropout(x, 0.25, True, seed=123)
        b = train_ops.dropout(x, 0.25, True, seed=123)
        counter += 1
        c = train_ops.dropout(x, 0.25, True, seed=123)
        c.sum().backward()
    finally:
        train_ops.set_dropout_step_counter(None)
    assert torch.equal(a, b) and not torch.equal(a, c)
    assert abs(float((c != 0).float().mean()) - 0.75) < 0.02
    assert torch.equal(x.grad, c.detach())
    with pytest.raises(ValueError):
        train_ops.set_dropout_step_counter(torch.zeros(1))


@pytest.mark.parametrize('n,c,h,w,stride', [(3, 24, 24, 24, 1), (2, 16, 40, 48, 1), (2, 8, 33, 128, 1), (2, 8, 30, 132, 1), (2, 12, 37, 41, 1),
                                            (2, 12, 37, 41, 2), (2, 16, 64, 96, 2)])
def test_depthwise_kernels_against_torch_and_run_to_run(n, c, h, w, stride):
    """Every depthwise kernel variant (float4 groups for stride 1 with widths that are multiples of 4 up to 128 -- 6, 12 and 32 lanes
    per row here --, the one-column walk for everything else, stride 2 with and without 8-byte rows) against torch.nn.functional.conv2d,
    and twice in a row: outputs and input gradients must be bit-identical between runs, weight gradients (double atomics, rounded to
    float once) equal to the last bit or one off."""
    import torch.nn.functional as F
    from fscnn_b200 import train_ops
    g = torch.Generator(device='cpu').manual_seed(1000 + h * w + stride)
    x0 = torch.randn(n, c, h, w, generator=g).to(DEV)
    w0 = torch.randn(c, 1, 3, 3, generator=g).to(DEV)
    ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
    dy = torch.randn(n, c, ho, wo, generator=g).to(DEV)
    runs = []
    for _ in range(2):
        x, wt = x0.clone().requires_grad_(True), w0.clone().requires_grad_(True)
        y = train_ops.depthwise_conv3x3(x, wt, stride)
        y.backward(dy)
        runs.append((y.detach(), x.grad, wt.grad))
    assert torch.equal(runs[0][0], runs[1][0]) and torch.equal(runs[0][1], runs[1][1])
    assert float((runs[0][2] - runs[1][2]).abs().max()) <= 2e-7 * float(runs[0][2].abs().max())
    xr, wr = x0.double().requires_grad_(True), w0.double().requires_grad_(True)
    yr = F.conv2d(xr, wr, None, stride, 1, 1, c)
    yr.backward(dy.double())
    for got, want in zip(runs[0], (yr.detach(), xr.grad, wr.grad)):
        assert rel_err(got.cpu().numpy(), want.float().cpu().numpy()) < 2e-6


def test_tf32_training_step_against_cudnn_tf32():
    """The whole training step with TF32 contractions (tcgen05 forward / data gradient, mma.sync weight gradient): loss within 5e-4 of
    the fp32 mode, and every parameter gradient as close to a float64 run of the same step (same kept pixels) as the reference
    operator sequence gets through cuDNN with TF32 convolutions on this GPU (torch's default allow_tf32; factor 2, floor 2e-2 of the
    gradient's norm).  Gradients of the first layers pass back through ~45 TF32 contractions and are cancelling sums: cuDNN's own
    TF32 run is tens of per cent off the truth there, which is why the bound is relative to it.  Then four steps from a CUDA graph:
    the loss must go down."""
    import fastscnn_oracle as fo
    from fscnn_b200 import Trainer, train_ops
    from models.fast_scnn import FastSCNN
    nc = 19
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in fo.make_state_dict(nc, True, 13).items()}
    x = torch.from_numpy(fo.make_input(2, 128, 192, 50)).to(DEV)
    target = torch.from_numpy(fo.make_labels(2, 128, 192, nc, 51)).to(DEV)

    def fresh():
        m = FastSCNN(nc, aux=True)
        m.load_state_dict(sd)
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.momentum = 0.0                     # several forwards below must see the same buffers
        return m.to(DEV).train()

    old_tf32 = torch.backends.cudnn.allow_tf32
    try:
        model = fresh()
        trainer = Trainer(model, aux_weight=0.4, fused_loss=False)
        outs = model(x)
        loss32 = trainer.loss(outs, target)
        kept = [(g_.abs().sum(1) > 0) for g_ in torch.autograd.grad(loss32, outs)]
        truth = _float64_truth(model, x, target, kept, trainer)
        torch.backends.cudnn.allow_tf32 = True
        yard = _float64_truth(model, x, target, kept, trainer, dtype=torch.float32)
        train_ops.set_matmul_precision('tf32')
        model.zero_grad()
        for p in model.parameters():
            p.grad = None
        outs = model(x)
        # the same kept pixels as the truth: plain weighted cross entropy on them, like _float64_truth
        import torch.nn.functional as F
        total = 0.0
        for i, (o, keep) in enumerate(zip(outs, kept)):
            li = F.cross_entropy(o, torch.where(keep, target, torch.full_like(target, -1)), weight=trainer.class_weight, ignore_index=-1)
            total = total + (li if i == 0 else trainer.aux_weight * li)
        assert abs(float(total.detach()) - float(loss32.detach())) <= 5e-4 * abs(float(loss32.detach()))
        total.backward()
        worst = 0.0
        for k, p in model.named_parameters():
            t = truth[k].ravel()
            scale = max(np.linalg.norm(t), 1e-4 * np.sqrt(t.size))
            d_ours = np.linalg.norm(p.grad.detach().cpu().numpy().ravel().astype(np.float64) - t) / scale
            d_yard = np.linalg.norm(yard[k].ravel().astype(np.float64) - t) / scale
            worst = max(worst, d_ours / max(d_yard, 1e-2))
            assert d_ours <= max(2e-2, 2.0 * d_yard), (k, d_ours, d_yard)
        print('worst ratio of our TF32 gradient error to cuDNN TF32 (floored at 1e-2):', worst)
        tr = Trainer(fresh(), base_lr=0.01, nepochs=1, iters_per_epoch=10, cuda_graph=True, graph_warmup=1, matmul_precision='tf32')
        losses = [float(tr.step(x, target)) for _ in range(4)]
        assert losses[-1] < losses[0], losses
    finally:
        torch.backends.cudnn.allow_tf32 = old_tf32
        train_ops.set_matmul_precision('fp32')


@pytest.mark.parametrize('n,h,w', [(2, 65, 97), (1, 96, 160), (2, 70, 131), (1, 35, 36)])
def test_stem_conv_direct_kernels_against_torch(n, h, w):
    """The direct stem kernels (Conv2d(3, 32, 3, stride 2, padding 0): window in registers, no column matrix) against F.conv2d in
    float64, forward and weight gradient, at even and odd widths (8-byte row loads or not), partial column / row blocks; and
    against the im2col path they replace (taken when the input wants a gradient)."""
    import torch.nn.functional as F
    from fscnn_b200 import train_ops
    g = torch.Generator(device='cpu').manual_seed(7 * h + w)
    x = torch.randn(n, 3, h, w, generator=g).to(DEV)
    wt0 = (torch.randn(32, 3, 3, 3, generator=g) / 5).to(DEV)
    ho, wo = (h - 3) // 2 + 1, (w - 3) // 2 + 1
    dy = torch.randn(n, 32, ho, wo, generator=g).to(DEV)
    wt = wt0.clone().requires_grad_(True)
    y = train_ops.conv3x3_dense(x, wt, 2, 0)
    assert y.grad_fn is not None and 'StemConv' in type(y.grad_fn).__name__
    y.backward(dy)
    wr = wt0.double().requires_grad_(True)
    yr = F.conv2d(x.double(), wr, None, 2, 0)
    yr.backward(dy.double())
    assert rel_err(y.detach().cpu().numpy(), yr.detach().float().cpu().numpy()) < 2e-6
    assert rel_err(wt.grad.cpu().numpy(), wr.grad.float().cpu().numpy()) < 5e-6
    xg, wt2 = x.clone().requires_grad_(True), wt0.clone().requires_grad_(True)
    y2 = train_ops.conv3x3_dense(xg, wt2, 2, 0)              # the im2col path (input gradient wanted)
    assert 'Conv3x3Dense' in type(y2.grad_fn).__name__
    y2.backward(dy)
    assert rel_err(y.detach().cpu().numpy(), y2.detach().cpu().numpy()) < 2e-6
    assert rel_err(wt.grad.cpu().numpy(), wt2.grad.cpu().numpy()) < 5e-6


# ---- the reference's other criteria: MixSoftmaxCrossEntropyLoss, DiceLoss / MixDiceLoss (train.py's default), FocalDiceLoss ----
def _loss_module(kind, kw, aux_weight):
    from utils import loss as L
    if kind == 'dice':
        return (L.MixDiceLoss(aux=True, aux_weight=aux_weight, **kw), False) if aux_weight is not None else (L.DiceLoss(**kw), True)
    if kind == 'ce':
        return L.MixSoftmaxCrossEntropyLoss(aux=aux_weight is not None, aux_weight=aux_weight or 0.2, **kw), False
    return L.FocalDiceLoss(**kw), True


def test_criteria_match_reference_fixtures():
    """utils/loss.py drop-ins (device kernels behind the reference's class names) against the loss and the gradients the unmodified
    reference classes produce under torch.autograd (oracle/gen_golden_loss.py); '*_low' cases hand over LOW-RESOLUTION logits, the
    reference side resizes them first (models/fast_scnn.py:40).  Loss 1e-5 relative, gradients 2e-5 of absmax."""
    from helpers import LOSS_CASES
    g = _load('train_loss_cases')
    for name, (kind, kw, aux_weight) in LOSS_CASES.items():
        crit, single = _loss_module(kind, kw, aux_weight)
        target = torch.from_numpy(g[name + '/target']).to(DEV)
        heads = [torch.from_numpy(g[f'{name}/logits{i}']).to(DEV).requires_grad_(True) for i in range(2 if aux_weight is not None else 1)]
        loss = crit(heads[0], target) if single else crit(tuple(heads), target)
        (3.0 * loss).backward()
        want = float(g[name + '/loss'])
        assert abs(float(loss.detach()) - want) <= 1e-5 * abs(want), (name, float(loss.detach()), want)
        for i, t in enumerate(heads):
            assert rel_err(t.grad.cpu().numpy(), 3.0 * g[f'{name}/grad{i}']) < 2e-5, (name, i)


@pytest.mark.parametrize('kind,n,c,hl,wl,h,w', [('ce', 2, 19, 12, 16, 96, 128), ('dice', 2, 2, 9, 13, 65, 97), ('focal_dice', 1, 2, 12, 12, 96, 96),
                                                 ('ce', 1, 7, 10, 20, 80, 160), ('dice', 2, 1, 6, 8, 48, 64), ('focal_dice', 2, 4, 5, 9, 40, 70),
                                                 ('ce', 1, 12, 6, 8, 48, 64), ('focal_dice', 1, 21, 6, 8, 48, 64), ('ce', 1, 40, 6, 8, 48, 64)])
def test_fused_upsample_criteria_equal_the_two_step_form_and_the_oracle(kind, n, c, hl, wl, h, w):
    """criterion(low) with the resize inside the loss kernels == criterion(bilinear_resize(low)) (same interpolation arithmetic: the
    loss agrees to rounding of the final sum), and both follow the float64 oracle."""
    import loss_oracle as lo
    from fscnn_b200 import train_ops
    gen = torch.Generator().manual_seed(100 + c + hl)
    low = (torch.randn((n, c, hl, wl), generator=gen) * 3.0)
    if kind == 'ce':
        target = torch.randint(-1, c, (n, h, w), generator=gen)
    elif kind == 'dice' or c <= 2:
        target = (torch.rand((n, h, w), generator=gen) < 0.3).long()
    else:
        target = torch.randint(0, c, (n, h, w), generator=gen)
        target[torch.rand((n, h, w), generator=gen) < 0.05] = -100
    a = low.clone().to(DEV).requires_grad_(True)
    b = low.clone().to(DEV).requires_grad_(True)
    t = target.to(DEV)
    fused = train_ops.criterion(a, t, kind)
    two = train_ops.criterion(train_ops.bilinear_resize(b, (h, w)), t, kind)
    fused.backward()
    two.backward()
    fused, two = fused.detach(), two.detach()
    assert abs(float(fused) - float(two)) <= 1e-6 * abs(float(two))
    assert rel_err(a.grad.cpu().numpy(), b.grad.cpu().numpy()) < 2e-5
    want, grad = lo.criterion_upsampled(kind, low.numpy(), target.numpy())
    assert abs(float(fused) - want) <= 1e-5 * abs(want)
    assert rel_err(a.grad.cpu().numpy(), grad.astype(np.float32)) < 2e-5


def test_trainer_with_the_default_dice_criterion():
    """train.py's default --loss-type (MixDiceLoss) on a 2-class model with the aux head: the Trainer's fused-resize loss equals
    utils.loss.MixDiceLoss on the model's full-resolution training outputs, and a few steps reduce it; 'ce' and 'focal_dice' step too."""
    import fastscnn_oracle as fo
    from fscnn_b200 import Trainer
    from models.fast_scnn import FastSCNN
    from utils.loss import MixDiceLoss
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in fo.make_state_dict(2, True, seed=9).items()}
    x = torch.from_numpy(fo.make_input(2, 64, 96, seed=3)).to(DEV)
    t = (torch.rand((2, 64, 96), generator=torch.Generator().manual_seed(1)) < 0.25).long().to(DEV)
    model = FastSCNN(2, aux=True)
    model.load_state_dict(sd)
    model.to(DEV).train()
    trainer = Trainer(model, base_lr=1e-2, loss_type='dice')
    for m in model.modules():      # identical forward twice: no dropout
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    with torch.no_grad():
        low = model._train_forward_lowres(x)
        full = model(x)
        fused = float(trainer.loss_from_lowres(low, t))
        ref = float(MixDiceLoss(aux=True, aux_weight=0.4)(full, t))
    assert abs(fused - ref) <= 1e-5 * abs(ref)
    losses = [float(trainer.step(x, t)) for _ in range(6)]
    assert all(np.isfinite(losses)) and losses[-1] < losses[0], losses
    for kind in ('ce', 'focal_dice'):
        m2 = _no_dropout(FastSCNN(2, aux=True))
        m2.load_state_dict(sd)
        m2.to(DEV).train()
        tr = Trainer(m2, base_lr=1e-2, loss_type=kind)
        ls = [float(tr.step(x, t)) for _ in range(4)]
        assert all(np.isfinite(ls)) and ls[-1] < ls[0], (kind, ls)


def test_adamw_step_matches_torch_adamw():
    """fscnn_train_adamw_step (train_bdd100k.py:183's optimizer) against torch.optim.AdamW over five steps on a flat buffer with changing
    learning rates, then through the Trainer on the network."""
    import fastscnn_oracle as fo
    from fscnn_b200 import Trainer, train_ops
    from models.fast_scnn import FastSCNN
    gen = torch.Generator().manual_seed(3)
    p0 = torch.randn(100_003, generator=gen)
    ours = p0.clone().to(DEV)
    ref = torch.nn.Parameter(p0.clone().to(DEV))
    opt = torch.optim.AdamW([ref], lr=1e-3, weight_decay=1e-4)
    m, v = torch.zeros_like(ours), torch.zeros_like(ours)
    for step in range(1, 6):
        g = torch.randn(p0.shape, generator=gen).to(DEV) * (0.1 if step % 2 else 3.0)
        lr = 1e-3 * (1.0 - step / 10.0) ** 0.9
        for gp in opt.param_groups:
            gp['lr'] = lr
        ref.grad = g.clone()
        opt.step()
        train_ops.adamw_step(ours, 2.0 * g, m, v, lr, step, weight_decay=1e-4, grad_scale=0.5)
    assert rel_err(ours.cpu().numpy(), ref.detach().cpu().numpy()) < 1e-6
    state = opt.state[ref]
    assert rel_err(m.cpu().numpy(), state['exp_avg'].cpu().numpy()) < 1e-6
    assert rel_err(v.cpu().numpy(), state['exp_avg_sq'].cpu().numpy()) < 1e-6
    model = _no_dropout(FastSCNN(2, aux=True))
    model.load_state_dict({k: torch.from_numpy(np.asarray(a)) for k, a in fo.make_state_dict(2, True, seed=9).items()})
    model.to(DEV).train()
    x = torch.from_numpy(fo.make_input(2, 64, 96, seed=3)).to(DEV)
    t = (torch.rand((2, 64, 96), generator=gen) < 0.25).long().to(DEV)
    trainer = Trainer(model, base_lr=1e-3, weight_decay=1e-4, loss_type='dice', optimizer='adamw')
    losses = [float(trainer.step(x, t)) for _ in range(6)]
    assert all(np.isfinite(losses)) and losses[-1] < losses[0], losses


@pytest.mark.parametrize('use_fp16,loss_type', [(True, 'dice'), (False, 'ce_ohem'), (True, 'ce_plain')])
def test_reference_training_loop_body_runs_on_the_drop_in_modules(use_fp16, loss_type):
    """The statements of the reference's training iteration (train.py:257-284: LRScheduler -> param_group['lr'] -> zero_grad ->
    [autocast] forward + criterion -> [GradScaler] backward + optimizer.step) with the objects train.py builds (:169, :182-198,
    :205-207), imported from this package instead: get_fast_scnn, utils.loss criteria, torch.optim.SGD on the module's parameters,
    utils.lr_scheduler.  The training operators keep fp32 under autocast; the loss must fall on a fixed batch."""
    import fastscnn_oracle as fo
    from torch.cuda.amp import GradScaler, autocast
    from models.fast_scnn import get_fast_scnn
    from utils.loss import MixDiceLoss, MixSoftmaxCrossEntropyLoss, MixSoftmaxCrossEntropyOHEMLoss
    from utils.lr_scheduler import LRScheduler
    model = _no_dropout(get_fast_scnn(dataset='citys', aux=True)).to(DEV)                      # train.py:169 (dropout off: see _no_dropout)
    if loss_type == 'dice':
        criterion = MixDiceLoss(aux=True, aux_weight=0.4).to(DEV)                              # train.py:184
        targets = (torch.rand((2, 96, 128), generator=torch.Generator().manual_seed(2)) < 0.3).long().to(DEV)
    elif loss_type == 'ce_ohem':
        criterion = MixSoftmaxCrossEntropyOHEMLoss(aux=True, aux_weight=0.4, ignore_index=-1).to(DEV)     # train.py:190-191
        targets = torch.from_numpy(fo.make_labels(2, 96, 128, 19, seed=4)).to(DEV)
    else:
        criterion = MixSoftmaxCrossEntropyLoss(True, 0.4, ignore_index=-1).to(DEV)             # train_custom_finetune.py:99
        targets = torch.from_numpy(fo.make_labels(2, 96, 128, 19, seed=4)).to(DEV)
    optimizer = torch.optim.SGD(model.parameters(), lr=1e-2, momentum=0.9, weight_decay=1e-4)  # train.py:195-198
    scaler = GradScaler() if use_fp16 else None                                                # train.py:201
    lr_scheduler = LRScheduler(mode='poly', base_lr=1e-2, nepochs=2, iters_per_epoch=4, power=0.9)
    images = torch.from_numpy(fo.make_input(2, 96, 128, seed=8)).to(DEV)
    model.train()
    losses = []
    for cur_iters in range(6):
        cur_lr = lr_scheduler(cur_iters)
        for param_group in optimizer.param_groups:
            param_group['lr'] = cur_lr
        optimizer.zero_grad()
        if use_fp16:
            with autocast():
                outputs = model(images)
                loss = criterion(outputs, targets)
            scaler.scale(loss).backward()
            scaler.step(optimizer)
            scaler.update()
        else:
            outputs = model(images)
            loss = criterion(outputs, targets)
            loss.backward()
            optimizer.step()
        losses.append(loss.item())
    assert outputs[0].shape == (2, 19, 96, 128) and outputs[0].dtype == torch.float32 and len(outputs) == 2
    assert all(np.isfinite(losses)) and min(losses[-2:]) < losses[0], losses
    assert all(torch.isfinite(p).all() for p in model.parameters())


@pytest.mark.parametrize('use_fp16,loss_type', [(True, 'dice'), (False, 'ce_ohem')])
def test_reference_validation_loop_body_runs_on_the_drop_in_modules(use_fp16, loss_type):
    """The statements of the reference's validation iteration (train.py:373-395: model.eval(), [autocast] forward + criterion,
    torch.argmax, SegmentationMetric.update on host arrays, metric.get) on the drop-in modules: the eval-mode forward is the
    inference engine (fp32 path), the criterion runs on its full-resolution logits; loss against the float64 oracles on the same
    logits, pixAcc / mIoU against the metric oracle."""
    import fastscnn_oracle as fo
    import loss_oracle as lo
    import metric_oracle as mo
    import ohem_oracle as oo
    from torch.cuda.amp import autocast
    from models.fast_scnn import FastSCNN
    from utils.loss import MixDiceLoss, MixSoftmaxCrossEntropyOHEMLoss
    from utils.metric import SegmentationMetric
    nc = 19
    x = fo.make_input(2, 96, 128, seed=8)
    sd = fo.calibrate_classifier_bias(fo.make_state_dict(nc, True, seed=7), x)
    model = FastSCNN(nc, aux=True)
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    model.to(DEV)
    if loss_type == 'dice':
        criterion = MixDiceLoss(aux=True, aux_weight=0.4).to(DEV)
        target = (torch.rand((2, 96, 128), generator=torch.Generator().manual_seed(2)) < 0.3).long().to(DEV)
    else:
        criterion = MixSoftmaxCrossEntropyOHEMLoss(aux=True, aux_weight=0.4, ignore_index=-1).to(DEV)
        target = torch.from_numpy(fo.make_labels(2, 96, 128, nc, seed=4)).to(DEV)
    image = torch.from_numpy(x).to(DEV)
    metric = SegmentationMetric(nc)
    metric.reset()
    model.eval()
    with torch.no_grad():
        if use_fp16:
            with autocast():
                outputs = model(image)
                loss = criterion(outputs, target)
        else:
            outputs = model(image)
            loss = criterion(outputs, target)
        val_loss = loss.item()
        pred = torch.argmax(outputs[0], 1)
        pred = pred.cpu().data.numpy()
        target_np = target.cpu().numpy()
        metric.update(pred, target_np)
    pix_acc, miou = metric.get()
    # (the reference's forward returns the aux prediction in eval mode too, fast_scnn.py:42-45)
    assert isinstance(outputs, tuple) and len(outputs) == 2 and outputs[0].shape == (2, nc, 96, 128) and outputs[1].shape == (2, nc, 96, 128)
    want = 0.0
    for scale, out in zip((1.0, 0.4), outputs):
        logits = out.cpu().numpy()
        if loss_type == 'dice':
            want += scale * lo.dice(logits, target_np)[0]
        else:
            want += scale * float(oo.ohem_loss_and_grad(logits, target_np, np.asarray(criterion.weight.cpu()), -1, 0.7, 256)[0])
    assert abs(val_loss - want) <= 1e-4 * abs(want), (val_loss, want)
    o = mo.SegmentationMetricOracle(nc)
    o.update(pred, target_np)
    assert (pix_acc, miou) == o.get()
    assert np.array_equal(pred, model.predict(image).cpu().numpy())      # the fused argmax agrees with torch.argmax of the logits
