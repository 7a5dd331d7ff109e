"""Generates the golden vectors of the training-step slice by running the UNMODIFIED reference
modules in train mode under torch.autograd on the CPU, in the build container (TEST INFRASTRUCTURE; /root/reference does not
exist on the GPU box, so the vectors are committed under tests/golden/ together with this script):

    python oracle/gen_golden_train.py

  train_dsconv_*.npz      models.fast_scnn._DSConv            (fast_scnn.py:64-79)
  train_bottleneck_*.npz  models.fast_scnn.LinearBottleneck   (fast_scnn.py:95-115)
  train_ohem_*.npz        utils.loss.SoftmaxCrossEntropyOHEMLoss (loss.py:127-182; its hard-coded .cuda() is patched to a no-op)

Each module fixture holds: the seeded parameters / buffers before the step, the input, the upstream gradient, and the
reference's output, input gradient, parameter gradients and updated BatchNorm buffers after ONE forward + backward."""
import os
import sys

import numpy as np
import torch

REF = '/root/reference'
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
from models.fast_scnn import LinearBottleneck, _DSConv  # noqa: E402


def seed_module(m, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, torch.nn.Conv2d):
                fan_in = mod.weight[0].numel()
                mod.weight.copy_(torch.randn(mod.weight.shape, generator=g) * (2.0 / fan_in) ** 0.5)
                if mod.bias is not None:
                    mod.bias.copy_(torch.randn(mod.bias.shape, generator=g) * 0.05)
            elif isinstance(mod, torch.nn.BatchNorm2d):
                mod.weight.copy_(torch.rand(mod.weight.shape, generator=g) * 0.4 + 0.8)
                mod.bias.copy_(torch.randn(mod.bias.shape, generator=g) * 0.1)
                mod.running_mean.copy_(torch.randn(mod.running_mean.shape, generator=g) * 0.1)
                mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=g) * 0.4 + 0.8)


def module_case(name, module, x_shape, seed):
    seed_module(module, seed)
    module.train()
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(x_shape, generator=g, requires_grad=True)
    before = {k: v.detach().clone().numpy() for k, v in module.state_dict().items()}
    y = module(x)
    gy = torch.randn(y.shape, generator=g)
    y.backward(gy)
    out = {'x': x.detach().numpy(), 'gy': gy.numpy(), 'y': y.detach().numpy(), 'dx': x.grad.numpy()}
    for k, v in before.items():
        out['before/' + k] = v
    for k, v in module.state_dict().items():
        if 'running' in k or 'num_batches' in k:
            out['after/' + k] = v.detach().numpy()
    for k, p in module.named_parameters():
        out['grad/' + k] = p.grad.numpy()
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
    print(name, 'y absmax', float(y.abs().max()), 'dx absmax', float(x.grad.abs().max()))


def ohem_case(name, logits, target, use_weight, min_kept=256, thresh=0.7):
    from utils.loss import SoftmaxCrossEntropyOHEMLoss
    torch.Tensor.cuda = lambda self, *a, **k: self          # loss.py:180 moves the rebuilt target to the GPU
    crit = SoftmaxCrossEntropyOHEMLoss(ignore_label=-1, thresh=thresh, min_kept=min_kept, use_weight=use_weight)
    lg = torch.from_numpy(logits).requires_grad_(True)
    loss = crit(lg, torch.from_numpy(target))
    loss.backward()
    weight = crit.criterion.weight.numpy() if use_weight else np.zeros(0, np.float32)
    np.savez_compressed(os.path.join(OUT, name + '.npz'), logits=logits, target=target, weight=weight,
                        loss=np.float32(loss.item()), dlogits=lg.grad.numpy(), min_kept=np.int64(min_kept), thresh=np.float32(thresh))
    print(name, 'loss', loss.item(), 'kept pixels', int((lg.grad.abs().sum(1) > 0).sum()))


def network_case(name, nc, aux, x_shape, seed):
    """One whole training step of the unmodified reference network (train.py:253-284 without the optimizer): FastSCNN in train
    mode (BatchNorm batch statistics; Dropout p set to 0 because its mask cannot be reproduced bit for bit), the reference's
    MixSoftmaxCrossEntropyOHEMLoss with aux weight 0.4, backward.  Per parameter the fixture keeps the gradient's L2 norm, sum and
    every 13th element; plus the loss, a strided sample of both outputs and the updated BatchNorm buffers."""
    from models.fast_scnn import FastSCNN
    from utils.loss import MixSoftmaxCrossEntropyOHEMLoss
    torch.Tensor.cuda = lambda self, *a, **k: self
    import fastscnn_oracle as fo
    model = FastSCNN(nc, aux=aux)
    # recipe-D2 weights from numpy's frozen RandomState streams: the test rebuilds them from the seed (nothing to store)
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in fo.make_state_dict(nc, aux, seed).items()})
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    model.train()
    x = torch.from_numpy(fo.make_input(x_shape[0], x_shape[2], x_shape[3], seed + 1))
    target = torch.from_numpy(fo.make_labels(x_shape[0], x_shape[2], x_shape[3], nc, seed + 2))
    crit = MixSoftmaxCrossEntropyOHEMLoss(aux=aux, aux_weight=0.4, ignore_index=-1)
    outputs = model(x)
    loss = crit(outputs, target)
    loss.backward()
    out = {'loss': np.float32(loss.item()), 'meta': np.array([nc, int(aux), seed, x_shape[0], x_shape[2], x_shape[3]])}
    for i, o in enumerate(outputs):
        out[f'out{i}_sample'] = o.detach().numpy()[:, :, ::5, ::7]
        out[f'out{i}_absmax'] = np.float32(o.detach().abs().max())
    for k, v in model.state_dict().items():
        if 'running' in k or 'num_batches' in k:
            out['after/' + k] = v.detach().numpy()
    for k, p in model.named_parameters():
        gr = p.grad.numpy().ravel()
        out['gstat/' + k] = np.array([np.sqrt((gr.astype(np.float64) ** 2).sum()), gr.astype(np.float64).sum()], np.float64)
        out['gsample/' + k] = gr[::13].copy()
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
    print(name, 'loss', loss.item(), 'params', sum(p.numel() for p in model.parameters()))


def main():
    os.makedirs(OUT, exist_ok=True)
    network_case('train_net_nc19_aux', 19, True, (2, 3, 96, 128), 21)
    module_case('train_dsconv_32_48_s2', _DSConv(32, 48, 2), (2, 32, 25, 33), 11)
    module_case('train_dsconv_16_16_s1', _DSConv(16, 16, 1), (3, 16, 18, 20), 12)
    module_case('train_bottleneck_16_16_s1', LinearBottleneck(16, 16, 6, 1), (2, 16, 14, 22), 13)
    module_case('train_bottleneck_16_24_s2', LinearBottleneck(16, 24, 6, 2), (2, 16, 15, 21), 14)
    rng = np.random.RandomState(5)
    n, c, h, w = 2, 19, 40, 56
    target = rng.randint(-1, c, size=(n, h, w)).astype(np.int64)
    onehot = np.eye(c, dtype=np.float32)[np.clip(target, 0, c - 1)].transpose(0, 3, 1, 2)
    # (a) confident and mostly right: fewer than min_kept pixels below 0.7 -> the threshold is the 256-th smallest probability
    ohem_case('train_ohem_kth', (rng.standard_normal((n, c, h, w)) * 1.5 + 9.0 * onehot).astype(np.float32), target, True)
    # (b) diffuse predictions: plenty of hard pixels -> the threshold stays at 0.7
    ohem_case('train_ohem_thresh', (rng.standard_normal((n, c, h, w)) * 3.0 + 2.0 * onehot).astype(np.float32), target, True)
    # (c) fewer valid pixels than min_kept -> everything valid is kept; no class weights
    sparse = np.full((n, h, w), -1, np.int64)
    sparse[:, ::9, ::7] = target[:, ::9, ::7]
    ohem_case('train_ohem_keepall', (rng.standard_normal((n, c, h, w)) * 2.0).astype(np.float32), sparse, False)
    # (d) 2 classes, no weights, min_kept larger than the easy set
    t2 = rng.randint(-1, 2, size=(1, 30, 44)).astype(np.int64)
    oh2 = np.eye(2, dtype=np.float32)[np.clip(t2, 0, 1)].transpose(0, 3, 1, 2)
    ohem_case('train_ohem_nc2', (rng.standard_normal((1, 2, 30, 44)) + 4.0 * oh2).astype(np.float32), t2, False, min_kept=300)


if __name__ == '__main__':
    main()
