"""Two ranks, two GPUs, NCCL: the product's sharded evaluation (FastSCNN.evaluate on each rank's shard, then
SegmentationMetric.all_reduce) against the oracle's confusion counts of the whole batch.  Skipped on a single-GPU box (NCCL
refuses two ranks on one device); run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import fastscnn_oracle as fo
import metric_oracle as mo

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, nc, sd, x, labels, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in ('fast-scnn-pytorch_b200', 'oracle', 'tests'):
        sys.path.insert(0, os.path.join(root, p))
    import torch.distributed as dist
    from helpers import build_model
    from utils.metric import SegmentationMetric
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    shard = slice(rank * x.shape[0] // world, (rank + 1) * x.shape[0] // world)
    model = build_model(sd, nc, False, dev)
    metric = SegmentationMetric(nc, device=dev)
    mask = torch.empty((shard.stop - shard.start,) + x.shape[2:], dtype=torch.uint8, device=dev)
    model.evaluate(torch.from_numpy(x[shard]).to(dev), torch.from_numpy(labels[shard]).to(dev), metric, mask=mask)
    metric.all_reduce()
    out.put((rank, shard.start, mask.cpu().numpy(), metric.device_confusion().cpu().numpy(), metric.get()))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_sharded_evaluation_and_nccl_all_reduce():
    nc, world, n, h, w = 19, 2, 6, 136, 200
    sd = fo.make_state_dict(nc, False, 7)
    x = fo.make_input(n, h, w, 8)
    sd = fo.calibrate_classifier_bias(sd, x)
    labels = fo.make_labels(n, h, w, nc, seed=9, adversarial=True)
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nc, sd, x, labels, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    mask = np.zeros((n, h, w), np.uint8)
    for rank, start, m, conf, scores in results:
        mask[start:start + m.shape[0]] = m
    want = mo.confusion_counts(mask, labels, nc)
    o = mo.SegmentationMetricOracle(nc)
    o.update(mask.astype(np.int64), labels)
    for rank, start, m, conf, scores in results:
        assert np.array_equal(conf, want), rank           # every rank holds the global counts after the all-reduce
        assert scores == o.get(), rank
    ref_mask = fo.argmax_classes(fo.forward(sd, x)[0])      # and the masks are the reference's
    near_tie = fo.top2_margin(fo.forward(sd, x)[0]) < 1e-4 * np.abs(fo.forward(sd, x)[0]).max()
    assert int(((mask != ref_mask) & ~near_tie).sum()) == 0


def _train_worker(rank, world, port, nc, sd, x, labels, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in ('fast-scnn-pytorch_b200', 'oracle', 'tests'):
        sys.path.insert(0, os.path.join(root, p))
    import torch.distributed as dist
    from fscnn_b200 import Trainer
    from models.fast_scnn import FastSCNN
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)

    def build():
        m = FastSCNN(nc, aux=True)
        m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
        return m.to(dev).train()

    shard = slice(rank * x.shape[0] // world, (rank + 1) * x.shape[0] // world)
    xs, ts = torch.from_numpy(x[shard]).to(dev), torch.from_numpy(labels[shard]).to(dev)
    # the DDP trainer: per-rank BatchNorm, ONE all-reduce of the flat gradient buffer, the same update on every rank
    # (fused_loss=False: deterministic loss backward, so the emulation below can be compared at rounding level)
    ddp = Trainer(build(), base_lr=0.02, nepochs=1, iters_per_epoch=10, fused_loss=False)
    loss = float(ddp.step(xs, ts))
    # emulation with plain collectives: a world-size-1 trainer's gradients of the local shard, summed over the ranks by hand
    dist.barrier()
    solo_model = build()
    solo = Trainer.__new__(Trainer)
    solo.__dict__.update(ddp.__dict__)
    solo.model, solo.world = solo_model, 1
    params = [p for p in solo_model.parameters() if p.requires_grad]
    for p in params:
        p.grad = None
    Trainer.loss(solo, solo_model(xs), ts).backward()
    flat_g = torch.cat([p.grad.reshape(-1) for p in params])
    gathered = [torch.empty_like(flat_g) for _ in range(world)]
    dist.all_gather(gathered, flat_g)
    mean_g = sum(gathered) / world
    flat_p = torch.cat([p.detach().reshape(-1) for p in params])
    want = flat_p - 0.02 * (mean_g + 1e-4 * flat_p)            # first SGD step: buf = g + wd p, p -= lr buf (poly lr at iteration 0 = base)
    out.put((rank, loss, ddp.flat_param.cpu().numpy(), want.cpu().numpy()))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_ddp_training_step_averages_gradients_over_two_ranks():
    """Trainer under torch.distributed (NCCL, two GPUs): after one step every rank holds the SAME parameters, and they equal the SGD
    update with the mean of the two ranks' local gradients (DDP semantics, train.py run under torch.distributed.launch)."""
    nc, world, n, h, w = 19, 2, 4, 96, 96
    sd = fo.make_state_dict(nc, True, 7)
    x = fo.make_input(n, h, w, 8)
    labels = fo.make_labels(n, h, w, nc, seed=9)
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_train_worker, args=(r, world, port, nc, sd, x, labels, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted([q.get(timeout=600) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    (_, loss0, got0, want0), (_, loss1, got1, want1) = results
    assert np.isfinite(loss0) and np.isfinite(loss1) and loss0 != loss1        # different shards
    assert np.array_equal(got0, got1)                                           # identical replicas after the step
    step = np.abs(got0 - want0).max() / np.abs(want0).max()
    assert step < 1e-6, step
    assert np.abs(got0 - want0).max() <= 1e-3 * np.abs(got0 - _flat_params(sd, nc)).max()     # a small fraction of the update itself


def _flat_params(sd, nc):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, 'fast-scnn-pytorch_b200'))
    from models.fast_scnn import FastSCNN
    m = FastSCNN(nc, aux=True)
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    return torch.cat([p.detach().reshape(-1) for p in m.parameters() if p.requires_grad]).numpy()
