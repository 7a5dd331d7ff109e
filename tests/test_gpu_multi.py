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
