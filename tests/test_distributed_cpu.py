"""N > 1 host logic on CPU with gloo (world_size 2): sharding a batch over ranks and summing the
per-rank confusion accumulators gives the single-process totals.  The GPU path does the same with
NCCL (SegmentationMetric.all_reduce)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import metric_oracle as mo


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, nc, pred, label, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    shard = slice(rank * pred.shape[0] // world, (rank + 1) * pred.shape[0] // world)   # contiguous batch split
    conf = torch.from_numpy(mo.confusion_counts(pred[shard], label[shard], nc))
    dist.all_reduce(conf, op=dist.ReduceOp.SUM)       # the path's only collective
    if rank == 0:
        out.put(conf.numpy())
    dist.destroy_process_group()


def test_sharded_confusion_all_reduce_matches_single_process():
    nc, world = 19, 2
    rng = np.random.RandomState(3)
    pred = rng.randint(0, nc, size=(6, 33, 41)).astype(np.int64)
    label = rng.randint(-1, nc + 1, size=(6, 33, 41)).astype(np.int64)
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nc, pred, label, q)) for r in range(world)]
    for p in procs:
        p.start()
    total = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert np.array_equal(total, mo.confusion_counts(pred, label, nc))
    inter, union, correct, labeled = mo.totals_from_confusion(total, nc)
    o = mo.SegmentationMetricOracle(nc)
    o.update(pred, label)
    assert np.array_equal(inter, o.total_inter) and np.array_equal(union, o.total_union)
    assert (correct, labeled) == (o.total_correct, o.total_label)


def _product_worker(rank, world, port, nc, pred, label, out):
    """The PRODUCT's SegmentationMetric on two gloo ranks: each rank's accumulator holds its shard's counts (the counting
    kernel needs a GPU, so the counts come from the oracle and are placed into the metric's state tensor); all_reduce(),
    the totals (fscnn_conf_to_totals in the shared library) and get() are the host logic under test."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'fast-scnn-pytorch_b200'))
    from utils.metric import SegmentationMetric
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    shard = slice(rank * pred.shape[0] // world, (rank + 1) * pred.shape[0] // world)
    metric = SegmentationMetric(nc)
    metric._conf = torch.from_numpy(mo.confusion_counts(pred[shard], label[shard], nc).astype(np.int64))
    metric._device = torch.device('cpu')
    assert metric._conf.numel() == metric.conf_len()
    metric.all_reduce()
    out.put((rank, metric.total_inter, metric.total_union, int(metric.total_correct), int(metric.total_label), metric.get()))
    dist.destroy_process_group()


def test_product_metric_all_reduce_on_two_ranks():
    nc, world = 5, 2
    rng = np.random.RandomState(11)
    pred = rng.randint(0, nc, size=(4, 21, 35)).astype(np.int64)
    label = rng.randint(-2, nc + 2, size=(4, 21, 35)).astype(np.int64)
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_product_worker, args=(r, world, port, nc, pred, label, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    o = mo.SegmentationMetricOracle(nc)
    o.update(pred, label)
    for rank, inter, union, correct, labeled, scores in results:      # every rank ends with the global result
        assert np.array_equal(inter, o.total_inter) and np.array_equal(union, o.total_union), rank
        assert (correct, labeled) == (o.total_correct, o.total_label), rank
        assert scores == o.get(), rank
