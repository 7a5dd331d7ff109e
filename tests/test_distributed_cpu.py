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
