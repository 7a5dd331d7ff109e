#!/usr/bin/env python
"""bench.py -- Cityscapes-size (19 classes, 1024x2048) batched evaluation throughput of the B200
Fast-SCNN forward path: forward + fused upsample/argmax + SegmentationMetric counting.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision fp32|bf16] [--batch B]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   (N > 1)
    python bench.py --impl reference ...      (the reference's CPU path, timed on the host cores)

One "step" = one pass of the hot path over a batch of B synthetic images per GPU.  Rank 0 prints ONE
JSON line (see the keys below).  `value` is device-resident throughput (inputs already in HBM);
`e2e` is the same metric through the public Python API with pinned HOST buffers, host->device
copies and the device->host read of the metric state inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, 'fast-scnn-pytorch_b200'))

METRIC, UNIT = 'cityscapes_1024x2048_images_per_sec', 'images/s'


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='native', choices=['native', 'reference'])
    ap.add_argument('--precision', default=os.environ.get('FSCNN_BENCH_PRECISION', 'bf16'), choices=['fp32', 'bf16'],
                    help='bf16 = tensor-core fast path (headline); fp32 = exactness path (also timed briefly as fp32_exact)')
    ap.add_argument('--batch', type=int, default=111,
                    help='images per GPU per step (111 x 16 coarsest-level tiles = 12 full waves of 148 persistent CTAs)')
    ap.add_argument('--micro-batch', type=int, default=0, help='0 = library default')
    ap.add_argument('--height', type=int, default=1024)
    ap.add_argument('--width', type=int, default=2048)
    ap.add_argument('--classes', type=int, default=19)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-stage-times', action='store_true')
    ap.add_argument('--no-fp32', action='store_true', help='skip the short fp32 exactness-path timing')
    ap.add_argument('--no-latency', dest='latency', action='store_false', help='skip the batch-1 CUDA-graph latency')
    ap.add_argument('--only', default='', help="run one auxiliary measurement alone and print its JSON: 'eager'")
    ap.add_argument('--no-train', action='store_true', help='skip the training-step measurement (BASELINE config 5)')
    ap.add_argument('--no-eager', action='store_true', help='skip the cuDNN-eager comparator on the same GPU')
    ap.add_argument('--no-extra', action='store_true', help='skip the secondary workloads (480x640, 360x640)')
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return {'hbm_gbs': float(p['hbm_gbs']), 'bf16_tflops': float(p.get('bf16_tflops_sustained', p['bf16_tflops'])),
                'sm_max_mhz': float(p.get('sm_max_mhz', 1965.0)), 'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1400.0, 'sm_max_mhz': 1965.0, 'source': 'fallback'}


# ---------------------------------------------------------------------------------------------
# shapes, algorithmic bytes and flops per stage ("plan-P")
# ---------------------------------------------------------------------------------------------
def stage_model(h, w, nc, es):
    """Per-image algorithmic HBM bytes and flops of every kernel of the plan (each stage tensor
    written once and read once; activations `es` bytes/element; input image fp32)."""
    c = lambda n: (n - 1) // 2 + 1
    h1, w1 = (h - 3) // 2 + 1, (w - 3) // 2 + 1
    h2, w2 = c(h1), c(w1)
    h3, w3 = c(h2), c(w2)
    h4, w4 = c(h3), c(w3)
    h5, w5 = c(h4), c(w4)
    st = []

    def add(name, rd, wr, flops):
        st.append({'stage': name, 'bytes': float(rd + wr), 'flops': float(flops)})

    add('stem', 3 * h * w * 4, h1 * w1 * 32 * es, 2 * h1 * w1 * 32 * 27)
    add('l2d.dsconv1', h1 * w1 * 32 * es, h2 * w2 * 48 * es, 2 * h2 * w2 * (32 * 9 + 32 * 48))
    add('l2d.dsconv2', h2 * w2 * 48 * es, h3 * w3 * 64 * es, 2 * h3 * w3 * (48 * 9 + 48 * 64))
    plan = [('gfe.bottleneck1.0', 64, 64, h3, w3, h4, w4), ('gfe.bottleneck1.1', 64, 64, h4, w4, h4, w4),
            ('gfe.bottleneck1.2', 64, 64, h4, w4, h4, w4), ('gfe.bottleneck2.0', 64, 96, h4, w4, h5, w5),
            ('gfe.bottleneck2.1', 96, 96, h5, w5, h5, w5), ('gfe.bottleneck2.2', 96, 96, h5, w5, h5, w5),
            ('gfe.bottleneck3.0', 96, 128, h5, w5, h5, w5), ('gfe.bottleneck3.1', 128, 128, h5, w5, h5, w5),
            ('gfe.bottleneck3.2', 128, 128, h5, w5, h5, w5)]
    for name, ci, co, hi, wi, ho, wo in plan:
        add(name, hi * wi * ci * es, ho * wo * co * es,
            2 * (hi * wi * ci * 6 * ci + ho * wo * 6 * ci * 9 + ho * wo * 6 * ci * co))
    add('gfe.ppm', 2 * h5 * w5 * 128 * es, h5 * w5 * 128 * es, 2 * (h5 * w5 * 128 * 128 + 50 * 128 * 32 + 50 * 32 * 128))
    add('ffm', h3 * w3 * 64 * es + h5 * w5 * 128 * es, h3 * w3 * 128 * es, 2 * h3 * w3 * (128 * 9 + 192 * 128))
    add('cls.dsconv1', h3 * w3 * 128 * es, h3 * w3 * 128 * es, 2 * h3 * w3 * (128 * 9 + 128 * 128))
    add('cls.dsconv2+head', h3 * w3 * 128 * es, h3 * w3 * nc * 4, 2 * h3 * w3 * (128 * 9 + 128 * 128 + 128 * nc))
    add('up8+argmax+metric', h3 * w3 * nc * 4, h * w * 1, 6 * h * w * nc)
    return st


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.proc, self.lines, self.windows, self._t0 = index, None, [], [], None

    def _reader(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln))

    def begin(self):      # a timed region starts (the caller has synchronised the device)
        self._t0 = time.time()

    def end(self):        # ... and ends
        if self._t0 is not None:
            self.windows.append((self._t0, time.time()))
            self._t0 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '20'], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._reader, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons, power = [], [], set(), []
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        inside = [ln for ts, ln in self.lines if any(a <= ts <= b for a, b in self.windows)]
        for ln in inside:
            f = [v.strip() for v in ln.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'power_w_max': max(power) if power else None, 'samples': len(sm),
                'sampled_s': round(sum(b - a for a, b in self.windows), 3), 'reasons': sorted(reasons)}


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference's own path (ATen-functional port + numpy metric oracle) on the host cores
# ---------------------------------------------------------------------------------------------
def cpu_reference_rate(h, w, nc, budget_s, warmup, steps=None):
    """images/s of forward + argmax + host SegmentationMetric on the CPU, all host threads."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import numpy as np
    import torch
    import fastscnn_oracle as fo
    import fastscnn_torch_port as tp
    import metric_oracle as mo
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = tp.to_torch_state_dict(fo.make_state_dict(nc, False, 7))
    rng = np.random.RandomState(0)
    x = torch.from_numpy(rng.standard_normal((1, 3, h, w)).astype(np.float32))
    labels = rng.randint(-1, nc, size=(1, h, w)).astype(np.int64)
    metric = mo.SegmentationMetricOracle(nc)
    for _ in range(max(1, warmup)):
        tp.eval_step(sd, x, labels, nc, metric)
    t0, done = time.perf_counter(), 0
    while True:
        tp.eval_step(sd, x, labels, nc, metric)
        done += 1
        el = time.perf_counter() - t0
        if (steps is not None and done >= steps) or (steps is None and el >= budget_s and done >= 3):
            break
    return done / el, done, el, cores, torch.get_num_threads()


def gpu_eager_baseline(dev, h, w, nc, batches=(1, 16), reps=10):
    """The honest GPU comparator: the reference's own op sequence (ATen-functional port, unfolded BN, NCHW
    fp32 tensors) run eagerly through cuDNN / ATen on the SAME B200, fp32 and bf16 autocast, forward + torch.argmax,
    synchronised.  Returns images/s per (precision, batch).  None of this repo's kernels are on this path."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import torch
    import fastscnn_oracle as fo
    import fastscnn_torch_port as tp
    sd = {k: v.to(dev) for k, v in tp.to_torch_state_dict(fo.make_state_dict(nc, False, 7)).items()}
    res = {}
    torch.backends.cudnn.benchmark = True
    for b in batches:
        x = torch.randn(b, 3, h, w, device=dev)
        for prec in ('fp32', 'bf16_autocast'):
            def step():
                if prec == 'fp32':
                    return torch.argmax(tp.forward(sd, x)[0], 1)
                with torch.autocast('cuda', dtype=torch.bfloat16):
                    return torch.argmax(tp.forward(sd, x)[0], 1)
            for _ in range(3):
                step()
            torch.cuda.synchronize()
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                step()
            e.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(e) / reps
            res[f'{prec}_b{b}'] = {'images_per_s': b / (ms / 1e3), 'ms_per_step': ms}
        del x
    # BASELINE config 5 on the same comparator: the reference op sequence in TRAINING mode (batch-statistics BatchNorm, aux head),
    # F.cross_entropy on both heads (the OHEM pixel selection is left out: it only removes work), backward, torch.optim.SGD;
    # fp32 and the fp16 autocast + GradScaler the reference actually trains with (train.py:73-74, :267-275)
    import torch.nn.functional as F
    try:
        psd = {k: (v.to(dev).requires_grad_(v.dtype.is_floating_point and 'running' not in k))
               for k, v in tp.to_torch_state_dict(fo.make_state_dict(nc, True, 7)).items()}
        params = [v for v in psd.values() if v.requires_grad]
        opt = torch.optim.SGD(params, lr=1e-2, momentum=0.9, weight_decay=1e-4)
        fwd = getattr(tp.forward, '__wrapped__', tp.forward)
        real_bn = tp._bn
        tp._bn = lambda sd_, p_, t_: F.batch_norm(t_, sd_[p_ + '.running_mean'], sd_[p_ + '.running_var'], sd_[p_ + '.weight'], sd_[p_ + '.bias'],
                                                  True, 0.1, 1e-5)
        xt = torch.randn(16, 3, 768, 768, device=dev)
        tt = torch.randint(-1, nc, (16, 768, 768), device=dev)
        for prec in ('fp32', 'fp16_autocast'):
            scaler = torch.amp.GradScaler('cuda', enabled=prec != 'fp32')

            def tstep():
                opt.zero_grad(set_to_none=True)
                with torch.enable_grad(), torch.autocast('cuda', dtype=torch.float16, enabled=prec != 'fp32'):
                    o = fwd(psd, xt, True)
                    loss = F.cross_entropy(o[0].float(), tt, ignore_index=-1) + 0.4 * F.cross_entropy(o[1].float(), tt, ignore_index=-1)
                scaler.scale(loss).backward()
                scaler.step(opt)
                scaler.update()
                return loss
            for _ in range(2):
                tstep()
            torch.cuda.synchronize()
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(3):
                float(tstep())
            e.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(e) / 3
            res[f'train_step_{prec}_crop768_b16'] = {'images_per_s': 16 / (ms / 1e3), 'ms_per_step': ms}
        tp._bn = real_bn
        del psd, params, opt, xt, tt
    except Exception as exc:      # the comparator must never take the bench line down
        res['train_step_error'] = repr(exc)[:200]
    torch.backends.cudnn.benchmark = False
    torch.cuda.empty_cache()
    return res


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    rate, done, el, cores, threads = cpu_reference_rate(args.height, args.width, args.classes, 0.0, args.warmup, args.steps)
    sample = f'{done} timed single-image steps of {args.height}x{args.width} after {max(1, args.warmup)} warm-up'
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': rate, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': done,
        'warmup': max(1, args.warmup), 'ms_per_step': 1e3 * el / done, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'cityscapes_eval_nc{args.classes}_{args.height}x{args.width}_fwd_argmax_metric',
                   'batch_per_step': 1, 'note': 'reference path = ATen-functional port of models/fast_scnn.py + host metric, CPU'},
        'cpu_baseline': {'value': rate, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': sample,
                         'host_cpus': cores},
        'e2e': {'value': rate, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }))


# ---------------------------------------------------------------------------------------------
# native arm
# ---------------------------------------------------------------------------------------------
def init_recipe_d2(model, seed):
    """Recipe D2 of SURVEY.md appendix D, restated with torch RNG (the oracle's numpy version is test infrastructure and is not
    imported here): variance-preserving conv weights N(0, sqrt(2 / fan_in)), conv bias N(0, .05), BN gamma U(.8, 1.2),
    beta / mean N(0, .1), var U(.8, 1.2).  Unlike plain random init (recipe D1: one class wins 99.997 % of the pixels) the
    logits then have class regions and boundaries, which is what the data-dependent tail kernel has to be timed on."""
    import math
    import torch
    import torch.nn as nn
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, nn.Conv2d):
                fan_in = m.weight[0].numel()
                m.weight.copy_(torch.randn(m.weight.shape, generator=g) * math.sqrt(2.0 / fan_in))
                if m.bias is not None:
                    m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.05)
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.copy_(torch.rand(m.weight.shape, generator=g) * 0.4 + 0.8)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) * 0.4 + 0.8)


def smooth_images(n, h, w, dev, seed, chunk=8):
    """Recipe D2's input: sum of bilinearly upsampled Gaussian noise at strides 64 / 16 / 4 / 1 (amplitudes 1 / .6 / .3 / .15)."""
    import torch
    import torch.nn.functional as F
    g = torch.Generator(device=dev).manual_seed(seed)
    x = torch.empty((n, 3, h, w), device=dev)
    for i0 in range(0, n, chunk):
        m = min(chunk, n - i0)
        acc = torch.zeros((m, 3, h, w), device=dev)
        for stride, amp in ((64, 1.0), (16, 0.6), (4, 0.3), (1, 0.15)):
            hs, ws = max(2, -(-h // stride) + 1), max(2, -(-w // stride) + 1)
            z = torch.randn((m, 3, hs, ws), device=dev, generator=g)
            acc += amp * (z if (hs, ws) == (h, w) else F.interpolate(z, size=(h, w), mode='bilinear', align_corners=True))
        x[i0:i0 + m] = acc
    return x


def lowres_logits(eng, x, h, w):
    """Runs the network up to its low-resolution logits and returns them as a contiguous [n, hl, wl, padded_classes] tensor
    (the layout fscnn_upsample_argmax takes)."""
    import ctypes as C
    import torch
    from fscnn_b200 import native
    names = eng.stage_names()
    eng.forward_range(x, 0, names.index('cls.dsconv2+head'))
    tap = native.Tap()
    native.check(eng.lib.fscnn_tap_info(eng._ctx, x.shape[0], h, w, b'cls.logits_lowres', C.byref(tap)))
    ws = eng._workspace(x.shape[0], h, w)
    count = tap.n * tap.h * tap.w * tap.c_stride
    return ws[tap.offset_bytes: tap.offset_bytes + count * 4].view(torch.float32).view(tap.n, tap.h, tap.w, tap.c_stride).clone()


def time_cuda(fn, reps, warm=2):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def time_cuda_trials(fn, reps, trials=5, warm=2):
    """Median over `trials` event-bracketed groups of `reps` calls (ms per call) and the per-trial values.  One group of five
    launches can land 30 % off on a fresh box (the round-2 record holds one such outlier for the tail kernel timed alone while the
    same kernel inside the full step was on its usual time), so the per-kernel table uses the median and keeps the spread."""
    ts = sorted(time_cuda(fn, reps, warm=warm if i == 0 else 0) for i in range(trials))
    return ts[len(ts) // 2], ts


def run_native_arm(args):
    import torch
    import torch.distributed as dist
    from models.fast_scnn import FastSCNN
    from utils.metric import SegmentationMetric

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    numa = None
    try:   # multi-GPU boxes: run this rank (and first-touch its pinned host buffers) on the CPUs closest to its GPU
        import pynvml
        pynvml.nvmlInit()
        vis = [v for v in os.environ.get('CUDA_VISIBLE_DEVICES', '').split(',') if v.strip().isdigit()]
        hnd = pynvml.nvmlDeviceGetHandleByIndex(int(vis[local]) if local < len(vis) else local)
        pynvml.nvmlDeviceSetCpuAffinity(hnd)
        numa = sorted(os.sched_getaffinity(0))
        numa = f'{len(numa)} cpus ({numa[0]}-{numa[-1]})'
    except Exception:   # no NVML / no permission: keep the inherited affinity
        numa = None
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    h, w, nc, B = args.height, args.width, args.classes, args.batch
    es = 4 if args.precision == 'fp32' else 2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ---- the model and the batch: recipe D2 (class regions and boundaries; all classes present) ----
    model = FastSCNN(nc, precision=args.precision).eval()
    init_recipe_d2(model, 7)
    model.to(dev)
    eng = model._engine(dev)
    if args.micro_batch:
        eng.set_micro_batch(args.micro_batch)
    x = smooth_images(B, h, w, dev, 1234 + rank)
    with torch.no_grad():      # D2's calibration: subtract the mean logit per class so that every class wins somewhere
        low = lowres_logits(eng, x[:min(B, 8)].contiguous(), h, w)
        model.classifier.conv[1].bias -= low[..., :nc].mean(dim=(0, 1, 2))
    if world > 1:              # every rank evaluates the same weights
        dist.broadcast(model.classifier.conv[1].bias.data, src=0)
    eng = model._engine(dev)
    g = torch.Generator(device=dev).manual_seed(99 + rank)
    labels = torch.randint(-1, nc, (B, h, w), device=dev, dtype=torch.int64, generator=g)   # -1 = ignore
    labeled_per_step = int((labels >= 0).sum().item())
    metric = SegmentationMetric(nc, device=dev)
    probe = model.predict(x[:2].contiguous())
    frac = torch.bincount(probe.flatten().long(), minlength=nc).float() / probe.numel()
    classes_present, largest_class = int((frac > 1e-4).sum().item()), float(frac.max().item())
    del probe

    def step_device():
        model.evaluate(x, labels, metric)

    # ---- device-resident throughput: K steps + the path's only collective + the metric read-back, all inside the timed region ----
    sampler = ClockSampler(local)      # started before the warm-up so that it is already sampling (every 20 ms) when the timed
    if rank == 0:                      # regions run; only samples taken inside a timed region (device-resident + e2e) are kept
        sampler.start()
    for _ in range(max(3, args.warmup)):
        step_device()
    metric.reset()
    barrier()
    sampler.begin()
    launches0 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step_device()
    metric.all_reduce()                # one NCCL sum of the int64 confusion state (no-op on one GPU)
    pix_acc, miou = metric.get()       # device -> host read of the reduced state, float64 pixAcc / mIoU as metric.py:42-54
    ev1.record()
    torch.cuda.synchronize()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    launches = eng.launch_count() - launches0
    sampler.end()
    expect = torch.tensor([labeled_per_step * args.steps], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(expect)
    if int(metric.total_label) != int(expect.item()):
        raise SystemExit(f'metric check failed: total_label {metric.total_label} != {int(expect.item())} labeled pixels evaluated')
    barrier()
    value = world * B * args.steps / (ms / 1e3)

    # ---- the same on recipe D1 (plain random init + randn images: one class wins everywhere -- what round 1 timed) ----
    torch.manual_seed(1234 + rank)
    model_d1 = FastSCNN(nc, precision=args.precision).eval()
    with torch.no_grad():
        for name, buf in model_d1.named_buffers():
            if name.endswith('running_mean'):
                buf.normal_(0, 0.1)
            elif name.endswith('running_var'):
                buf.uniform_(0.8, 1.2)
    model_d1.to(dev)
    x_d1 = torch.randn(min(B, 37), 3, h, w, device=dev)
    met_d1 = SegmentationMetric(nc, device=dev)
    lab_d1 = labels[:x_d1.shape[0]].contiguous()
    barrier()
    ms_d1 = max_over_ranks(time_cuda(lambda: model_d1.evaluate(x_d1, lab_d1, met_d1), 5, warm=3))
    value_d1 = world * x_d1.shape[0] / (ms_d1 / 1e3)
    low_d1 = lowres_logits(model_d1._engine(dev), x_d1, h, w) if rank == 0 else None
    del model_d1, x_d1, met_d1

    # ---- BASELINE config 3: 256 images sharded over the ranks (strong scaling), fused confusion + one all-reduce ----
    shard = 256 // world
    xs3 = x[:shard] if shard <= B else torch.cat([x] * (-(-shard // B)))[:shard]
    ls3 = labels[:shard] if shard <= B else torch.cat([labels] * (-(-shard // B)))[:shard]
    xs3, ls3 = xs3.contiguous(), ls3.contiguous()
    met3 = SegmentationMetric(nc, device=dev)

    def strong_step():
        met3.reset()
        model.evaluate(xs3, ls3, met3)
        met3.all_reduce()
        return met3.get()

    barrier()
    strong_ms = max_over_ranks(time_cuda(strong_step, 3, warm=2))
    strong = {'workload': f'strong_256: 256 images of {h}x{w} sharded over {world} GPU(s), fused metric + all-reduce + read-back',
              'images_per_gpu': shard, 'value': shard * world / (strong_ms / 1e3), 'unit': UNIT, 'ms': strong_ms, 'n_gpus': world}
    del xs3, ls3, met3

    # ---- BASELINE config 4: TuSimple 2-class 480x640, batch 512 per GPU (weak scaling) ----
    tus = {}
    if not args.no_extra:
        m2 = FastSCNN(2, precision=args.precision).eval()
        init_recipe_d2(m2, 9)
        m2.to(dev)
        met2 = SegmentationMetric(2, device=dev)
        x2 = smooth_images(512, 480, 640, dev, 77 + rank, chunk=64)
        l2 = torch.randint(-1, 2, (512, 480, 640), device=dev, dtype=torch.int64)
        barrier()
        t_f32 = max_over_ranks(time_cuda(lambda: m2.evaluate(x2, l2, met2), 5, warm=2))
        xu2 = torch.randint(0, 256, (512, 480, 640, 3), dtype=torch.uint8, device=dev)
        lu2 = torch.randint(0, 3, (512, 480, 640), dtype=torch.uint8, device=dev)
        barrier()
        t_u8 = max_over_ranks(time_cuda(lambda: m2.evaluate(xu2, lu2, met2), 5, warm=2))
        tus = {'workload': 'tusimple_480x640_b512: 2 classes, 512 images per GPU per step, forward + argmax + metric',
               'value': 512 * world / (t_f32 / 1e3), 'unit': UNIT, 'n_gpus': world, 'scaling': 'weak',
               'value_uint8_inputs': 512 * world / (t_u8 / 1e3),
               'note': 'value: fp32 NCHW images + int64 labels resident in HBM; value_uint8_inputs: uint8 HWC images + uint8 labels'}
        del m2, met2, x2, l2, xu2, lu2
        torch.cuda.empty_cache()

    # ---- BASELINE config 5: one training step (forward + backward + OHEM loss + gradient all-reduce + SGD), crop 768, batch 16/GPU ----
    train = {}
    if not args.no_train:
        from fscnn_b200 import Trainer, train_ops
        torch.cuda.empty_cache()
        tb, crop = 16, 768
        xt = smooth_images(tb, crop, crop, dev, 500 + rank, chunk=16)
        tt = torch.randint(-1, nc, (tb, crop, crop), device=dev, dtype=torch.int64)

        def time_trainer(precision, graph, tsteps, classes=nc, tt=tt, **trainer_kw):
            mt = FastSCNN(classes, aux=True).train()
            init_recipe_d2(mt, 3)
            mt.to(dev)
            trainer = Trainer(mt, base_lr=1e-2, aux_weight=0.4, cuda_graph=graph, graph_warmup=2, matmul_precision=precision, **trainer_kw)
            l0 = float(trainer.step(xt, tt))
            for _ in range(3):             # the third of these captures the graph (graph mode)
                trainer.step(xt, tt)
            barrier()
            ev0t, ev1t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0t.record()
            for _ in range(tsteps):
                last = trainer.step(xt, tt)
            l1 = float(last)           # device -> host read of the loss, like the reference's loss.item() (train.py:283)
            ev1t.record()
            torch.cuda.synchronize()
            t_ms = max_over_ranks(ev0t.elapsed_time(ev1t)) / tsteps
            del trainer, mt
            torch.cuda.empty_cache()
            return t_ms, l0, l1

        try:
            eager_ms, _, _ = time_trainer('fp32', False, 3)
            tf32_ms, _, tf32_l1 = time_trainer('tf32', True, 5)
            t_ms, l0, l1 = time_trainer('fp32', True, 5)
            # train.py's other criteria / train_bdd100k.py's optimizer on the same step (TF32, CUDA graph): lane labels for the 2-class
            # dice settings (train.py's default --loss-type), plain cross entropy on the 19-class labels
            lane = (torch.rand((tb, crop, crop), device=dev) < 0.1).long()
            other = {'dice_nc2_sgd_ms': time_trainer('tf32', True, 4, classes=2, tt=lane, loss_type='dice')[0],
                     'dice_nc2_adamw_ms': time_trainer('tf32', True, 4, classes=2, tt=lane, loss_type='dice', optimizer='adamw')[0],
                     'focal_dice_nc2_sgd_ms': time_trainer('tf32', True, 4, classes=2, tt=lane, loss_type='focal_dice')[0],
                     'ce_nc19_sgd_ms': time_trainer('tf32', True, 4, loss_type='ce')[0]}
            del lane
        finally:
            train_ops.set_matmul_precision('fp32')
        train = {'workload': f'train_step_nc{nc}_aux_crop{crop}_b{tb}_per_gpu: forward + backward + MixSoftmaxCrossEntropyOHEMLoss + '
                             'gradient all-reduce + SGD(momentum, weight decay)',
                 'value': tb * world / (tf32_ms / 1e3), 'unit': UNIT, 'ms_per_step': tf32_ms, 'n_gpus': world, 'scaling': 'weak',
                 'dtype': 'tf32',
                 'dtype_note': 'fscnn_train_set_math(1): TF32 operands (fp32 accumulate) for the pointwise / dense 3x3 contractions -- tcgen05.mma.kind::tf32 '
                               'for forward and data gradient, mma.sync for the weight gradient; everything else fp32.  This is what cuDNN does '
                               'under torch allow_tf32 (the comparator\'s fp32 line) and at least the precision of the reference\'s fp16 autocast',
                 'loss_after_9_steps': tf32_l1,
                 'launch_mode': 'CUDA graph of zero_grad + forward + loss + backward, replayed; all-reduce + SGD outside the graph',
                 'fp32': {'value': tb * world / (t_ms / 1e3), 'ms_per_step': t_ms, 'loss_first_step': l0, 'loss_after_9_steps': l1,
                          'eager_launch_ms_per_step': eager_ms,
                          'note': 'default mode of the library (fp32 FMA contractions): the mode the 1e-4 parity tests are stated for'},
                 'other_criteria_ms_per_step': other,
                 'note': 'CUDA kernels of csrc/train.cu + train_tc.cu behind autograd wrappers; DDP semantics (per-rank BatchNorm, one NCCL '
                         'all-reduce of the flat 4.6 MB gradient buffer)'}
        del xt, tt
        torch.cuda.empty_cache()

    # ---- the node's host->device ceiling: plain pinned copies of the e2e buffers, all ranks at once ----
    from fscnn_b200 import StreamingEvaluator
    e2e_steps = max(3, args.steps)
    nbuf = 3
    img_h = [torch.randint(0, 256, (B, h, w, 3), dtype=torch.uint8).pin_memory() for _ in range(nbuf)]
    lab_h = [torch.randint(0, nc + 1, (B, h, w), dtype=torch.uint8).pin_memory() for _ in range(nbuf)]   # nc = "ignore"-like overflow label
    img_d, lab_d = torch.empty_like(img_h[0], device=dev), torch.empty_like(lab_h[0], device=dev)

    def h2d_only():
        for i in range(nbuf):
            img_d.copy_(img_h[i], non_blocking=True)
            lab_d.copy_(lab_h[i], non_blocking=True)

    barrier()
    h2d_ms = max_over_ranks(time_cuda(h2d_only, 3, warm=1))
    h2d_bytes = img_h[0].numel() + lab_h[0].numel()
    h2d_ceiling_gbs = world * nbuf * h2d_bytes / (h2d_ms / 1e3) / 1e9
    del img_d, lab_d

    # ---- end to end through the public API with host buffers ----
    # (a) the streaming evaluator (fscnn_b200.StreamingEvaluator): raw uint8 HWC images + uint8 labels in pinned
    #     host memory, H2D on a copy stream overlapped with compute, metric state read back every step;
    # (b) the reference's own tensor layout (normalised fp32 NCHW + int64 labels), copied and evaluated step by step.
    metric.reset()
    ev = StreamingEvaluator(model, metric, img_h[0], lab_h[0], device=dev)
    for i in range(3):
        ev.submit(img_h[i % nbuf], lab_h[i % nbuf])
    ev.result()
    barrier()
    sampler.begin()
    ev0.record()
    for i in range(e2e_steps):
        ev.submit(img_h[i % nbuf], lab_h[i % nbuf])
    e2e_metric = ev.result()
    ev1.record()
    torch.cuda.synchronize()
    sampler.end()
    clocks = sampler.stop() if rank == 0 else None
    e2e_ms = max_over_ranks(ev0.elapsed_time(ev1))
    e2e_value = world * B * e2e_steps / (e2e_ms / 1e3)
    h2d = h2d_bytes
    d2h = metric.conf_len() * 8
    del ev, img_h, lab_h

    Br = min(B, 32)     # secondary number, PCIe-bound at 42 MB per image: keep the pinned host buffers at 1.3 GB per rank
    xh = torch.empty((Br, 3, h, w), dtype=torch.float32).pin_memory()
    xh.copy_(x[:Br])
    lh = torch.empty((Br, h, w), dtype=torch.int64).pin_memory()
    lh.copy_(labels[:Br])
    xd, ld = torch.empty_like(x[:Br]), torch.empty_like(labels[:Br])
    metric.reset()

    def step_ref_layout():
        xd.copy_(xh, non_blocking=True)
        ld.copy_(lh, non_blocking=True)
        model.evaluate(xd, ld, metric)
        return metric.get()            # device -> host read of the metric state (synchronises)

    ref_steps = max(3, min(args.steps, 6))
    barrier()
    ref_ms = max_over_ranks(time_cuda(step_ref_layout, ref_steps, warm=2))
    e2e_ref_layout = {'value': world * Br / (ref_ms / 1e3), 'unit': UNIT, 'batch_per_gpu': Br,
                      'h2d_bytes_per_step': int(xh.numel() * 4 + lh.numel() * 8), 'd2h_bytes_per_step': int(d2h),
                      'steps': ref_steps, 'host_buffers': 'pinned fp32 NCHW images + int64 labels, no copy/compute overlap'}
    del xh, lh, xd, ld

    out = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': max(3, args.warmup), 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32' if args.precision == 'fp32' else 'bf16', 'data': 'synthetic',
        'config': {'workload': f'cityscapes_eval_nc{nc}_{h}x{w}_fwd_argmax_metric', 'batch_per_gpu': B,
                   'global_batch': B * world, 'precision': args.precision, 'l2_policy': 'inputs_exceed_l2',
                   'input_bytes_per_step_per_gpu': int(x.numel() * 4 + labels.numel() * 8),
                   'parallelism': f'dp{world}', 'cpu_affinity': numa,
                   'weights': 'recipe D2: variance-preserving random weights, randomised BN statistics, calibrated classifier bias',
                   'inputs': 'multi-scale smooth noise images (fp32 NCHW), int64 labels uniform in [-1, nc)',
                   'classes_present': classes_present, 'largest_class_fraction': largest_class,
                   'timed_region': 'K x (forward + fused upsample/argmax/metric) + NCCL all-reduce of the confusion state + '
                                   'device->host read + pixAcc/mIoU'},
        'value_recipe_d1': value_d1,
        'pixAcc_mIoU': [float(pix_acc), float(miou)],
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                'steps': e2e_steps, 'ms_per_step': e2e_ms / e2e_steps, 'api': 'fscnn_b200.StreamingEvaluator.submit',
                'h2d_gbs': e2e_value * h2d / B / 1e9, 'h2d_ceiling_gbs': h2d_ceiling_gbs,
                'frac_of_h2d_ceiling': (e2e_value * h2d / B / 1e9) / h2d_ceiling_gbs,
                'host_buffers': 'pinned uint8 HWC images (ToTensor+Normalize fused into the stem) + uint8 labels; '
                                'H2D on a copy stream overlaps compute; metric state copied to the host every step; '
                                'h2d_ceiling_gbs = the same buffers copied with plain cudaMemcpyAsync by all ranks at once, no compute'},
        'e2e_reference_layout': e2e_ref_layout,
        'strong_256': strong,
        'gpu_launches': int(launches),
        'clocks': clocks,
    }
    if tus:
        out['tusimple_480x640_b512'] = tus
    if train:
        out['training_step'] = train
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = measured_peaks()
    stages = stage_model(h, w, nc, es)
    plan_bytes = sum(s['bytes'] for s in stages)
    plan_flops = sum(s['flops'] for s in stages)
    per_img_s = (ms / 1e3) / (B * args.steps)
    out['pipeline'] = {'plan_bytes_per_image': plan_bytes, 'flops_per_image': plan_flops,
                       'achieved_gbs': plan_bytes / per_img_s / 1e9, 'hbm_frac': plan_bytes / per_img_s / 1e9 / peaks['hbm_gbs'],
                       'achieved_tflops': plan_flops / per_img_s / 1e12}

    # ---- per-kernel timing with CUDA events on the launch stream (one micro-batch per launch) ----
    if not args.no_stage_times:
        names = eng.stage_names()
        import ctypes as C
        from fscnn_b200 import native
        tap = native.Tap()
        native.check(eng.lib.fscnn_tap_info(eng._ctx, B, h, w, b'l2d.conv', C.byref(tap)))
        mb = tap.n                                   # images per launch (the library's micro-batch)
        xs = x[:mb].contiguous()
        eng.forward_range(xs, 0, len(names) - 1)     # fill every stage tensor once
        reps = 5
        table = []
        fused_front = args.precision == 'bf16'      # bf16: stem + dsconv1 run as ONE kernel
        for i, name in enumerate(names):
            if fused_front and name == 'l2d.dsconv1':
                continue
            j = i + 1 if (fused_front and name == 'stem') else i
            t, _ = time_cuda_trials(lambda: eng.forward_range(xs, i, j), reps, trials=3, warm=1)
            table.append(('stem+l2d.dsconv1' if j != i else name, t / mb * 1e3))   # microseconds per image
        # the tail kernel (x8 upsample + argmax + metric counting) with its OWN event pair, on the logits of this batch; its time
        # depends on the logits (exact class pruning) and on the labels (run-length aggregated counting), so both extremes are timed
        lab_mb = labels[:mb].contiguous()
        conf = torch.zeros(eng.conf_len(), dtype=torch.int64, device=dev)
        low_d2 = lowres_logits(eng, xs, h, w)
        hl_, wl_ = low_d2.shape[1], low_d2.shape[2]

        tail_trials = {}

        def tail_us(low, lab, key=None, **kw):
            n_ = low.shape[0]
            lab_n = lab[:n_]
            med, ts = time_cuda_trials(lambda: eng.upsample_argmax(low, h, w, labels=lab_n, conf=conf, want_mask=False, **kw), reps,
                                       trials=5 if key else 3, warm=2)
            if key:
                tail_trials[key] = [t / n_ * 1e3 for t in ts]
            return med / n_ * 1e3

        tail = tail_us(low_d2, lab_mb, key='us_per_image_trials')
        checker = ((torch.arange(hl_, device=dev)[:, None] + torch.arange(wl_, device=dev)[None, :]) % 2 * 2 - 1).float()
        low_adv = torch.zeros_like(low_d2)
        low_adv[..., :nc] = checker[None, :, :, None] * torch.arange(nc, device=dev).float()
        dom = torch.zeros_like(lab_mb)
        dom[:, ::16, ::16] = -1
        tail_extra = {
            'us_per_image_recipe_d1_logits': tail_us(low_d1, lab_mb),
            'us_per_image_no_class_can_be_pruned': tail_us(low_adv, lab_mb),       # class order flips between neighbouring taps
            'us_per_image_exhaustive_flag': tail_us(low_d2, lab_mb, exhaustive=True),
            'us_per_image_dominant_class_labels': tail_us(low_d2, dom),
            'us_per_image_mask_only_uint8': time_cuda(lambda: eng.upsample_argmax(low_d2, h, w), reps, warm=1) / mb * 1e3,
        }
        table.append(('up8+argmax+metric', tail))
        del low_adv, dom
        full_us = time_cuda_trials(lambda: model.evaluate(xs, lab_mb, metric), reps, trials=3, warm=1)[0] / mb * 1e3
        model_by_name = {s['stage']: s for s in stages}
        if fused_front:   # plan-P bytes of both stages (the denominator is not changed); the fused kernel's own traffic beside it
            a_, b_ = model_by_name['stem'], model_by_name['l2d.dsconv1']
            model_by_name['stem+l2d.dsconv1'] = {'stage': 'stem+l2d.dsconv1', 'bytes': a_['bytes'] + b_['bytes'], 'flops': a_['flops'] + b_['flops'],
                                                 'fused_bytes': 3.0 * h * w * 4 + ((((h - 3) // 2 + 1) - 1) // 2 + 1) * ((((w - 3) // 2 + 1) - 1) // 2 + 1) * 48.0 * es}
        rows = []
        for name, us in table:
            m = model_by_name.get(name)
            if m is None:
                continue
            rows.append({'stage': name, 'us_per_image': us, 'gbs': m['bytes'] / us / 1e3, 'tflops': m['flops'] / us / 1e6,
                         'bytes': m['bytes'], 'flops': m['flops']})
            if name == 'up8+argmax+metric':
                rows[-1].update(tail_extra)
                rows[-1].update(tail_trials)
                # plan-P counts the low-res logits + a uint8 mask; this run is the fused-metric mode, which reads the
                # caller's labels instead of writing a mask (int64 in the reference tensor layout): the bytes this launch must move
                run_bytes = float(m['bytes'] - h * w * 1 + h * w * labels.element_size())
                rows[-1]['bytes_fused_metric_mode'] = run_bytes
                rows[-1]['gbs_fused_metric_mode'] = run_bytes / us / 1e3
            if 'fused_bytes' in m:
                rows[-1]['fused_bytes'] = m['fused_bytes']
                rows[-1]['fused_gbs'] = m['fused_bytes'] / us / 1e3
        out['stages'] = rows
        out['stage_sum_us_per_image'] = sum(r['us_per_image'] for r in rows)
        out['full_step_us_per_image_one_launch_set'] = full_us
        top = dict(max(rows, key=lambda r: r['us_per_image']))
        note = ''
        if 'bytes_fused_metric_mode' in top:   # the tail kernel: algorithmic bytes of the mode that was run (labels read, no mask written)
            top['bytes'], top['gbs'] = top['bytes_fused_metric_mode'], top['gbs_fused_metric_mode']
            note = ('algorithmic bytes = low-res logits + the int64 labels the fused-metric mode reads (plan-P counts a uint8 mask '
                    'instead: see stages[].bytes); timed with its own CUDA-event pair on this batch\'s logits')
        elif 'fused_bytes' in top:
            # one kernel for two plan-P stages: ITS algorithmic bytes are the image read + the dsconv1 tensor written (the stem
            # output never reaches HBM); the plan-P figure of the two unfused stages stays in stages[].bytes
            note = (f"algorithmic bytes of the fused kernel = input image + dsconv1 output = {top['fused_bytes'] / 1e6:.1f} MB per image; "
                    f"plan-P counts the two unfused stages ({top['bytes'] / 1e6:.1f} MB, stages[].bytes / .gbs), which this kernel beats by not "
                    'writing the stem output')
            top['bytes'], top['gbs'] = top['fused_bytes'], top['fused_gbs']
        hbm_time = top['bytes'] / (peaks['hbm_gbs'] * 1e9)
        tens_time = top['flops'] / (peaks['bf16_tflops'] * 1e12)
        if args.precision == 'bf16' and tens_time > hbm_time:
            roof = {'bound': 'tensor', 'achieved': top['tflops'], 'peak': peaks['bf16_tflops'], 'unit': 'TFLOP/s',
                    'frac': top['tflops'] / peaks['bf16_tflops']}
        else:
            roof = {'bound': 'hbm', 'achieved': top['gbs'], 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                    'frac': top['gbs'] / peaks['hbm_gbs']}
        traffic = None   # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed ncu --set full capture
        tpath = os.path.join(ROOT, 'profiles', 'kernel_traffic.json')
        if os.path.exists(tpath):
            rec = json.load(open(tpath)).get(args.precision, {}).get(top['stage'])
            if rec:
                traffic = rec['dram_bytes_per_image'] * mb
        roof.update({'kernel': top['stage'], 'traffic': traffic, 'traffic_source': 'profiles/kernel_traffic.json (ncu --set full capture, scaled to this launch)',
                     'algorithmic_bytes_per_launch': top['bytes'] * mb,
                     'peak_source': peaks['source'], 'images_per_launch': mb,
                     'us_per_launch': top['us_per_image'] * mb,
                     'note': 'fp32 path: this kernel is FP32-FMA bound on CUDA cores; fma_frac = achieved fp32 TFLOP/s / '
                             '(148 SM x 128 lanes x 2 x sm_mhz)' if args.precision == 'fp32' else note})
        if args.precision == 'fp32':
            mhz = (clocks or {}).get('sm_mhz') or peaks['sm_max_mhz']
            roof['fma_frac'] = top['tflops'] / (148 * 128 * 2 * mhz * 1e6 / 1e12)
        out['roofline'] = roof

    if not args.no_extra:
        # same workload with the camera / dataset byte layout resident in HBM (uint8 HWC images, ToTensor+Normalize fused into
        # the first kernel, uint8 labels): 8.4 MB per image instead of the reference tensor layout's 42 MB
        xu = torch.randint(0, 256, (B, h, w, 3), dtype=torch.uint8, device=dev)
        lu = torch.randint(0, nc + 1, (B, h, w), dtype=torch.uint8, device=dev)
        metu = SegmentationMetric(nc, device=dev)
        t = time_cuda(lambda: model.evaluate(xu, lu, metu), 10, warm=3)
        out['device_resident_uint8_inputs'] = {'value': B / (t / 1e3), 'unit': UNIT, 'steps': 10, 'n_gpus': 1,
                                               'note': 'uint8 HWC images + uint8 labels resident in HBM, rank 0'}
        del xu, lu, metu
        model.evaluate(x, labels, metric)      # back to the fp32 NCHW input format for what follows

    if args.precision != 'fp32' and not args.no_fp32:
        # the fp32 exactness path (parity 1e-4 vs the reference) on the same inputs, short run
        m32 = FastSCNN(nc, precision='fp32').eval()
        m32.load_state_dict(model.state_dict())
        m32.to(dev)
        met32 = SegmentationMetric(nc, device=dev)
        t = time_cuda(lambda: m32.evaluate(x, labels, met32), 5, warm=2)
        out['fp32_exact'] = {'value': B / (t / 1e3), 'unit': UNIT, 'steps': 5, 'n_gpus': 1,
                             'note': 'fp32 storage + fp32 FMA path on rank 0 (logits within 1e-4 of the reference)'}
        del m32, met32

    if args.latency:
        xs1 = x[:1].contiguous()
        mask1 = torch.empty((1, h, w), dtype=torch.uint8, device=dev)
        for _ in range(3):
            model.predict(xs1, out=mask1)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            model.predict(xs1, out=mask1)
        out['latency_batch1_ms'] = time_cuda(graph.replay, 1000, warm=200)        # 200 warm + 1000 timed replays
        # what the reference's callers run (eval.py:43-45, demo.py:46-48): eager model(image) then torch.argmax, batch 1, no graph;
        # and the fused eager call.  Host-side cost per call (parameter fingerprint, ctypes, tensor-map encodes) is inside.
        out['latency_batch1_eager_ms'] = {
            'forward_then_torch_argmax': time_cuda(lambda: torch.argmax(model(xs1)[0], 1), 100, warm=20),
            'predict': time_cuda(lambda: model.predict(xs1, out=mask1), 200, warm=20),
            'note': 'eager (no CUDA graph), synchronised at both ends; forward_then_torch_argmax materialises the 159 MB full-resolution logits like the reference',
        }

    if not args.no_extra:
        # BASELINE.json's other inference shape (2-class drivable-area 360x640), device-resident, rank 0
        extra = []
        for enc, eh, ew, eb in ((2, 360, 640, 128),):
            em = FastSCNN(enc, precision=args.precision).eval()
            init_recipe_d2(em, 11)
            em.to(dev)
            ex = smooth_images(eb, eh, ew, dev, 5, chunk=64)
            el = torch.randint(-1, enc, (eb, eh, ew), device=dev, dtype=torch.int64)
            emet = SegmentationMetric(enc, device=dev)
            t = time_cuda(lambda: em.evaluate(ex, el, emet), 10, warm=3)
            extra.append({'workload': f'eval_nc{enc}_{eh}x{ew}_fwd_argmax_metric', 'batch': eb, 'value': eb / (t / 1e3),
                          'unit': UNIT, 'n_gpus': 1, 'precision': args.precision})
            del em, ex, el, emet
        # the deployed camera pipeline (reference export_onnx_fixed.EndToEndFastSCNN): 640x360 uint8 frames -> 1024x1024
        # -> 2-class network -> probabilities at 640x360, frames resident in HBM
        from models.end_to_end import EndToEndFastSCNN
        cam = EndToEndFastSCNN(FastSCNN(2, precision=args.precision).eval().to(dev), input_size=(640, 360), base_size=1024).eval()
        frames = torch.randint(0, 256, (64, 3, 360, 640), dtype=torch.uint8, device=dev)
        t = time_cuda(lambda: cam(frames), 5, warm=2)
        extra.append({'workload': 'camera_frames_640x360_via_1024x1024_softmax_nc2', 'batch': 64, 'value': 64 / (t / 1e3),
                      'unit': 'frames/s', 'n_gpus': 1, 'precision': args.precision})
        del cam, frames
        out['other_workloads'] = extra

    if world == 1 and not args.no_eager:
        # the honest GPU comparator: the reference's op sequence run eagerly through cuDNN / ATen on this same GPU
        torch.cuda.empty_cache()
        out['gpu_eager_baseline'] = gpu_eager_baseline(dev, h, w, nc)
        out['gpu_eager_baseline']['note'] = ('reference op sequence (ATen-functional port, unfolded BN, NCHW fp32 tensors) + torch.argmax, '
                                             'cuDNN eager on this GPU, cudnn.benchmark on, synchronised; none of this repo\'s kernels')
        best = max(v['images_per_s'] for k, v in out['gpu_eager_baseline'].items() if isinstance(v, dict))
        out['gpu_eager_baseline']['speedup_device_resident_vs_best_eager'] = value / best

    if world == 1 and not args.no_cpu_baseline:
        rate, done, el, cores, threads = cpu_reference_rate(h, w, nc, 12.0, 1)
        out['cpu_baseline'] = {'value': rate, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'host_cpus': cores,
                               'sample': f'{done} single-image {h}x{w} steps in {el:.1f} s (forward+argmax+host metric), '
                                         'ATen-functional port of the reference on all host threads'}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.only == 'eager':
        import torch
        print(json.dumps({'gpu_eager_baseline': gpu_eager_baseline(torch.device('cuda', 0), args.height, args.width, args.classes)}))
        return
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_native_arm(args)


if __name__ == '__main__':
    main()
