"""Golden learning-rate sequences of the reference's utils/lr_scheduler.LRScheduler (utils/lr_scheduler.py:6-91), produced by the
UNMODIFIED reference class in the build container (TEST INFRASTRUCTURE; /root/reference does not travel to the GPU box):

    python oracle/gen_golden_lr.py        ->  tests/golden/lr_scheduler_cases.json

Every mode ('poly' as train.py:205-207 builds it, 'cosine', 'linear', 'constant', 'step' by iteration and by epoch), offsets and
iterations past the end of the schedule; doubles are stored exactly (json round-trips repr)."""
import importlib.util
import json
import os

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'lr_scheduler_cases.json')
spec = importlib.util.spec_from_file_location('ref_lr_scheduler', '/root/reference/utils/lr_scheduler.py')
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

CASES = [dict(mode='poly', base_lr=0.01, nepochs=60, iters_per_epoch=176, power=0.9),
         dict(mode='poly', base_lr=0.045, target_lr=1e-4, niters=500, power=2, offset=20),
         dict(mode='cosine', base_lr=0.1, target_lr=0.001, niters=300),
         dict(mode='linear', base_lr=0.1, niters=100, offset=5),
         dict(mode='constant', base_lr=0.02, target_lr=0.5, niters=50),
         dict(mode='step', base_lr=0.1, niters=200, step_iter=[50, 120, 180], step_factor=0.5),
         dict(mode='step', base_lr=0.1, nepochs=10, iters_per_epoch=20, step_epoch=[3, 6]),
         dict(mode='step', base_lr=0.1, nepochs=10, iters_per_epoch=20, step_iter=[30])]

out = []
for kw in CASES:
    sched = ref.LRScheduler(**kw)
    n = sched.niters + 30
    iters = list(range(0, n, max(1, n // 40))) + [n]
    out.append({'kwargs': kw, 'iters': iters, 'lr': [sched(i) for i in iters]})
json.dump(out, open(OUT, 'w'))
print(len(out), 'cases,', sum(len(c['iters']) for c in out), 'values')
