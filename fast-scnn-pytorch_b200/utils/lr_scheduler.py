"""Drop-in for the reference ``utils/lr_scheduler.LRScheduler`` (utils/lr_scheduler.py:6-91; built by train.py:205-207 and
train_bdd100k.py:193 with mode='poly', power=0.9): a host-side function of the iteration number.  ``fscnn_b200.Trainer`` applies the
poly mode itself (``poly_lr``); this class is for callers that keep the reference's loop and set the optimizer's ``lr`` themselves."""
import math

# decay factor in [0, 1] at iteration t of (last + 1); lr = target + (base - target) * factor.  The operation order is the
# reference's (:70-74), so the values are bit-identical doubles
_FACTORS = {
    'linear': lambda t, last, power: 1 - t / last,
    'poly': lambda t, last, power: pow(1 - t / last, power),
    'cosine': lambda t, last, power: (1 + math.cos(math.pi * t / last)) / 2,
}


class LRScheduler(object):
    """Same constructor arguments, attributes (``learning_rate`` after ``update``) and arithmetic as the reference class."""

    def __init__(self, mode, base_lr=0.01, target_lr=0, niters=0, nepochs=0, iters_per_epoch=0, offset=0, power=2, step_iter=None,
                 step_epoch=None, step_factor=0.1):
        assert mode in ('constant', 'step', 'linear', 'poly', 'cosine')
        if mode == 'step':
            assert step_iter is not None or step_epoch is not None
        self.mode, self.base_lr, self.offset, self.power, self.step_factor = mode, base_lr, offset, power, step_factor
        self.target_lr = base_lr if mode == 'constant' else target_lr
        self.niters, self.step = niters, step_iter
        if nepochs * iters_per_epoch > 0:          # an epoch count overrides niters, and step_epoch overrides step_iter (:55-59)
            self.niters = nepochs * iters_per_epoch
            if step_epoch is not None:
                self.step = [e * iters_per_epoch for e in step_epoch]

    def __call__(self, num_update):
        self.update(num_update)
        return self.learning_rate

    def update(self, num_update):
        last = self.niters - 1
        t = min(max(0, num_update - self.offset), last)
        if self.mode == 'step':
            passed = 0 if self.step is None else sum(1 for s in self.step if s <= t)
            self.learning_rate = self.base_lr * pow(self.step_factor, passed)
            return
        factor = 0 if self.mode == 'constant' else _FACTORS[self.mode](t, last, self.power)
        self.learning_rate = self.target_lr + (self.base_lr - self.target_lr) * factor
