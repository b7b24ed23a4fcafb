"""ExponentialWarmup ramp (host-side scalar per step).  Mirror of desed_task/utils/schedulers.py:8-104: same
constructor, `step()`, `state_dict()/load_state_dict()`, `_get_scaling_factor()` (also the consistency-loss ramp at
recipes/dcase2023_task4_baseline/local/sed_trainer.py:329-332) and `_get_lr()`."""
import math


class BaseScheduler(object):
    def __init__(self, optimizer):
        self.optimizer = optimizer
        self.step_num = 0

    def zero_grad(self):
        self.optimizer.zero_grad()

    def _get_lr(self):
        raise NotImplementedError

    def _set_lr(self, lr):
        for param_group in self.optimizer.param_groups:
            param_group["lr"] = lr

    def step(self, metrics=None, epoch=None):
        self.step_num += 1
        lr = self._get_lr()
        self._set_lr(lr)

    def load_state_dict(self, state_dict):
        self.__dict__.update(state_dict)

    def state_dict(self):
        return {key: value for key, value in self.__dict__.items() if key != "optimizer"}


class ExponentialWarmup(BaseScheduler):
    def __init__(self, optimizer, max_lr, rampup_length, exponent=-5.0, start_annealing=None, max_steps=None,
                 min_lr=1e-8):
        super().__init__(optimizer)
        self.rampup_len = rampup_length
        self.max_lr = max_lr
        self.step_num = 1
        self.exponent = exponent
        self.start_annealing = start_annealing
        self.max_steps = max_steps
        self.min_lr = min_lr

    def _ramp(self):
        current = min(max(float(self.step_num), 0.0), float(self.rampup_len))
        phase = 1.0 - current / self.rampup_len
        return float(math.exp(self.exponent * phase * phase))

    def _get_scaling_factor(self):
        if self.rampup_len == 0:
            return 1.0
        if self.start_annealing is None or self.step_num < self.start_annealing:
            return self._ramp()
        one_steps = self.step_num - self.start_annealing
        zero_steps = self.max_steps - self.start_annealing
        return max(self.min_lr / self.max_lr, math.cos(one_steps * math.pi / (2 * zero_steps)))

    def _get_lr(self):
        return self.max_lr * self._get_scaling_factor()
