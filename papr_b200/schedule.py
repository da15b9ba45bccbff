"""Learning-rate schedules of reference models/utils.py:260-322 (warm-up + decay SequentialLR), in closed form.

The reference builds SequentialLR([LinearLR, <decay>]) and, after every prune/add, re-creates the optimisers and
fast-forwards each scheduler by calling .step() `step` times in Python (models/model.py:175-179) -- O(step) host work.
The same learning rates are produced here by a LambdaLR whose factor is the closed form of that composition, so a
fast-forward is O(1).  tests/test_host_logic.py checks the values against torch's own SequentialLR.
"""
import math

import torch.optim.lr_scheduler as lr_scheduler


class WarmupDecay:
    def __init__(self, kind, warmup, max_steps, gamma=1.0):
        self.kind, self.warmup, self.max_steps, self.gamma = kind, int(warmup), int(max_steps), gamma

    def _cos(self, u):
        T = max(self.max_steps - self.warmup, 1) * (2 if self.kind == "cosine-hlfperiod" else 1)
        return 0.5 * (1.0 + math.cos(math.pi * u / T))

    def __call__(self, t):
        w = self.warmup
        if t < w:   # LinearLR(start_factor=1e-16, end_factor=1, total_iters=warmup)
            s = 1e-16
            return s + (1.0 - s) * t / w
        u = t - w
        if w == 0:
            # With milestone 0, SequentialLR never calls the decay scheduler's step(0): its __init__ leaves that
            # scheduler at last_epoch = -1 and every step() then applies the *recursive* update.  For the cosine law
            # this gives lr_t = base * c(t-1)/c(1) (slightly above base at t = 1); the other laws lag one step.
            # The shipped configs hit this with lr.points (cosine, warmup 0); reproduced as is.
            if t == 0:
                return 1.0
            if self.kind in ("cosine", "cosine-hlfperiod"):
                return self._cos(t - 1) / self._cos(1)
            u = t - 1
        if self.kind == "linear":
            total = self.max_steps - w
            return 1.0 - min(u, total) / total
        if self.kind in ("cosine", "cosine-hlfperiod"):
            return self._cos(u)
        if self.kind == "exp":
            return self.gamma ** u
        if self.kind == "stop":
            return 1.0 if u == 0 else 0.0
        raise NotImplementedError(self.kind)


def reposition(sched, step):
    """Jump a closed-form scheduler to `step` in O(1) (the reference calls .step() `step` times)."""
    if step <= 0:
        return
    sched.last_epoch = step - 1
    sched._step_count = step
    sched.optimizer._opt_called = True     # silence the "scheduler before optimizer" warning on the jump
    sched.step()


def create_learning_rate_fn(optimizer, max_steps, args, start_step=0):
    """Scheduler positioned at `start_step` (reference: create + step() x start_step)."""
    if args.type == "none":
        return None
    fn = WarmupDecay(args.type, args.warmup, max_steps, args.get("gamma", 1.0))
    sched = lr_scheduler.LambdaLR(optimizer, lr_lambda=fn)
    reposition(sched, start_step)
    return sched
