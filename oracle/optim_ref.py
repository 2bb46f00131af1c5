"""TEST INFRASTRUCTURE ONLY — CPU restatement (torch float64 on the host) of the optimiser half of the reference's
training step: `optimizer.step()` with torch.optim.AdamW as configured in train.py:176-181 (lr, weight_decay=1e-3,
betas=[0.9, 0.999], default eps=1e-8, no amsgrad) followed by `StandardEMA.update()` (training_loop.py:381-390,
src/thor/ema.py:24-27).  Only tests/ may import this module.

Pinned (tests/test_optim.py) against torch.optim.AdamW itself and against the reference's own thor/ema.py where the
reference checkout is present.
"""
import torch


class AdamWEMARef:
    def __init__(self, params, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, ema_rate=None):
        self.p = [p.detach().clone().double() for p in params]
        self.m = [torch.zeros_like(p) for p in self.p]
        self.v = [torch.zeros_like(p) for p in self.p]
        self.ema = [p.clone() for p in self.p] if ema_rate is not None else None  # copy.deepcopy(net), ema.py:14
        self.lr, self.betas, self.eps, self.wd, self.rate = lr, betas, eps, weight_decay, ema_rate
        self.t = 0

    def step(self, grads, grad_scale=1.0):
        self.t += 1
        b1, b2 = self.betas
        bias1 = 1.0 - b1 ** self.t
        bias2_sqrt = (1.0 - b2 ** self.t) ** 0.5
        for i, g in enumerate(grads):
            g = g.double() * grad_scale
            self.p[i] = self.p[i] * (1.0 - self.lr * self.wd)          # decoupled weight decay
            self.m[i] = self.m[i] + (1.0 - b1) * (g - self.m[i])       # lerp
            self.v[i] = b2 * self.v[i] + (1.0 - b2) * g * g
            denom = self.v[i].sqrt() / bias2_sqrt + self.eps
            self.p[i] = self.p[i] - (self.lr / bias1) * self.m[i] / denom
            if self.ema is not None:                                   # ema.py:24-27, after the parameter update
                self.ema[i] = self.ema[i] * self.rate + self.p[i] * (1.0 - self.rate)
