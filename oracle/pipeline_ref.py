"""Oracle: cosine VP schedule, predictor-corrector sampler, DSM loss (TEST INFRASTRUCTURE ONLY).

Restates src/thor/pipelines.py: alpha/mu/sigma (:13-20), forward/loss (:22-35), _sample_step (:41-46),
sample (:48-97).  fp32 on the CPU, like the reference's default (`device=None` -> CPU, :59-60).
"""
from __future__ import annotations

import math
from typing import Callable, Optional

import torch

Tensor = torch.Tensor


class RefPipeline:
    def __init__(self, eta: float = 1e-3):
        self.eta = eta

    def alpha(self, t: Tensor) -> Tensor:
        return torch.cos(math.acos(math.sqrt(self.eta)) * t) ** 2

    def mu(self, t: Tensor) -> Tensor:
        return self.alpha(t)

    def sigma(self, t: Tensor) -> Tensor:
        return (1 - self.alpha(t) ** 2 + self.eta ** 2).sqrt()

    def pred_eps(self, score_fn: Callable, x: Tensor, t: Tensor) -> Tensor:
        return score_fn(x, t)

    def loss(self, net: Callable, x: Tensor, t: Optional[Tensor] = None, eps: Optional[Tensor] = None) -> Tensor:
        """:27-35 — unreduced squared error.  t / eps may be injected for parity tests; drawn like the
        reference (torch.rand, then torch.randn_like) otherwise."""
        if t is None:
            t = torch.rand(x.shape[0], 1, 1, 1, dtype=x.dtype)
        if eps is None:
            eps = torch.randn_like(x)
        xt = self.mu(t) * x + self.sigma(t) * eps
        return (net(xt, t) - eps) ** 2

    def predictor(self, score_fn: Callable, x: Tensor, t: Tensor, dt: float) -> Tensor:
        """:41-46."""
        eps = score_fn(x, t)
        x0 = (x - self.sigma(t) * eps) / self.mu(t)
        return self.mu(t - dt) * x0 + self.sigma(t - dt) * eps

    def sample(self, score_fn: Callable, noise: Tensor, steps: int = 64, corrections: int = 0, tau: float = 1.0,
               z_draw: Optional[Callable[[Tensor], Tensor]] = None, trace: Optional[list] = None) -> Tensor:
        """:48-97.  `z_draw(z)` fills the corrector noise buffer; default `z.normal_()` from the global CPU
        generator exactly as the reference (:82).  `trace`, if given, collects x after every step."""
        x = noise.clone()
        dims = tuple(range(-x.dim(), 0))
        time_steps = torch.linspace(1, 0, steps + 1).to(dtype=x.dtype)
        dt = 1 / steps
        z = torch.empty_like(x) if corrections > 0 else None
        with torch.no_grad():
            for t in time_steps[:-1]:
                x = self.predictor(score_fn, x, t, dt)
                for _ in range(corrections):
                    if z_draw is None:
                        z.normal_()
                    else:
                        z = z_draw(z)
                    eps = score_fn(x, t - dt)
                    delta = tau / eps.square().mean(dim=dims, keepdim=True)
                    x = x - (delta * eps + torch.sqrt(2 * delta) * z) * self.sigma(t - dt)
                if torch.isnan(x).any():
                    raise ValueError("NaN detected in sample")
                if trace is not None:
                    trace.append(x.clone())
        return x.reshape(noise.shape)
