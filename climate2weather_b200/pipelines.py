"""Host-side mirror of `thor.pipelines.SDAPipeline` (reference src/thor/pipelines.py:8-97): cosine VP schedule,
DSM loss, predictor-corrector sampler — same method names and signatures.

`sample()` with one of this package's score functions runs device-resident: the trajectory never leaves HBM
between steps, each step is  window score (K0/K1/K2/K4/K5)  ->  fused guidance + predictor (K6)  [-> corrector
(K7)]  [-> halo exchange when time-sharded].  The reference keeps x on the CPU and round-trips every window batch
(src/thor/score.py:170-181).
"""
from __future__ import annotations

import math
import time
from typing import Optional

import torch
from torch import Tensor

from . import _lib
from .score import AbstractScoreFunction, _mu_sigma
from .sharding import all_gather_frames, gather_frames


class SDAPipeline:
    #: corrector noise: "device" = on-chip Philox keyed by the global pixel index (default, sharding-invariant);
    #: "reference" = z.normal_() on a HOST tensor from the global CPU generator — what src/thor/pipelines.py:82 does when
    #: sample() runs on its default device (CPU, :59-60; the shipped driver calls it that way) — then uploaded (parity runs).
    rng: str = "device"
    #: read the device NaN flag every this many steps (0 = only at the end); the reference syncs every step (:90).
    #: The per-step read is asynchronous and inspected one check later, so a NaN raises one step after it appeared.
    nan_check_every: int = 0
    #: optional debugging / parity hook: sampler steps (1-based) after which the trajectory is copied out; the copies
    #: land in `self.traces[step]` (NCHW, host).  None = off (no copies, no synchronisation).
    trace_at = None

    def __init__(self, eta=1e-3):
        self.eta = eta  # src/thor/pipelines.py:9-11
        self.traces = {}

    # ---------------------------------------------------------------- schedule (src/thor/pipelines.py:13-20)
    def alpha(self, t):
        return torch.cos(math.acos(math.sqrt(self.eta)) * t) ** 2

    def mu(self, t):
        return self.alpha(t)

    def sigma(self, t):
        return (1 - self.alpha(t) ** 2 + self.eta ** 2).sqrt()

    # ---------------------------------------------------------------- training objective (:22-35)
    def forward(self, x, t):
        eps = torch.randn_like(x)
        # mu x + sigma eps in two passes over the batch instead of three (mul, then fused multiply-add)
        return torch.addcmul(self.mu(t) * x, self.sigma(t).expand_as(x), eps), eps

    def loss(self, net, x, forcing=None):
        t = torch.rand(x.shape[0], 1, 1, 1, dtype=x.dtype, device=x.device)
        xt, eps = self.forward(x, t)
        # (net(xt, t) - eps) ** 2, elementwise (:35), as ONE kernel forward and one backward
        return torch.nn.functional.mse_loss(net(xt, t, forcing=forcing), eps, reduction="none")

    def pred_eps(self, score_fn, x, t):
        return score_fn(x, t)

    def _sample_step(self, score_fn, x, t, dt, proc_x0=None):
        """:41-46, the reference's formula kept for API completeness (`sample` itself runs the fused kernels)."""
        eps_pred = self.pred_eps(score_fn, x, t)
        pred_x0 = (x - self.sigma(t) * eps_pred) / self.mu(t)
        if proc_x0 is not None:
            pred_x0 = proc_x0(pred_x0)
        return self.mu(t - dt) * pred_x0 + self.sigma(t - dt) * eps_pred

    # ---------------------------------------------------------------- sampler (:48-97)
    def sample(self, score_fn, noise, steps: int = 64, corrections: int = 0, tau: float = 1.0, proc_x0=None,
               device=None, show_progressbar=True, seed: Optional[int] = None):
        if not isinstance(score_fn, AbstractScoreFunction):
            raise TypeError("SDAPipeline.sample drives this package's score functions (Default/BatchedScoreFunction); "
                            f"got {type(score_fn).__name__}.  There is no generic torch-op sampling loop here.")
        return self._sample_resident(score_fn, noise, steps, corrections, tau, device, show_progressbar, seed, proc_x0)

    def _sample_resident(self, sf: AbstractScoreFunction, noise: Tensor, steps: int, corrections: int, tau: float,
                         device, show_progressbar: bool, seed: Optional[int], proc_x0=None) -> Optional[Tensor]:
        if device is not None and torch.device(device).type == "cuda":
            sf.device = torch.device(device)
        rt = sf.runtime(noise)
        group = sf.shard[2] if sf.shard else None
        rt.reset_finite_check()
        rt.load(noise)
        time_steps = torch.linspace(1, 0, steps + 1).to(dtype=torch.float32)
        dt = 1 / steps
        z_dev = None
        if seed is None:
            # rng == "reference" replays the reference's draws (z.normal_() only, src/thor/pipelines.py:82): nothing else
            # may touch the global CPU generator, or the stream is shifted by one draw
            device_rng = corrections > 0 and self.rng != "reference"
            seed = int(torch.empty((), dtype=torch.int64).random_().item()) if device_rng else 0
            if sf.shard and device_rng:
                # the on-chip Philox stream is keyed by (seed, step, GLOBAL pixel): every rank must use rank 0's draw,
                # otherwise the result would depend on the sharding
                s_t = torch.tensor([seed], dtype=torch.int64, device=rt.device)
                torch.distributed.broadcast(s_t, src=torch.distributed.get_global_rank(group, 0) if group else 0,
                                            group=group)
                seed = int(s_t.item())
        self.traces = {}
        trace_at = set(int(v) for v in self.trace_at) if self.trace_at else ()
        iterator = time_steps[:-1]
        if show_progressbar:
            try:
                from tqdm.auto import tqdm
                iterator = tqdm(iterator, desc="Sampling")
            except Exception:  # pragma: no cover
                pass
        total_start_time = time.time()
        z_host = torch.empty(noise.shape, dtype=torch.float32) if (corrections > 0 and self.rng == "reference") else None
        for istep, t in enumerate(iterator):
            t_next = t - dt
            mu, sigma = _mu_sigma(self, t)
            mu_n, sigma_n = _mu_sigma(self, t_next)
            # predictor
            rt.score(float(t), group)
            if proc_x0 is None:
                rt.predictor(mu, sigma, mu_n, sigma_n)
            else:
                rt.predictor_with_hook(mu, sigma, mu_n, sigma_n, proc_x0)
            rt.halo(group)
            # corrector
            for ic in range(corrections):
                if z_host is not None:
                    z_host.normal_()
                    p = rt.plan
                    if z_dev is None:
                        z_dev = torch.empty_like(rt.x)
                    src = z_host[p.frame_lo:p.frame_hi].to(rt.device, non_blocking=False).contiguous()
                    with torch.cuda.device(rt.device):
                        _lib.check(rt.lib.c2w_traj_pack(src.data_ptr(), z_dev.data_ptr(), p.n_local, rt.C, rt.H * rt.W,
                                                        rt.stream), "c2w_traj_pack")
                rt.score(float(t_next), group)
                rt.guided_eps(mu_n, sigma_n)
                rt.corrector(tau, sigma_n, z_dev, seed, istep * max(corrections, 1) + ic, group)
                rt.halo(group)
            if self.nan_check_every and (istep + 1) % self.nan_check_every == 0:
                rt.check_finite_lagged()  # reads the flag every check, one check behind: no launch-queue stall
            if (istep + 1) in trace_at:
                tr = rt.owned(rt.x)
                if sf.shard:
                    tr = all_gather_frames(tr, rt.plan, group)
                self.traces[istep + 1] = tr.cpu()
        rt.check_finite()
        out = rt.owned(rt.x)
        if sf.shard:
            # time-sharded: the trajectory is assembled on the ranks that asked for it (default: rank 0 of the group;
            # `enable_time_sharding(gather="all")` for every rank) — the others return None
            if sf.shard_gather == "all":
                out = all_gather_frames(out, rt.plan, group)
            else:
                out = gather_frames(out, rt.plan, group, dst=int(sf.shard_gather))
        total_time = time.time() - total_start_time
        print(f"Total sampling time: {total_time:.2f} s  = {total_time / 60:.3f} min = {total_time / 3600:.4f} h")
        if out is None:
            return None
        target = noise.device if device is None else torch.device(device)
        if target.type == "cpu":
            # one asynchronous copy into pinned memory (a pageable destination makes the driver stage it in chunks)
            host = torch.empty(out.shape, dtype=noise.dtype, pin_memory=True)  # a fresh tensor per call: the caller owns it
            host.copy_(out.to(noise.dtype), non_blocking=True)
            torch.cuda.current_stream(rt.device).synchronize()
            return host.reshape(noise.shape)
        return out.to(device=target, dtype=noise.dtype).reshape(noise.shape)
