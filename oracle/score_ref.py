"""Oracle: Markov-window score composition, Gaussian likelihood guidance (TEST INFRASTRUCTURE ONLY).

Restates src/thor/score.py:
  * unfold / fold            (:68-88)      -> integer index maps (numpy) + tensor versions
  * batched window compose   (:111-154, :156-185)
  * condition_on / log_p     (:44-60)  and the guided score  eps - sigma * d(log p)/dx  (:24-35)
and the observation operator of exp/downscaling.py:129-132 (every t_step-th frame, s_step x s_step mean).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ----------------------------------------------------------------------------- integer index maps
def unfold_index(L: int, k: int, C: int) -> np.ndarray:
    """src/thor/score.py:68-74 — window j, window-channel tau*C + c reads frame j + tau, channel c.
    Returns int64 [Nw, w*C, 2] holding (frame, channel)."""
    w = 2 * k + 1
    nw = L - w + 1
    if nw < 1:
        raise ValueError("trajectory shorter than one window")
    j = np.arange(nw)[:, None]
    wc = np.arange(w * C)[None, :]
    frame = j + wc // C
    chan = np.broadcast_to(wc % C, frame.shape)
    return np.stack([frame, chan], axis=-1).astype(np.int64)


def fold_index(L: int, k: int, C: int) -> np.ndarray:
    """src/thor/score.py:76-88 — output frame i, channel c is taken from (window, window-channel):
    i < k: window 0, slot i;  k <= i < L-k: window i-k, slot k;  i >= L-k: last window, slot i-(L-w).
    Returns int64 [L, C, 2]."""
    w = 2 * k + 1
    nw = L - w + 1
    out = np.zeros((L, C, 2), dtype=np.int64)
    for i in range(L):
        if i < k:
            win, slot = 0, i
        elif i < L - k:
            win, slot = i - k, k
        else:
            win, slot = nw - 1, i - (L - w)
        out[i, :, 0] = win
        out[i, :, 1] = slot * C + np.arange(C)
    return out


def batched_fold_plan(L: int, k: int, batch_size: int) -> List[Tuple[int, int, List[Tuple[int, int]]]]:
    """src/thor/score.py:111-154,165-185 — per batch (first window, n windows, [(window, slot) in output
    order]).  The head slots come only from the first batch, the tail only from the last."""
    w = 2 * k + 1
    nw = L - w + 1
    plan = []
    starts = list(range(0, nw, batch_size))
    for bi, s in enumerate(starts):
        n = min(batch_size, nw - s)
        items: List[Tuple[int, int]] = []
        if bi == 0:
            items += [(s, tau) for tau in range(k)]
        items += [(s + r, k) for r in range(n)]
        if bi == len(starts) - 1:
            items += [(s + n - 1, tau) for tau in range(k + 1, w)]
        plan.append((s, n, items))
    return plan


# ----------------------------------------------------------------------------- tensor versions
def unfold(x: Tensor, k: int) -> Tensor:
    L, C = x.shape[:2]
    idx = torch.from_numpy(unfold_index(L, k, C))
    return x[idx[..., 0], idx[..., 1]]


def fold(n: Tensor, k: int, C: int) -> Tensor:
    nw = n.shape[0]
    L = nw + 2 * k
    idx = torch.from_numpy(fold_index(L, k, C))
    return n[idx[..., 0], idx[..., 1]]


def window_score(net: Callable, x: Tensor, t: Tensor, k: int, batch_size: Optional[int] = None) -> Tensor:
    """DefaultScoreFunction.score_fn (batch_size None, :90-93) or BatchedScoreFunction.score_fn (:156-185)."""
    L, C = x.shape[:2]
    u = unfold(x, k)
    if batch_size is None:
        return fold(net(u, t), k, C)
    pieces = []
    for s, n, items in batched_fold_plan(L, k, batch_size):
        out = net(u[s:s + n], t)
        for win, slot in items:
            pieces.append(out[win - s, slot * C:(slot + 1) * C])
    return torch.stack(pieces, dim=0)


# ----------------------------------------------------------------------------- observation operator
def coarse_grain(x: Tensor, t_step: int, s_step: int) -> Tensor:
    """exp/downscaling.py:129-132 — AvgPool2d(s_step, stride=s_step)(x[::t_step])."""
    return F.avg_pool2d(x[::t_step], s_step, stride=s_step)


def coarse_grain_adjoint(r: Tensor, L: int, t_step: int, s_step: int) -> Tensor:
    """A^T r: r[m, c, p, q] / s^2 spread over tile (p, q) of frame m * t_step, zero elsewhere."""
    up = r.repeat_interleave(s_step, dim=-2).repeat_interleave(s_step, dim=-1) / float(s_step * s_step)
    out = torch.zeros((L,) + tuple(up.shape[1:]), dtype=r.dtype)
    out[::t_step] = up
    return out


def mu_sigma(t: Tensor, eta: float = 1e-3) -> Tuple[Tensor, Tensor]:
    """src/thor/pipelines.py:13-20."""
    import math
    alpha = torch.cos(math.acos(math.sqrt(eta)) * t) ** 2
    return alpha, (1 - alpha ** 2 + eta ** 2).sqrt()


def guided_score(net: Callable, x: Tensor, t: Tensor, k: int, y: Tensor, std, gamma, t_step: int, s_step: int,
                 exact_grad: bool, batch_size: Optional[int] = None, eta: float = 1e-3) -> Tensor:
    """AbstractScoreFunction.__call__ with a likelihood (src/thor/score.py:24-35,44-60), by autograd exactly as
    the reference does it (jacrev of a scalar == one VJP)."""
    mu, sigma = mu_sigma(t, eta)
    xg = x.detach().clone().requires_grad_(True)
    with torch.set_grad_enabled(exact_grad):
        eps = window_score(net, xg, t, k, batch_size)
    with torch.enable_grad():
        x0 = (xg - sigma * eps) / mu
        err = y - coarse_grain(x0, t_step, s_step)
        var = std ** 2 + gamma * (sigma / mu) ** 2
        logp = -(err ** 2 / var).sum() / 2
        (J,) = torch.autograd.grad(logp, xg)
    return (eps - sigma * J).detach()


def guided_score_closed_form(eps: Tensor, x: Tensor, t: Tensor, y: Tensor, std, gamma, t_step: int, s_step: int,
                             eta: float = 1e-3) -> Tensor:
    """exact_grad=False in closed form: J = A^T((y - A x0)/var) / mu."""
    mu, sigma = mu_sigma(t, eta)
    x0 = (x - sigma * eps) / mu
    err = y - coarse_grain(x0, t_step, s_step)
    var = std ** 2 + gamma * (sigma / mu) ** 2
    J = coarse_grain_adjoint(err / var, x.shape[0], t_step, s_step) / mu
    return eps - sigma * J
