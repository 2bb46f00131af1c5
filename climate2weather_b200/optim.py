"""N2, optimiser half: `optimizer.step()` + `ema.update()` of the reference's training step
(training_loop.py:381-390) as one fused pass over a flat parameter buffer (`c2w_adamw_ema_step`).

`AdamW` keeps the part of `torch.optim.AdamW`'s interface the training loop touches — `param_groups[i]["lr"]` is
re-read every step (training_loop.py:380-382), `zero_grad()`, `step()` — and `StandardEMA` the reference class's
(`src/thor/ema.py`): `update()`, `reset()`, `get()`, `emas`, `rates`.  Fusing is opt-in and explicit:
`optimizer.fuse_ema(ema)`; the next `ema.update()` after a fused `step()` is then a no-op, exactly the reference's
order of operations (nothing touches the parameters between the two calls).

What this does NOT provide yet: parameter gradients of the ScoreUNet from this package's kernels (the wgrad half of N2);
gradients may come from any autograd graph over the parameters.  No CPU path.
"""
from __future__ import annotations

import copy
import ctypes
from typing import Iterable, List, Optional

import torch

from . import _lib


def _flatten(tensors: List[torch.Tensor]):
    """One contiguous fp32 buffer holding `tensors` back to back, each start padded to 4 elements (16 B)."""
    dev = tensors[0].device
    offs, n = [], 0
    for t in tensors:
        offs.append(n)
        n += (t.numel() + 3) // 4 * 4
    flat = torch.zeros(n, dtype=torch.float32, device=dev)
    return flat, offs


class AdamW:
    """torch.optim.AdamW (decoupled weight decay, bias correction; no amsgrad / maximize / foreach options)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, loss_scaling: float = 1.0, data_parallel_group="none",
                 direct_grads: bool = False):
        self.lib = _lib.load()
        ps = [p for p in params if p.requires_grad]
        if not ps:
            raise ValueError("optimizer got an empty parameter list")
        if any((not p.is_cuda) or p.dtype != torch.float32 or p.device != ps[0].device for p in ps):
            raise _lib.C2WError("AdamW needs fp32 parameters on one CUDA device: climate2weather_b200 has no CPU path")
        self.params = ps
        self.param_groups = [dict(params=ps, lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)]
        self.loss_scaling = loss_scaling
        self.flat, self.offsets = _flatten(ps)
        self.grad = torch.zeros_like(self.flat)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        for p, o in zip(ps, self.offsets):  # parameters and their gradients become views of the flat buffers
            self.flat[o:o + p.numel()].view_as(p).copy_(p.data)
            p.data = self.flat[o:o + p.numel()].view_as(p)
            p.grad = self.grad[o:o + p.numel()].view_as(p)
        self.step_count = 0
        # data-parallel training without a DDP wrapper (training_loop.py:116,375-378): the gradients of all ranks are
        # averaged with ONE in-place NCCL all-reduce of the flat gradient buffer at the start of step() — no bucket
        # copies.  "none": gradients are taken as they are (single process, or a DDP wrapper already averaged them);
        # None: the default process group; or a torch.distributed group.
        self.dp_group = data_parallel_group
        self._ema: Optional["StandardEMA"] = None
        self._owners: List[torch.nn.Module] = []  # modules whose packed-weight engines go stale when we step
        from .model import owner_of
        for p in ps:
            owner = owner_of(p)
            if owner is not None:
                self.track(owner)
        # direct_grads: a ScoreUNet whose parameters are exactly this optimizer's (same order) lets its backward kernels
        # accumulate straight into self.grad instead of handing autograd 228 tensors to add.  Autograd's per-parameter
        # hooks do not fire on that route — do not combine with a DDP wrapper (use data_parallel_group instead).
        if direct_grads:
            owners = [m for m in self._owners if [id(q) for q in m.parameters()] == [id(q) for q in ps]]
            if not owners:
                raise ValueError("direct_grads needs the parameters of exactly one climate2weather_b200.ScoreUNet")
            for m in owners:
                m._direct_grad = self.grad

    def track(self, *modules: torch.nn.Module) -> "AdamW":
        """Modules holding these parameters (ScoreUNet instances): `step()` updates the flat buffer through raw
        pointers, which does not bump `Tensor._version`, so their cached packed-weight engines are invalidated
        explicitly (`ScoreUNet.invalidate_engines`)."""
        for m in modules:
            if hasattr(m, "invalidate_engines") and all(m is not o for o in self._owners):
                self._owners.append(m)
        return self

    def _bind_views(self) -> None:
        """Parameters must alias the flat buffer (a later `.to()`, `load_state_dict(assign=True)` or manual `p.data =`
        detaches them): re-adopt the current values and rebind, so a step never updates memory nobody reads."""
        for p, o in zip(self.params, self.offsets):
            if p.data_ptr() != self.flat.data_ptr() + 4 * o:
                if p.device != self.flat.device or p.dtype != torch.float32:
                    raise _lib.C2WError("a parameter left the optimizer's device/dtype; rebuild the optimizer")
                self.flat[o:o + p.numel()].view_as(p).copy_(p.data)
                p.data = self.flat[o:o + p.numel()].view_as(p)

    # ---- checkpointing (training_loop.py:131-138 saves/restores the optimiser through CheckpointIO): torch.optim layout
    def state_dict(self) -> dict:
        state = {}
        for i, (p, o) in enumerate(zip(self.params, self.offsets)):
            sl = slice(o, o + p.numel())
            state[i] = dict(step=torch.tensor(float(self.step_count)),
                            exp_avg=self.exp_avg[sl].view_as(p).clone(), exp_avg_sq=self.exp_avg_sq[sl].view_as(p).clone())
        groups = [{k: v for k, v in g.items() if k != "params"} | {"params": list(range(len(self.params)))}
                  for g in self.param_groups]
        return dict(state=state if self.step_count > 0 else {}, param_groups=groups, loss_scaling=self.loss_scaling)

    def load_state_dict(self, sd: dict) -> None:
        groups = sd["param_groups"]
        if len(groups) != 1 or len(groups[0]["params"]) != len(self.params):
            raise ValueError("optimizer state does not match this parameter list")
        for k, v in groups[0].items():
            if k != "params":
                self.param_groups[0][k] = tuple(v) if k == "betas" else v
        self.loss_scaling = sd.get("loss_scaling", self.loss_scaling)
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        self.step_count = 0
        for i, st in sd.get("state", {}).items():
            i = int(i)
            p, o = self.params[i], self.offsets[i]
            sl = slice(o, o + p.numel())
            self.exp_avg[sl].view_as(p).copy_(st["exp_avg"])
            self.exp_avg_sq[sl].view_as(p).copy_(st["exp_avg_sq"])
            self.step_count = max(self.step_count, int(float(st["step"])))

    def broadcast_parameters(self, src: int = 0) -> None:
        """Every rank starts from rank `src`'s parameters (what a DDP wrapper does at construction)."""
        import torch.distributed as dist

        dist.broadcast(self.flat, src=src, group=None if isinstance(self.dp_group, str) else self.dp_group)
        for m in self._owners:
            m.invalidate_engines()

    def __getstate__(self):
        d = self.__dict__.copy()
        d.pop("lib", None)  # a ctypes.CDLL cannot be pickled; reloaded on demand
        if not isinstance(d.get("dp_group"), str):
            d["dp_group"] = None  # process groups do not pickle; the default group is re-resolved
        return d

    def __setstate__(self, d):
        self.__dict__.update(d)
        self.lib = _lib.load()

    def fuse_ema(self, ema: "StandardEMA") -> None:
        """Update `ema`'s first rate in the same pass as the parameters."""
        if [id(p) for p in ema.net.parameters() if p.requires_grad] != [id(p) for p in self.params]:
            raise ValueError("the EMA tracks a different parameter list than this optimizer")
        ema._bind(self)
        self._ema = ema
        self.track(ema.net)

    def zero_grad(self, set_to_none: bool = False) -> None:
        self.grad.zero_()
        for p, o in zip(self.params, self.offsets):  # keep the views bound even if someone dropped .grad
            if p.grad is None or p.grad.data_ptr() != self.grad.data_ptr() + 4 * o:
                p.grad = self.grad[o:o + p.numel()].view_as(p)

    @torch.no_grad()
    def step(self) -> None:
        self._bind_views()
        for p, o in zip(self.params, self.offsets):
            if p.grad is None:
                raise RuntimeError("a parameter has no gradient; call backward() before step()")
            if p.grad.data_ptr() != self.grad.data_ptr() + 4 * o:  # autograd installed a fresh tensor
                self.grad[o:o + p.numel()].view_as(p).copy_(p.grad)
        if not (isinstance(self.dp_group, str) and self.dp_group == "none"):
            import torch.distributed as dist

            dist.all_reduce(self.grad, op=dist.ReduceOp.AVG, group=self.dp_group)
        g = self.param_groups[0]
        self.step_count += 1
        hp = _lib.AdamW(float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]),
                        float(g["weight_decay"]), 0.0, 1.0 / float(self.loss_scaling), self.step_count)
        ema_ptr = None
        if self._ema is not None:
            hp.ema_rate = float(self._ema.rates[0])
            ema_ptr = self._ema._flat[0].data_ptr()
            self._ema._fused_pending = True
        dev = self.flat.device
        with torch.cuda.device(dev):
            st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(self.lib.c2w_adamw_ema_step(self.flat.data_ptr(), self.grad.data_ptr(), self.exp_avg.data_ptr(),
                                                   self.exp_avg_sq.data_ptr(), ema_ptr, self.flat.numel(),
                                                   ctypes.byref(hp), st), "c2w_adamw_ema_step")
        for m in self._owners:
            m.invalidate_engines()
        if self._ema is not None:
            self._ema._invalidate(0)


class StandardEMA:
    """src/thor/ema.py: exponential moving average(s) of the weights; `emas` are deep copies of `net`."""

    @torch.no_grad()
    def __init__(self, net: torch.nn.Module, rates=[0.9999]):
        self.net = net
        self.rates = list(rates)
        self.emas = [copy.deepcopy(net) for _ in self.rates]
        self._flat: List[torch.Tensor] = []
        self._opt: Optional[AdamW] = None
        self._fused_pending = False

    def _bind(self, opt: AdamW) -> None:
        """Lay every EMA copy's parameters out like the optimizer's flat buffer."""
        self._opt = opt
        self._flat = []
        for ema in self.emas:
            ps = [p for p in ema.parameters() if p.requires_grad]
            flat = torch.zeros_like(opt.flat)
            for p, o in zip(ps, opt.offsets):
                flat[o:o + p.numel()].view_as(p).copy_(p.data)
                p.data = flat[o:o + p.numel()].view_as(p)
            self._flat.append(flat)
            self._invalidate(len(self._flat) - 1)

    def _invalidate(self, i: int) -> None:
        """EMA copy i changed through its flat buffer: its cached packed-weight engine (validation sampling runs on
        the EMA network, training_loop.py:277-313) must be rebuilt."""
        if hasattr(self.emas[i], "invalidate_engines"):
            self.emas[i].invalidate_engines()

    @torch.no_grad()
    def reset(self):
        for i, ema in enumerate(self.emas):
            for p_net, p_ema in zip(self.net.parameters(), ema.parameters()):
                p_ema.copy_(p_net)
            self._invalidate(i)

    @torch.no_grad()
    def update(self, **kwargs):
        first = 0
        if self._fused_pending:  # rate 0 was applied inside the optimizer's pass
            self._fused_pending = False
            first = 1
        for i in range(first, len(self.rates)):
            rate, ema = self.rates[i], self.emas[i]
            if self._opt is not None:
                self._flat[i].mul_(rate).add_(self._opt.flat, alpha=1 - rate)
            else:
                for p_net, p_ema in zip(self.net.parameters(), ema.parameters()):
                    p_ema.detach().mul_(rate).add_(p_net, alpha=1 - rate)
            self._invalidate(i)

    @torch.no_grad()
    def get(self):
        for ema in self.emas:
            for b_net, b_ema in zip(self.net.buffers(), ema.buffers()):
                b_ema.copy_(b_net)
        return [(ema, f"-{rate:.6f}") for rate, ema in zip(self.rates, self.emas)]

    def state_dict(self):
        return dict(rates=self.rates, emas=[ema.state_dict() for ema in self.emas])

    def load_state_dict(self, state):
        self.rates = state["rates"]
        for ema, s_ema in zip(self.emas, state["emas"]):
            ema.load_state_dict(s_ema)
