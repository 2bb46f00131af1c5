"""Time-axis sharding of a trajectory across ranks (new in this build).

Frame i's score depends only on frames i-k .. i+k (src/thor/score.py:68-93), so the Nw = L - 2k window CENTRES
are split evenly over the ranks; the edge ranks additionally own the k edge frames (which cost no extra windows,
src/thor/score.py:76-88).  Before every score evaluation each rank needs the k boundary frames of its neighbours'
current state: one send/recv pair per neighbour (NCCL over NVLink on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass
from typing import List, Optional

import torch
import torch.distributed as dist


@dataclass(frozen=True)
class ShardPlan:
    L: int          # trajectory length (frames)
    k: int          # Markov order
    rank: int
    world: int
    win_lo: int     # windows [win_lo, win_hi) are evaluated on this rank (global window indices)
    win_hi: int

    @property
    def n_win_global(self) -> int:
        return self.L - 2 * self.k

    @property
    def frame_lo(self) -> int:
        """first global frame held locally (including the left halo)"""
        return self.win_lo

    @property
    def frame_hi(self) -> int:
        """one past the last global frame held locally (including the right halo)"""
        return self.win_hi + 2 * self.k

    @property
    def n_local(self) -> int:
        return self.frame_hi - self.frame_lo

    @property
    def own_lo(self) -> int:
        """first OWNED global frame: window centres, plus the k leading frames on rank 0"""
        return 0 if self.rank == 0 else self.win_lo + self.k

    @property
    def own_hi(self) -> int:
        return self.L if self.rank == self.world - 1 else self.win_hi + self.k

    @property
    def own_n(self) -> int:
        return self.own_hi - self.own_lo


def make_plan(L: int, k: int, rank: int = 0, world: int = 1) -> ShardPlan:
    nw = L - 2 * k
    if nw < 1:
        raise ValueError(f"trajectory of {L} frames is shorter than one window of {2 * k + 1}")
    if world > 1 and nw // world < max(k, 1):
        raise ValueError(f"{nw} windows over {world} ranks leaves fewer than k={k} windows per rank; use fewer ranks")
    base, rem = divmod(nw, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return ShardPlan(L, k, rank, world, lo, hi)


def exchange_halos(x_local: torch.Tensor, plan: ShardPlan, group: Optional[dist.ProcessGroup] = None) -> None:
    """Refreshes the k halo frames on each side of `x_local` ([n_local, ...], frames outermost, contiguous) from the
    neighbours' owned boundary frames.  No-op for a single rank."""
    if plan.world == 1:
        return
    k = plan.k
    ops: List[dist.P2POp] = []
    n = plan.n_local
    if plan.rank > 0:  # left neighbour: my first k owned frames -> its right halo; its last k owned -> my left halo
        ops.append(dist.P2POp(dist.isend, x_local[k:2 * k], plan.rank - 1, group))
        ops.append(dist.P2POp(dist.irecv, x_local[0:k], plan.rank - 1, group))
    if plan.rank < plan.world - 1:
        ops.append(dist.P2POp(dist.isend, x_local[n - 2 * k:n - k], plan.rank + 1, group))
        ops.append(dist.P2POp(dist.irecv, x_local[n - k:n], plan.rank + 1, group))
    for req in dist.batch_isend_irecv(ops):
        req.wait()


def exchange_halos_adjoint(v_local: torch.Tensor, plan: ShardPlan, group: Optional[dist.ProcessGroup] = None) -> None:
    """Adjoint of `exchange_halos`: what this rank accumulated in its k halo frames belongs to the neighbour that
    owns them — send it there and ADD what the neighbours accumulated for this rank's boundary frames (the UNet
    vector-Jacobian product reaches k frames beyond the owned range).  No-op for a single rank."""
    if plan.world == 1:
        return
    k = plan.k
    n = plan.n_local
    ops: List[dist.P2POp] = []
    recv_l = recv_r = None
    if plan.rank > 0:
        recv_l = torch.empty_like(v_local[k:2 * k])
        ops.append(dist.P2POp(dist.isend, v_local[0:k].contiguous(), plan.rank - 1, group))
        ops.append(dist.P2POp(dist.irecv, recv_l, plan.rank - 1, group))
    if plan.rank < plan.world - 1:
        recv_r = torch.empty_like(v_local[n - 2 * k:n - k])
        ops.append(dist.P2POp(dist.isend, v_local[n - k:n].contiguous(), plan.rank + 1, group))
        ops.append(dist.P2POp(dist.irecv, recv_r, plan.rank + 1, group))
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    if recv_l is not None:
        v_local[k:2 * k] += recv_l
    if recv_r is not None:
        v_local[n - 2 * k:n - k] += recv_r


def all_gather_frames(x_owned: torch.Tensor, plan: ShardPlan, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Concatenates every rank's owned frames into the full [L, ...] trajectory (end of sampling only)."""
    if plan.world == 1:
        return x_owned
    counts = [make_plan(plan.L, plan.k, r, plan.world).own_n for r in range(plan.world)]
    cmax = max(counts)
    tail = tuple(x_owned.shape[1:])
    mine = torch.zeros((cmax,) + tail, dtype=x_owned.dtype, device=x_owned.device)
    mine[:plan.own_n] = x_owned
    bufs = [torch.empty_like(mine) for _ in counts]
    dist.all_gather(bufs, mine, group=group)  # equal-sized buffers: valid on both nccl and gloo
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)


def gather_frames(x_owned: torch.Tensor, plan: ShardPlan, group: Optional[dist.ProcessGroup] = None,
                  dst: int = 0) -> Optional[torch.Tensor]:
    """Collects every rank's owned frames on rank `dst` only (end of sampling): the other ranks send their frames
    straight into the matching slice of dst's [L, ...] tensor — no padding, no copy on the ranks that did not ask for
    the result.  Returns the full trajectory on `dst`, None elsewhere.  (`all_gather_frames` is the every-rank form.)"""
    if plan.world == 1:
        return x_owned
    x_owned = x_owned.contiguous()
    if plan.rank != dst:
        for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, x_owned, dst, group)]):
            req.wait()
        return None
    full = torch.empty((plan.L,) + tuple(x_owned.shape[1:]), dtype=x_owned.dtype, device=x_owned.device)
    ops: List[dist.P2POp] = []
    for r in range(plan.world):
        pr = make_plan(plan.L, plan.k, r, plan.world)
        if r == dst:
            full[pr.own_lo:pr.own_hi] = x_owned
        else:
            ops.append(dist.P2POp(dist.irecv, full[pr.own_lo:pr.own_hi], r, group))
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    return full


class PeerHalo:
    """The halo exchange over NVLink peer memory (csrc/halo.cu behind c2w_halo_*): every rank's k boundary frames are
    stored straight into the neighbours' mailboxes by a kernel on the compute stream and picked up by a second one —
    no NCCL launch, no host synchronisation.  The 64-byte IPC handles travel once, at construction, through
    torch.distributed.  Same result as `exchange_halos` (bit-exact: it is a copy)."""

    def __init__(self, plan: ShardPlan, frame_shape, device: torch.device, group: Optional[dist.ProcessGroup] = None):
        from . import _lib

        self.lib = _lib.load()
        self.plan, self.device = plan, torch.device(device)
        self.frame_floats = 1
        for d in frame_shape:
            self.frame_floats *= int(d)
        self.handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.c2w_halo_create(plan.k * self.frame_floats * 4, ctypes.byref(self.handle)), "c2w_halo_create")
            mine = (ctypes.c_uint8 * 64)()
            _lib.check(self.lib.c2w_halo_handle(self.handle, mine), "c2w_halo_handle")
            t = torch.tensor(list(mine), dtype=torch.uint8, device=self.device)
            every = [torch.empty_like(t) for _ in range(plan.world)]
            dist.all_gather(every, t, group=group)
            raw = [bytes(e.cpu().tolist()) for e in every]
            left = ctypes.create_string_buffer(raw[plan.rank - 1], 64) if plan.rank > 0 else None
            right = ctypes.create_string_buffer(raw[plan.rank + 1], 64) if plan.rank < plan.world - 1 else None
            rc = self.lib.c2w_halo_connect(self.handle, left, right)
            msg = self.lib.c2w_last_error() if rc != 0 else b""
            ok = torch.tensor([1 if rc == 0 else 0], dtype=torch.int32, device=self.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)  # also the barrier: every mailbox is mapped before a push
            if int(ok.item()) == 0:  # all ranks agree to fall back (a one-sided failure would deadlock the exchange)
                raise _lib.C2WError(f"c2w_halo_connect failed on at least one rank ({msg.decode() if msg else 'a peer'})")

    def exchange(self, x_local: torch.Tensor) -> None:
        from . import _lib

        assert x_local.is_contiguous() and x_local.dtype == torch.float32 and x_local.shape[0] == self.plan.n_local
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream(self.device).cuda_stream
            _lib.check(self.lib.c2w_halo_exchange(self.handle, x_local.data_ptr(), self.plan.n_local, self.frame_floats,
                                                  self.plan.k, st), "c2w_halo_exchange")

    def pull(self, x_local: torch.Tensor) -> None:
        """Second half of an exchange whose push was fused into the kernel that updated x (c2w_guided_step)."""
        from . import _lib

        assert x_local.is_contiguous() and x_local.dtype == torch.float32 and x_local.shape[0] == self.plan.n_local
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream(self.device).cuda_stream
            _lib.check(self.lib.c2w_halo_pull(self.handle, x_local.data_ptr(), self.plan.n_local, self.frame_floats,
                                              self.plan.k, st), "c2w_halo_pull")

    def __del__(self):
        try:
            if getattr(self, "handle", None) and self.handle.value:
                self.lib.c2w_halo_destroy(self.handle)
                self.handle = ctypes.c_void_p()
        except Exception:
            pass


def halo_transport() -> str:
    """C2W_HALO=nccl keeps the torch.distributed send/recv pairs (A/B runs, and the only choice on gloo); default on
    CUDA: peer memory."""
    return os.environ.get("C2W_HALO", "p2p").lower()
