"""Multi-GPU use of the solve: one process per GPU, the batch of independent solves partitioned
contiguously across ranks.

The solve itself needs NO data-path collective: every batch element is an independent problem (quirk Q11).
The only exchanges are
  * quirk Q4 under ``compat="reference"`` with per-element costs: every rank's line search must use the
    cost of GLOBAL element 0, so rank 0 broadcasts ``(C[0], c[0], C_final[0], c_final[0])`` once per call;
  * gradients of parameters shared by all ranks (a shared ``C``, or the actor/critic weights that produced
    per-element costs): one flat all-reduce (NCCL over NVLink on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Iterable, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(B: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of a batch of B for ``rank``; sizes differ by at most one."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard(t: torch.Tensor, rank: int, world: int, B: int | None = None) -> torch.Tensor:
    """Slice the leading (batch) dimension of ``t`` for this rank."""
    lo, hi = shard_range(t.shape[0] if B is None else B, rank, world)
    return t[lo:hi]


def broadcast_line_search_cost(controller, group=None, src: int = 0) -> None:
    """Install global element 0's cost as the line-search cost on every rank (quirk Q4).

    No-op when both the running and the terminal cost are shared (then the line search already uses the
    same cost everywhere) or when ``compat="exact"``."""
    cm = controller.cost_module
    if getattr(controller, "compat", "reference") != "reference":
        controller.ls_cost_override = None
        return
    if hasattr(cm, "q") and hasattr(cm, "q_final"):
        # DiagQuadCost: always per-element; element 0's DIAGONALS travel ([T,n], [T,n], [n], [n]).  The controller
        # expands them if a solve ends up on the dense kernels (controller._costs), so the override is valid on both paths.
        parts = [cm.q[0], cm.p[0], cm.q_final[0], cm.p_final[0]]
    else:
        C, c, Cf, cf = cm.C, cm.c, cm.C_final, cm.c_final
        if C.ndim == 3 and Cf.ndim == 2:
            controller.ls_cost_override = None
            return
        parts = [(C[0] if C.ndim == 4 else C), (c[0] if c.ndim == 3 else c), (Cf[0] if Cf.ndim == 3 else Cf),
                 (cf[0] if cf.ndim == 2 else cf)]
    parts = [p.detach().clone().contiguous() for p in parts]
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        flat = torch.cat([p.reshape(-1) for p in parts])
        dist.broadcast(flat, src=src, group=group)
        out, o = [], 0
        for p in parts:
            out.append(flat[o:o + p.numel()].view_as(p))
            o += p.numel()
        parts = out
    controller.ls_cost_override = tuple(parts)


def allreduce_sum_(tensors: Iterable[torch.Tensor], group=None) -> None:
    """Sum the given tensors across ranks in place with ONE flat bucket (latency-bound sizes: a shared
    cost gradient is ~2 KB, actor+critic gradients ~3.8 MB)."""
    ts = [t for t in tensors if t is not None]
    if not ts or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    flat = torch.cat([t.reshape(-1) for t in ts])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    o = 0
    for t in ts:
        t.copy_(flat[o:o + t.numel()].view_as(t))
        o += t.numel()


def allreduce_max_scalar(x: float, device, group=None) -> float:
    """max over ranks of a python float (device timings)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


class GradBucket:
    """ONE flat buffer whose named views are handed to the backward kernels as their output tensors
    (``ops.ilqr_backward(..., out=bucket.views)``), so the cross-rank sum of shared-parameter gradients is a single
    all-reduce with no gather (`torch.cat`) before it and no scatter (`copy_`) after it, issued on a COMMUNICATION stream:
    the caller's stream is not made to wait — the next solve overlaps the collective — until ``join()`` is called where the
    reduced gradient is consumed (the collective sat on the critical path of every step)."""

    def __init__(self, shapes: dict, device, dtype=torch.float32, group=None):
        self.group = group
        n = sum(int(torch.Size(s).numel()) for s in shapes.values())
        self.flat = torch.zeros(n, device=device, dtype=dtype)
        self.views, o = {}, 0
        for k, s in shapes.items():
            m = int(torch.Size(s).numel())
            self.views[k] = self.flat[o:o + m].view(s)
            o += m
        self.cuda = self.flat.is_cuda
        self._comm = torch.cuda.Stream(device=device) if self.cuda else None
        self._ready = torch.cuda.Event() if self.cuda else None
        self.done = torch.cuda.Event(enable_timing=True) if self.cuda else None
        self._pending = False

    def _active(self):
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def all_reduce_async(self):
        """Sum across ranks once the work queued so far on the current stream has produced the bucket."""
        if not self.cuda:
            if self._active():
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            return
        cur = torch.cuda.current_stream(self.flat.device)
        self._ready.record(cur)
        with torch.cuda.stream(self._comm):
            self._comm.wait_event(self._ready)
            if self._active():
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.done.record(self._comm)
        self._pending = True

    def join(self):
        """Make the current stream wait for the last ``all_reduce_async`` (no host synchronisation)."""
        if self.cuda and self._pending:
            torch.cuda.current_stream(self.flat.device).wait_event(self.done)
            self._pending = False
