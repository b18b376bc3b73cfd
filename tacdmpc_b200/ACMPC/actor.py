"""ActorMPC (ACMPC/actor.py:13-120) with the cost kept as diagonals end to end.

The reference maps an observation through an MLP to ``q_diag, p`` of shape ``[B, T, n]``, expands
``C = diag_embed(q_diag)`` to ``[B, T, n, n]`` and differentiates back through the dense tensor
(actor.py:58-81).  Here the same numbers reach the kernels as ``q_diag, p`` (``DiagQuadCost`` /
``tacd_problem.C_diag``) and come back as ``grad_q_diag, grad_p``; constructor, ``forward`` and
``evaluate_actions`` keep the reference's signatures and return values.  ``fused=False`` runs the
reference's dense formulation through the same kernels (used by the parity tests).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor
from torch.distributions import Normal

from ..DifferentialMPC import DiagQuadCost, DifferentiableMPCController, GeneralQuadCost


class ActorMPC(nn.Module):
    def __init__(self, nx: int, nu: int, horizon: int, dt: float, f_dyn, f_dyn_jac=None, device: str = "cuda",
                 grad_method: str = "auto_diff", observation_dim: Optional[int] = None, fused: bool = True):
        super().__init__()
        self.nx, self.nu, self.horizon, self.dt = nx, nu, horizon, dt
        self.device = torch.device(device)
        self.dtype = torch.float32
        self.fused = fused
        n = nx + nu
        input_dim = observation_dim if observation_dim is not None else nx
        self.cost_map_net = nn.Sequential(
            nn.Linear(input_dim, 512), nn.ReLU(),
            nn.Linear(512, 512), nn.ReLU(),
            nn.Linear(512, horizon * n * 2),
        ).to(self.device, dtype=self.dtype)
        z = lambda *s: torch.zeros(*s, device=self.device, dtype=self.dtype)
        if fused:
            cost_module = DiagQuadCost(nx, nu, z(1, horizon, n), z(1, horizon, n), z(1, n), z(1, n), device=device)
        else:
            cost_module = GeneralQuadCost(nx=nx, nu=nu, C=z(horizon, n, n), c=z(horizon, n), C_final=z(n, n),
                                          c_final=z(n), device=device)
        self.mpc = DifferentiableMPCController(
            f_dyn=f_dyn, total_time=horizon, step_size=dt, horizon=horizon, cost_module=cost_module,
            f_dyn_jac=f_dyn_jac, device=device, reg_eps=1e-2, grad_method=grad_method)
        self.log_std = nn.Parameter(torch.full((nu,), 0.0, device=self.device, dtype=self.dtype))

    def _update_cost_module(self, x: Tensor):
        """softplus(raw) + 1e-2 on the diagonal, raw linear term, C_final = C[:, -1] (actor.py:58-81)."""
        B, n = x.shape[0], self.nx + self.nu
        params = self.cost_map_net(x)
        q_raw, p_raw = params.split([self.horizon * n, self.horizon * n], dim=-1)
        q_diag = (F.softplus(q_raw) + 1e-2).view(B, self.horizon, n)
        p = p_raw.view(B, self.horizon, n)
        cm = self.mpc.cost_module
        if self.fused:
            cm.set_diag(q_diag, p, q_diag[:, -1].clone(), p[:, -1].clone())
        else:
            C = torch.diag_embed(q_diag)
            cm.C, cm.c = C, p
            cm.C_final, cm.c_final = C[:, -1].clone(), p[:, -1].clone()
        cm.set_reference(torch.zeros(B, self.horizon + 1, self.nx, device=self.device, dtype=self.dtype),
                         torch.zeros(B, self.horizon, self.nu, device=self.device, dtype=self.dtype))

    def forward(self, x: Tensor, deterministic: bool = False, *, _validate_args=None) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
        if x.ndim == 1:
            x = x.unsqueeze(0)
        self._update_cost_module(x)
        U_init = torch.randn(x.shape[0], self.horizon, self.nu, device=self.device, dtype=self.dtype) * 0.01
        predicted_states, predicted_actions = self.mpc(x[:, :self.nx], U_init)
        u_mpc_mean = predicted_actions[:, 0]
        dist = Normal(u_mpc_mean, self.log_std.exp(), validate_args=_validate_args)
        action = u_mpc_mean if deterministic else dist.rsample()
        log_prob = dist.log_prob(action).sum(dim=-1)
        return action, log_prob, predicted_states, predicted_actions

    def graphed_update(self, loss_fn, obs, actions, *extras) -> "GraphedUpdate":
        """``evaluate_actions`` + ``loss_fn`` + backward for the mini-batch shape of the example tensors as ONE
        CUDA-graph replay — see ``GraphedUpdate``."""
        return GraphedUpdate(self, loss_fn, obs, actions, *extras)

    def graphed(self, batch_size: int, deterministic: bool = False) -> "GraphedPolicy":
        """The rollout-phase policy call (no gradient) for a fixed batch size as ONE CUDA-graph replay — see
        ``GraphedPolicy``."""
        return GraphedPolicy(self, batch_size, deterministic)

    def _line_search_cost_of(self, x0: Tensor):
        """The cost the network assigns to ONE observation, as the controller's line-search override (quirk Q4: under
        ``compat="reference"`` every element's candidates are scored with batch element 0's cost).  A rank that holds a
        shard of a mini-batch evaluates the network on GLOBAL element 0's observation itself — same weights on every
        rank — instead of waiting for a broadcast of 2(T+1)n numbers inside the step."""
        n = self.nx + self.nu
        with torch.no_grad():
            params = self.cost_map_net(x0.reshape(1, -1))
            q_raw, p_raw = params.split([self.horizon * n, self.horizon * n], dim=-1)
            q = (F.softplus(q_raw) + 1e-2).view(self.horizon, n)
            p = p_raw.view(self.horizon, n)
        return (q, p, q[-1].clone(), p[-1].clone())

    def evaluate_actions(self, x: Tensor, actions: Tensor, *, ls_state: Optional[Tensor] = None, _validate_args=None) -> Tuple[Tensor, Tensor]:
        """``ls_state`` (not in the reference): observation of the GLOBAL mini-batch's element 0 when ``x`` is one rank's
        shard of it (``ShardedPPOUpdate``); None = ``x[0]``, the reference's behaviour."""
        if x.ndim == 1:
            x = x.unsqueeze(0)
        self._update_cost_module(x)
        U_init = torch.zeros(x.shape[0], self.horizon, self.nu, device=self.device, dtype=self.dtype)
        if ls_state is not None:
            self.mpc.ls_cost_override = self._line_search_cost_of(ls_state)
        try:
            _, predicted_actions = self.mpc(x[:, :self.nx], U_init)
        finally:
            if ls_state is not None:
                self.mpc.ls_cost_override = None
        dist = Normal(predicted_actions[:, 0], self.log_std.exp(), validate_args=_validate_args)
        return dist.log_prob(actions).sum(dim=-1), dist.entropy().sum(dim=-1)


class GraphedPolicy:
    """``actor(obs, deterministic)`` under ``torch.no_grad()`` captured in a CUDA graph (at PPO batch
    sizes the policy call is host-bound — ~1 ms of Python / launch overhead around ~0.3 ms of kernels).

    The cost network, the whole iLQR solve and the action sampling replay as one graph launch on new observations;
    nothing in the solve synchronises the host or allocates outside torch's allocator, which is what makes the capture
    legal.  ``torch.distributions`` argument validation (a device->host ``.all()``) is switched off
    inside the graph.  Outputs are fresh tensors (copies of the graph's static outputs).  Weights are read at replay
    time, so optimiser steps between calls are seen; the batch size is fixed.
    """

    def __init__(self, actor: ActorMPC, batch_size: int, deterministic: bool = False):
        self.actor, self.batch_size = actor, int(batch_size)
        in_dim = actor.cost_map_net[0].in_features
        self._obs = torch.zeros(self.batch_size, in_dim, device=actor.device, dtype=actor.dtype)
        side = torch.cuda.Stream(device=actor.device)
        side.wait_stream(torch.cuda.current_stream(actor.device))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(3):   # warm-up outside the capture: workspace / allocator state, lazy initialisations
                actor(self._obs, deterministic, _validate_args=False)
        torch.cuda.current_stream(actor.device).wait_stream(side)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph), torch.no_grad():
            self._out = actor(self._obs, deterministic, _validate_args=False)

    @torch.no_grad()
    def __call__(self, obs: Tensor, deterministic=None) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
        if obs.shape != self._obs.shape:
            raise RuntimeError(f"graphed policy was captured for observations {tuple(self._obs.shape)}, got {tuple(obs.shape)}")
        self._obs.copy_(obs)
        self._graph.replay()
        return tuple(t.clone() for t in self._out)


class GraphedUpdate:
    """The PPO update's device work for a fixed mini-batch shape — ``log_probs, entropy = actor.evaluate_actions(obs,
    actions)``, ``loss = loss_fn(log_probs, entropy, *extras)``, ``loss.backward()`` (``training_loop.py:140-183``) — as one
    CUDA-graph replay.  The forward solve, the autograd backward through ``ILQRSolve`` and the cost network are all
    capturable (no host synchronisation, torch-allocator memory only), and at the reference's mini-batch of 512 the
    eager step is host-bound (≈ 2 ms around ≈ 0.5 ms of kernels).

    After a call the gradients are in the actor parameters' ``.grad`` (the graph's static output tensors, attached on
    every call, so an ``optimizer.zero_grad(set_to_none=True)`` in between is harmless); the optimiser step stays with
    the caller.  Weights are read at replay time.  Gradients are overwritten, not accumulated.  As for any whole-step
    CUDA-graph capture in PyTorch, build it before eager backward passes through THIS actor's parameters have run on the
    default stream (their autograd state is bound to that stream and a capture may not depend on it); eager steps
    afterwards are fine.
    """

    def __init__(self, actor: ActorMPC, loss_fn, obs: Tensor, actions: Tensor, *extras: Tensor):
        self.actor, self.loss_fn = actor, loss_fn
        self._in = [t.detach().clone() for t in (obs, actions, *extras)]
        self._params = [p for p in actor.parameters() if p.requires_grad]
        side = torch.cuda.Stream(device=actor.device)
        side.wait_stream(torch.cuda.current_stream(actor.device))
        with torch.cuda.stream(side):
            for _ in range(3):
                self._step()
        torch.cuda.current_stream(actor.device).wait_stream(side)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._loss, self._grads = self._step()

    def _step(self):
        log_probs, entropy = self.actor.evaluate_actions(self._in[0], self._in[1], _validate_args=False)
        loss = self.loss_fn(log_probs, entropy, *self._in[2:])
        # torch.autograd.grad, not loss.backward(): no AccumulateGrad nodes take part, so gradient accumulators left
        # over from eager steps on the default stream cannot pull that stream into the capture
        grads = torch.autograd.grad(loss, self._params, allow_unused=True)
        return loss.detach(), grads

    def __call__(self, obs: Tensor, actions: Tensor, *extras: Tensor) -> Tensor:
        new = (obs, actions, *extras)
        if len(new) != len(self._in) or any(a.shape != b.shape for a, b in zip(new, self._in)):
            raise RuntimeError("graphed update was captured for inputs of shapes " + str([tuple(t.shape) for t in self._in]))
        with torch.no_grad():
            for dst, src in zip(self._in, new):
                dst.copy_(src)
        self._graph.replay()
        for p, g in zip(self._params, self._grads):
            p.grad = g
        return self._loss.clone()
