"""The numerical pieces of the reference's PPO loop (ACMPC/training_loop.py) that sit next to the MPC solve, on the
device — SURVEY §8f, row N4: advantage estimation, value-expansion targets and the rollout phase.  The optimiser
loop, critic and checkpointing are plain torch in the reference and stay out of scope.
"""
from __future__ import annotations

from typing import Callable, Dict

import torch

from .. import ops


def compute_gae_and_returns(rewards: torch.Tensor, values: torch.Tensor, *, gamma: float = 0.99, lam: float = 1.0):
    """``(returns, advantages)`` with the reference's signature (training_loop.py:30-41).  ``rewards``/``values`` are
    ``[T]`` as there, or ``[N, T]`` — the ``vmap(compute_gae_and_returns, in_dims=(0, 0))`` train() builds (:76) —
    in one kernel.  Bit-identical to the reference on the CPU (same operation order, no fused multiply-adds),
    including its bootstrap with the LAST value (``values_extended = cat(values, values[-1:])``)."""
    if rewards.ndim == 1:
        ret, adv = ops.gae(rewards[None], values[None], gamma, lam)
        return ret[0], adv[0]
    return ops.gae(rewards, values, gamma, lam)


def mpve_targets(pred_rewards: torch.Tensor, v_h: torch.Tensor, gamma: float = 0.99) -> torch.Tensor:
    """Value-expansion targets along the MPC prediction (training_loop.py:162-170):
    ``target_t = pred_rewards[:, t] + gamma * target_{t+1}``, ``target_H = v_h``."""
    return ops.discounted_targets(pred_rewards, v_h, gamma)


@torch.no_grad()
def collect_rollout(actor: Callable, envs, episode_len: int) -> Dict[str, torch.Tensor]:
    """Phase 1 of train() (training_loop.py:86-108) against a ``DeviceEnvManager``: no per-environment Python loop and
    no host round trip per step.  ``actor(states, deterministic=False)`` returns
    ``(actions, log_probs, pred_states, pred_actions)`` as ``ActorMPC.forward`` does."""
    cur = envs.reset()
    bufs: Dict[str, list] = {k: [] for k in ("states", "actions", "rewards", "log_probs", "pred_states", "pred_actions")}
    for _ in range(episode_len):
        actions, log_probs, pred_states, pred_actions = actor(cur, deterministic=False)
        nxt, rewards, _, _, _ = envs.step_autoreset(actions)
        for k, v in zip(bufs, (cur, actions, rewards, log_probs, pred_states, pred_actions)):
            bufs[k].append(v)
        cur = nxt
    return {k: torch.stack(v, dim=1) for k, v in bufs.items()}
