"""The numerical pieces of the reference's PPO loop (ACMPC/training_loop.py) that sit next to the MPC solve, on the
device — SURVEY §8f, row N4: advantage estimation, value-expansion targets and the rollout phase — and the mini-batch
update with its samples partitioned over ranks (one flat all-reduce of the actor + critic gradients).  The critic
model, the outer loop and checkpointing are plain torch in the reference and stay out of scope.
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


class ShardedPPOUpdate:
    """One PPO mini-batch update (training_loop.py:134-199) with the mini-batch partitioned over ranks — SURVEY §8(e):
    the MPC solves of a mini-batch are independent, so each rank evaluates its own samples and the only exchange is ONE
    flat all-reduce of the actor + critic gradients (NCCL over NVLink on GPUs).

    Per call, on this rank's shard ``(mb_states, mb_actions, mb_log_probs_old, mb_advantages, mb_returns, ls_state)`` —
    ``ls_state`` [1, obs_dim] is the observation of the GLOBAL mini-batch's element 0, the same on every rank: the reference
    scores every element's line search with batch element 0's cost (quirk Q4), and each rank evaluates the cost network on that
    observation itself (``ActorMPC.evaluate_actions(ls_state=)``), so no rank waits for another inside the step:
      1. one CUDA-graph replay of ``actor.evaluate_actions`` (cost network, forward iLQR solve), the clipped surrogate,
         the entropy bonus and the critic's value loss (:140-152; ``critic(mb_states.unsqueeze(1)).squeeze()``), the
         autograd backward through ``ILQRSolve`` and both networks, and the copy of every gradient — already weighted by
         ``n_local / n_global`` — into one flat bucket;
      2. the all-reduce of that bucket on a communication stream;
      3. ``optimizer.step()`` of the two optimisers (:197-198) on gradients that are views of the bucket.
    The loss of the global mini-batch is the sample mean, so the reduced bucket holds exactly the gradient a single
    process computes on the whole mini-batch (tests/test_gpu_ppo_sharded.py).  The value-expansion term (:155-172) needs
    the critic on predicted trajectories and enters through ``extra_loss_fn(critic, *extras)`` when given.  AMP / GradScaler
    of the reference are not used: the solve is fp32 either way.
    """

    def __init__(self, actor, critic, example_batch, *, clip_param: float = 0.2, entropy_coeff: float = 0.01,
                 actor_optimizer=None, critic_optimizer=None, group=None, graph: bool = True, extra_loss_fn=None):
        import torch.distributed as dist
        import torch.nn.functional as F

        from ..sharding import GradBucket
        self.actor, self.critic, self.group = actor, critic, group
        self.clip_param, self.entropy_coeff, self.extra_loss_fn = float(clip_param), float(entropy_coeff), extra_loss_fn
        self.actor_optimizer, self.critic_optimizer = actor_optimizer, critic_optimizer
        self._F = F
        self._in = [t.detach().clone() for t in example_batch]
        dev = self._in[0].device
        self.params = [p for p in actor.parameters() if p.requires_grad] + [p for p in critic.parameters() if p.requires_grad]
        n_local = self._in[0].shape[0]
        n_global = torch.tensor([float(n_local)], device=dev)
        self.world = 1
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(n_global, group=group)
            self.world = dist.get_world_size(group)
        self.weight = n_local / float(n_global.item())
        self.bucket = GradBucket({str(i): p.shape for i, p in enumerate(self.params)}, dev, self.params[0].dtype, group)
        self._views = [self.bucket.views[str(i)] for i in range(len(self.params))]
        self._graph = None
        if graph:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(3):
                    self._step()
            torch.cuda.current_stream(dev).wait_stream(side)
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                self._loss = self._step()

    def _step(self):
        F = self._F
        st, ac, lp_old, adv, ret, ls_state = self._in[:6]
        new_lp, entropy = self.actor.evaluate_actions(st, ac, ls_state=ls_state, _validate_args=False)
        ratio = torch.exp(new_lp - lp_old)
        surr1, surr2 = ratio * adv, torch.clamp(ratio, 1.0 - self.clip_param, 1.0 + self.clip_param) * adv
        actor_loss, entropy_loss = -torch.min(surr1, surr2).mean(), -entropy.mean()
        new_values = self.critic(st.unsqueeze(1)).reshape(-1)
        critic_loss = F.mse_loss(new_values, ret)
        if self.extra_loss_fn is not None:
            critic_loss = critic_loss + self.extra_loss_fn(self.critic, *self._in[6:])
        total = actor_loss + self.entropy_coeff * entropy_loss + critic_loss
        grads = torch.autograd.grad(total, self.params, allow_unused=True)
        with torch.no_grad():
            for v, g in zip(self._views, grads):
                if g is None:
                    v.zero_()
                else:
                    torch.mul(g, self.weight, out=v)
        return total.detach()

    def __call__(self, *batch):
        if len(batch) != len(self._in) or any(a.shape != b.shape for a, b in zip(batch, self._in)):
            raise RuntimeError("sharded PPO update was built for inputs of shapes " + str([tuple(t.shape) for t in self._in]))
        with torch.no_grad():
            for dst, src in zip(self._in, batch):
                dst.copy_(src)
        if self._graph is not None:
            self._graph.replay()
            loss = self._loss
        else:
            loss = self._step()
        self.bucket.all_reduce_async()
        self.bucket.join()
        for p, v in zip(self.params, self._views):
            p.grad = v
        if self.actor_optimizer is not None:
            self.actor_optimizer.step()
        if self.critic_optimizer is not None:
            self.critic_optimizer.step()
        return loss
