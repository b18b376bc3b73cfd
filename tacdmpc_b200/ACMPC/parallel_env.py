"""On-device batched environments with the interface of the reference's ``ParallelEnvManager``
(ACMPC/parallel_env.py:6-58) — SURVEY §8f, row N4.

The reference steps ``num_envs`` Python gym environments one after the other and moves every action and
observation across the host/device boundary (``actions.cpu().numpy()`` ... ``torch.from_numpy(...).to(device)``);
at 4096 environments that loop dominates a PPO iteration.  ``DeviceEnvManager`` keeps all environment states in one
device tensor and advances them with one kernel (``tacd_env_step``): dynamics plugin, termination test, step
counter / truncation, quadratic reward and — for ``step_autoreset`` — the reset-on-done of
``training_loop.py:103-107``.  CUDA only; there is no CPU path.
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Sequence, Tuple

import torch

from .. import _abi, ops
from ..dynamics import CartPole, DoubleIntegrator, NativeDynamics


class DeviceEnvManager:
    """``num_envs`` copies of one environment whose transition is a shipped dynamics plugin.

    reward = -(sum_i w_x[i] s_i^2 + sum_j w_u[j] a_j^2) on the next (``reward_on_next``) or current state;
    terminated = any |x'_i| > term_limit[i]; truncated = steps >= max_steps.  ``reset_fn(n, device, dtype)`` draws
    ``n`` start states (the reference's ``env.reset()``).
    """

    def __init__(self, dynamics: NativeDynamics, num_envs: int, device, *, dt: float, reset_fn: Callable,
                 max_steps: int = 0, term_limit: Optional[Sequence[float]] = None, w_x: Optional[Sequence[float]] = None,
                 w_u: Optional[Sequence[float]] = None, clip_action: bool = False, reward_on_next: bool = True,
                 u_min: Optional[Sequence[float]] = None, u_max: Optional[Sequence[float]] = None, dtype=torch.float32):
        if not isinstance(dynamics, NativeDynamics):
            raise TypeError("DeviceEnvManager needs a dynamics plugin with a device twin (tacdmpc_b200.dynamics)")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DeviceEnvManager runs on CUDA devices only (there is no CPU path)")
        self.num_envs, self.dtype, self.dt = int(num_envs), dtype, float(dt)
        self.dynamics, self.reset_fn = dynamics, reset_fn
        self.max_steps, self.term_limit, self.w_x, self.w_u = int(max_steps), term_limit, w_x, w_u
        self.clip_action, self.reward_on_next = bool(clip_action), bool(reward_on_next)
        dynA, dynB = dynamics.device_matrices(self.device, dtype)
        self._spec = ops.Spec(nx=dynamics.nx, nu=dynamics.nu, T=1, dt=self.dt, dyn_id=dynamics.dyn_id,
                              dyn_params=tuple(getattr(dynamics, "params", ())), dyn_A=dynA, dyn_B=dynB, u_min=u_min, u_max=u_max)
        self.single_observation_space_shape = (dynamics.nx,)
        self.single_action_space_shape = (dynamics.nu,)
        self.state = None
        self.steps = torch.zeros(self.num_envs, dtype=torch.int32, device=self.device)

    def reset(self) -> torch.Tensor:
        """Resets all environments; returns the stacked observations (parallel_env.py:26-32)."""
        self.state = self.reset_fn(self.num_envs, self.device, self.dtype).contiguous()
        self.steps.zero_()
        return self.state.clone()

    def _step(self, actions, reset_state):
        if self.state is None:
            raise RuntimeError("call reset() before step()")
        obs, nxt, rew, term, trunc = ops.env_step(
            self._spec, self.state, actions, steps=self.steps, reset_state=reset_state, max_steps=self.max_steps,
            term_limit=self.term_limit, w_x=self.w_x, w_u=self.w_u, clip_action=self.clip_action,
            reward_on_next=self.reward_on_next)
        self.state = nxt
        return obs, nxt, rew, term, trunc

    def step(self, actions: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, List[dict]]:
        """``(next_states, rewards, terminateds, truncateds, infos)`` as parallel_env.py:34-54; finished environments
        stay where they are until ``reset_done`` (or use ``step_autoreset``)."""
        obs, _, rew, term, trunc = self._step(actions, None)
        return obs, rew, term, trunc, [{}] * self.num_envs

    def reset_done(self, done: torch.Tensor) -> torch.Tensor:
        """Restart the environments flagged in ``done`` (bool [num_envs]); returns the current states."""
        fresh = self.reset_fn(self.num_envs, self.device, self.dtype)
        self.state = torch.where(done[:, None], fresh, self.state).contiguous()
        self.steps.masked_fill_(done, 0)
        return self.state.clone()

    def step_autoreset(self, actions: torch.Tensor, reset_state: Optional[torch.Tensor] = None):
        """Step + the reset-on-done of training_loop.py:103-107 in the same kernel.  Returns
        ``(current_states, rewards, terminateds, truncateds, obs)`` where ``current_states`` are what the next policy
        call sees (``obs`` with finished environments replaced by fresh starts)."""
        if reset_state is None:
            reset_state = self.reset_fn(self.num_envs, self.device, self.dtype)
        obs, nxt, rew, term, trunc = self._step(actions, reset_state)
        return nxt.clone(), rew, term, trunc, obs

    def close(self):
        self.state = None


def uniform_reset(low: float, high: float, nx: int, generator: Optional[torch.Generator] = None) -> Callable:
    """``np.random.uniform(low, high, size=(nx,))`` of the example environments, batched on the device."""
    def fn(n, device, dtype):
        return low + (high - low) * torch.rand(n, nx, device=device, dtype=dtype, generator=generator)
    return fn


def cartpole_envs(num_envs: int, device, *, dt: float = 0.05, m_c: float = 1.0, m_p: float = 0.1, l: float = 0.5, g: float = 9.8,
                  max_steps: int = 500, generator=None, dtype=torch.float32) -> DeviceEnvManager:
    """``CustomCartPoleEnv`` (examples/02_demo_cartpole_AC.py:54-110): thresholds 2.4 m / 12 deg, 500 steps, the
    weights of ``cartpole_reward_fn`` (:43-52) on the next state, start states U(-0.2, 0.25)."""
    return DeviceEnvManager(CartPole(m_c, m_p, l, g), num_envs, device, dt=dt, reset_fn=uniform_reset(-0.2, 0.25, 4, generator),
                            max_steps=max_steps, term_limit=(2.4, math.inf, 12 * 2 * math.pi / 360, math.inf),
                            w_x=(10.0, 10.0, 100.0, 10.0), w_u=(0.01,), reward_on_next=True, dtype=dtype)


def double_integrator_envs(num_envs: int, device, *, dt: float = 0.05, max_steps: int = 250, generator=None,
                           dtype=torch.float32) -> DeviceEnvManager:
    """``DoubleIntegratorEnv`` (examples/01_demo_linear_AC.py:12-51): force clipped to +-5, reward on the state the
    step starts from, episode ends after ``max_steps`` (reported as ``truncated`` here; the example calls it
    ``terminated``), start states U(-2, 2)."""
    return DeviceEnvManager(DoubleIntegrator(dt, semi_implicit=True), num_envs, device, dt=dt,
                            reset_fn=uniform_reset(-2.0, 2.0, 2, generator), max_steps=max_steps, w_x=(1.0, 0.1), w_u=(0.01,),
                            clip_action=True, reward_on_next=False, u_min=(-5.0,), u_max=(5.0,), dtype=dtype)
