"""CPU restatement (numpy) of the reference's PPO glue — TEST INFRASTRUCTURE ONLY.

Nothing under tacdmpc_b200/ imports this module; tests/ use it as the checker of the CUDA kernels
(tacd_env_step, tacd_gae, tacd_discounted_targets).  Pinned against the unmodified reference by
tests/golden/ppo_glue.npz (oracle/gen_golden_ppo.py; tests/test_oracle_ppo.py).

Every function works in the dtype of its inputs, one rounding per operation, in the reference's operation order.
"""
from __future__ import annotations

import numpy as np


def gae(rewards, values, gamma=0.99, lam=1.0):
    """compute_gae_and_returns (ACMPC/training_loop.py:30-41) for [N, T] arrays (the vmap of :76, :119)."""
    dt = rewards.dtype.type
    g, gl = dt(gamma), dt(gamma * lam)  # gamma * lam is a Python-float product in the reference, rounded once
    N, T = rewards.shape
    adv = np.zeros_like(rewards)
    last = np.zeros(N, dtype=rewards.dtype)
    vnext = values[:, T - 1]  # values_extended = cat(values, values[-1:])  (:34)
    for t in range(T - 1, -1, -1):
        delta = (rewards[:, t] + g * vnext) - values[:, t]
        last = delta + gl * last
        adv[:, t] = last
        vnext = values[:, t]
    return adv + values, adv


def discounted_targets(rewards, bootstrap, gamma=0.99):
    """The value-expansion recursion of ACMPC/training_loop.py:165-169."""
    dt = rewards.dtype.type
    g = dt(gamma)
    N, H = rewards.shape
    out = np.zeros_like(rewards)
    nxt = np.zeros(N, dtype=rewards.dtype) if bootstrap is None else bootstrap.astype(rewards.dtype)
    for t in range(H - 1, -1, -1):
        nxt = rewards[:, t] + g * nxt
        out[:, t] = nxt
    return out


def cartpole_f(x, u, dt, m_c, m_p, l, g):
    """f_cartpole_torch (examples/02_demo_cartpole_AC.py:25-41; the same formula as 02_demo_cartpole_regulation.py)."""
    R = x.dtype.type
    vel, theta, omega = x[..., 1], x[..., 2], x[..., 3]
    force = u[..., 0]
    s, c = np.sin(theta), np.cos(theta)
    total_mass = R(m_c + m_p)
    m_p_l = R(m_p * l)
    temp = (force + m_p_l * omega ** 2 * s) / total_mass
    th_num = R(g) * s - c * temp
    th_den = R(l) * (R(4.0 / 3.0) - R(m_p) * c ** 2 / total_mass)
    theta_dd = th_num / th_den
    vel_dd = temp - m_p_l * theta_dd * c / total_mass
    xdot = np.stack([vel, vel_dd, omega, theta_dd], axis=-1)
    return x + xdot * R(dt)


def double_integrator_f(x, u, dt):
    """DoubleIntegratorEnv.step (examples/01_demo_linear_AC.py:31-35): semi-implicit Euler."""
    R = x.dtype.type
    new_vel = x[..., 1] + u[..., 0] * R(dt)
    new_pos = x[..., 0] + new_vel * R(dt)
    return np.stack([new_pos, new_vel], axis=-1)


def env_step(f, state, action, steps, reset_state, *, max_steps=0, term_limit=None, w_x=None, w_u=None,
             clip_action=False, reward_on_next=True, u_min=None, u_max=None):
    """The contract of tacd_env_step (include/tacd_ilqr.h): ParallelEnvManager.step (ACMPC/parallel_env.py:34-54) over
    the example environments plus the reset-on-done of training_loop.py:103-107.  Returns
    (obs, next_state, reward, terminated, truncated, steps)."""
    R = state.dtype.type
    B, nx = state.shape
    a = action.copy()
    if clip_action:
        if u_min is not None:
            a = np.maximum(a, np.asarray(u_min, dtype=state.dtype))
        if u_max is not None:
            a = np.minimum(a, np.asarray(u_max, dtype=state.dtype))
    xn = f(state, a)
    term = np.zeros(B, dtype=bool)
    if term_limit is not None:
        term = (np.abs(xn) > np.asarray(term_limit, dtype=state.dtype)).any(axis=1)
    st = steps + 1
    trunc = (st >= max_steps) if max_steps > 0 else np.zeros(B, dtype=bool)
    s = xn if reward_on_next else state
    cost = np.zeros(B, dtype=state.dtype)
    for i in range(nx):
        cost = cost + R(0.0 if w_x is None else w_x[i]) * (s[:, i] * s[:, i])
    acost = np.zeros(B, dtype=state.dtype)
    for j in range(action.shape[1]):
        acost = acost + R(0.0 if w_u is None else w_u[j]) * (action[:, j] * action[:, j])
    reward = -(cost + acost)
    done = (term | trunc) if reset_state is not None else np.zeros(B, dtype=bool)
    nxt = np.where(done[:, None], reset_state if reset_state is not None else xn, xn)
    st = np.where(done, 0, st)
    return xn, nxt, reward, term, trunc, st.astype(np.int32)
