"""Generate tests/golden/ppo_glue.npz by running the UNMODIFIED reference's PPO glue.

Run in the build container only (needs /root/reference):

    python oracle/gen_golden_ppo.py

Recorded, all from the reference's own code on seeded inputs:
  * compute_gae_and_returns (ACMPC/training_loop.py:30-41) under vmap as train() calls it (:119), fp32 and fp64,
    for (gamma, lam) = (0.99, 1.0) — what train() actually passes — and (0.99, 0.97);
  * the value-expansion target recursion of train() (:165-169).  It is inline code there, not a function, so the
    five lines are re-executed here verbatim on random inputs (torch, CPU);
  * CustomCartPoleEnv (examples/02_demo_cartpole_AC.py:54-110) and DoubleIntegratorEnv
    (examples/01_demo_linear_AC.py:12-51) stepped through ParallelEnvManager (ACMPC/parallel_env.py) with the
    reset-on-done of train() (:103-107): per step the states before, the actions, what step() returned, the reset
    states that were drawn and the states the next step starts from.
gym and matplotlib are absent from the image; a minimal stand-in for gym.Env / gym.spaces.Box / gym.register is
installed before the imports (the reference files themselves are untouched).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types
import warnings
from unittest.mock import MagicMock

import numpy as np
import torch

REF = os.environ.get("TACD_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


class _Env:
    """What the example environments use of gym.Env: reset(seed=) seeding self.np_random."""
    np_random = None

    def reset(self, seed=None, options=None):
        if seed is not None or self.np_random is None:
            self.np_random = np.random.default_rng(seed)

    def close(self):
        pass


class _Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.low = np.broadcast_to(np.asarray(low, dtype=dtype), shape if shape is not None else np.shape(low)).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=dtype), shape if shape is not None else np.shape(high)).copy()
        self.shape = self.low.shape
        self.dtype = dtype


gym = types.ModuleType("gym")
gym.Env = _Env
gym.register = lambda **kw: None
gym.make = lambda *a, **k: None
gym.spaces = types.ModuleType("gym.spaces")
gym.spaces.Box = _Box
sys.modules["gym"] = gym
sys.modules["gym.spaces"] = gym.spaces
for _m in ["matplotlib", "matplotlib.pyplot"]:
    sys.modules.setdefault(_m, MagicMock())
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")

from torch.func import vmap  # noqa: E402
from ACMPC.parallel_env import ParallelEnvManager  # noqa: E402
from ACMPC.training_loop import compute_gae_and_returns  # noqa: E402


def _load(name):
    spec = importlib.util.spec_from_file_location("ex_" + name, f"{REF}/examples/{name}.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def gae_cases(out):
    batched = vmap(compute_gae_and_returns, in_dims=(0, 0))  # training_loop.py:76
    g = torch.Generator().manual_seed(11)
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        r = torch.randn(7, 53, generator=g, dtype=dt) * 3
        v = torch.randn(7, 53, generator=g, dtype=dt) * 5
        out[f"gae_{tag}_rewards"], out[f"gae_{tag}_values"] = r.numpy(), v.numpy()
        ret, adv = batched(r, v)  # as train() calls it: the defaults gamma=0.99, lam=1.0
        out[f"gae_{tag}_ret_default"], out[f"gae_{tag}_adv_default"] = ret.numpy(), adv.numpy()
        ret, adv = vmap(lambda a, b: compute_gae_and_returns(a, b, gamma=0.99, lam=0.97), in_dims=(0, 0))(r, v)
        out[f"gae_{tag}_ret_097"], out[f"gae_{tag}_adv_097"] = ret.numpy(), adv.numpy()
        # the un-vmapped [T] signature
        ret1, adv1 = compute_gae_and_returns(r[0], v[0], gamma=0.9, lam=0.5)
        out[f"gae_{tag}_ret_row0"], out[f"gae_{tag}_adv_row0"] = ret1.numpy(), adv1.numpy()


def mpve_cases(out):
    g = torch.Generator().manual_seed(12)
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        pred_rewards = torch.randn(9, 20, generator=g, dtype=dt) * 4
        v_h = torch.randn(9, generator=g, dtype=dt) * 10
        gamma, horizon = 0.99, pred_rewards.shape[1]
        # ---- ACMPC/training_loop.py:165-169, verbatim (actor.horizon -> horizon)
        mpve_targets = torch.zeros_like(pred_rewards)
        next_return = v_h
        for t in reversed(range(horizon)):
            mpve_targets[:, t] = pred_rewards[:, t] + gamma * next_return
            next_return = mpve_targets[:, t]
        # ----
        out[f"mpve_{tag}_rewards"], out[f"mpve_{tag}_vh"], out[f"mpve_{tag}_targets"] = pred_rewards.numpy(), v_h.numpy(), mpve_targets.numpy()


def env_trace(env_fn, num_envs, n_steps, action_scale, seed, tweak=None):
    """ParallelEnvManager + the reset-on-done of training_loop.py:98-108 (the env's global numpy RNG seeded once)."""
    np.random.seed(seed)
    mgr = ParallelEnvManager(env_fn, num_envs, torch.device("cpu"))
    if tweak:
        for e in mgr.envs:
            tweak(e)
    for i, e in enumerate(mgr.envs):
        _Env.reset(e, seed=seed + i)  # seeds self.np_random for envs that draw from it
    g = torch.Generator().manual_seed(seed)
    cur = mgr.reset()
    rec = {k: [] for k in ("state", "action", "obs", "reward", "terminated", "truncated", "reset_state", "next_state")}
    for _ in range(n_steps):
        actions = (torch.randn(num_envs, 1, generator=g) * action_scale).float()
        nxt, rew, term, trunc, _ = mgr.step(actions)
        rec["state"].append(cur.numpy().copy())
        rec["action"].append(actions.numpy().copy())
        rec["obs"].append(nxt.numpy().copy())
        rec["reward"].append(rew.numpy().copy())
        rec["terminated"].append(term.numpy().copy())
        rec["truncated"].append(trunc.numpy().copy())
        resets = np.zeros_like(nxt.numpy())
        for i in range(num_envs):
            if term[i] or trunc[i]:
                reset_obs, _ = mgr.envs[i].reset()
                resets[i] = reset_obs
                nxt[i] = torch.from_numpy(reset_obs).float()
        rec["reset_state"].append(resets)
        rec["next_state"].append(nxt.numpy().copy())
        cur = nxt
    return {k: np.stack(v) for k, v in rec.items()}


def env_cases(out):
    ex2 = _load("02_demo_cartpole_AC")
    params = ex2.CartPoleParams(m_c=1.0, m_p=0.1, l=0.5, g=9.8)

    def short(e):
        e.max_steps = 9  # so truncation shows up inside a short trace

    tr = env_trace(lambda: ex2.CustomCartPoleEnv(params, dt=0.05), 6, 40, 25.0, 21, tweak=short)
    for k, v in tr.items():
        out["cartpole_env_" + k] = v
    out["cartpole_env_meta"] = np.array([0.05, 2.4, 12 * 2 * np.pi / 360, 9, 1.0, 0.1, 0.5, 9.8])
    # cartpole_reward_fn on a batch of predicted observations (what the MPVE block feeds it, training_loop.py:160)
    g = torch.Generator().manual_seed(22)
    s, a = torch.randn(5, 20, 4, generator=g), torch.randn(5, 20, 1, generator=g)
    out["cartpole_reward_states"], out["cartpole_reward_actions"] = s.numpy(), a.numpy()
    out["cartpole_reward_values"] = vmap(ex2.cartpole_reward_fn)(s, a).numpy()

    ex1 = _load("01_demo_linear_AC")

    def short_di(e):
        e.max_steps = 6

    tr = env_trace(lambda: ex1.DoubleIntegratorEnv(dt=0.05), 5, 20, 4.0, 31, tweak=short_di)
    for k, v in tr.items():
        out["di_env_" + k] = v


def main():
    out = {}
    gae_cases(out)
    mpve_cases(out)
    env_cases(out)
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "ppo_glue.npz"), **out)
    print("wrote", os.path.join(OUT, "ppo_glue.npz"), {k: v.shape for k, v in out.items() if k.endswith("state") or "adv" in k})


if __name__ == "__main__":
    main()
