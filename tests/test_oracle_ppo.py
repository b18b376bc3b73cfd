"""CPU: the numpy restatement of the PPO glue (oracle/ppo_oracle.py) against the golden vectors recorded from the
unmodified reference (oracle/gen_golden_ppo.py -> tests/golden/ppo_glue.npz).  SURVEY §8f, row N4."""
import math
import os

import numpy as np
import pytest

from oracle import ppo_oracle as po

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ppo_glue.npz"))


@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_gae_bit_exact(tag):
    r, v = G[f"gae_{tag}_rewards"], G[f"gae_{tag}_values"]
    ret, adv = po.gae(r, v, 0.99, 1.0)
    assert np.array_equal(ret, G[f"gae_{tag}_ret_default"]) and np.array_equal(adv, G[f"gae_{tag}_adv_default"])
    ret, adv = po.gae(r, v, 0.99, 0.97)
    assert np.array_equal(ret, G[f"gae_{tag}_ret_097"]) and np.array_equal(adv, G[f"gae_{tag}_adv_097"])
    ret, adv = po.gae(r[:1], v[:1], 0.9, 0.5)
    assert np.array_equal(ret[0], G[f"gae_{tag}_ret_row0"]) and np.array_equal(adv[0], G[f"gae_{tag}_adv_row0"])


@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_mpve_targets_bit_exact(tag):
    out = po.discounted_targets(G[f"mpve_{tag}_rewards"], G[f"mpve_{tag}_vh"], 0.99)
    assert np.array_equal(out, G[f"mpve_{tag}_targets"])


CARTPOLE = dict(term_limit=(2.4, math.inf, 12 * 2 * math.pi / 360, math.inf), w_x=(10.0, 10.0, 100.0, 10.0), w_u=(0.01,),
                reward_on_next=True)
DI = dict(w_x=(1.0, 0.1), w_u=(0.01,), clip_action=True, reward_on_next=False, u_min=(-5.0,), u_max=(5.0,))


def replay_env(prefix, f, max_steps, step_fn, **kw):
    """Drive `step_fn` with the recorded states / actions / reset draws; compare every recorded output."""
    S, A = G[prefix + "state"], G[prefix + "action"]
    steps = np.zeros(S.shape[1], dtype=np.int32)
    for k in range(S.shape[0]):
        obs, nxt, rew, term, trunc, steps = step_fn(f, S[k], A[k], steps, G[prefix + "reset_state"][k], max_steps=max_steps, **kw)
        np.testing.assert_allclose(obs, G[prefix + "obs"][k], rtol=2e-5, atol=2e-6)
        np.testing.assert_allclose(rew, G[prefix + "reward"][k], rtol=2e-5, atol=2e-6)
        done_ref = G[prefix + "terminated"][k] | G[prefix + "truncated"][k]
        assert np.array_equal(term | trunc, done_ref), k
        if prefix.startswith("cartpole"):
            assert np.array_equal(term, G[prefix + "terminated"][k]) and np.array_equal(trunc, G[prefix + "truncated"][k])
        np.testing.assert_allclose(nxt, G[prefix + "next_state"][k], rtol=2e-5, atol=2e-6)


def test_cartpole_env_trace():
    dt, _, _, max_steps, m_c, m_p, l, g = G["cartpole_env_meta"]
    replay_env("cartpole_env_", lambda x, u: po.cartpole_f(x, u, dt, m_c, m_p, l, g), int(max_steps), po.env_step, **CARTPOLE)
    assert G["cartpole_env_terminated"].sum() > 10 and G["cartpole_env_truncated"].sum() > 0


def test_double_integrator_env_trace():
    replay_env("di_env_", lambda x, u: po.double_integrator_f(x, u, 0.05), 6, po.env_step, **DI)


def test_cartpole_reward_fn():
    s, a = G["cartpole_reward_states"], G["cartpole_reward_actions"]
    B, H, _ = s.shape
    zero = np.zeros(B * H, dtype=np.int32)
    _, _, rew, _, _, _ = po.env_step(lambda x, u: x, s.reshape(B * H, 4), a.reshape(B * H, 1), zero, None, w_x=(10.0, 10.0, 100.0, 10.0),
                                      w_u=(0.01,), reward_on_next=False)
    np.testing.assert_allclose(rew.reshape(B, H), G["cartpole_reward_values"], rtol=1e-5, atol=1e-5)
