"""-m gpu parity tests of the PPO glue kernels (tacd_env_step, tacd_gae, tacd_discounted_targets):
CUDA against the golden vectors recorded from the unmodified reference (tests/golden/ppo_glue.npz), against the numpy
oracle at PPO-sized inputs, and through the host mirrors (DeviceEnvManager, compute_gae_and_returns, mpve_targets,
collect_rollout)."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import ppo_oracle as po
from test_oracle_ppo import CARTPOLE, DI, G, replay_env

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_gae_matches_reference_bit_for_bit(tag):
    from tacdmpc_b200.ACMPC import compute_gae_and_returns
    r, v = cu(G[f"gae_{tag}_rewards"]), cu(G[f"gae_{tag}_values"])
    ret, adv = compute_gae_and_returns(r, v)                        # train()'s call: gamma 0.99, lam 1.0
    assert np.array_equal(ret.cpu().numpy(), G[f"gae_{tag}_ret_default"]) and np.array_equal(adv.cpu().numpy(), G[f"gae_{tag}_adv_default"])
    ret, adv = compute_gae_and_returns(r, v, gamma=0.99, lam=0.97)
    assert np.array_equal(ret.cpu().numpy(), G[f"gae_{tag}_ret_097"]) and np.array_equal(adv.cpu().numpy(), G[f"gae_{tag}_adv_097"])
    ret, adv = compute_gae_and_returns(r[0], v[0], gamma=0.9, lam=0.5)   # the reference's own [T] signature
    assert np.array_equal(ret.cpu().numpy(), G[f"gae_{tag}_ret_row0"]) and np.array_equal(adv.cpu().numpy(), G[f"gae_{tag}_adv_row0"])


@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_mpve_targets_match_reference_bit_for_bit(tag):
    from tacdmpc_b200.ACMPC import mpve_targets
    out = mpve_targets(cu(G[f"mpve_{tag}_rewards"]), cu(G[f"mpve_{tag}_vh"]), 0.99)
    assert np.array_equal(out.cpu().numpy(), G[f"mpve_{tag}_targets"])


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_gae_and_targets_ppo_sized_vs_oracle(dtype):
    """4096 environments x 250 steps (train()'s buffers at SURVEY config 4's batch), bit-exact against the oracle;
    lam = 1 with zero values degenerates to the discounted-return recursion."""
    from tacdmpc_b200 import ops
    rng = np.random.default_rng(5)
    r = rng.standard_normal((4096, 250)).astype(dtype)
    v = rng.standard_normal((4096, 250)).astype(dtype) * 3
    ret, adv = ops.gae(cu(r), cu(v), 0.99, 0.95)
    ret_o, adv_o = po.gae(r, v, 0.99, 0.95)
    assert np.array_equal(ret.cpu().numpy(), ret_o) and np.array_equal(adv.cpu().numpy(), adv_o)
    boot = rng.standard_normal(4096).astype(dtype)
    tg = ops.discounted_targets(cu(r), cu(boot), 0.97)
    assert np.array_equal(tg.cpu().numpy(), po.discounted_targets(r, boot, 0.97))
    z = np.zeros_like(r)
    _, adv0 = ops.gae(cu(r), cu(z), 0.97, 1.0)
    assert np.array_equal(adv0.cpu().numpy(), ops.discounted_targets(cu(r), None, 0.97).cpu().numpy())
    # ragged / degenerate shapes
    assert ops.gae(cu(r[:0]), cu(v[:0]))[0].shape == (0, 250)
    r1, v1 = r[:3, :1], v[:3, :1]
    ret1, adv1 = ops.gae(cu(r1), cu(v1), 0.9, 0.9)
    assert np.array_equal(adv1.cpu().numpy(), po.gae(r1, v1, 0.9, 0.9)[1])


def _gpu_step(spec):
    from tacdmpc_b200 import ops

    def step(f, state, action, steps, reset_state, *, max_steps, term_limit=None, w_x=None, w_u=None, clip_action=False,
             reward_on_next=True, u_min=None, u_max=None):
        st = cu(steps.astype(np.int32))
        obs, nxt, rew, term, trunc = ops.env_step(spec, cu(state), cu(action), steps=st, reset_state=None if reset_state is None else cu(reset_state),
                                                  max_steps=max_steps, term_limit=term_limit, w_x=w_x, w_u=w_u, clip_action=clip_action,
                                                  reward_on_next=reward_on_next)
        return (obs.cpu().numpy(), nxt.cpu().numpy(), rew.cpu().numpy(), term.cpu().numpy(), trunc.cpu().numpy(), st.cpu().numpy())
    return step


def test_cartpole_env_trace_matches_reference():
    from tacdmpc_b200 import _abi, ops
    dt, _, _, max_steps, m_c, m_p, l, g = G["cartpole_env_meta"]
    spec = ops.Spec(nx=4, nu=1, T=1, dt=float(dt), dyn_id=_abi.DYN_CARTPOLE, dyn_params=(m_c, m_p, l, g))
    replay_env("cartpole_env_", None, int(max_steps), _gpu_step(spec), **CARTPOLE)


def test_double_integrator_env_trace_matches_reference():
    from tacdmpc_b200 import ops
    from tacdmpc_b200.dynamics import DoubleIntegrator
    f = DoubleIntegrator(0.05, semi_implicit=True)
    A, B = f.device_matrices(torch.device("cuda"), torch.float32)
    spec = ops.Spec(nx=2, nu=1, T=1, dt=0.05, dyn_id=f.dyn_id, dyn_A=A, dyn_B=B, u_min=(-5.0,), u_max=(5.0,))
    kw = {k: v for k, v in DI.items() if k not in ("u_min", "u_max")}
    replay_env("di_env_", None, 6, _gpu_step(spec), **kw)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_env_step_4096_envs_vs_oracle(dtype):
    from tacdmpc_b200 import _abi, ops
    rng = np.random.default_rng(9)
    B = 4096
    x = rng.uniform(-0.3, 0.3, (B, 4)).astype(dtype)
    x[:, 0] *= 7                                   # a good share beyond the 2.4 m threshold
    a = (rng.standard_normal((B, 1)) * 20).astype(dtype)
    steps = rng.integers(0, 12, B).astype(np.int32)
    resets = rng.uniform(-0.2, 0.25, (B, 4)).astype(dtype)
    spec = ops.Spec(nx=4, nu=1, T=1, dt=0.05, dyn_id=_abi.DYN_CARTPOLE, dyn_params=(1.0, 0.1, 0.5, 9.8))
    got = _gpu_step(spec)(None, x, a, steps, resets, max_steps=10, **CARTPOLE)
    exp = po.env_step(lambda s, u: po.cartpole_f(s, u, 0.05, 1.0, 0.1, 0.5, 9.8), x, a, steps, resets, max_steps=10, **CARTPOLE)
    tol = dict(rtol=2e-5, atol=2e-6) if dtype == np.float32 else dict(rtol=1e-12, atol=1e-13)
    margin = np.abs(np.abs(exp[0][:, [0, 2]]) - np.array([2.4, 12 * 2 * math.pi / 360])).min(axis=1) > 1e-5   # not at a threshold
    assert margin.mean() > 0.99
    np.testing.assert_allclose(got[0], exp[0], **tol)
    np.testing.assert_allclose(got[2], exp[2], rtol=tol["rtol"] * 5, atol=1e-4 if dtype == np.float32 else 1e-10)
    assert np.array_equal(got[3][margin], exp[3][margin]) and np.array_equal(got[4], exp[4])
    np.testing.assert_allclose(got[1][margin], exp[1][margin], **tol)
    assert np.array_equal(got[5][margin], exp[5][margin])
    assert 0.2 < exp[3].mean() < 0.9 and exp[4].sum() > 100


def test_device_env_manager_and_rollout():
    from tacdmpc_b200.ACMPC import cartpole_envs, collect_rollout, compute_gae_and_returns
    gen = torch.Generator(device="cuda").manual_seed(3)
    envs = cartpole_envs(256, "cuda", max_steps=7, generator=gen)
    obs = envs.reset()
    assert obs.shape == (256, 4) and float(obs.min()) >= -0.2 and float(obs.max()) <= 0.25
    assert envs.single_observation_space_shape == (4,) and envs.single_action_space_shape == (1,)
    # step(): the reference's tuple; no reset happens behind the caller's back
    nxt, rew, term, trunc, infos = envs.step(torch.zeros(256, 1, device="cuda"))
    assert nxt.shape == (256, 4) and rew.shape == (256,) and term.dtype == torch.bool and len(infos) == 256
    assert torch.equal(envs.state, nxt) and int(envs.steps.max()) == 1
    # a dummy policy with ActorMPC.forward's return signature
    def actor(s, deterministic=False):
        a = -20.0 * s[:, 2:3] - 2.0 * s[:, 3:4]
        return a, torch.zeros(s.shape[0], device=s.device), s[:, None, :].expand(-1, 6, -1), a[:, None, :].expand(-1, 5, -1)
    buf = collect_rollout(actor, envs, 20)
    assert buf["states"].shape == (256, 20, 4) and buf["rewards"].shape == (256, 20) and buf["pred_states"].shape == (256, 20, 6, 4)
    assert int(envs.steps.max()) < 7                      # every environment was restarted at 7 steps at the latest
    # without a termination test only truncation ends an episode: the states at t = 7 and 14 are fresh draws
    envs.term_limit = None
    buf = collect_rollout(actor, envs, 20)
    for t in (7, 14):
        assert float(buf["states"][:, t].min()) >= -0.2 and float(buf["states"][:, t].max()) <= 0.25
    assert float(buf["states"][:, 6].abs().max()) > 0.25
    assert torch.equal(envs.steps, torch.full_like(envs.steps, 20 - 14))
    ret, adv = compute_gae_and_returns(buf["rewards"], torch.zeros_like(buf["rewards"]))
    assert torch.isfinite(ret).all() and (adv <= 0).all()   # rewards are non-positive


def test_graphed_policy_replays_on_new_observations():
    """ActorMPC.graphed: cost network + whole solve + sampling as one CUDA-graph replay; every replay must be a valid
    policy call on ITS observations (the plan starts at the observation, is the rollout of its own inputs) and agree with
    the eager call up to what the random 0.01-sigma warm start moves the solution."""
    from tacdmpc_b200 import ops
    from tacdmpc_b200.ACMPC import ActorMPC, cartpole_envs, collect_rollout
    from tacdmpc_b200.dynamics import CartPole
    torch.manual_seed(0)
    f = CartPole()
    actor = ActorMPC(4, 1, 12, 0.05, f_dyn=f, f_dyn_jac=f.jac, device="cuda", grad_method="analytic")
    pol = actor.graphed(64, deterministic=True)
    spec = actor.mpc._last_spec
    for seed in (1, 2, 3):
        obs = torch.rand(64, 4, device="cuda", generator=torch.Generator(device="cuda").manual_seed(seed)) * 0.45 - 0.2
        a, lp, Xp, Up = pol(obs)
        assert a.shape == (64, 1) and lp.shape == (64,) and Xp.shape == (64, 13, 4) and Up.shape == (64, 12, 1)
        assert torch.equal(Xp[:, 0], obs) and torch.equal(a, Up[:, 0])
        assert float((ops.rollout(spec, obs, Up) - Xp).abs().max()) < 1e-5
        with torch.no_grad():
            ae, _, _, Ue = actor(obs, deterministic=True)
        assert float((Ue - Up).abs().max()) < 5e-2 * max(1.0, float(Ue.abs().max()))
    with pytest.raises(RuntimeError, match="captured"):
        pol(torch.zeros(3, 4, device="cuda"))
    # an optimiser step between replays is seen (weights are read at replay time)
    with torch.no_grad():
        for p_ in actor.cost_map_net.parameters():
            p_.add_(0.05 * torch.randn_like(p_))
    a2, _, _, _ = pol(obs)
    assert not torch.equal(a2, a)
    # the rollout phase with the graphed policy
    envs = cartpole_envs(64, "cuda", max_steps=50)
    buf = collect_rollout(actor.graphed(64), envs, 10)
    assert buf["states"].shape == (64, 10, 4) and torch.isfinite(buf["rewards"]).all() and buf["pred_actions"].shape == (64, 10, 12, 1)


def test_graphed_update_matches_eager_update():
    """ActorMPC.graphed_update: evaluate_actions + PPO-style loss + backward as one CUDA-graph replay gives the eager
    step's loss and parameter gradients on new mini-batches, also after an optimiser step and a zero_grad(None)."""
    from tacdmpc_b200.ACMPC import ActorMPC
    from tacdmpc_b200.dynamics import CartPole
    torch.manual_seed(0)
    f = CartPole()
    actor = ActorMPC(4, 1, 10, 0.05, f_dyn=f, f_dyn_jac=f.jac, device="cuda", grad_method="analytic")
    opt = torch.optim.SGD(actor.parameters(), lr=1e-3)
    gen = torch.Generator(device="cuda").manual_seed(5)
    mk = lambda: (torch.rand(48, 4, device="cuda", generator=gen) * 0.45 - 0.2, torch.randn(48, 1, device="cuda", generator=gen),
                  torch.randn(48, device="cuda", generator=gen), torch.randn(48, device="cuda", generator=gen) - 1.0)

    def loss_fn(log_probs, entropy, adv, old_lp):                 # training_loop.py:142-147
        ratio = torch.exp(log_probs - old_lp)
        return -torch.min(ratio * adv, torch.clamp(ratio, 0.8, 1.2) * adv).mean() - 0.01 * entropy.mean()

    upd = actor.graphed_update(loss_fn, *mk())                 # (captured before any eager backward, see GraphedUpdate)
    for it in range(3):
        obs, act, adv, old = mk()
        opt.zero_grad(set_to_none=True)
        lp, ent = actor.evaluate_actions(obs, act)
        le = loss_fn(lp, ent, adv, old)
        le.backward()
        ge = [p.grad.clone() for p in actor.parameters()]
        opt.zero_grad(set_to_none=True)
        lg = upd(obs, act, adv, old)
        gg = [p.grad for p in actor.parameters()]
        assert abs(float(lg) - float(le.detach())) <= 1e-5 * max(1.0, abs(float(le.detach())))
        for a, b in zip(gg, ge):
            assert float((a - b).abs().max()) <= 1e-4 * max(1e-3, float(b.abs().max()))
        opt.step()                                              # weights change between replays
    with pytest.raises(RuntimeError, match="captured"):
        upd(*[t[:3] for t in mk()])


def test_ppo_glue_rejects_cpu_tensors():
    from tacdmpc_b200 import _lib, ops
    with pytest.raises(_lib.TacdError, match="CUDA"):
        ops.gae(torch.zeros(2, 3), torch.zeros(2, 3))
    with pytest.raises(RuntimeError):
        ops.gae(torch.zeros(2, 3, device="cuda"), torch.zeros(2, 4, device="cuda"))
