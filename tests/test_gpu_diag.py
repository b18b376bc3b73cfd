"""-m gpu tests of the compact diagonal cost path (tacd_problem.C_diag, DiagQuadCost,
tacdmpc_b200.ACMPC.ActorMPC).  The golden case ``cartpole_actor_*`` was recorded from the unmodified reference
with the dense ``diag_embed`` tensors ACMPC/actor.py builds; the compact path must reproduce it to the same
tolerance as the dense path, and agree with the dense path of this library to rounding."""
import numpy as np
import pytest
import torch

import parity as P

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    import gpu_common
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return gpu_common


def _diag(a):
    return np.ascontiguousarray(np.diagonal(a, axis1=-2, axis2=-1))


def _forward_diag(G, g, replay):
    from tacdmpc_b200 import ops
    spec = P.spec_of(g)
    s = G.gpu_spec(spec, g["x0"].dtype)
    d = G.dev
    ls = P.ls_cost_of(g)
    ls = None if ls is None else (d(_diag(ls[0])), d(ls[1]), d(_diag(ls[2])), d(ls[3]))
    rp = dict(iters=d(g["iters"]), alpha_trace=d(g["alpha_trace"]), flags=d(g["flags"])) if replay else None
    out = ops.ilqr_forward(s, d(g["x0"]), d(_diag(g["C"])), d(g["c"]), d(_diag(g["Cf"])), d(g["cf"]), d(g["xref"]),
                           d(g["uref"]), d(g["Uinit"]), ls_cost=ls, replay=rp, want_AB=True, C_diag=True)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items()}


@pytest.mark.parametrize("name", ["cartpole_actor_f32", "cartpole_actor_f64"])
def test_compact_forward_matches_reference_and_dense_path(G, name):
    g = P.load(name)
    off = g["C"].copy()
    idx = np.arange(off.shape[-1])
    off[..., idx, idx] = 0
    assert np.abs(off).max() == 0, "the golden ActorMPC cost is diagonal"
    out = _forward_diag(G, g, replay=True)
    assert np.array_equal(out["iters"], g["iters"])
    tx, tu = P.elem_tol(g, "X"), P.elem_tol(g, "U")
    rep = P.reproducible(g)
    assert (P.rel_err(out["X"], g["X"])[rep] <= tx[rep]).all() and (P.rel_err(out["U"], g["U"])[rep] <= tu[rep]).all()
    assert (P.rel_err(out["cost"][:, None], g["cost"][:, None])[rep] <= 10 * np.maximum(tx, tu)[rep]).all()
    dense = G.forward(P.spec_of(g), g, replay=True)
    # same kernel, same arithmetic on the non-zero entries: identical results
    for k in ("X", "U", "cost", "A", "Bm"):
        assert np.array_equal(out[k], dense[k]), k
    assert np.array_equal(out["tight"], dense["tight"]) and np.array_equal(out["flags"], dense["flags"])
    # free-running: identical decisions as the dense path
    free, dfree = _forward_diag(G, g, replay=False), G.forward(P.spec_of(g), g)
    assert np.array_equal(free["iters"], dfree["iters"]) and np.array_equal(free["alpha_trace"], dfree["alpha_trace"])
    assert np.array_equal(free["U"], dfree["U"])


@pytest.mark.parametrize("name", ["cartpole_actor_f32", "cartpole_actor_f64"])
@pytest.mark.parametrize("loss", ["U0sum", "weighted"])
def test_compact_backward_matches_reference(G, name, loss):
    from tacdmpc_b200 import ops
    g = P.load(name)
    s = G.gpu_spec(P.spec_of(g), g["x0"].dtype)
    d = G.dev
    X, U = g["X"], g["U"]
    if loss == "U0sum":
        gX, gU = np.zeros_like(X), np.zeros_like(U)
        gU[:, 0] = 1
    else:
        gX, gU = g["wX"], g["wU"]
    want = ("grad_C", "grad_c", "grad_Cf", "grad_cf", "grad_x0", "grad_xref", "grad_uref")
    out = ops.ilqr_backward(s, d(X), d(U), d(_diag(g["C"])), d(g["xref"]), d(g["uref"]), d(g["tight"]), d(gX), d(gU),
                            want=want, C_diag=True)
    dense = ops.ilqr_backward(s, d(X), d(U), d(g["C"]), d(g["xref"]), d(g["uref"]), d(g["tight"]), d(gX), d(gU),
                              Cf_batched=True, want=want)
    torch.cuda.synchronize()
    tol = P.case_tol(g)
    for nm in ("C", "c", "Cf", "cf", "x0", "xref", "uref"):
        got = out[f"grad_{nm}"].cpu().numpy()
        ref = dense[f"grad_{nm}"].cpu().numpy()
        gold = g[f"grad_{loss}_{nm}"]
        if nm in ("C", "Cf"):
            ref, gold = _diag(ref), _diag(gold)
        assert np.array_equal(got, ref.reshape(got.shape)), nm            # vs the dense path of this library
        scale = max(np.abs(gold).max(), 1.0)
        floor = 0.0
        if f"grad64_{loss}_{nm}" in g:
            g64 = g[f"grad64_{loss}_{nm}"]
            floor = 4 * np.abs((_diag(g64) if nm in ("C", "Cf") else g64) - gold).max() / scale
        if floor > 4 * P.REPRO_LIMIT:
            continue
        assert np.abs(got.astype(np.float64) - gold.reshape(got.shape)).max() / scale <= 50 * tol + floor, nm


def test_unicycle_compact_equals_dense(G):
    """n = 5 with two inputs (2x2 Cholesky gains, box constraints): compact vs dense on synthetic diagonal costs."""
    from tacdmpc_b200 import _abi, ops
    rng = np.random.default_rng(0)
    B, T, nx, nu = 96, 30, 3, 2
    n = nx + nu
    q = (0.5 + rng.random((B, T, n))).astype(np.float32) * np.array([20, 20, 5, 0.1, 0.1], np.float32)
    p = (0.1 * rng.standard_normal((B, T, n))).astype(np.float32)
    qf, pf = 10 * q[:, -1], p[:, -1].copy()
    x0 = rng.standard_normal((B, nx)).astype(np.float32)
    xref = np.cumsum(0.1 * rng.standard_normal((B, T + 1, nx)), axis=1).astype(np.float32)
    uref = np.zeros((B, T, nu), np.float32)
    U0 = np.zeros((B, T, nu), np.float32)
    s = ops.Spec(nx=nx, nu=nu, T=T, dt=0.1, dyn_id=_abi.DYN_UNICYCLE, u_min=[-2.0, -1.0], u_max=[2.0, 1.0])
    d = G.dev
    dg = lambda a: np.ascontiguousarray(np.einsum("...i,ij->...ij", a, np.eye(n, dtype=a.dtype)))
    a = ops.ilqr_forward(s, d(x0), d(q), d(p), d(qf), d(pf), d(xref), d(uref), d(U0),
                         ls_cost=(d(q[0]), d(p[0]), d(qf[0]), d(pf[0])), C_diag=True)
    b = ops.ilqr_forward(s, d(x0), d(dg(q)), d(p), d(dg(qf)), d(pf), d(xref), d(uref), d(U0),
                         ls_cost=(d(dg(q[0])), d(p[0]), d(dg(qf[0])), d(pf[0])))
    for k in ("X", "U", "iters", "alpha_trace", "tight", "cost"):
        assert torch.equal(a[k], b[k]), k
    assert int(a["tight"].sum()) > 0 and int(a["iters"].max()) > 2
    gU = torch.ones_like(a["U"])
    gX = torch.zeros_like(a["X"])
    ga = ops.ilqr_backward(s, a["X"], a["U"], d(q), d(xref), d(uref), a["tight"], gX, gU, C_diag=True)
    gb = ops.ilqr_backward(s, a["X"], a["U"], d(dg(q)), d(xref), d(uref), a["tight"], gX, gU, Cf_batched=True)
    assert torch.equal(ga["grad_C"], torch.diagonal(gb["grad_C"], dim1=-2, dim2=-1))
    assert torch.equal(ga["grad_Cf"], torch.diagonal(gb["grad_Cf"], dim1=-2, dim2=-1))
    for k in ("grad_c", "grad_cf", "grad_xref", "grad_uref"):
        assert torch.equal(ga[k], gb[k]), k


def test_actor_mpc_fused_equals_dense_formulation(G):
    """tacdmpc_b200.ACMPC.ActorMPC (compact q_diag, p end to end) against the reference's dense formulation run
    through the same kernels: same actions, same log-probs, same parameter gradients."""
    from tacdmpc_b200.ACMPC import ActorMPC
    from tacdmpc_b200.dynamics import CartPole
    f = CartPole()
    torch.manual_seed(0)
    fused = ActorMPC(4, 1, 15, 0.05, f, device="cuda", grad_method="auto_diff", fused=True)
    dense = ActorMPC(4, 1, 15, 0.05, f, device="cuda", grad_method="auto_diff", fused=False)
    dense.load_state_dict(fused.state_dict())
    x = torch.rand(256, 4, device="cuda") * 0.45 - 0.2
    actions = torch.zeros(256, 1, device="cuda")
    outs = []
    for m in (fused, dense):
        lp, ent = m.evaluate_actions(x, actions)
        (-(lp.mean()) - 0.01 * ent.mean()).backward()
        outs.append((lp.detach(), [p.grad.clone() for p in m.parameters()], m.mpc.iters_last.clone()))
    assert torch.equal(outs[0][2], outs[1][2])
    assert torch.equal(outs[0][0], outs[1][0])
    for a, b in zip(outs[0][1], outs[1][1]):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6 * float(b.abs().max()) + 1e-12)
    assert sum(float(g.abs().sum()) for g in outs[0][1]) > 0
    # forward() (sampling path): shapes and the reference's return signature
    action, log_prob, Xp, Up = fused(x[:8], deterministic=True)
    assert action.shape == (8, 1) and log_prob.shape == (8,) and Xp.shape == (8, 16, 4) and Up.shape == (8, 15, 1)
