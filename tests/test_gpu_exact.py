"""-m gpu tests of compat="exact" (SURVEY §8f row N3): exact implicit differentiation in the backward kernel —
terminal value from C_final / grad_X[T] (fixes Q1, Q2), grad_x0 = -lambda_0 (Q3), frozen active inputs in the
gains (Q6), grad_ref = C dtau — checked against tests/exact_ref.py (pinned by finite differences in
tests/test_oracle_exact.py) and by torch.autograd.gradcheck, which is what the reference's tests/test_gradcheck.py
asks for and the reference itself fails."""
import numpy as np
import pytest
import torch

import parity as P
from exact_ref import exact_backward

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    import gpu_common
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return gpu_common


@pytest.mark.parametrize("name", ["di_f64", "cartpole_shared_f64", "cartpole_actor_f64", "unicycle_tight_f64", "unicycle_loose_f64",
                                  "lti8_f64", "lti8_T50_f64", "lti8_mixed_f64"])   # lti8_*: the generic (run-time dimension) kernel
def test_exact_kernel_matches_numpy_restatement(G, name):
    from tacdmpc_b200 import ops
    g = P.load(name)
    spec = P.spec_of(g)
    s = G.gpu_spec(spec, g["x0"].dtype)
    d = G.dev
    B, nx, nu = g["x0"].shape[0], spec["nx"], spec["nu"]
    X, U, gX, gU = g["X"], g["U"], g["wX"], g["wU"]
    want = ("grad_C", "grad_c", "grad_Cf", "grad_cf", "grad_x0", "grad_xref", "grad_uref", "dX", "dU")
    out = ops.ilqr_backward(s, d(X), d(U), d(g["C"]), d(g["xref"]), d(g["uref"]), d(g["tight"]), d(gX), d(gU),
                            Cf_batched=g["Cf"].ndim == 3, want=want, exact=True, Cf=d(g["Cf"]))
    torch.cuda.synchronize()
    Cb = g["C"] if g["C"].ndim == 4 else np.broadcast_to(g["C"], (B,) + g["C"].shape)
    Cfb = g["Cf"] if g["Cf"].ndim == 3 else np.broadcast_to(g["Cf"], (B,) + g["Cf"].shape)
    xr = np.broadcast_to(g["xref"], (B,) + g["xref"].shape[1:])
    ur = np.broadcast_to(g["uref"], (B,) + g["uref"].shape[1:])
    ref = exact_backward(nx, nu, X, U, Cb, g["c"], Cfb, g["cf"], xr, ur, g["tight"], gX, gU, g["A"], g["Bm"], spec.get("reg_eps", 1e-6))
    for nm in want:
        got = out[nm].cpu().numpy().astype(np.float64)
        exp = ref[nm]
        if got.shape != exp.shape:            # shared inputs: the kernel returns the sum over the batch
            exp = exp.sum(0).reshape(got.shape) if got.size * B == exp.size else exp.reshape(got.shape)
        scale = max(np.abs(exp).max(), 1.0)
        # T = 50 on the (unstable) random LTI system: the value function grows to 1e6 and the two LU orders part at 1e-7
        tol = 1e-6 if name == "lti8_T50_f64" else 1e-9
        assert np.abs(got - exp).max() / scale <= tol, (nm, np.abs(got - exp).max() / scale)
    assert float(out["grad_x0"].abs().max()) > 0     # no longer identically zero (quirk Q3)


def test_exact_compact_diagonal_equals_dense(G):
    from tacdmpc_b200 import ops
    g = P.load("cartpole_actor_f64")
    s = G.gpu_spec(P.spec_of(g), g["x0"].dtype)
    d = G.dev
    dg = lambda a: np.ascontiguousarray(np.diagonal(a, axis1=-2, axis2=-1))
    args = (d(g["X"]), d(g["U"]))
    rest = (d(g["xref"]), d(g["uref"]), d(g["tight"]), d(g["wX"]), d(g["wU"]))
    a = ops.ilqr_backward(s, *args, d(dg(g["C"])), *rest, C_diag=True, exact=True, Cf=d(dg(g["Cf"])))
    b = ops.ilqr_backward(s, *args, d(g["C"]), *rest, Cf_batched=True, exact=True, Cf=d(g["Cf"]))
    assert torch.equal(a["grad_C"], torch.diagonal(b["grad_C"], dim1=-2, dim2=-1))
    assert torch.equal(a["grad_Cf"], torch.diagonal(b["grad_Cf"], dim1=-2, dim2=-1))
    for k in ("grad_c", "grad_cf", "grad_x0", "grad_xref", "grad_uref"):
        assert torch.equal(a[k], b[k]), k


@pytest.mark.parametrize("native", [True, False])
def test_reference_gradcheck_passes_in_exact_mode(G, native):
    """tests/test_gradcheck.py:7-51 as written (1-D integrator, H=4, all-zero cost, gradcheck of X*.sum() w.r.t. x0):
    fails at reference HEAD (grad_x0 = 0, quirk Q3), passes with compat='exact' — with the shipped LTI plugin and
    with the test's own Python callable (generic path, A/B from autograd)."""
    from tacdmpc_b200.DifferentialMPC import DifferentiableMPCController, GeneralQuadCost, GradMethod
    from tacdmpc_b200.dynamics import LTI
    torch.manual_seed(0)
    nx = nu = 1
    H, dt = 4, 0.1
    z = lambda *s: torch.zeros(*s, device="cuda", dtype=torch.double)
    cost = GeneralQuadCost(nx, nu, C=z(H, 2, 2), c=z(H, 2), C_final=z(2, 2), c_final=z(2), device="cuda")
    f = LTI(torch.tensor([[1.0]], dtype=torch.double), torch.tensor([[dt]], dtype=torch.double)) if native else (lambda x, u, dt_: x + dt_ * u)
    mpc = DifferentiableMPCController(f_dyn=f, total_time=H * dt, step_size=dt, horizon=H, cost_module=cost,
                                      grad_method=GradMethod.AUTO_DIFF, device="cuda", compat="exact")
    x0 = torch.randn(1, nx, device="cuda", dtype=torch.double, requires_grad=True)

    def func(x):
        X_opt, _ = mpc.forward(x)
        return X_opt.sum()

    assert torch.autograd.gradcheck(func, (x0,), eps=1e-6, atol=1e-4, rtol=1e-3)


def test_gradcheck_linear_quadratic_all_inputs(G):
    """Double integrator (examples/01_demo_linear.py) with a real cost: gradcheck w.r.t. x0, c, c_final, x_ref, u_ref
    and the diagonal of C.  One Newton step (max_iter=1) is the exact optimum of an LQ problem, so the forward is
    smooth in its inputs (free-running iLQR adds 1e-8 jitter from its tolerance-free stopping rule)."""
    from tacdmpc_b200.DifferentialMPC import DifferentiableMPCController, GeneralQuadCost
    from tacdmpc_b200.dynamics import DoubleIntegrator
    torch.manual_seed(1)
    nx, nu, T, dt, B = 2, 1, 6, 0.1, 2
    n = nx + nu
    dd = dict(device="cuda", dtype=torch.double)
    qd = torch.tensor([3.0, 0.5, 0.2], **dd)
    f = DoubleIntegrator(dt)

    def solve(x0, qdiag, c, cf, xref, uref):
        C = torch.diag_embed(qdiag).expand(T, n, n).contiguous()
        Cf = torch.diag_embed(torch.cat([5 * qdiag[:nx], torch.zeros(nu, **dd)]))
        cm = GeneralQuadCost(nx, nu, C, c, Cf, cf, device="cuda", x_ref=xref, u_ref=uref)
        mpc = DifferentiableMPCController(f, T * dt, dt, T, cm, grad_method="analytic", f_dyn_jac=f.jac, device="cuda",
                                          compat="exact", max_iter=1, reg_eps=1e-10)   # (a regularised step is not the optimum)
        X, U = mpc(x0)
        w = torch.linspace(0.5, 1.5, T + 1, **dd).view(1, -1, 1)
        return (w * X).sum() + (U ** 2).sum() + U[:, 0].sum()

    ins = (torch.randn(B, nx, **dd), qd.clone(), 0.1 * torch.randn(T, n, **dd), 0.1 * torch.randn(n, **dd) * torch.tensor([1.0, 1.0, 0.0], **dd),
           0.2 * torch.randn(B, T + 1, nx, **dd), 0.1 * torch.randn(B, T, nu, **dd))
    ins = tuple(t.requires_grad_(True) for t in ins)
    assert torch.autograd.gradcheck(solve, ins, eps=1e-6, atol=1e-6, rtol=1e-5, nondet_tol=1e-9)
