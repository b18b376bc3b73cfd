"""-m gpu regression tests for the round-1 advisor findings: each fails on the round-1 host code.

 * DiagQuadCost / ActorMPC(fused=True) with a problem the compact-diagonal kernels do not take ((nx, nu) without an
   instance, horizon beyond shared memory) must solve on the dense expansion instead of raising;
 * simulate(): u_ref_full of any admissible length means the same thing on the device loop and on the step-by-step loop;
   warm_start=False uses U_init for the first solve only on both;
 * a ConstrainedQuadCost is constrained with and without requires_grad;
 * a cost with a leading batch of 1 and B > 1 back-propagates (the gradient is summed down to the input's shape).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def _lti(nx, nu, seed=0):
    from tacdmpc_b200.dynamics import LTI
    g = torch.Generator().manual_seed(seed)
    A = torch.eye(nx, dtype=torch.float64) + 0.05 * torch.randn(nx, nx, generator=g, dtype=torch.float64)
    B = 0.3 * torch.randn(nx, nu, generator=g, dtype=torch.float64)
    return LTI(A, B)


@pytest.mark.parametrize("nx,nu,T", [(2, 2, 10), (3, 1, 10), (4, 1, 600)])
def test_diag_cost_falls_back_to_dense_where_the_compact_kernels_do_not_apply(nx, nu, T):
    from tacdmpc_b200.DifferentialMPC import DiagQuadCost, DifferentiableMPCController, GeneralQuadCost
    from tacdmpc_b200.dynamics import CartPole
    f = CartPole() if (nx, nu) == (4, 1) else _lti(nx, nu)
    B, n, dt = 5, nx + nu, 0.05
    g = torch.Generator().manual_seed(1)
    q = (0.5 + torch.rand(B, T, n, generator=g, dtype=torch.float64)).cuda().requires_grad_(True)
    p = (0.1 * torch.randn(B, T, n, generator=g, dtype=torch.float64)).cuda().requires_grad_(True)
    x0 = (0.1 * torch.randn(B, nx, generator=g, dtype=torch.float64)).cuda()
    cm = DiagQuadCost(nx, nu, q, p, q[:, -1], p[:, -1], device="cuda")
    mpc = DifferentiableMPCController(f, T * dt, dt, T, cm, device="cuda", grad_method="analytic", f_dyn_jac=f.jac, reg_eps=1e-3, max_iter=5)
    X, U = mpc(x0)          # round 1: TacdError "C_diag: not eligible ... pass dense costs"
    assert not mpc._last_diag
    (U[:, 0].sum() + X.sum()).backward()
    assert q.grad.shape == q.shape and p.grad.shape == p.shape
    # the same problem through GeneralQuadCost
    qd, pd = q.detach().clone().requires_grad_(True), p.detach().clone().requires_grad_(True)
    C = torch.diag_embed(qd)
    cm2 = GeneralQuadCost(nx, nu, C, pd, C[:, -1], pd[:, -1], device="cuda")
    mpc2 = DifferentiableMPCController(f, T * dt, dt, T, cm2, device="cuda", grad_method="analytic", f_dyn_jac=f.jac, reg_eps=1e-3, max_iter=5)
    X2, U2 = mpc2(x0)
    assert torch.equal(X, X2) and torch.equal(U, U2)
    (U2[:, 0].sum() + X2.sum()).backward()
    assert torch.allclose(p.grad, pd.grad, rtol=1e-9, atol=1e-12) and torch.allclose(q.grad, qd.grad, rtol=1e-9, atol=1e-12)


def _unicycle_loop(N_sim=6, T=8, B=3):
    from tacdmpc_b200.DifferentialMPC import DifferentiableMPCController, GeneralQuadCost
    from tacdmpc_b200.dynamics import Unicycle
    dt, nx, nu = 0.1, 3, 2
    n = nx + nu
    C = torch.zeros(T, n, n, dtype=torch.float64)
    C[:, :nx, :nx] = torch.diag(torch.tensor([20.0, 20.0, 5.0], dtype=torch.float64))
    C[:, nx:, nx:] = 0.1 * torch.eye(nu, dtype=torch.float64)
    cm = GeneralQuadCost(nx, nu, C.cuda(), torch.zeros(T, n, dtype=torch.float64).cuda(), (10 * C[0]).cuda(),
                         torch.zeros(n, dtype=torch.float64).cuda(), device="cuda")
    f = Unicycle()
    mpc = DifferentiableMPCController(f, T * dt, dt, T, cm, u_min=torch.tensor([-10.0, -6.0]), u_max=torch.tensor([10.0, 6.0]),
                                      grad_method="analytic", f_dyn_jac=f.jac, device="cuda", N_sim=N_sim)
    g = torch.Generator().manual_seed(3)
    L = N_sim + T + 1
    ts = torch.arange(L + 7, dtype=torch.float64) * dt
    xr = torch.stack([torch.stack([a * torch.cos(0.4 * ts + ph), a * torch.sin(0.4 * ts + ph), 0.4 * ts + ph + np.pi / 2], -1)
                      for a, ph in ((3.0, 0.0), (4.0, 1.0), (2.0, 2.0))])[:B]
    ur = 0.05 * torch.randn(B, L + 6, nu, generator=g, dtype=torch.float64)     # per-element, non-zero, LONGER than ref_len - 1
    x0 = xr[:, 0] + 0.3 * torch.randn(B, nx, generator=g, dtype=torch.float64)
    return mpc, x0.cuda(), xr.cuda(), ur.cuda(), L


@pytest.mark.parametrize("warm", [True, False])
def test_simulate_device_loop_equals_step_loop_for_any_reference_length(warm):
    mpc, x0, xr, ur, L = _unicycle_loop()
    U0 = 0.1 * torch.ones(x0.shape[0], mpc.horizon, 2, dtype=torch.float64, device="cuda")
    outs = []
    for xr_, ur_ in ((xr[:, :L], ur[:, :L - 1]), (xr, ur), (xr[:, :L + 3], ur[:, :L + 5]), (xr[:, :L + 5], ur[:, :L - 1])):
        with torch.no_grad():
            Xd, Ud = mpc(x0, U0, x_ref_full=xr_, u_ref_full=ur_, warm_start=warm)          # device loop
        x0g = x0.clone().requires_grad_(True)
        Xg, Ug = mpc(x0g, U0, x_ref_full=xr_, u_ref_full=ur_, warm_start=warm)             # autograd loop, step by step
        assert torch.allclose(Xd, Xg.detach(), rtol=0, atol=1e-9) and torch.allclose(Ud, Ug.detach(), rtol=0, atol=1e-9)
        outs.append((Xd, Ud))
    for Xd, Ud in outs[1:]:
        assert torch.equal(Xd, outs[0][0]) and torch.equal(Ud, outs[0][1])     # the extra reference steps are never read


def test_constrained_cost_is_constrained_without_grad_too():
    from tacdmpc_b200.DifferentialMPC import AugmentedLagrangianManager, ConstrainedQuadCost, DifferentiableMPCController
    from tacdmpc_b200.dynamics import DoubleIntegrator
    # (the velocity-limited double integrator of tests/test_gpu_al.py, whose plans the dense QP confirms)
    nx, nu, T, dt, N_sim = 2, 1, 20, 0.1, 4
    n = nx + nu
    C = torch.zeros(T, n, n, dtype=torch.float64)
    C[:, 0, 0], C[:, 1, 1], C[:, 2, 2] = 10.0, 1.0, 0.1
    al = AugmentedLagrangianManager(rho_init=10.0, rho_growth=4.0, rho_max=1e6, n_outer=10, tol=0.0)
    vmax = 1.0
    cm = ConstrainedQuadCost(al, torch.tensor([float("inf"), vmax], dtype=torch.float64), torch.tensor([-float("inf"), -vmax], dtype=torch.float64),
                             nx, nu, C.cuda(), torch.zeros(T, n, dtype=torch.float64).cuda(), (10 * C[0]).cuda(),
                             torch.zeros(n, dtype=torch.float64).cuda(), device="cuda")
    f = DoubleIntegrator(dt)
    mpc = DifferentiableMPCController(f, T * dt, dt, T, cm, grad_method="analytic", f_dyn_jac=f.jac, device="cuda", N_sim=N_sim)
    x0 = torch.tensor([[4.0, 0.0], [-3.0, 0.5]], dtype=torch.float64).cuda()
    xr = torch.zeros(1, N_sim + T + 1, nx, dtype=torch.float64).cuda()
    with torch.no_grad():
        Xs, _ = mpc(x0, x_ref_full=xr)        # round 1: took the device loop, which ignored x_min / x_max
    assert float(Xs[:, :, 1].abs().max()) <= vmax + 5e-3
    # the unconstrained loop does exceed the limit, so the test is not vacuous
    from tacdmpc_b200.DifferentialMPC import GeneralQuadCost
    cm0 = GeneralQuadCost(nx, nu, C.cuda(), torch.zeros(T, n, dtype=torch.float64).cuda(), (10 * C[0]).cuda(),
                          torch.zeros(n, dtype=torch.float64).cuda(), device="cuda")
    mpc0 = DifferentiableMPCController(f, T * dt, dt, T, cm0, grad_method="analytic", f_dyn_jac=f.jac, device="cuda", N_sim=N_sim)
    with torch.no_grad():
        X0, _ = mpc0(x0, x_ref_full=xr)
    assert float(X0[:, :, 1].abs().max()) > 1.3 * vmax


@pytest.mark.parametrize("diag", [False, True])
def test_cost_with_leading_batch_of_one_backpropagates(diag):
    from tacdmpc_b200.DifferentialMPC import DiagQuadCost, DifferentiableMPCController, GeneralQuadCost
    from tacdmpc_b200.dynamics import CartPole
    nx, nu, T, dt, B = 4, 1, 12, 0.05, 7
    n = nx + nu
    g = torch.Generator().manual_seed(5)
    q = (0.5 + torch.rand(1, T, n, generator=g, dtype=torch.float64)).cuda()
    p = (0.1 * torch.randn(1, T, n, generator=g, dtype=torch.float64)).cuda()
    x0 = (0.2 * torch.randn(B, nx, generator=g, dtype=torch.float64)).cuda()
    f = CartPole()

    def run(make):
        leaves = [t.clone().requires_grad_(True) for t in (q, p, q[:, -1], p[:, -1])]
        cm = make(*leaves)
        mpc = DifferentiableMPCController(f, T * dt, dt, T, cm, device="cuda", grad_method="analytic", f_dyn_jac=f.jac, reg_eps=1e-2)
        X, U = mpc(x0)
        (U[:, 0].sum() + 0.1 * X.sum()).backward()   # round 1: g.reshape((1, n, n)) raised on a [B, n, n] gradient
        return X, U, [t.grad for t in leaves]

    if diag:
        X, U, gr = run(lambda a, b, c_, d: DiagQuadCost(nx, nu, a, b, c_, d, device="cuda"))
    else:
        X, U, gr = run(lambda a, b, c_, d: GeneralQuadCost(nx, nu, torch.diag_embed(a), b, torch.diag_embed(c_), d, device="cuda"))
    assert all(g_ is not None and g_.shape == t.shape for g_, t in zip(gr, (q, p, q[:, -1], p[:, -1])))
    # same numbers as the genuinely shared ([T,n,n]) cost
    leaves = [t.clone().requires_grad_(True) for t in (q[0], p[0], q[0, -1], p[0, -1])]
    cm = GeneralQuadCost(nx, nu, torch.diag_embed(leaves[0]), leaves[1], torch.diag_embed(leaves[2]), leaves[3], device="cuda")
    mpc = DifferentiableMPCController(f, T * dt, dt, T, cm, device="cuda", grad_method="analytic", f_dyn_jac=f.jac, reg_eps=1e-2)
    X2, U2 = mpc(x0)
    (U2[:, 0].sum() + 0.1 * X2.sum()).backward()
    assert torch.equal(X, X2) and torch.equal(U, U2)
    for a, b in zip(gr, leaves):
        assert torch.allclose(a.reshape(b.shape), b.grad, rtol=1e-10, atol=1e-12)
