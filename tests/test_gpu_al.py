"""-m gpu tests of the state box constraints by augmented Lagrangian.  The
reference file is un-importable, so there is no golden vector: the DEFINED schedule (AugmentedLagrangianManager.py)
is checked against a dense QP solution of a state-constrained linear-quadratic problem (scipy SLSQP) and by its
own invariants on the cart-pole."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def _qp_reference(A, Bm, Q, R, Qf, x0, T, xmin, xmax):
    """min sum 1/2 x'Qx + 1/2 u'Ru + 1/2 xT'Qf xT  s.t. x+ = Ax + Bu, xmin <= x_t <= xmax (t >= 1): condensed to a dense
    QP in U (states are affine in U) and solved with exact derivatives."""
    from scipy.optimize import LinearConstraint, minimize
    nx, nu = Bm.shape
    # X[1:] = S U + s0
    S = np.zeros((T * nx, T * nu))
    s0 = np.zeros(T * nx)
    Ak = np.eye(nx)
    for t in range(T):
        Ak = A @ Ak
        s0[t * nx:(t + 1) * nx] = Ak @ x0
        for k in range(t + 1):
            S[t * nx:(t + 1) * nx, k * nu:(k + 1) * nu] = np.linalg.matrix_power(A, t - k) @ Bm
    Qbar = np.kron(np.eye(T), Q)
    Qbar[-nx:, -nx:] = Qf                      # x_T carries the terminal weight (x_1..x_{T-1} the running one)
    H = S.T @ Qbar @ S + np.kron(np.eye(T), R)
    g = S.T @ Qbar @ s0
    const = 0.5 * s0 @ Qbar @ s0 + 0.5 * x0 @ Q @ x0
    hi, lo = np.tile(xmax, T), np.tile(xmin, T)
    rows = np.isfinite(hi) | np.isfinite(lo)
    con = LinearConstraint(S[rows], lo[rows] - s0[rows], hi[rows] - s0[rows])
    r = minimize(lambda u: 0.5 * u @ H @ u + g @ u, np.zeros(T * nu), jac=lambda u: H @ u + g, hess=lambda u: H,
                 method="trust-constr", constraints=[con], options=dict(gtol=1e-12, xtol=1e-14, maxiter=3000))
    U = r.x.reshape(T, nu)
    X = np.vstack([x0[None], (S @ r.x + s0).reshape(T, nx)])
    return X, U, r.fun + const


def test_state_constrained_lq_matches_dense_qp():
    from tacdmpc_b200.DifferentialMPC import AugmentedLagrangianManager, ConstrainedQuadCost, DifferentiableMPCController
    from tacdmpc_b200.dynamics import DoubleIntegrator
    nx, nu, T, dt = 2, 1, 20, 0.1
    n = nx + nu
    f = DoubleIntegrator(dt)
    A = np.array([[1.0, dt], [0.0, 1.0]])
    Bm = np.array([[0.0], [dt]])
    Q, R = np.diag([10.0, 1.0]), np.diag([0.1])
    Qf = 10 * Q
    xmax, xmin = np.array([np.inf, 1.0]), np.array([-np.inf, -1.0])       # |velocity| <= 1, position free
    dd = dict(device="cuda", dtype=torch.float64)
    C = torch.zeros(T, n, n, **dd)
    C[:, :nx, :nx] = torch.from_numpy(Q).cuda()
    C[:, nx:, nx:] = torch.from_numpy(R).cuda()
    Cf = torch.zeros(n, n, **dd)
    Cf[:nx, :nx] = torch.from_numpy(Qf).cuda()
    al = AugmentedLagrangianManager(rho_init=10.0, rho_growth=4.0, rho_max=1e6, n_outer=10, tol=0.0)
    cm = ConstrainedQuadCost(al, torch.from_numpy(xmax), torch.from_numpy(xmin), nx, nu, C, torch.zeros(T, n, **dd), Cf,
                             torch.zeros(n, **dd), device="cuda")
    mpc = DifferentiableMPCController(f, T * dt, dt, T, cm, grad_method="analytic", f_dyn_jac=f.jac, device="cuda", reg_eps=1e-9)
    x0 = np.array([[4.0, 0.0], [-3.0, 0.5], [2.0, -0.8], [0.5, 0.9]])
    X, U = mpc(torch.from_numpy(x0).cuda())
    X, U = X.cpu().numpy(), U.cpu().numpy()
    # the unconstrained solution of the first element violates the bound by a wide margin
    cm0 = ConstrainedQuadCost(AugmentedLagrangianManager(n_outer=1), torch.from_numpy(xmax), torch.from_numpy(xmin), nx, nu, C,
                              torch.zeros(T, n, **dd), Cf, torch.zeros(n, **dd), device="cuda")
    mpc0 = DifferentiableMPCController(f, T * dt, dt, T, cm0, grad_method="analytic", f_dyn_jac=f.jac, device="cuda", reg_eps=1e-9)
    X0, _ = mpc0(torch.from_numpy(x0).cuda())
    assert float(X0[0, :, 1].abs().max()) > 2.0
    assert float(al.violation.max()) <= 1e-5 and al.outer_iters == 10
    for b in range(x0.shape[0]):
        Xq, Uq, Jq = _qp_reference(A, Bm, Q, R, Qf, x0[b], T, xmin, xmax)
        J = 0.5 * sum(X[b, t] @ Q @ X[b, t] + U[b, t] @ R @ U[b, t] for t in range(T)) + 0.5 * X[b, T] @ Qf @ X[b, T]
        assert abs(J - Jq) <= 1e-4 * max(1.0, abs(Jq)), (b, J, Jq)
        assert np.abs(X[b] - Xq).max() <= 2e-3 and np.abs(U[b] - Uq).max() <= 2e-2, (b, np.abs(X[b] - Xq).max(), np.abs(U[b] - Uq).max())
        assert (X[b, :, 1] <= 1.0 + 1e-5).all() and (X[b, :, 1] >= -1.0 - 1e-5).all()
    # multipliers are non-negative and non-zero only where the bound is (nearly) attained
    lam = torch.maximum(al.lam_hi, al.lam_lo).cpu().numpy()
    assert (lam >= 0).all() and lam.max() > 0
    assert (np.abs(np.abs(X[..., 1]) - 1.0)[lam[..., 1] > 1e-6] <= 1e-3).all()


def test_cartpole_track_limit_large_batch():
    """Cart-pole regulation with a track limit |x_cart| <= 0.5 (fp32, B = 4096): the constrained plans respect the
    limit, the unconstrained ones do not, the inputs stay inside their box, and early exit on `tol` works."""
    from tacdmpc_b200.DifferentialMPC import AugmentedLagrangianManager, ConstrainedQuadCost, DifferentiableMPCController
    from tacdmpc_b200.dynamics import CartPole
    B, T, dt, nx, nu = 4096, 20, 0.05, 4, 1
    n = nx + nu
    C = torch.zeros(T, n, n, device="cuda")
    C[:, :nx, :nx] = torch.diag(torch.tensor([10.0, 1.0, 100.0, 1.0]))
    C[:, nx:, nx:] = 0.1
    inf = float("inf")
    xmax = torch.tensor([0.5, inf, inf, inf])
    f = CartPole()
    torch.manual_seed(0)
    x0 = torch.tensor([0.0, 0.0, 0.2, 0.0], device="cuda") + 0.03 * torch.randn(B, nx, device="cuda")

    def solve(n_outer):
        al = AugmentedLagrangianManager(rho_init=50.0, rho_growth=3.0, n_outer=n_outer, tol=2e-3)
        cm = ConstrainedQuadCost(al, xmax, -xmax, nx, nu, C, torch.zeros(T, n, device="cuda"), C[0] * 10, torch.zeros(n, device="cuda"), device="cuda")
        mpc = DifferentiableMPCController(f, T * dt, dt, T, cm, u_min=torch.tensor([-20.0]), u_max=torch.tensor([20.0]),
                                          grad_method="analytic", f_dyn_jac=f.jac, device="cuda")
        X, U = mpc(x0)
        return X, U, al

    X0, U0, _ = solve(1)
    X1, U1, al = solve(10)
    assert float(X0[:, :, 0].abs().max()) > 0.6                      # the free solve overshoots the track limit
    assert float(X1[:, :, 0].abs().max()) <= 0.5 + 5e-3
    assert float(al.violation.max()) <= 5e-3 and float(al.violation.mean()) <= 1e-4 and 2 <= al.outer_iters <= 10
    assert torch.isfinite(X1).all() and float(U1.abs().max()) <= 20.0
