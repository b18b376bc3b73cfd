"""CPU-only checks of the host-side mirror added in the N1-N4 rows: cost-module semantics that need no kernel
(DiagQuadCost dense views, ConstrainedQuadCost constraint values), argument validation, and that nothing falls back
to a CPU solve."""
import pytest
import torch

from tacdmpc_b200.DifferentialMPC import (AugmentedLagrangianManager, ConstrainedQuadCost, DiagQuadCost, DifferentiableMPCController,
                                          GeneralQuadCost)
from tacdmpc_b200.dynamics import CartPole, DoubleIntegrator


def test_diag_cost_is_the_dense_cost():
    torch.manual_seed(0)
    B, T, nx, nu = 3, 4, 2, 1
    n = nx + nu
    q, p = torch.rand(B, T, n) + 0.1, torch.randn(B, T, n)
    dq = DiagQuadCost(nx, nu, q, p, q[:, -1].clone(), p[:, -1].clone(), device="cpu")
    gq = GeneralQuadCost(nx, nu, torch.diag_embed(q), p, torch.diag_embed(q[:, -1]), p[:, -1].clone(), device="cpu")
    assert torch.equal(dq.C, gq.C) and torch.equal(dq.C_final, gq.C_final) and torch.equal(dq.c, gq.c)
    X, U = torch.randn(B, T + 1, nx), torch.randn(B, T, nu)
    assert torch.allclose(dq.objective(X, U), gq.objective(X, U))
    for a, b in zip(dq.quadraticize(X, U), gq.quadraticize(X, U)):
        assert torch.equal(a, b)
    with pytest.raises(RuntimeError):
        dq.set_diag(q[:, :, :2], p)            # wrong width


def test_constraint_values_follow_the_reference_helper():
    """utils.py:31-35: g = [x - x_max ; x_min - x]."""
    al = AugmentedLagrangianManager()
    T, nx, nu = 3, 2, 1
    n = nx + nu
    cm = ConstrainedQuadCost(al, torch.tensor([1.0, 2.0]), torch.tensor([-1.0, -2.0]), nx, nu, torch.zeros(T, n, n), torch.zeros(T, n),
                             torch.zeros(n, n), torch.zeros(n), device="cpu")
    X = torch.tensor([[0.5, 3.0]])
    assert torch.equal(cm.constraint_values(X), torch.tensor([[-0.5, 1.0, -1.5, -5.0]]))
    al.reset(2, T, nx, "cpu", torch.float32)
    assert al.lam_lo.shape == (2, T + 1, nx) and al.rho == al.rho_init


def test_no_cpu_solve_and_argument_errors():
    T, nx, nu, dt = 3, 4, 1, 0.05
    n = nx + nu
    cm = GeneralQuadCost(nx, nu, torch.zeros(T, n, n), torch.zeros(T, n), torch.zeros(n, n), torch.zeros(n), device="cpu")
    f = CartPole()
    mpc = DifferentiableMPCController(f, T * dt, dt, T, cm, device="cpu", grad_method="analytic", f_dyn_jac=f.jac, N_sim=5)
    with pytest.raises(RuntimeError, match="CUDA"):
        mpc(torch.zeros(2, nx))                                       # there is no CPU path
    with pytest.raises(RuntimeError, match="cover"):
        mpc(torch.zeros(2, nx), x_ref_full=torch.zeros(2, 5, nx))     # needs N_sim + T + 1 reference states
    with pytest.raises(ValueError):
        DifferentiableMPCController(f, T * dt, dt, T, cm, device="cpu", compat="fast")
    al = AugmentedLagrangianManager()
    ccm = ConstrainedQuadCost(al, torch.ones(nx), -torch.ones(nx), nx, nu, torch.zeros(T, n, n), torch.zeros(T, n), torch.zeros(n, n),
                              torch.zeros(n), device="cpu")
    mpc2 = DifferentiableMPCController(DoubleIntegrator(dt), T * dt, dt, T, ccm, device="cpu")
    with pytest.raises(RuntimeError, match="CUDA"):
        mpc2(torch.zeros(2, 2))
