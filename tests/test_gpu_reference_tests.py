"""-m gpu: the reference's own tests for this path, as written (tests/test_forward.py, tests/test_jacobian_fallback.py;
`from DifferentialMPC import ...` replaced by the drop-in package, `CartPoleParams.from_gym()` by the gym CartPole-v1
constants it reads).  Both are stale at reference HEAD (the closed-loop keywords were removed); they state
what a user of the API expects, and they run here."""
from dataclasses import dataclass

import pytest
import torch

from tacdmpc_b200.DifferentialMPC import DifferentiableMPCController, GeneralQuadCost, GradMethod

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    yield
    torch.set_default_dtype(torch.float32)


@dataclass(frozen=True)
class CartPoleParams:
    m_c: float
    m_p: float
    l: float
    g: float


def f_cartpole(x, u, dt, p):
    """tests/test_forward.py:29-44"""
    pos, vel, theta, omega = torch.unbind(x, dim=-1)
    force = u.squeeze(-1)
    sin_t, cos_t = torch.sin(theta), torch.cos(theta)
    tot_mass, m_p_l = p.m_c + p.m_p, p.m_p * p.l
    temp = (force + m_p_l * omega ** 2 * sin_t) / tot_mass
    theta_dd = (p.g * sin_t - cos_t * temp) / (p.l * (4 / 3 - p.m_p * cos_t ** 2 / tot_mass))
    vel_dd = temp - m_p_l * theta_dd * cos_t / tot_mass
    return torch.stack((pos + vel * dt, vel + vel_dd * dt, theta + omega * dt, omega + theta_dd * dt), dim=-1)


def test_batch_cartpole_mpc_forward_backward():
    """tests/test_forward.py:59-110: closed loop of 150 steps from a Python dynamics callable (AUTO_DIFF), shapes, and
    a backward through the whole loop."""
    torch.set_default_dtype(torch.float64)
    device = torch.device("cuda")
    BATCH = 8
    DT, H, N = 0.05, 20, 150
    nx, nu = 4, 1
    params = CartPoleParams(1.0, 0.1, 0.5, 9.8)
    dyn = lambda x, u, dt: f_cartpole(x, u, dt, params)
    Q = torch.diag(torch.tensor([10., 1., 100., 1.], device=device))
    R = torch.diag(torch.tensor([0.1], device=device))
    C = torch.zeros(H, nx + nu, nx + nu, device=device)
    C[:, :nx, :nx] = Q
    C[:, nx:, nx:] = R
    c = torch.zeros_like(C[..., 0])
    C_final, c_final = C[0] * 10, torch.zeros(nx + nu, device=device)
    cost = GeneralQuadCost(nx, nu, C, c, C_final, c_final, device=str(device))
    mpc = DifferentiableMPCController(
        f_dyn=dyn, total_time=H * DT, step_size=DT, horizon=H, cost_module=cost,
        u_min=torch.tensor([-20.], device=device), u_max=torch.tensor([20.], device=device),
        grad_method=GradMethod.AUTO_DIFF, N_sim=N, device=str(device))
    x0 = torch.tensor([0., 0., 0.2, 0.], device=device).repeat(BATCH, 1).requires_grad_()
    x_ref = torch.zeros(BATCH, N + H + 1, nx, device=device)
    u_ref = torch.zeros(BATCH, N + H, nu, device=device)
    X_star, U_star = mpc.forward(x0, x_ref_full=x_ref, u_ref_full=u_ref)
    assert X_star.shape == (BATCH, N + 1, nx)
    assert U_star.shape == (BATCH, N, nu)
    loss = X_star[:, -1].pow(2).sum()
    loss.backward()
    # beyond the reference's assertions: the regulator does its job and the gradient is a number
    assert float(X_star.detach()[:, -1].abs().max()) < 1e-2 and torch.isfinite(x0.grad).all()


def f_simple(x, u, dt):
    return x + dt * u


def f_simple_nondiff(x, u, dt):
    return torch.round(x + dt * u)


@pytest.mark.parametrize("nondiff", [False, True])
def test_linearize_dynamics_fallback(nondiff):
    """tests/test_jacobian_fallback.py:18-63."""
    torch.set_default_dtype(torch.float64)
    device = torch.device("cuda")
    nx = nu = 1
    H, dt = 5, 0.1
    dyn = f_simple_nondiff if nondiff else f_simple
    cost = GeneralQuadCost(nx=nx, nu=nu, C=torch.zeros(H, nx + nu, nx + nu, device=device), c=torch.zeros(H, nx + nu, device=device),
                           C_final=torch.zeros(nx + nu, nx + nu, device=device), c_final=torch.zeros(nx + nu, device=device),
                           device=str(device))
    mpc = DifferentiableMPCController(f_dyn=dyn, total_time=H * dt, step_size=dt, horizon=H, cost_module=cost,
                                      grad_method=GradMethod.AUTO_DIFF, device=str(device))
    BATCH = 4
    x0 = torch.zeros(BATCH, nx, device=device, requires_grad=True)
    X_opt, U_opt = mpc.forward(x0)
    loss = X_opt[:, -1].pow(2).sum()
    loss.backward()
    assert x0.grad is not None
    assert torch.isfinite(x0.grad).all()
