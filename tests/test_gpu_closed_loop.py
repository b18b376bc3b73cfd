"""-m gpu tests of the closed-loop receding-horizon API: ``mpc(x0, x_ref_full=, u_ref_full=)``
as the reference's stale tests call it (tests/test_batch_tracking.py:76-84, tests/test_forward.py:92-110), checked
against closed loops recorded from the unmodified reference driven step by step (oracle/gen_golden.py)."""
import numpy as np
import pytest
import torch

import parity as P

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_batch_tracking_api_reproduces_reference_closed_loop():
    """tests/test_batch_tracking.py as written (cold start every step): shape, distinct agents, reference states."""
    from tacdmpc_b200.DifferentialMPC import DifferentiableMPCController, GeneralQuadCost
    from tacdmpc_b200.dynamics import DoubleIntegrator
    g = P.load("closed_loop_batch_tracking")
    T, N_sim, dt = int(g["meta_T"]), int(g["meta_N_sim"]), float(g["meta_dt"])
    cm = GeneralQuadCost(2, 1, _t(g["C"]), _t(g["c"]), _t(g["Cf"]), _t(g["cf"]), device="cuda")
    f = DoubleIntegrator(dt)
    mpc = DifferentiableMPCController(f, T * dt, dt, T, cm, grad_method="analytic", f_dyn_jac=f.jac, device="cuda", N_sim=N_sim)
    Xs, Us = mpc.forward(_t(g["x0"]), x_ref_full=_t(g["xref_full"]), u_ref_full=_t(g["uref_full"]), warm_start=False)
    assert Xs.shape == (2, N_sim + 1, 2) and Us.shape == (2, N_sim, 1)
    assert not torch.allclose(Xs[0], Xs[1])
    assert np.abs(Xs.cpu().numpy() - g["Xs"]).max() <= 1e-7
    assert np.abs(Us.cpu().numpy() - g["Us"]).max() <= 1e-6


def _unicycle(g):
    from tacdmpc_b200.DifferentialMPC import DifferentiableMPCController, GeneralQuadCost
    from tacdmpc_b200.dynamics import Unicycle
    T, N_sim, dt = int(g["meta_T"]), int(g["meta_N_sim"]), float(g["meta_dt"])
    cm = GeneralQuadCost(3, 2, _t(g["C"]), _t(g["c"]), _t(g["Cf"]), _t(g["cf"]), device="cuda")
    f = Unicycle()
    mpc = DifferentiableMPCController(f, T * dt, dt, T, cm, u_min=torch.tensor([-10.0, -2 * np.pi]),
                                      u_max=torch.tensor([10.0, 2 * np.pi]), grad_method="analytic", f_dyn_jac=f.jac,
                                      reg_eps=1e-6, device="cuda", N_sim=N_sim)
    return mpc, cm, N_sim


def test_warm_started_unicycle_tracking_matches_reference():
    """examples/03_demo_unicycle_tracking.py:119-141 (shifted warm start, input box) in fp64."""
    g = P.load("closed_loop_unicycle_warm")
    mpc, cm, N_sim = _unicycle(g)
    Xs, Us = mpc(_t(g["x0"]), x_ref_full=_t(g["xref_full"]), u_ref_full=_t(g["uref_full"]))
    assert Xs.shape == g["Xs"].shape and Us.shape == g["Us"].shape
    # 25 chained solves: per-solve agreement is ~1e-9 (fp64); the closed loop is contractive around the reference
    assert np.abs(Xs.cpu().numpy() - g["Xs"]).max() <= 1e-6
    assert np.abs(Us.cpu().numpy() - g["Us"]).max() <= 1e-5
    # the tracking error at the end is the reference's
    err = float((Xs[:, -1, :2].cpu() - torch.from_numpy(g["xref_full"][:, N_sim, :2])).abs().max())
    assert err < 0.1


def test_autograd_loop_equals_device_loop_and_differentiates():
    """The device-resident loop (no gradient) and the autograd loop (one ILQRSolve node per step) give the same
    closed loop; tests/test_forward.py:92-110: backward through the final state must run and reach the cost."""
    g = P.load("closed_loop_unicycle_warm")
    mpc, cm, N_sim = _unicycle(g)
    x0, xr, ur = _t(g["x0"]), _t(g["xref_full"]), _t(g["uref_full"])
    with torch.no_grad():
        Xd, Ud = mpc(x0, x_ref_full=xr, u_ref_full=ur)
    cm.C = cm.C.clone().requires_grad_(True)
    Xa, Ua = mpc(x0, x_ref_full=xr, u_ref_full=ur)
    # (the autograd loop steps the state with the torch twin of the dynamics, the device loop with the kernel's
    # sincos: last-bit differences, carried through 25 chained solves — same bound as against the reference)
    assert float((Xa.detach() - Xd).abs().max()) <= 2e-6 and float((Ua.detach() - Ud).abs().max()) <= 2e-5
    Xa[:, -1].pow(2).sum().backward()
    assert cm.C.grad is not None and torch.isfinite(cm.C.grad).all() and float(cm.C.grad.abs().max()) > 0


def test_cartpole_regulation_closed_loop_fp32_large_batch():
    """tests/test_forward.py:60-110 shapes at a batch the device loop is meant for; regulation must stabilise."""
    from tacdmpc_b200.DifferentialMPC import DifferentiableMPCController, GeneralQuadCost
    from tacdmpc_b200.dynamics import CartPole
    B, H, N, dt, nx, nu = 2048, 20, 60, 0.05, 4, 1
    n = nx + nu
    C = torch.zeros(H, n, n, device="cuda")
    C[:, :nx, :nx] = torch.diag(torch.tensor([10.0, 1.0, 100.0, 1.0]))
    C[:, nx:, nx:] = 0.1
    cm = GeneralQuadCost(nx, nu, C, torch.zeros(H, n, device="cuda"), C[0] * 10, torch.zeros(n, device="cuda"), device="cuda")
    f = CartPole()
    mpc = DifferentiableMPCController(f, H * dt, dt, H, cm, u_min=torch.tensor([-20.0]), u_max=torch.tensor([20.0]),
                                      grad_method="auto_diff", N_sim=N, device="cuda")
    torch.manual_seed(0)
    x0 = torch.tensor([0.0, 0.0, 0.2, 0.0], device="cuda") + 0.05 * torch.randn(B, nx, device="cuda")
    X, U = mpc(x0, x_ref_full=torch.zeros(B, N + H + 1, nx, device="cuda"), u_ref_full=torch.zeros(B, N + H, nu, device="cuda"))
    assert X.shape == (B, N + 1, nx) and U.shape == (B, N, nu)
    assert torch.isfinite(X).all() and float(X[:, -1, 2].abs().max()) < 0.05      # pole upright after 3 s
    assert float(U.abs().max()) <= 20.0


@pytest.mark.parametrize("grad_method", ["analytic", "auto_diff"])
def test_config1_double_integrator_demo_known_answer(grad_method):
    """SURVEY §8d config 1 / §8c soft KAT: examples/01_demo_linear.py as the script runs it (B = 50, T = 20, 150
    closed-loop steps, fp64, references drawn after torch.manual_seed(0), cold start every step): position MSE 0.1296
    for both Jacobian modes, and the recorded closed loop of the unmodified reference state for state."""
    from tacdmpc_b200.DifferentialMPC import DifferentiableMPCController, GeneralQuadCost
    from tacdmpc_b200.dynamics import DoubleIntegrator
    g = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "config1_di_closed_loop.npz"))
    B, DT, H, N_SIM, nx, nu = 50, 0.05, 20, 150, 2, 1
    dd = dict(device="cuda", dtype=torch.float64)
    C = torch.zeros(H, 3, 3, **dd)
    C[:, 0, 0], C[:, 1, 1], C[:, 2, 2] = 100.0, 1.0, 0.01
    Cf = C[0].clone() * 10
    cm = GeneralQuadCost(nx, nu, C, torch.zeros(H, 3, **dd), Cf, torch.zeros(3, **dd), device="cuda")
    ref_len = N_SIM + H + 1
    t_full = torch.arange(ref_len, **dd) * DT
    amp, om = _t(g["amp"]), _t(g["om"])
    xref = torch.stack([amp * torch.sin(om * t_full[None]), amp * om * torch.cos(om * t_full[None])], dim=2)
    uref = torch.zeros(B, ref_len - 1, nu, **dd)
    f = DoubleIntegrator(DT)
    mpc = DifferentiableMPCController(f, H * DT, DT, H, cm, u_min=torch.tensor([-50.0]), u_max=torch.tensor([50.0]),
                                      grad_method=grad_method, f_dyn_jac=f.jac if grad_method == "analytic" else None, device="cuda",
                                      N_sim=N_SIM)
    Xs, Us = mpc.forward(torch.zeros(B, nx, **dd), x_ref_full=xref, u_ref_full=uref, warm_start=False)
    mse = float(torch.mean((Xs[:, :, 0] - xref[:, :N_SIM + 1, 0]) ** 2))
    assert abs(mse - 0.1296) < 5e-5 and abs(mse - float(g["mse_" + grad_method])) < 1e-9
    assert np.abs(Xs.cpu().numpy() - g["Xs"]).max() < 1e-6 and np.abs(Us.cpu().numpy() - g["Us"]).max() < 1e-5
