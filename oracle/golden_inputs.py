"""Seeded INPUT generators of the golden solve cases (tests/golden/<name>.npz).

TEST INFRASTRUCTURE ONLY.  torch-only: this module does not import the reference, so it also runs where
/root/reference does not exist.  Two users:

  * oracle/gen_golden.py attaches the reference's dynamics callables to these inputs, runs the unmodified
    reference on them and records the outputs;
  * tests/test_golden_inputs.py re-derives every fixture's inputs from here and compares them with what the
    committed .npz holds, so a fixture that was not produced by the committed generator (a stale file) fails.

Every generator draws its random numbers in an EXPLICIT `gen_dtype` (set as torch's default dtype for the
duration of the call and restored afterwards).  Round 1 generated the inputs under whatever default dtype the
previous case of gen_golden.main() had left behind, so a fixture's bytes depended on the order the cases ran in
(lti8_T50_f64 could not be reproduced by main()).  CASES below records, per fixture, the
generator, its arguments and the dtype its inputs are drawn in; the inputs are then cast to the case's `dtype`.
"""
from __future__ import annotations

import contextlib

import numpy as np
import torch

F32, F64 = torch.float32, torch.float64


@contextlib.contextmanager
def default_dtype(dt):
    old = torch.get_default_dtype()
    torch.set_default_dtype(dt)
    try:
        yield
    finally:
        torch.set_default_dtype(old)


def quad_cost(Q, R, T, final_scale=10.0, final_full=True):
    nx, nu = Q.shape[0], R.shape[0]
    n = nx + nu
    C = torch.zeros(T, n, n)
    C[:, :nx, :nx] = Q
    C[:, nx:, nx:] = R
    c = torch.zeros(T, n)
    if final_full:
        Cf = C[0].clone() * final_scale
    else:
        Cf = torch.zeros(n, n)
        Cf[:nx, :nx] = Q * final_scale
    cf = torch.zeros(n)
    return C, c, Cf, cf


def case_di(dtype, B=6, seed=0):
    """SURVEY 8d config 1 shaped: double integrator (examples/01_demo_linear.py:148-169), sinusoid references."""
    g = torch.Generator().manual_seed(seed)
    nx, nu, T, dt = 2, 1, 20, 0.05
    C, c, Cf, cf = quad_cost(torch.diag(torch.tensor([100.0, 1.0])), torch.diag(torch.tensor([0.01])), T)
    t = torch.arange(T + 1) * dt
    amp = 5.0 + torch.rand(B, 1, generator=g) * 5.0
    om = 1.0 + torch.rand(B, 1, generator=g)
    ph = torch.rand(B, 1, generator=g) * 3.0
    xref = torch.stack([amp * torch.sin(om * t + ph), amp * om * torch.cos(om * t + ph)], dim=2)
    uref = torch.zeros(B, T, nu)
    x0 = torch.randn(B, nx, generator=g) * torch.tensor([2.0, 1.0])
    return dict(nx=nx, nu=nu, T=T, dt=dt, dyn_kind="di", grad_method="analytic", dtype=dtype, x0=x0, C=C, c=c, Cf=Cf, cf=cf,
                xref=xref, uref=uref, Uinit=torch.zeros(B, T, nu), u_min=torch.tensor([-50.0]), u_max=torch.tensor([50.0]),
                spec_extra=dict(dyn="lti", lti_A=np.array([[1.0, dt], [0.0, 1.0]]), lti_B=np.array([[0.0], [dt]])))


def case_cartpole_shared(dtype, B=32, seed=0, grad_method="analytic", T=20):
    """SURVEY 8d config 2a (examples/02_demo_cartpole_regulation.py:124-135,179-181)."""
    g = torch.Generator().manual_seed(seed)
    nx, nu, dt = 4, 1, 0.05
    C, c, Cf, cf = quad_cost(torch.diag(torch.tensor([10.0, 1.0, 100.0, 1.0])), torch.diag(torch.tensor([0.1])), T,
                             final_full=False)
    x0 = torch.tensor([0.0, 0.0, 0.2, 0.0]) + torch.randn(B, nx, generator=g) * torch.tensor([0.5, 0.5, 0.2, 0.2])
    return dict(nx=nx, nu=nu, T=T, dt=dt, dyn_kind="cartpole", grad_method=grad_method, dtype=dtype, x0=x0, C=C, c=c, Cf=Cf, cf=cf,
                xref=torch.zeros(B, T + 1, nx), uref=torch.zeros(B, T, nu), Uinit=torch.zeros(B, T, nu),
                u_min=torch.tensor([-20.0]), u_max=torch.tensor([20.0]),
                spec_extra=dict(dyn="cartpole", dyn_params=np.array([1.0, 0.1, 0.5, 9.8])))


def case_cartpole_actor(dtype, B=16, seed=1, T=20):
    """ActorMPC-shaped problem (ACMPC/actor.py:58-81): per-element diagonal C, c; C_final=C[:,-1];
    reg_eps=1e-2; no bounds; U_init = 0.01 randn.  Exercises quirk Q4."""
    g = torch.Generator().manual_seed(seed)
    nx, nu, dt = 4, 1, 0.05
    n = nx + nu
    q = torch.nn.functional.softplus(torch.randn(B, T, n, generator=g)) + 1e-2
    C = torch.diag_embed(q)
    c = 0.3 * torch.randn(B, T, n, generator=g)
    Cf, cf = C[:, -1].clone(), c[:, -1].clone()
    x0 = torch.rand(B, nx, generator=g) * 0.45 - 0.2
    return dict(nx=nx, nu=nu, T=T, dt=dt, dyn_kind="cartpole", grad_method="analytic", dtype=dtype,
                x0=x0, C=C, c=c, Cf=Cf, cf=cf, xref=torch.zeros(B, T + 1, nx), uref=torch.zeros(B, T, nu),
                Uinit=0.01 * torch.randn(B, T, nu, generator=g), reg_eps=1e-2,
                spec_extra=dict(dyn="cartpole", dyn_params=np.array([1.0, 0.1, 0.5, 9.8])))


def case_unicycle(dtype, B=8, seed=2, tight=True, wrapped=False, T=30):
    """SURVEY 8d config 3 (examples/03_demo_unicycle_tracking.py:8-49): lemniscate references, two input boxes."""
    g = torch.Generator().manual_seed(seed)
    nx, nu, dt = 3, 2, 0.1
    C, c, Cf, cf = quad_cost(torch.diag(torch.tensor([20.0, 20.0, 5.0])), torch.diag(torch.tensor([0.1, 0.1])), T)
    ph = torch.rand(B, 1, generator=g) * 2 * torch.pi
    amp = 3.0 + 2.0 * torch.rand(B, 1, generator=g)
    tt = ph + torch.arange(T + 1) * dt * 0.5
    xr = amp * torch.sin(tt)
    yr = amp * torch.sin(tt) * torch.cos(tt)
    dxr = amp * torch.cos(tt)
    dyr = amp * torch.cos(2 * tt)
    th = torch.atan2(dyr, dxr)
    xref = torch.stack([xr, yr, th], dim=2)
    uref = torch.zeros(B, T, nu)
    x0 = xref[:, 0] + torch.randn(B, nx, generator=g) * torch.tensor([1.0, 1.0, 0.5])
    if tight:
        umin, umax = torch.tensor([-2.0, -1.0]), torch.tensor([2.0, 1.0])
    else:
        umin, umax = torch.tensor([-10.0, -2 * np.pi]), torch.tensor([10.0, 2 * np.pi])
    return dict(nx=nx, nu=nu, T=T, dt=dt, dyn_kind="unicycle_wrapped" if wrapped else "unicycle", grad_method="analytic",
                dtype=dtype, x0=x0, C=C, c=c, Cf=Cf, cf=cf, xref=xref, uref=uref, Uinit=torch.zeros(B, T, nu),
                u_min=umin, u_max=umax, spec_extra=dict(dyn="unicycle_wrapped" if wrapped else "unicycle"))


def case_lti(dtype, nx=8, nu=4, T=50, B=4, seed=0, mixed=False):
    """shared LTI dynamics A = I + 0.1 randn / sqrt(nx), B = 0.1 randn, per-element diagonal costs."""
    g = torch.Generator().manual_seed(seed)
    n = nx + nu
    A = torch.eye(nx) + 0.1 * torch.randn(nx, nx, generator=g) / np.sqrt(nx)
    Bm = 0.1 * torch.randn(nx, nu, generator=g)
    d = 0.5 + torch.rand(B, T, n, generator=g)
    d[..., nx:] *= 0.1
    C = torch.diag_embed(d)
    c = 0.1 * torch.randn(B, T, n, generator=g)
    Cf, cf = C[:, -1].clone(), c[:, -1].clone()
    if mixed:  # shared running cost, per-element terminal cost
        C, c = C[0].clone(), c[0].clone()
    x0 = torch.randn(B, nx, generator=g)
    return dict(nx=nx, nu=nu, T=T, dt=0.1, dyn_kind="lti", lti=(A, Bm), grad_method="analytic", dtype=dtype,
                x0=x0, C=C, c=c, Cf=Cf, cf=cf, xref=torch.zeros(B, T + 1, nx), uref=torch.zeros(B, T, nu),
                Uinit=torch.zeros(B, T, nu), spec_extra=dict(dyn="lti", lti_A=A.numpy().astype(np.float64),
                                                             lti_B=Bm.numpy().astype(np.float64)))


def case_gradcheck(dtype=F64):
    """tests/test_gradcheck.py:7-51 adapted to the current API: nx=nu=1, H=4, all-zero cost, f = x + dt u."""
    nx = nu = 1
    T, dt = 4, 0.1
    torch.manual_seed(0)
    x0 = torch.randn(1, nx, dtype=torch.float64)
    n = nx + nu
    return dict(nx=nx, nu=nu, T=T, dt=dt, dyn_kind="gradcheck", grad_method="auto_diff", dtype=torch.float64,
                x0=x0, C=torch.zeros(T, n, n), c=torch.zeros(T, n), Cf=torch.zeros(n, n), cf=torch.zeros(n),
                xref=torch.zeros(1, T + 1, nx), uref=torch.zeros(1, T, nu), Uinit=torch.zeros(1, T, nu),
                spec_extra=dict(dyn="lti", lti_A=np.array([[1.0]]), lti_B=np.array([[dt]])))


# fixture name -> (generator, keyword arguments, dtype the inputs are DRAWN in).  The third column is what round 1's
# gen_golden.main() produced implicitly through its case order (the default dtype left by the previous case); it is
# explicit now, so any case can be regenerated alone (`python oracle/gen_golden.py --only NAME`) bit-identically.
CASES = {
    "di_f64": (case_di, dict(dtype=F64), F32),
    "di_f32": (case_di, dict(dtype=F32), F64),
    "cartpole_shared_f32": (case_cartpole_shared, dict(dtype=F32), F32),
    "cartpole_shared_f64": (case_cartpole_shared, dict(dtype=F64, B=16), F32),
    "cartpole_actor_f32": (case_cartpole_actor, dict(dtype=F32), F64),
    "cartpole_actor_f64": (case_cartpole_actor, dict(dtype=F64, B=8), F32),
    "cartpole_ad_f64": (case_cartpole_shared, dict(dtype=F64, B=4, grad_method="auto_diff"), F64),
    "unicycle_tight_f32": (case_unicycle, dict(dtype=F32), F64),
    "unicycle_tight_f64": (case_unicycle, dict(dtype=F64), F32),
    "unicycle_loose_f64": (case_unicycle, dict(dtype=F64, B=4, tight=False), F64),
    "unicycle_loose_f32": (case_unicycle, dict(dtype=F32, B=8, tight=False, seed=5, T=20), F64),
    "unicycle_wrapped_f64": (case_unicycle, dict(dtype=F64, B=4, tight=True, wrapped=True, T=15), F32),
    "lti8_f32": (case_lti, dict(dtype=F32), F64),
    "lti8_T20_f32": (case_lti, dict(dtype=F32, T=20, B=4, seed=3), F32),
    "lti8_f64": (case_lti, dict(dtype=F64, T=20), F32),
    "lti8_T50_f64": (case_lti, dict(dtype=F64, T=50, B=2, seed=4), F32),
    "lti8_mixed_f64": (case_lti, dict(dtype=F64, T=12, mixed=True), F64),
    "lti16_f32": (case_lti, dict(dtype=F32, nx=16, nu=4, T=20, B=2), F64),
    "lti32_f32": (case_lti, dict(dtype=F32, nx=32, nu=8, T=10, B=2), F32),
    "gradcheck_f64": (case_gradcheck, dict(), F64),
}

INPUT_KEYS = ("x0", "C", "c", "Cf", "cf", "xref", "uref", "Uinit")


def make(name):
    """The inputs of fixture `name` exactly as gen_golden.py hands them to the reference (before the cast to the
    case dtype that run_case applies)."""
    fn, kw, gen_dt = CASES[name]
    with default_dtype(gen_dt):
        return fn(**kw)


def as_numpy(case):
    """What run_case stores: every input cast to the case dtype."""
    dt = case["dtype"]
    out = {k: case[k].clone().detach().to(dt).numpy() for k in INPUT_KEYS}
    if case.get("u_min") is not None:
        out["u_min"], out["u_max"] = case["u_min"].to(dt).numpy(), case["u_max"].to(dt).numpy()
    return out
