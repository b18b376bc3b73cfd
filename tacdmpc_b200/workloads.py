"""Synthetic, seeded workloads of the shapes BASELINE.json / SURVEY §8(d) name (configs 2-5).

One definition shared by bench.py (`extra_configs`), tools/bench_configs.py and the BASELINE-size parity tests, so the
numbers and the parity evidence are about the same inputs.  CPU tensors come back; `spec` is the oracle-style dict
(oracle/oracle.py) — `gpu_spec()` turns it into the ops.Spec of the CUDA path.  Nothing here computes a solve.
"""
from __future__ import annotations

import math

import torch


def _quad(Q, R, T, final_full, final_scale=10.0):
    nx, nu = len(Q), len(R)
    n = nx + nu
    C = torch.zeros(T, n, n)
    C[:, :nx, :nx] = torch.diag(torch.tensor(Q))
    C[:, nx:, nx:] = torch.diag(torch.tensor(R))
    if final_full:
        Cf = final_scale * C[0].clone()
    else:
        Cf = torch.zeros(n, n)
        Cf[:nx, :nx] = final_scale * C[0, :nx, :nx]
    return C, torch.zeros(T, n), Cf, torch.zeros(n)


def cartpole(B, layout="shared", seed=0, T=20):
    """Config 2 (examples/02_demo_cartpole_regulation.py:124-135,179-181): Q=diag(10,1,100,1), R=0.1, C_final[:4,:4]=10Q,
    refs 0, x0=[0,0,0.2,0]+N(0,1)*[0.5,0.5,0.2,0.2], U_init=0, u in [-20,20].  layout 'shared' = 2a; 'per_element' = 2b
    (every element's cost scaled by U(0.5,1.5)); 'diag' = 2b's costs as compact diagonals (8f N1)."""
    g = torch.Generator().manual_seed(seed)
    nx, nu = 4, 1
    n = nx + nu
    C, c, Cf, cf = _quad([10.0, 1.0, 100.0, 1.0], [0.1], T, final_full=False)
    x0 = torch.tensor([0.0, 0.0, 0.2, 0.0]) + torch.randn(B, nx, generator=g) * torch.tensor([0.5, 0.5, 0.2, 0.2])
    ls = None
    if layout in ("per_element", "diag"):
        s = 0.5 + torch.rand(B, 1, 1, 1, generator=g)
        C = (C.unsqueeze(0) * s).contiguous()
        c = c.unsqueeze(0).expand(B, T, n).contiguous()
        Cf = (Cf.unsqueeze(0) * s[:, 0]).contiguous()
        cf = cf.unsqueeze(0).expand(B, n).contiguous()
        if layout == "diag":
            C = torch.diagonal(C, dim1=-2, dim2=-1).contiguous()
            Cf = torch.diagonal(Cf, dim1=-2, dim2=-1).contiguous()
        ls = (C[0].clone(), c[0].clone(), Cf[0].clone(), cf[0].clone())   # quirk Q4: element 0's cost scores candidates
    inp = dict(x0=x0, C=C, c=c, Cf=Cf, cf=cf, xref=torch.zeros(B, T + 1, nx), uref=torch.zeros(B, T, nu), Uinit=torch.zeros(B, T, nu))
    spec = dict(nx=nx, nu=nu, T=T, dt=0.05, dyn="cartpole", dyn_params=(1.0, 0.1, 0.5, 9.8), reg_eps=1e-6, max_iter=40,
                u_min=[-20.0], u_max=[20.0])
    return dict(name=f"cartpole_{layout}", spec=spec, inp=inp, ls=ls, C_diag=layout == "diag", Cf_batched=layout != "shared")


def unicycle(B=16384, tight=False, seed=0, T=30):
    """Config 3 (examples/03_demo_unicycle_tracking.py:8-49): lemniscate tracking, dt=0.1, Q=diag(20,20,5), R=diag(.1,.1),
    C_final=10 C[0], per-element references (phase U(0,2pi), amplitude U(3,5)), x0=ref_0+N(0,1)*[1,1,.5]; input box (i) +-[10,2pi]
    or (ii) +-[2,1] (clamp + active set + box QP)."""
    dt = 0.1
    g = torch.Generator().manual_seed(seed)
    nx, nu = 3, 2
    C, c, Cf, cf = _quad([20.0, 20.0, 5.0], [0.1, 0.1], T, final_full=True)
    ph = torch.rand(B, 1, generator=g) * 2 * math.pi
    amp = 3 + 2 * torch.rand(B, 1, generator=g)
    t = torch.arange(T + 1).view(1, -1) * dt
    th = 0.5 * t + ph
    xr = torch.stack([amp * torch.cos(th), amp * torch.sin(th) * torch.cos(th), torch.zeros_like(th)], -1)
    dx, dy = xr[:, 1:, 0] - xr[:, :-1, 0], xr[:, 1:, 1] - xr[:, :-1, 1]
    xr[:, :-1, 2] = torch.atan2(dy, dx)
    xr[:, -1, 2] = xr[:, -2, 2]
    x0 = xr[:, 0] + torch.randn(B, nx, generator=g) * torch.tensor([1.0, 1.0, 0.5])
    inp = dict(x0=x0, C=C, c=c, Cf=Cf, cf=cf, xref=xr.contiguous(), uref=torch.zeros(B, T, nu), Uinit=torch.zeros(B, T, nu))
    b = [2.0, 1.0] if tight else [10.0, 2 * math.pi]
    spec = dict(nx=nx, nu=nu, T=T, dt=dt, dyn="unicycle", reg_eps=1e-6, max_iter=40, u_min=[-b[0], -b[1]], u_max=b)
    return dict(name="unicycle_" + ("tight" if tight else "loose"), spec=spec, inp=inp, ls=None, C_diag=False, Cf_batched=False)


def ppo_actor(B, T=20, seed=0, diag=True):
    """Config 4 (ACMPC/actor.py:58-81): per-element diagonal costs softplus(randn)+1e-2, linear terms 0.1 randn, C_final = C[:,-1],
    reg_eps=1e-2, no bounds, U_init = 0.01 randn, obs U(-0.2,0.25)^4 (examples/02_demo_cartpole_AC.py:100)."""
    g = torch.Generator().manual_seed(seed)
    nx, nu = 4, 1
    n = nx + nu
    q = torch.nn.functional.softplus(torch.randn(B, T, n, generator=g)) + 1e-2
    p = 0.1 * torch.randn(B, T, n, generator=g)
    x0 = torch.rand(B, nx, generator=g) * 0.45 - 0.2
    Ui = 0.01 * torch.randn(B, T, nu, generator=g)
    if diag:
        C, Cf = q, q[:, -1].clone()
    else:
        C, Cf = torch.diag_embed(q), torch.diag_embed(q[:, -1])
    cf = p[:, -1].clone()
    inp = dict(x0=x0, C=C, c=p, Cf=Cf, cf=cf, xref=torch.zeros(B, T + 1, nx), uref=torch.zeros(B, T, nu), Uinit=Ui)
    spec = dict(nx=nx, nu=nu, T=T, dt=0.05, dyn="cartpole", dyn_params=(1.0, 0.1, 0.5, 9.8), reg_eps=1e-2, max_iter=40, jac="auto_diff")
    ls = (C[0].clone(), p[0].clone(), Cf[0].clone(), cf[0].clone())
    return dict(name=f"ppo_actor_{'diag' if diag else 'dense'}", spec=spec, inp=inp, ls=ls, C_diag=diag, Cf_batched=True)


def lti(nx, nu, B=4096, T=50, seed=0, rho=None):
    """Config 5: A = I + 0.1 randn/sqrt(nx), B = 0.1 randn (shared), per-element C = diag(U(0.5,1.5)) with the input block x0.1,
    c = 0.1 randn, C_final = C[:,-1], x0 ~ N(0,I), fp32.  As SURVEY states it the system is open-loop unstable over T=50 (every
    element meets a non-PD Q_uu in fp32 at nx >= 16 and stops after one iteration); `rho` rescales A to that spectral radius —
    the well-conditioned variant (rho < 1) that exercises several iLQR iterations."""
    g = torch.Generator().manual_seed(seed)
    n = nx + nu
    A = torch.eye(nx) + 0.1 * torch.randn(nx, nx, generator=g) / math.sqrt(nx)
    Bm = 0.1 * torch.randn(nx, nu, generator=g)
    dgl = 0.5 + torch.rand(B, T, n, generator=g)
    dgl[..., nx:] *= 0.1
    C = torch.diag_embed(dgl)
    c = 0.1 * torch.randn(B, T, n, generator=g)
    x0 = torch.randn(B, nx, generator=g)
    if rho is not None:
        A = A * (rho / float(torch.linalg.eigvals(A.double()).abs().max()))
    inp = dict(x0=x0, C=C, c=c, Cf=C[:, -1].clone(), cf=c[:, -1].clone(), xref=torch.zeros(B, T + 1, nx), uref=torch.zeros(B, T, nu),
               Uinit=torch.zeros(B, T, nu))
    spec = dict(nx=nx, nu=nu, T=T, dt=0.05, dyn="lti", A=A.numpy(), B=Bm.numpy(), reg_eps=1e-6, max_iter=40)
    ls = (C[0].clone(), c[0].clone(), C[0, -1].clone(), c[0, -1].clone())
    return dict(name=f"lti{nx}x{nu}" + ("" if rho is None else f"_rho{rho:g}"), spec=spec, inp=inp, ls=ls, C_diag=False, Cf_batched=True)


def gpu_spec(spec: dict, device, dtype=torch.float32):
    """oracle-style spec dict -> tacdmpc_b200.ops.Spec (LTI matrices as device tensors)."""
    from . import _abi, ops
    dyn = {"lti": _abi.DYN_LTI, "cartpole": _abi.DYN_CARTPOLE, "unicycle": _abi.DYN_UNICYCLE,
           "unicycle_wrapped": _abi.DYN_UNICYCLE_WRAPPED}[spec["dyn"]]
    jac = {"analytic": _abi.JAC_ANALYTIC, "auto_diff": _abi.JAC_AUTO_DIFF, "finite_diff": _abi.JAC_FINITE_DIFF}[spec.get("jac", "analytic")]
    A = Bm = None
    if spec["dyn"] == "lti":
        A = torch.as_tensor(spec["A"]).to(device=device, dtype=dtype).contiguous()
        Bm = torch.as_tensor(spec["B"]).to(device=device, dtype=dtype).contiguous()
    return ops.Spec(nx=spec["nx"], nu=spec["nu"], T=spec["T"], dt=spec["dt"], dyn_id=dyn, dyn_params=spec.get("dyn_params", ()),
                    dyn_A=A, dyn_B=Bm, reg_eps=spec.get("reg_eps", 1e-6), max_iter=spec.get("max_iter", 40), jac_mode=jac,
                    u_min=spec.get("u_min"), u_max=spec.get("u_max"))


def bytes_per_solve(spec: dict, layout: str, es: int = 4):
    """SURVEY §8(d) algorithmic bytes per solve (inputs read once, outputs written once, A/B recomputed): (forward, backward).
    layout: 'shared' | 'per_element' | 'diag'."""
    nx, nu, T = spec["nx"], spec["nu"], spec["T"]
    n = nx + nu
    xu = (T + 1) * nx + T * nu
    cost = T * n * n + T * n + n * n + n
    fwd = es * (nx + xu + T * nu + T * nu) + es * xu + T * nu
    bwd = es * (3 * xu) + T * nu + es * (xu + nx)
    if layout == "per_element":
        fwd += es * cost
        bwd += es * T * n * n + es * cost
    elif layout == "diag":
        fwd += es * (2 * T * n + 2 * n)
        bwd += es * T * n + es * (2 * T * n + 2 * n)
    return fwd, bwd


def flops_per_solve(spec: dict, mean_iters: float):
    """SURVEY §8(d) flop model (FMA = 2): (forward incl. the final re-linearisation, backward)."""
    nx, nu, T = spec["nx"], spec["nu"], spec["T"]
    n = nx + nu
    F_dyn, F_jac = {"cartpole": (45, 90), "unicycle": (10, 8), "unicycle_wrapped": (10, 8), "lti": (2 * nx * n, 0)}[spec["dyn"]]
    if spec["dyn"] == "lti" and nx == 2:
        F_dyn = 8
    F_quad = T * (2 * n * n + n) + 2 * n * n + n
    F_ric = T * (4 * nx ** 3 + 10 * nx * nx * nu + 4 * nx * nu * nu + 2 * nx * nx + 8 * nx * nu + nu ** 3 / 3 + 4 * nu * nu + nu)
    F_ls = 5 * (T * (2 * nx * nu + 4 * nu + nx + F_dyn) + T * (2 * n * n + 3 * n) + 2 * nx * nx + 3 * nx)
    F_acc = T * (2 * n * n + 3 * n) + 2 * nx * nx + 3 * nx
    F_iter = F_quad + T * F_jac + F_ric + F_ls + F_acc
    F_final = F_quad + T * F_jac
    F_bwd = F_ric + T * (2 * nx * nx + 6 * nx * nu) + 3 * T * n * n
    return mean_iters * F_iter + F_final, F_bwd
