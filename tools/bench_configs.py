"""Device-timed forward / backward of the other BASELINE.json / SURVEY §8d configurations (developer tool;
the driver's contract line is bench.py).  Prints one line per case: solves/s, ms, mean iterations.

    python tools/bench_configs.py [case ...]      cases: unicycle, unicycle_tight, ppo512, ppo4096, sweep, lti8, lti16, lti32
"""
import math
import sys
import time

import torch

sys.path.insert(0, ".")
from tacdmpc_b200 import _abi, ops  # noqa: E402
from tacdmpc_b200.dynamics import CartPole, LTI, Unicycle  # noqa: E402

dev = torch.device("cuda")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        out = fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        flush.fill_(1)
        a.record()
        out = fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2], out


def run(name, spec, inp, ls=None, C_diag=False, Cf_batched=False):
    d = {k: v.to(dev) for k, v in inp.items()}
    lsd = None if ls is None else tuple(t.to(dev) for t in ls)
    B, T = d["x0"].shape[0], spec.T
    fwd = lambda: ops.ilqr_forward(spec, d["x0"], d["C"], d["c"], d["Cf"], d["cf"], d["xref"], d["uref"], d["Uinit"],
                                   ls_cost=lsd, want_trace=False, C_diag=C_diag)
    t_f, fw = timed(fwd)
    gX = torch.zeros(B, T + 1, spec.nx, device=dev)
    gU = torch.zeros(B, T, spec.nu, device=dev)
    gU[:, 0] = 1
    bwd = lambda: ops.ilqr_backward(spec, fw["X"], fw["U"], d["C"], d["xref"], d["uref"], fw["tight"], gX, gU,
                                    Cf_batched=Cf_batched, C_diag=C_diag)
    t_b, _ = timed(bwd)
    it = fw["iters"].float()
    tight = float(fw["tight"].float().mean())
    nonpd = float((fw["flags"] & 1).float().mean())
    print(f"{name:28s} B={B:6d} T={T:3d} nx={spec.nx:2d} nu={spec.nu:2d}  fwd {t_f:8.3f} ms  bwd {t_b:7.3f} ms  "
          f"{B / (t_f + t_b) * 1e3:12.4g} solves/s  iters mean {float(it.mean()):5.2f} max {int(it.max()):2d}  tight {tight:.2f}  nonpd {nonpd:.3f}", flush=True)


def cartpole_inputs(B, T, seed=0):
    g = torch.Generator().manual_seed(seed)
    n = 5
    C = torch.zeros(T, n, n)
    C[:, :4, :4] = torch.diag(torch.tensor([10.0, 1.0, 100.0, 1.0]))
    C[:, 4:, 4:] = 0.1
    Cf = torch.zeros(n, n)
    Cf[:4, :4] = 10 * C[0, :4, :4]
    x0 = torch.tensor([0.0, 0.0, 0.2, 0.0]) + torch.randn(B, 4, generator=g) * torch.tensor([0.5, 0.5, 0.2, 0.2])
    return dict(x0=x0, C=C, c=torch.zeros(T, n), Cf=Cf, cf=torch.zeros(n), xref=torch.zeros(B, T + 1, 4),
                uref=torch.zeros(B, T, 1), Uinit=torch.zeros(B, T, 1))


def case_sweep():
    f = CartPole()
    for B in (1024, 4096, 16384, 65536):
        spec = ops.Spec(nx=4, nu=1, T=20, dt=0.05, dyn_id=f.dyn_id, dyn_params=f.params, u_min=[-20.0], u_max=[20.0])
        run(f"cartpole 2a B={B}", spec, cartpole_inputs(B, 20))


def case_unicycle(tight):
    """lemniscate tracking, T=30, B=16384, Q=diag(20,20,5), R=diag(.1,.1)."""
    B, T, dt = 16384, 30, 0.1
    g = torch.Generator().manual_seed(0)
    n = 5
    C = torch.zeros(T, n, n)
    C[:, :3, :3] = torch.diag(torch.tensor([20.0, 20.0, 5.0]))
    C[:, 3:, 3:] = torch.diag(torch.tensor([0.1, 0.1]))
    Cf = 10 * C[0]
    ph = torch.rand(B, 1, generator=g) * 2 * math.pi
    amp = 3 + 2 * torch.rand(B, 1, generator=g)
    t = torch.arange(T + 1).view(1, -1) * dt
    th = 0.5 * t + ph
    xr = torch.stack([amp * torch.cos(th), amp * torch.sin(th) * torch.cos(th), torch.zeros_like(th)], -1)
    dx, dy = xr[:, 1:, 0] - xr[:, :-1, 0], xr[:, 1:, 1] - xr[:, :-1, 1]
    xr[:, :-1, 2] = torch.atan2(dy, dx)
    xr[:, -1, 2] = xr[:, -2, 2]
    x0 = xr[:, 0] + torch.randn(B, 3, generator=g) * torch.tensor([1.0, 1.0, 0.5])
    inp = dict(x0=x0, C=C, c=torch.zeros(T, n), Cf=Cf, cf=torch.zeros(n), xref=xr.contiguous(), uref=torch.zeros(B, T, 2),
               Uinit=torch.zeros(B, T, 2))
    f = Unicycle()
    b = [2.0, 1.0] if tight else [10.0, 2 * math.pi]
    spec = ops.Spec(nx=3, nu=2, T=T, dt=dt, dyn_id=f.dyn_id, u_min=[-b[0], -b[1]], u_max=b)
    run("unicycle 3" + ("(ii) tight" if tight else "(i) loose"), spec, inp)


def case_ppo(B, T=20):
    """ActorMPC-shaped per-element diagonal costs, reg 1e-2, no bounds, U_init 0.01 randn."""
    g = torch.Generator().manual_seed(0)
    n = 5
    q = torch.nn.functional.softplus(torch.randn(B, T, n, generator=g)) + 1e-2
    p = 0.1 * torch.randn(B, T, n, generator=g)
    x0 = torch.rand(B, 4, generator=g) * 0.45 - 0.2
    inp = dict(x0=x0, C=q, c=p, Cf=q[:, -1].clone(), cf=p[:, -1].clone(), xref=torch.zeros(B, T + 1, 4), uref=torch.zeros(B, T, 1),
               Uinit=0.01 * torch.randn(B, T, 1, generator=g))
    f = CartPole()
    spec = ops.Spec(nx=4, nu=1, T=T, dt=0.05, dyn_id=f.dyn_id, dyn_params=f.params, reg_eps=1e-2, jac_mode=_abi.JAC_AUTO_DIFF)
    ls = (q[0], p[0], q[0, -1], p[0, -1])
    run(f"ppo 4 diag B={B} T={T}", spec, inp, ls=ls, C_diag=True, Cf_batched=True)
    dense = dict(inp, C=torch.diag_embed(q), Cf=torch.diag_embed(q[:, -1]))
    lsd = (torch.diag_embed(q[0]), p[0], torch.diag_embed(q[0, -1]), p[0, -1])
    run(f"ppo 4 dense B={B} T={T}", spec, dense, ls=lsd, Cf_batched=True)


def case_lti(nx, nu, B=4096, T=50):
    """random stable-ish LTI, per-element diagonal C (dense layout), one 8th of B=32768."""
    g = torch.Generator().manual_seed(0)
    n = nx + nu
    A = torch.eye(nx) + 0.1 * torch.randn(nx, nx, generator=g) / math.sqrt(nx)
    Bm = 0.1 * torch.randn(nx, nu, generator=g)
    dgl = 0.5 + torch.rand(B, T, n, generator=g)
    dgl[..., nx:] *= 0.1
    C = torch.diag_embed(dgl)
    c = 0.1 * torch.randn(B, T, n, generator=g)
    inp = dict(x0=torch.randn(B, nx, generator=g), C=C, c=c, Cf=C[:, -1].clone(), cf=c[:, -1].clone(),
               xref=torch.zeros(B, T + 1, nx), uref=torch.zeros(B, T, nu), Uinit=torch.zeros(B, T, nu))
    f = LTI(A, Bm)
    dA, dB = f.device_matrices(dev, torch.float32)
    spec = ops.Spec(nx=nx, nu=nu, T=T, dt=0.05, dyn_id=_abi.DYN_LTI, dyn_A=dA, dyn_B=dB)
    ls = (C[0], c[0], C[0, -1], c[0, -1])
    run(f"lti 5 nx={nx} nu={nu}", spec, inp, ls=ls, Cf_batched=True)


if __name__ == "__main__":
    cases = sys.argv[1:] or ["sweep", "unicycle", "unicycle_tight", "ppo512", "ppo4096", "lti8", "lti16", "lti32"]
    for c in cases:
        if c == "sweep": case_sweep()
        elif c == "unicycle": case_unicycle(False)
        elif c == "unicycle_tight": case_unicycle(True)
        elif c.startswith("ppo"): case_ppo(int(c[3:]))
        elif c == "lti8": case_lti(8, 4)
        elif c == "lti16": case_lti(16, 4); case_lti(16, 8)
        elif c == "lti32": case_lti(32, 8)
