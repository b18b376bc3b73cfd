"""Generate tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python oracle/gen_golden.py            # writes tests/golden/*.npz

The reference ships no golden vectors for the DMPC path and its own tests are
stale, so parity is pinned by executing the reference's
DifferentialMPC package (torch, CPU) on seeded inputs and recording
  * the solve outputs X*, U*, the active-set mask and the final Jacobians,
  * a lock-step decision trace (argmin alpha index, improved flag, candidate /
    new / best costs per iteration and element) obtained by re-driving
    solve_step with the controller's OWN methods (asserted bit-identical to
    solve_step itself),
  * first-iteration intermediates (A, B, l_*, K, k, candidates) for stage parity,
  * the seven gradients for three losses,
  * pnqp known answers.
The dynamics plugins are the shipped example functions, imported from
/root/reference/examples with matplotlib/gym stubbed.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import warnings
from unittest.mock import MagicMock

import numpy as np
import torch

REF = os.environ.get("TACD_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

for _m in ["matplotlib", "matplotlib.pyplot", "gym", "gymnasium", "gym.spaces", "gymnasium.spaces"]:
    sys.modules.setdefault(_m, MagicMock())
sys.path.insert(0, REF)
sys.path.insert(0, HERE)
warnings.filterwarnings("ignore")

from DifferentialMPC import DifferentiableMPCController, GeneralQuadCost, pnqp  # noqa: E402
from DifferentialMPC.controller import _vmap  # noqa: E402


def _load(name):
    spec = importlib.util.spec_from_file_location("ex_" + name, f"{REF}/examples/{name}.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


EX1 = _load("01_demo_linear")
EX2 = _load("02_demo_cartpole_regulation")
EX3 = _load("03_demo_unicycle_tracking")
EX3AC = _load("03_demo_unicycle_AC")
CP_PARAMS = EX2.CartPoleParams(m_c=1.0, m_p=0.1, l=0.5, g=9.8)  # gym CartPole-v1 constants


# ----------------------------------------------------------------------------
# lock-step tracer: solve_step (controller.py:275-315) re-driven with the
# controller's own methods
# ----------------------------------------------------------------------------
def traced_solve(ctrl, x0, U_init, first_iter_n=4, force=None):
    """force=(alpha[B,G], improved[B,G]) replays a recorded decision sequence instead of deciding."""
    B = x0.shape[0]
    cm = ctrl.cost_module
    U, X = U_init.clone(), ctrl.rollout_trajectory(x0, U_init)
    xr, ur = cm.x_ref, cm.u_ref
    best = cm.objective(X, U, x_ref_override=xr, u_ref_override=ur)
    tr = dict(alpha=[], improved=[], cand=[], newc=[], best=[])
    first = None
    active = torch.ones(B, dtype=torch.bool)
    for i in range(ctrl.max_iter):
        l = cm.quadraticize(X, U)
        A, Bm = ctrl.linearize_dynamics(X, U)
        K, k = _vmap(ctrl.backward_lqr)(A, Bm, *l)
        Xc, Uc, costs = _vmap(ctrl.evaluate_alphas, in_dims=(0, 0, 0, 0, 0, 0, 0))(x0, X, U, K, k, xr, ur)
        ai = torch.argmin(costs, dim=1)
        if force is not None:
            if i >= force[0].shape[1]:
                break
            ai = force[0][:, i].long()
        Xn = torch.gather(Xc, 1, ai.view(B, 1, 1, 1).expand(-1, 1, X.shape[1], X.shape[2])).squeeze(1)
        Un = torch.gather(Uc, 1, ai.view(B, 1, 1, 1).expand(-1, 1, U.shape[1], U.shape[2])).squeeze(1)
        newc = cm.objective(Xn, Un, x_ref_override=xr, u_ref_override=ur)
        imp = newc < best
        if force is not None:
            imp = force[1][:, i].bool()
        if first is None:
            n = min(first_iter_n, B)
            first = dict(A=A[:n], Bm=Bm[:n], lx=l[0][:n], lu=l[1][:n],
                         lxx=l[2].expand(B, *l[2].shape[-3:])[:n] if l[2].ndim == 4 else l[2][:n],
                         lxu=l[3].expand(B, *l[3].shape[-3:])[:n], luu=l[4].expand(B, *l[4].shape[-3:])[:n],
                         lxN=l[5][:n], lxxN=l[6].expand(B, *l[6].shape[-2:])[:n],
                         K=K[:n], k=k[:n], Xc=Xc[:n], Uc=Uc[:n], costs=costs[:n], X=X[:n], U=U[:n])
        tr["alpha"].append(ai.clone()); tr["improved"].append(imp.clone() & active)
        tr["cand"].append(costs.clone()); tr["newc"].append(newc.clone()); tr["best"].append(best.clone())
        tr["active"] = tr.get("active", []) + [active.clone()]
        if not imp.any():
            break
        # an element that fails to improve is frozen for good (quirk Q11)
        active = active & imp
        best = torch.where(imp, newc, best)
        X = torch.where(imp.view(B, 1, 1), Xn, X)
        U = torch.where(imp.view(B, 1, 1), Un, U)
    out = {k_: torch.stack(v, dim=1) for k_, v in tr.items()}
    return X, U, out, first


def per_element_trace(tr, max_iter):
    """Convert the batch-global trace to the per-element format of the C ABI."""
    act = tr["active"].numpy()            # [B, G] element still live at global iteration g
    imp = tr["improved"].numpy()
    alpha = tr["alpha"].numpy().astype(np.uint8)
    B, G = act.shape
    iters = act.sum(1).astype(np.int32)   # executed iterations while live
    at = np.full((B, max_iter), 255, np.uint8)
    flags = np.zeros(B, np.uint8)
    for b in range(B):
        at[b, :iters[b]] = alpha[b, :iters[b]]
        if iters[b] == max_iter and imp[b, iters[b] - 1]:
            flags[b] |= 2  # TACD_FLAG_MAXITER
    return iters, at, flags


def _np(t):
    return t.detach().cpu().numpy()


def run_case(name, *, nx, nu, T, dt, f_dyn, f_dyn_jac, grad_method, dtype, x0, C, c, Cf, cf, xref, uref, Uinit,
             u_min=None, u_max=None, reg_eps=1e-6, max_iter=40, spec_extra=None, grads=True):
    torch.set_default_dtype(dtype)
    leaves = [t.clone().detach().to(dtype) for t in (x0, C, c, Cf, cf, xref, uref)]
    x0, C, c, Cf, cf, xref, uref = leaves
    Uinit = Uinit.to(dtype)
    if u_min is not None:
        u_min, u_max = u_min.to(dtype), u_max.to(dtype)

    def make():
        cm = GeneralQuadCost(nx, nu, C, c, Cf, cf, device="cpu", x_ref=xref, u_ref=uref)
        ctrl = DifferentiableMPCController(f_dyn, T * dt, dt, T, cm, u_min=u_min, u_max=u_max, reg_eps=reg_eps,
                                           device="cpu", grad_method=grad_method, f_dyn_jac=f_dyn_jac,
                                           max_iter=max_iter)
        return cm, ctrl

    cm, ctrl = make()
    with torch.no_grad():
        Xt, Ut, tr, first = traced_solve(ctrl, x0, Uinit)
        cm2, ctrl2 = make()
        Xs, Us = ctrl2.solve_step(x0, Uinit)
        assert torch.equal(Xs, Xt) and torch.equal(Us, Ut), f"{name}: tracer is not bit-identical to solve_step"
        A_fin, B_fin = ctrl2.F_last
        tight = ctrl2.tight_mask_last
        final_cost = cm2.objective(Xs, Us, x_ref_override=cm2.x_ref, u_ref_override=cm2.u_ref)
    iters, at, flags = per_element_trace(tr, max_iter)
    force = (tr["alpha"], tr["improved"])

    def rel(a, b):
        a = a.double().reshape(a.shape[0], -1)
        b = b.double().reshape(b.shape[0], -1)
        return ((a - b).abs().amax(1) / b.abs().amax(1).clamp_min(1.0)).numpy()

    # (1) the reference's own sensitivity: x0 moved by one ulp, decisions held fixed
    noise_X = np.zeros(x0.shape[0])
    noise_U = np.zeros(x0.shape[0])
    gp = torch.Generator().manual_seed(99)
    with torch.no_grad():
        for trial in range(3):
            sgn = torch.where(torch.rand(x0.shape, generator=gp) < 0.5, -1.0, 1.0).to(dtype)
            x0p = torch.nextafter(x0, x0 + sgn)
            _, ctrlp = make()
            Xp, Up, _, _ = traced_solve(ctrlp, x0p, Uinit, force=force)
            noise_X = np.maximum(noise_X, rel(Xp, Xs))
            noise_U = np.maximum(noise_U, rel(Up, Us))
    # (2) the same decision sequence carried out in fp64 (the "exact" answer the fp32 run approximates)
    X64 = U64 = None
    if dtype == torch.float32:
        torch.set_default_dtype(torch.float64)
        with torch.no_grad():
            cm64 = GeneralQuadCost(nx, nu, C.double(), c.double(), Cf.double(), cf.double(), device="cpu",
                                   x_ref=xref.double(), u_ref=uref.double())
            ctrl64 = DifferentiableMPCController(
                f_dyn, T * dt, dt, T, cm64, u_min=None if u_min is None else u_min.double(),
                u_max=None if u_max is None else u_max.double(), reg_eps=reg_eps, device="cpu",
                grad_method=grad_method, f_dyn_jac=f_dyn_jac, max_iter=max_iter)
            X64, U64, _, first64 = traced_solve(ctrl64, x0.double(), Uinit.double(), force=force)
        torch.set_default_dtype(dtype)
    d = dict(
        meta_nx=nx, meta_nu=nu, meta_T=T, meta_dt=dt, meta_reg_eps=reg_eps, meta_max_iter=max_iter,
        meta_grad_method=grad_method, meta_dtype=str(dtype).replace("torch.", ""),
        x0=_np(x0), C=_np(C), c=_np(c), Cf=_np(Cf), cf=_np(cf), xref=_np(xref), uref=_np(uref), Uinit=_np(Uinit),
        X=_np(Xs), U=_np(Us), A=_np(A_fin), Bm=_np(B_fin), tight=_np(tight).astype(np.uint8), cost=_np(final_cost),
        iters=iters, alpha_trace=at, flags=flags, global_iters=np.int32(tr["alpha"].shape[1]),
        tr_cand=_np(tr["cand"]), tr_newc=_np(tr["newc"]), tr_best=_np(tr["best"]),
        tr_improved=_np(tr["improved"]).astype(np.uint8), tr_active=_np(tr["active"]).astype(np.uint8),
    )
    d["noise_X"], d["noise_U"] = noise_X, noise_U
    if X64 is not None:
        d["X64"], d["U64"] = _np(X64), _np(U64)
        d["it0_K64"], d["it0_k64"] = _np(first64["K"]), _np(first64["k"])
    if u_min is not None:
        d["u_min"] = _np(u_min)
        d["u_max"] = _np(u_max)
    for k_, v in (spec_extra or {}).items():
        d["meta_" + k_] = v
    for k_, v in first.items():
        d["it0_" + k_] = _np(v.contiguous())
    if grads:
        g = torch.Generator().manual_seed(1234)
        wX = torch.randn(Xs.shape, generator=g, dtype=dtype)
        wU = torch.randn(Us.shape, generator=g, dtype=dtype)
        d["wX"], d["wU"] = _np(wX), _np(wU)
        losses = {
            "Usum": lambda X_, U_: U_.sum(),
            "U0sum": lambda X_, U_: U_[:, 0].sum(),
            "weighted": lambda X_, U_: (X_ * wX).sum() + (U_ * wU).sum(),
        }
        for ln, lf in losses.items():
            lv = [t.clone().detach().requires_grad_(True) for t in leaves]
            cm3 = GeneralQuadCost(nx, nu, lv[1], lv[2], lv[3], lv[4], device="cpu", x_ref=lv[5], u_ref=lv[6])
            # registered buffers are not leaves any more; route the leaf tensors in again
            cm3.x_ref, cm3.u_ref = lv[5], lv[6]
            ctrl3 = DifferentiableMPCController(f_dyn, T * dt, dt, T, cm3, u_min=u_min, u_max=u_max, reg_eps=reg_eps,
                                                device="cpu", grad_method=grad_method, f_dyn_jac=f_dyn_jac,
                                                max_iter=max_iter)
            Xo, Uo = ctrl3(lv[0], Uinit)
            assert torch.equal(Xo.detach(), Xs) and torch.equal(Uo.detach(), Us)
            lf(Xo, Uo).backward()
            for nm, t in zip(("x0", "C", "c", "Cf", "cf", "xref", "uref"), lv):
                d[f"grad_{ln}_{nm}"] = _np(t.grad if t.grad is not None else torch.zeros_like(t))
            if dtype == torch.float32:
                # the same backward (ILQRSolve.backward, controller.py:63-115) carried out in fp64 on the
                # SAME saved tensors: how far the reference's fp32 gradients are from exact arithmetic
                from types import SimpleNamespace
                from DifferentialMPC import ILQRSolve
                torch.set_default_dtype(torch.float64)
                H = ctrl3.H_last
                ctrl3.u_min = None if u_min is None else u_min.double()
                ctrl3.u_max = None if u_max is None else u_max.double()
                ctx = SimpleNamespace(controller=ctrl3, saved_tensors=tuple(
                    t.detach().double() if t.dtype.is_floating_point else t.detach() for t in
                    (Xo, Uo, H[0], H[1], H[2], ctrl3.F_last[0], ctrl3.F_last[1], ctrl3.tight_mask_last, lv[5], lv[6])))
                gXo = torch.autograd.grad(lf(Xo.detach().requires_grad_(True), Uo.detach()), [], allow_unused=True) \
                    if False else None
                Xd, Ud = Xo.detach().double().requires_grad_(True), Uo.detach().double().requires_grad_(True)
                lval = (Ud.sum() if ln == "Usum" else Ud[:, 0].sum() if ln == "U0sum"
                        else (Xd * wX.double()).sum() + (Ud * wU.double()).sum())
                gXd, gUd = torch.autograd.grad(lval, [Xd, Ud], allow_unused=True)
                gXd = torch.zeros_like(Xd) if gXd is None else gXd
                gUd = torch.zeros_like(Ud) if gUd is None else gUd
                with torch.no_grad():
                    g64 = ILQRSolve.backward(ctx, gXd, gUd)
                shapes = [t.shape for t in lv]
                for nm, t, sh in zip(("x0", "C", "c", "Cf", "cf", "xref", "uref"), g64[:7], shapes):
                    d[f"grad64_{ln}_{nm}"] = _np(t.sum_to_size(sh) if t.shape != sh else t)
                torch.set_default_dtype(dtype)
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **d)
    print(f"{name:32s} B={x0.shape[0]:3d} global_iters={int(d['global_iters']):2d} "
          f"iters[min,max]=[{iters.min()},{iters.max()}] tight={d['tight'].mean():.2f} noiseU={noise_U.max():.1e} "
          + (f"ref32-vs-64={rel(Us, U64).max():.1e} " if X64 is not None else "") +
          f"{os.path.getsize(path) / 1024:.0f} KB")
    return d


# ----------------------------------------------------------------------------
# configurations
# ----------------------------------------------------------------------------
from golden_inputs import CASES, make as make_inputs, quad_cost  # noqa: E402  (seeded input generators, torch only)


def cartpole_fns():
    return (lambda x, u, dt: EX2.f_cartpole(x, u, dt, CP_PARAMS),
            lambda x, u, dt: EX2.f_cartpole_jac_batched(x, u, dt, CP_PARAMS))


def attach_dynamics(case):
    """golden_inputs case (tensors only) -> run_case keyword arguments: the reference's shipped example dynamics."""
    case = dict(case)
    kind = case.pop("dyn_kind")
    if kind == "di":
        f, fj = EX1.f_dyn_linear, EX1.f_dyn_jac_linear_batched
    elif kind == "cartpole":
        f, fj = cartpole_fns()
    elif kind == "unicycle":
        f, fj = EX3.f_dyn_unicycle, EX3.f_dyn_jac_unicycle
    elif kind == "unicycle_wrapped":
        f, fj = EX3AC.f_dyn_torch, EX3AC.f_dyn_jac_torch
    elif kind == "lti":
        A, Bm = case.pop("lti")

        def f(x, u, dt, A=A, Bm=Bm):
            return x @ A.to(x.dtype).T + u @ Bm.to(x.dtype).T

        def fj(x, u, dt, A=A, Bm=Bm):
            N = x.shape[0]
            return A.to(x.dtype).unsqueeze(0).expand(N, -1, -1), Bm.to(x.dtype).unsqueeze(0).expand(N, -1, -1)
    elif kind == "gradcheck":
        def f(x, u, dt):
            return x + dt * u
        fj = None
    else:
        raise KeyError(kind)
    case["f_dyn"] = f
    case["f_dyn_jac"] = fj if case["grad_method"] == "analytic" else None
    return case


def solve_case(name):
    return run_case(name, **attach_dynamics(make_inputs(name)))


def closed_loop_batch_tracking():
    """tests/test_batch_tracking.py:25-90 adapted to the current API (per-step set_reference + forward,
    as examples/01_demo_linear.py:114-125 does): B=2 double integrator, T=5, N_sim=20."""
    torch.set_default_dtype(torch.float64)
    nx, nu, T, N_sim, dt, B = 2, 1, 5, 20, 0.1, 2
    C, c, Cf, cf = quad_cost(torch.diag(torch.tensor([1.0, 0.1])), torch.diag(torch.tensor([0.01])), T)
    C, c, Cf, cf = (t.double() for t in (C, c, Cf, cf))
    cm = GeneralQuadCost(nx, nu, C, c, Cf, cf, device="cpu")
    Tm = __import__("importlib").import_module  # noqa: F841

    def f(x, u, dt):
        A = torch.tensor([[1.0, dt], [0.0, 1.0]], dtype=x.dtype)
        Bm = torch.tensor([[0.0], [dt]], dtype=x.dtype)
        return torch.einsum("...ij,...j->...i", A, x) + torch.einsum("...ij,...j->...i", Bm, u)

    def fj(x, u, dt):
        N = x.shape[0]
        A = torch.tensor([[1.0, dt], [0.0, 1.0]], dtype=x.dtype)
        Bm = torch.tensor([[0.0], [dt]], dtype=x.dtype)
        return A.expand(N, -1, -1), Bm.expand(N, -1, -1)

    mpc = DifferentiableMPCController(f, T * dt, dt, T, cm, grad_method="analytic", f_dyn_jac=fj, device="cpu",
                                      N_sim=N_sim)
    x0 = torch.tensor([[4.0, 0.0], [-4.0, 0.0]])
    ref_len = N_sim + T + 1
    ts = torch.linspace(0, 10, ref_len)
    xrA = torch.stack([torch.sin(ts), torch.cos(ts)], dim=1)
    xrB = torch.stack([0.5 * torch.cos(ts), -0.5 * torch.sin(ts)], dim=1)
    xref_full = torch.stack([xrA, xrB], dim=0)
    uref_full = torch.zeros(B, ref_len - 1, nu)
    Xs = [x0]
    Us = []
    x = x0
    with torch.no_grad():
        for k in range(N_sim):
            cm.set_reference(xref_full[:, k:k + T + 1], uref_full[:, k:k + T])
            _, U = mpc.forward(x)
            u = U[:, 0]
            x = f(x, u, dt)
            Xs.append(x)
            Us.append(u)
    d = dict(x0=_np(x0), xref_full=_np(xref_full), uref_full=_np(uref_full), C=_np(C), c=_np(c), Cf=_np(Cf), cf=_np(cf),
             Xs=_np(torch.stack(Xs, 1)), Us=_np(torch.stack(Us, 1)), meta_T=T, meta_N_sim=N_sim, meta_dt=dt)
    np.savez_compressed(os.path.join(OUT, "closed_loop_batch_tracking.npz"), **d)
    print("closed_loop_batch_tracking       Xs", d["Xs"].shape)


def config1_di_closed_loop():
    """SURVEY 8d config 1 = examples/01_demo_linear.py:94-169 as the script runs it (main() + run_simulation, plotting
    left out): double integrator, B = 50, T = 20, 150 closed-loop steps, fp64, sinusoid references with amplitude
    U(5,10) and frequency U(1,2) drawn after torch.manual_seed(0).  Position MSE 0.1296 for analytic and auto_diff
    (SURVEY 8c "soft KAT")."""
    torch.set_default_dtype(torch.float64)
    torch.manual_seed(0)
    B, DT, H, N_SIM, nx, nu = 50, 0.05, 20, 150, 2, 1
    Q, R = torch.diag(torch.tensor([100.0, 1.0])), torch.diag(torch.tensor([0.01]))
    C = torch.zeros(H, nx + nu, nx + nu)
    C[:, :nx, :nx], C[:, nx:, nx:] = Q, R
    c = torch.zeros(H, nx + nu)
    Cf, cf = C[0].clone() * 10, torch.zeros(nx + nu)
    cm = GeneralQuadCost(nx, nu, C, c, Cf, cf, device="cpu")
    ref_len = N_SIM + H + 1
    t_full = torch.arange(ref_len) * DT
    amp = 5.0 + torch.rand(B, 1) * 5.0
    om = 1.0 + torch.rand(B, 1) * 1.0
    xref = torch.stack([amp * torch.sin(om * t_full.unsqueeze(0)), amp * om * torch.cos(om * t_full.unsqueeze(0))], dim=2)
    uref = torch.zeros(B, ref_len - 1, nu)
    res = {}
    for gm, jac in (("analytic", EX1.f_dyn_jac_linear_batched), ("auto_diff", None)):
        mpc = DifferentiableMPCController(f_dyn=EX1.f_dyn_linear, total_time=H * DT, step_size=DT, horizon=H, cost_module=cm,
                                          u_min=torch.tensor([-50.0]), u_max=torch.tensor([50.0]), grad_method=gm, f_dyn_jac=jac,
                                          device="cpu")
        x = torch.zeros(B, nx)
        xs, us = [x], []
        for k in range(N_SIM):
            mpc.cost_module.set_reference(x_ref=xref[:, k:k + H + 1], u_ref=uref[:, k:k + H])
            _, U = mpc.forward(x)
            us.append(U[:, 0, :])
            x = EX1.f_dyn_linear(x, U[:, 0, :], DT)
            xs.append(x)
        Xs, Us = torch.stack(xs, 1), torch.stack(us, 1)
        res[gm] = (Xs, Us, torch.mean((Xs[:, :, 0] - xref[:, :N_SIM + 1, 0]) ** 2).item())
    np.savez_compressed(os.path.join(OUT, "config1_di_closed_loop.npz"), amp=_np(amp), om=_np(om), Xs=_np(res["analytic"][0]),
                        Us=_np(res["analytic"][1]), mse_analytic=res["analytic"][2], mse_auto_diff=res["auto_diff"][2])
    print("config1_di_closed_loop           MSE analytic", res["analytic"][2], "auto_diff", res["auto_diff"][2])


def closed_loop_unicycle_warm():
    """examples/03_demo_unicycle_tracking.py:119-141: two unicycle agents track lemniscates in closed loop, the
    solve warm-started with the previous plan shifted by one step (torch.roll, last input zeroed)."""
    torch.set_default_dtype(torch.float64)
    nx, nu, T, N_sim, dt, B = 3, 2, 10, 25, 0.1, 2
    C, c, Cf, cf = quad_cost(torch.diag(torch.tensor([20.0, 20.0, 5.0])), torch.diag(torch.tensor([0.1, 0.1])), T)
    C, c, Cf, cf = (t.double() for t in (C, c, Cf, cf))
    cm = GeneralQuadCost(nx, nu, C, c, Cf, cf, device="cpu")
    mpc = DifferentiableMPCController(EX3.f_dyn_unicycle, T * dt, dt, T, cm,
                                      u_min=torch.tensor([-10.0, -2 * np.pi]), u_max=torch.tensor([10.0, 2 * np.pi]),
                                      grad_method="analytic", f_dyn_jac=EX3.f_dyn_jac_unicycle, reg_eps=1e-6, device="cpu",
                                      N_sim=N_sim)
    ref_len = N_sim + T + 1
    ts = torch.arange(ref_len, dtype=torch.float64) * dt
    refs = []
    for amp, ph in ((4.0, 0.0), (3.0, 1.0)):
        th = 0.4 * ts + ph
        xr, yr = amp * torch.cos(th), amp * torch.sin(th) * torch.cos(th)
        hd = torch.atan2(torch.gradient(yr)[0], torch.gradient(xr)[0])
        refs.append(torch.stack([xr, yr, hd], dim=1))
    xref_full = torch.stack(refs, dim=0)
    uref_full = torch.zeros(B, ref_len - 1, nu)
    x0 = xref_full[:, 0].clone() + torch.tensor([[1.0, -1.0, 0.3], [-1.0, 1.0, -0.3]])
    Xs, Us, its = [x0], [], []
    x, warm = x0, torch.zeros(B, T, nu)
    with torch.no_grad():
        for k in range(N_sim):
            cm.set_reference(xref_full[:, k:k + T + 1], uref_full[:, k:k + T])
            _, U = mpc.forward(x, U_init=warm)
            u = U[:, 0]
            x = EX3.f_dyn_unicycle(x, u, dt)
            Xs.append(x)
            Us.append(u)
            warm = torch.roll(U, shifts=-1, dims=1)
            warm[:, -1] = 0.0
    d = dict(x0=_np(x0), xref_full=_np(xref_full), uref_full=_np(uref_full), C=_np(C), c=_np(c), Cf=_np(Cf), cf=_np(cf),
             Xs=_np(torch.stack(Xs, 1)), Us=_np(torch.stack(Us, 1)), meta_T=T, meta_N_sim=N_sim, meta_dt=dt)
    np.savez_compressed(os.path.join(OUT, "closed_loop_unicycle_warm.npz"), **d)
    print("closed_loop_unicycle_warm        Xs", d["Xs"].shape, "max |x - ref| at the end",
          float((d["Xs"][:, -1, :2] - d["xref_full"][:, N_sim, :2]).__abs__().max()))


def pnqp_kats():
    torch.set_default_dtype(torch.float64)   # (0.3 * torch.eye(m) below is formed in the default dtype)
    d = {}
    for dtype, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
        g = torch.Generator().manual_seed(7)
        for m in (1, 2, 3, 4):
            N = 48
            M = torch.randn(N, m, m, generator=g, dtype=torch.float64)
            H = (M @ M.mT + 0.3 * torch.eye(m)).to(dtype)
            q = (3.0 * torch.randn(N, m, generator=g, dtype=torch.float64)).to(dtype)
            lo = (-torch.rand(N, m, generator=g, dtype=torch.float64) * 1.5).to(dtype)
            hi = (torch.rand(N, m, generator=g, dtype=torch.float64) * 1.5).to(dtype)
            # a third of the problems unbounded, as in the un-constrained backward (controller.py:739-740)
            lo[::3] = -float("inf")
            hi[::3] = float("inf")
            xs, its = [], []
            for i in range(N):
                x, info = pnqp(H[i:i + 1], q[i:i + 1], lo[i:i + 1], hi[i:i + 1])
                xs.append(x[0])
                its.append(info["iters"])
            d[f"{tag}_m{m}_H"], d[f"{tag}_m{m}_q"] = _np(H), _np(q)
            d[f"{tag}_m{m}_lo"], d[f"{tag}_m{m}_hi"] = _np(lo), _np(hi)
            d[f"{tag}_m{m}_x"] = _np(torch.stack(xs))
            d[f"{tag}_m{m}_iters"] = np.array(its, np.int32)
    np.savez_compressed(os.path.join(OUT, "pnqp_kat.npz"), **d)
    print("pnqp_kat                         ", len(d), "arrays")


def fd_jacobian_kats():
    """grad_method="finite_diff" cannot be run end to end at reference HEAD: linearize_dynamics
    shadows the batch size B with the returned Jacobian (controller.py:468-471) and raises TypeError.
    The stage it calls, jacobian_finite_diff_batched (utils.py:48-64, eps=1e-4), works and is pinned here."""
    from DifferentialMPC import jacobian_finite_diff_batched
    d = {}
    f_cp, _ = cartpole_fns()
    for dtype, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
        g = torch.Generator().manual_seed(11)
        x = torch.randn(32, 4, generator=g, dtype=torch.float64).to(dtype)
        u = (3 * torch.randn(32, 1, generator=g, dtype=torch.float64)).to(dtype)
        A, Bm = jacobian_finite_diff_batched(f_cp, x, u, dt=0.05)
        d[f"cartpole_{tag}_x"], d[f"cartpole_{tag}_u"] = _np(x), _np(u)
        d[f"cartpole_{tag}_A"], d[f"cartpole_{tag}_B"] = _np(A), _np(Bm)
        x = torch.randn(32, 3, generator=g, dtype=torch.float64).to(dtype)
        u = torch.randn(32, 2, generator=g, dtype=torch.float64).to(dtype)
        A, Bm = jacobian_finite_diff_batched(EX3.f_dyn_unicycle, x, u, dt=0.1)
        d[f"unicycle_{tag}_x"], d[f"unicycle_{tag}_u"] = _np(x), _np(u)
        d[f"unicycle_{tag}_A"], d[f"unicycle_{tag}_B"] = _np(A), _np(Bm)
    np.savez_compressed(os.path.join(OUT, "fd_jacobian_kat.npz"), **d)
    print("fd_jacobian_kat                  ", len(d), "arrays")


def check_shipped_dynamics_match_example_variants():
    """The three cart-pole f's shipped in the reference (examples/02_demo_cartpole_regulation.py,
    02_demo_cartpole_AC.py, tests/test_forward.py) are the same function; record that."""
    ac = _load("02_demo_cartpole_AC")
    x = torch.randn(64, 4, dtype=torch.float64)
    u = torch.randn(64, 1, dtype=torch.float64)
    p2 = ac.CartPoleParams(m_c=1.0, m_p=0.1, l=0.5, g=9.8)
    a = EX2.f_cartpole(x, u, 0.05, CP_PARAMS)
    b = ac.f_cartpole_torch(x, u, 0.05, p2)
    assert torch.allclose(a, b, rtol=0, atol=1e-15), (a - b).abs().max()


SOLVE_ORDER = ["di_f64", "di_f32", "cartpole_shared_f32", "cartpole_shared_f64", "cartpole_actor_f32", "cartpole_actor_f64",
               "cartpole_ad_f64", "unicycle_tight_f32", "unicycle_tight_f64", "unicycle_loose_f64", "unicycle_loose_f32",
               "unicycle_wrapped_f64", "lti8_f32", "lti8_T20_f32", "lti8_f64", "lti8_T50_f64", "lti8_mixed_f64", "lti16_f32",
               "lti32_f32", "gradcheck_f64"]
EXTRA = {"fd_jacobian_kat": fd_jacobian_kats, "closed_loop_batch_tracking": closed_loop_batch_tracking,
         "closed_loop_unicycle_warm": closed_loop_unicycle_warm, "config1_di_closed_loop": config1_di_closed_loop,
         "pnqp_kat": pnqp_kats}


def main(argv=None):
    """python oracle/gen_golden.py [--only NAME ...] [--out DIR]: every fixture can be regenerated on its own; the inputs
    do not depend on which cases ran before (golden_inputs.CASES fixes the dtype they are drawn in)."""
    import argparse
    global OUT
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", nargs="*", default=None)
    ap.add_argument("--out", default=OUT)
    a = ap.parse_args(argv)
    OUT = a.out
    assert set(SOLVE_ORDER) == set(CASES)
    names = a.only if a.only else SOLVE_ORDER + list(EXTRA)
    torch.set_num_threads(4)
    os.makedirs(OUT, exist_ok=True)
    check_shipped_dynamics_match_example_variants()
    for nm in names:
        torch.manual_seed(0)
        torch.set_default_dtype(torch.float32)
        if nm in CASES:
            solve_case(nm)
        else:
            EXTRA[nm]()
    torch.set_default_dtype(torch.float32)


if __name__ == "__main__":
    main()
