"""Helpers for the -m gpu parity tests: golden/oracle numpy <-> CUDA tensors through tacdmpc_b200.ops."""
from __future__ import annotations

import numpy as np
import torch

import parity as P
from tacdmpc_b200 import _abi, ops

DYN = {"lti": _abi.DYN_LTI, "cartpole": _abi.DYN_CARTPOLE, "unicycle": _abi.DYN_UNICYCLE,
       "unicycle_wrapped": _abi.DYN_UNICYCLE_WRAPPED, "external": _abi.DYN_EXTERNAL}
JAC = {"analytic": _abi.JAC_ANALYTIC, "auto_diff": _abi.JAC_AUTO_DIFF, "finite_diff": _abi.JAC_FINITE_DIFF}


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def gpu_spec(spec: dict, dtype) -> ops.Spec:
    """oracle-style spec dict -> ops.Spec (device matrices for LTI)."""
    tdt = torch.float32 if np.dtype(dtype) == np.float32 else torch.float64
    A = B = None
    if spec["dyn"] == "lti":
        A, B = dev(spec["A"], tdt), dev(spec["B"], tdt)
    um, uM = spec.get("u_min"), spec.get("u_max")
    bc = lambda b: None if b is None else [float(v) for v in np.broadcast_to(np.asarray(b, np.float64), (spec["nu"],))]
    return ops.Spec(nx=spec["nx"], nu=spec["nu"], T=spec["T"], dt=spec["dt"], dyn_id=DYN[spec["dyn"]],
                    dyn_params=spec.get("dyn_params", ()), dyn_A=A, dyn_B=B, reg_eps=spec.get("reg_eps", 1e-6),
                    max_iter=spec.get("max_iter", 40), jac_mode=JAC[spec.get("jac", "analytic")],
                    u_min=bc(um), u_max=bc(uM))


def forward(spec: dict, g, *, replay=False, want_AB=True, cost_trace=False):
    dt = g["x0"].dtype
    s = gpu_spec(spec, dt)
    ls = P.ls_cost_of(g)
    rp = None
    if replay:
        rp = dict(iters=dev(g["iters"]), alpha_trace=dev(g["alpha_trace"]), flags=dev(g["flags"]))
    out = ops.ilqr_forward(s, dev(g["x0"]), dev(g["C"]), dev(g["c"]), dev(g["Cf"]), dev(g["cf"]), dev(g["xref"]),
                           dev(g["uref"]), dev(g["Uinit"]), ls_cost=None if ls is None else tuple(dev(a) for a in ls),
                           replay=rp, want_AB=want_AB, want_trace=True, want_cost_trace=cost_trace)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items()}


def backward(spec: dict, g, gX, gU, want=("grad_C", "grad_c", "grad_Cf", "grad_cf", "grad_x0", "grad_xref", "grad_uref")):
    s = gpu_spec(spec, g["x0"].dtype)
    out = ops.ilqr_backward(s, dev(g["X"]), dev(g["U"]), dev(g["C"]), dev(g["xref"]), dev(g["uref"]), dev(g["tight"]), dev(gX), dev(gU),
                            Cf_batched=g["Cf"].ndim == 3, want=want)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items()}


def golden_report(names=None):
    """One row per golden case (recorded from the unmodified reference): what the CUDA path achieves against it.
    Used by __graft_entry__.smoke() so the driver's record (smoke tail) carries the parity table, and by the tests.
      held      elements whose reference answer is itself reproducible to the north-star tolerance (tests/parity.py::elem_tol)
      X, U      max relative error in REPLAY (the reference's decisions) over those elements
      grads     max over the seven gradients and three losses of the error relative to the gradient's largest entry,
                beyond the reference's own fp32-vs-fp64 floor
      traces    free-running: elements whose whole decision trace (iteration count + alpha indices) equals the reference's
    """
    rows = []
    for name in names or P.SOLVE_CASES:
        g = P.load(name)
        spec = P.spec_of(g)
        tol = P.case_tol(g)
        out = forward(spec, g, replay=True, want_AB=False)
        ex, eu = P.rel_err(out["X"], g["X"]), P.rel_err(out["U"], g["U"])
        well = (P.elem_tol(g, "U") <= tol) & (P.elem_tol(g, "X") <= tol)
        free = forward(spec, g, want_AB=False)
        same, below, bad = P.lockstep(g, free["iters"], free["alpha_trace"], free["flags"], P.MARGIN[g["x0"].dtype])
        eg = 0.0
        gskip = False
        for loss in ("Usum", "U0sum", "weighted"):
            X, U = g["X"], g["U"]
            if loss == "Usum":
                gX, gU = np.zeros_like(X), np.ones_like(U)
            elif loss == "U0sum":
                gX, gU = np.zeros_like(X), np.zeros_like(U)
                gU[:, 0] = 1
            else:
                gX, gU = g["wX"], g["wU"]
            ob = backward(spec, g, gX, gU)
            for nm in ("C", "c", "Cf", "cf", "x0", "xref", "uref"):
                gold = g[f"grad_{loss}_{nm}"]
                scale = max(np.abs(gold).max(), 1.0)
                err = np.abs(ob[f"grad_{nm}"].reshape(gold.shape).astype(np.float64) - gold).max() / scale
                floor = 0.0
                if f"grad64_{loss}_{nm}" in g:
                    floor = 4 * np.abs(g[f"grad64_{loss}_{nm}"] - gold).max() / scale
                if floor > 4 * P.REPRO_LIMIT:
                    gskip = True
                    continue
                eg = max(eg, err - floor)
        B = len(well)
        rows.append(dict(case=name, B=B, tol=tol, held=int(well.sum()), X=float(ex[well].max()) if well.any() else float("nan"),
                         U=float(eu[well].max()) if well.any() else float("nan"), grads=max(eg, 0.0), grads_partial=gskip,
                         traces=int(same), below_margin=int(below), above_margin=len(bad)))
    return rows


def format_report(rows):
    lines = ["golden case (reference)            tol     held   X relerr  U relerr  grads     identical traces (departed below / above margin)"]
    for r in rows:
        lines.append(f"{r['case']:32s} {r['tol']:.0e} {r['held']:3d}/{r['B']:<3d} {r['X']:9.2e} {r['U']:9.2e} {r['grads']:9.2e}"
                     f"{'*' if r['grads_partial'] else ' '} {r['traces']:3d}/{r['B']:<3d} ({r['below_margin']} / {r['above_margin']})")
    return "\n".join(lines)
