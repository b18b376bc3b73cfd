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
