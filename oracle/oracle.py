"""numpy front-end of the CPU oracle (oracle/libtacd_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(tacdmpc_b200/) never imports this module.

A *spec* is a plain dict describing the problem family:
    dict(nx, nu, T, dt, dyn="lti"|"cartpole"|"unicycle"|"unicycle_wrapped"|"external",
         dyn_params=(...), A=..., B=... (lti), reg_eps, max_iter, u_min, u_max,
         jac="analytic"|"auto_diff"|"finite_diff", alphas=(...))
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from tacdmpc_b200 import _abi  # noqa: E402  (struct layout of include/tacd_ilqr.h)

_LIB = None

DYN_IDS = {"lti": _abi.DYN_LTI, "cartpole": _abi.DYN_CARTPOLE, "unicycle": _abi.DYN_UNICYCLE,
           "unicycle_wrapped": _abi.DYN_UNICYCLE_WRAPPED, "external": _abi.DYN_EXTERNAL}
JAC_IDS = {"analytic": _abi.JAC_ANALYTIC, "auto_diff": _abi.JAC_AUTO_DIFF, "finite_diff": _abi.JAC_FINITE_DIFF}


def build(force: bool = False) -> str:
    """Compile the oracle with oracle/Makefile (gcc, a second or two)."""
    so = os.path.join(_HERE, "libtacd_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("ilqr_oracle.c", "ilqr_oracle_impl.h")] + \
           [os.path.join(_ROOT, "include", "tacd_ilqr.h")]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "libtacd_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        assert _LIB.tacd_oracle_abi_version() == _abi.ABI_VERSION
    return _LIB


def num_threads() -> int:
    return int(lib().tacd_oracle_num_threads())


def set_num_threads(n: int) -> None:
    lib().tacd_oracle_set_num_threads(int(n))


def _dt(a):
    return {np.dtype(np.float32): _abi.F32, np.dtype(np.float64): _abi.F64}[np.dtype(a.dtype)]


def _c(a, dtype=None):
    if a is None:
        return None
    return np.ascontiguousarray(a, dtype=dtype)


def _p(a):
    return None if a is None else a.ctypes.data


def problem(spec: dict, B: int, dtype, *, C_batched=0, Cf_batched=0, xref_batched=1, uref_batched=1,
            ls_cost_shared=0, keep=None):
    """Build a tacd_problem from a spec.  ``keep`` collects arrays whose memory the struct points at."""
    keep = keep if keep is not None else []
    npdt = np.dtype(dtype)
    dyn = spec["dyn"]
    A = B_ = None
    if dyn == "lti":
        A = _c(spec["A"], npdt)
        B_ = _c(spec["B"], npdt)
        keep += [A, B_]
    umin = spec.get("u_min")
    umax = spec.get("u_max")
    return _abi.make_problem(
        B=B, T=spec["T"], nx=spec["nx"], nu=spec["nu"], dtype=_abi.F32 if npdt == np.float32 else _abi.F64,
        dyn_id=DYN_IDS[dyn], dt=spec["dt"], reg_eps=spec.get("reg_eps", 1e-6), max_iter=spec.get("max_iter", 40),
        C_batched=C_batched, Cf_batched=Cf_batched, xref_batched=xref_batched, uref_batched=uref_batched,
        ls_cost_shared=ls_cost_shared, jac_mode=JAC_IDS[spec.get("jac", "analytic")], fd_eps=spec.get("fd_eps", 1e-4),
        dyn_params=spec.get("dyn_params", ()), dyn_A=_p(A), dyn_B=_p(B_),
        alphas=spec.get("alphas", _abi.DEFAULT_ALPHAS),
        umin=None if umin is None else np.broadcast_to(np.asarray(umin, dtype=np.float64), (spec["nu"],)),
        umax=None if umax is None else np.broadcast_to(np.asarray(umax, dtype=np.float64), (spec["nu"],)))


def _layout(spec, x0, Cm, Cf, xref, uref):
    B = x0.shape[0]
    return dict(C_batched=int(Cm.ndim == 4), Cf_batched=int(Cf.ndim == 3),
                xref_batched=int(xref.ndim == 3 and xref.shape[0] == B and B > 1 or (xref.ndim == 3 and B == 1)),
                uref_batched=int(uref.ndim == 3 and uref.shape[0] == B and B > 1 or (uref.ndim == 3 and B == 1)))


def forward(spec, x0, Cm, c, Cf, cf, xref, uref, Uinit, *, ls_cost=None, replay=None, want_AB=True,
            want_trace=True):
    """solve_step on the CPU oracle.  Returns a dict of numpy arrays."""
    dt = x0.dtype
    x0, Cm, c, Cf, cf, xref, uref, Uinit = (_c(a, dt) for a in (x0, Cm, c, Cf, cf, xref, uref, Uinit))
    B, T, nx, nu = x0.shape[0], spec["T"], spec["nx"], spec["nu"]
    keep = []
    lay = _layout(spec, x0, Cm, Cf, xref, uref)
    p = problem(spec, B, dt, ls_cost_shared=int(ls_cost is not None), keep=keep, **lay)
    io = _abi.ForwardIO()
    out = dict(X=np.empty((B, T + 1, nx), dt), U=np.empty((B, T, nu), dt),
               tight=np.zeros((B, T, nu), np.uint8), iters=np.zeros(B, np.int32),
               alpha_trace=np.full((B, p.max_iter), 255, np.uint8), flags=np.zeros(B, np.uint8),
               cost=np.empty(B, dt))
    if want_AB:
        out["A"] = np.empty((B, T, nx, nx), dt)
        out["Bm"] = np.empty((B, T, nx, nu), dt)
    if want_trace:
        out["cost_trace"] = np.full((B, p.max_iter, p.n_alpha + 2), np.nan, dt)
    io.x0, io.C, io.c, io.Cf, io.cf = _p(x0), _p(Cm), _p(c), _p(Cf), _p(cf)
    if ls_cost is not None:
        ls = [_c(a, dt) for a in ls_cost]
        keep += ls
        io.C_ls, io.c_ls, io.Cf_ls, io.cf_ls = (_p(a) for a in ls)
    io.xref, io.uref, io.Uinit = _p(xref), _p(uref), _p(Uinit)
    if replay is not None:
        r = [_c(replay["iters"], np.int32), _c(replay["alpha_trace"], np.uint8), _c(replay["flags"], np.uint8)]
        keep += r
        io.replay_iters, io.replay_alpha, io.replay_flags = (_p(a) for a in r)
    for k_, v in out.items():
        setattr(io, k_, _p(v))
    rc = lib().tacd_oracle_ilqr_forward(C.byref(p), C.byref(io))
    if rc != 0:
        raise RuntimeError(f"tacd_oracle_ilqr_forward rc={rc}")
    return out


def backward(spec, X, U, Cm, xref, uref, tight, gX, gU, *, A=None, Bm=None, Cf_batched=0,
             want=("grad_C", "grad_c", "grad_Cf", "grad_cf", "grad_x0", "grad_xref", "grad_uref", "dX", "dU")):
    """ILQRSolve.backward on the CPU oracle."""
    dt = X.dtype
    X, U, Cm, xref, uref, gX, gU = (_c(a, dt) for a in (X, U, Cm, xref, uref, gX, gU))
    tight = _c(tight, np.uint8)
    B, T, nx, nu = X.shape[0], spec["T"], spec["nx"], spec["nu"]
    n = nx + nu
    keep = []
    Cb = int(Cm.ndim == 4)
    xb = int(xref.shape[0] == B)
    ub = int(uref.shape[0] == B)
    p = problem(spec, B, dt, C_batched=Cb, Cf_batched=Cf_batched, xref_batched=xb, uref_batched=ub, keep=keep)
    io = _abi.BackwardIO()
    io.X, io.U, io.C, io.xref, io.uref, io.tight, io.gX, io.gU = (_p(a) for a in (X, U, Cm, xref, uref, tight, gX, gU))
    if A is not None:
        A, Bm = _c(A, dt), _c(Bm, dt)
        io.A, io.Bm = _p(A), _p(Bm)
    shapes = dict(grad_C=((B,) if Cb else ()) + (T, n, n), grad_c=((B,) if Cb else ()) + (T, n),
                  grad_Cf=((B,) if Cf_batched else ()) + (n, n), grad_cf=((B,) if Cf_batched else ()) + (n,),
                  grad_x0=(B, nx), grad_xref=((B,) if xb else (1,)) + (T + 1, nx),
                  grad_uref=((B,) if ub else (1,)) + (T, nu), dX=(B, T + 1, nx), dU=(B, T, nu))
    out = {k_: np.zeros(shapes[k_], dt) for k_ in want}
    for k_, v in out.items():
        setattr(io, k_, _p(v))
    rc = lib().tacd_oracle_ilqr_backward(C.byref(p), C.byref(io))
    if rc != 0:
        raise RuntimeError(f"tacd_oracle_ilqr_backward rc={rc}")
    return out


def rollout(spec, x0, U):
    dt = x0.dtype
    x0, U = _c(x0, dt), _c(U, dt)
    B = x0.shape[0]
    keep = []
    p = problem(spec, B, dt, keep=keep)
    X = np.empty((B, spec["T"] + 1, spec["nx"]), dt)
    assert lib().tacd_oracle_rollout(C.byref(p), C.c_void_p(_p(x0)), C.c_void_p(_p(U)), C.c_void_p(_p(X))) == 0
    return X


def linearize(spec, X, U):
    dt = X.dtype
    X, U = _c(X, dt), _c(U, dt)
    B, T, nx, nu = X.shape[0], spec["T"], spec["nx"], spec["nu"]
    keep = []
    p = problem(spec, B, dt, keep=keep)
    A = np.empty((B, T, nx, nx), dt)
    Bm = np.empty((B, T, nx, nu), dt)
    assert lib().tacd_oracle_linearize(C.byref(p), C.c_void_p(_p(X)), C.c_void_p(_p(U)), C.c_void_p(_p(A)),
                                       C.c_void_p(_p(Bm))) == 0
    return A, Bm


def objective(spec, X, U, Cm, c, Cf, cf, xref, uref):
    dt = X.dtype
    X, U, Cm, c, Cf, cf, xref, uref = (_c(a, dt) for a in (X, U, Cm, c, Cf, cf, xref, uref))
    B = X.shape[0]
    keep = []
    p = problem(spec, B, dt, keep=keep, C_batched=int(Cm.ndim == 4), Cf_batched=int(Cf.ndim == 3),
                xref_batched=int(xref.shape[0] == B), uref_batched=int(uref.shape[0] == B))
    cost = np.empty(B, dt)
    args = [C.c_void_p(_p(a)) for a in (X, U, Cm, c, Cf, cf, xref, uref, cost)]
    assert lib().tacd_oracle_objective(C.byref(p), *args) == 0
    return cost


def riccati(spec, A, Bm, lx, lu, lxx, lxu, luu, lxN, lxxN):
    dt = A.dtype
    arrs = [_c(a, dt) for a in (A, Bm, lx, lu, lxx, lxu, luu, lxN, lxxN)]
    B, T, nx, nu = A.shape[0], spec["T"], spec["nx"], spec["nu"]
    keep = []
    p = problem(spec, B, dt, keep=keep)
    K = np.empty((B, T, nu, nx), dt)
    k = np.empty((B, T, nu), dt)
    flags = np.zeros(B, np.uint8)
    args = [C.c_void_p(_p(a)) for a in arrs + [K, k, flags]]
    assert lib().tacd_oracle_riccati(C.byref(p), *args) == 0
    return K, k, flags


def linesearch(spec, x0, X, U, K, k, Cm, c, Cf, cf, xref, uref):
    dt = x0.dtype
    x0, X, U, K, k, Cm, c, Cf, cf, xref, uref = (_c(a, dt) for a in (x0, X, U, K, k, Cm, c, Cf, cf, xref, uref))
    B, T, nx, nu = x0.shape[0], spec["T"], spec["nx"], spec["nu"]
    keep = []
    p = problem(spec, B, dt, keep=keep, C_batched=int(Cm.ndim == 4), Cf_batched=int(Cf.ndim == 3),
                xref_batched=int(xref.shape[0] == B), uref_batched=int(uref.shape[0] == B))
    na = p.n_alpha
    Xc = np.empty((B, na, T + 1, nx), dt)
    Uc = np.empty((B, na, T, nu), dt)
    costs = np.empty((B, na), dt)
    args = [C.c_void_p(_p(a)) for a in (x0, X, U, K, k, Cm, c, Cf, cf, xref, uref, Xc, Uc, costs)]
    assert lib().tacd_oracle_linesearch(C.byref(p), *args) == 0
    return Xc, Uc, costs


def pnqp(H, q, lo, hi):
    dt = H.dtype
    H, q, lo, hi = (_c(a, dt) for a in (H, q, lo, hi))
    N, m = q.shape
    x = np.empty((N, m), dt)
    iters = np.zeros(N, np.int32)
    args = [C.c_void_p(_p(a)) for a in (H, q, lo, hi, x, iters)]
    assert lib().tacd_oracle_pnqp(_dt(H), N, m, *args) == 0
    return x, iters
