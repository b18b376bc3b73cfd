"""ctypes mirror of include/tacd_ilqr.h (ABI version 2).

Only POD structs and constants live here; the CUDA library is loaded by
``tacdmpc_b200._lib`` and the CPU oracle (test infrastructure) by
``oracle/oracle.py`` — both bind the same structs.
"""
from __future__ import annotations

import ctypes as C

ABI_VERSION = 4
MAX_ALPHA = 8
MAX_NU = 16
MAX_NX = 64

OK, EINVAL, EUNSUPPORTED, ECUDA, EWORKSPACE = 0, -1, -2, -3, -4

F32, F64 = 0, 1
DYN_LTI, DYN_CARTPOLE, DYN_UNICYCLE, DYN_UNICYCLE_WRAPPED, DYN_EXTERNAL = 0, 1, 2, 3, 4
JAC_ANALYTIC, JAC_AUTO_DIFF, JAC_FINITE_DIFF = 0, 1, 2
FLAG_NONPD, FLAG_MAXITER = 1, 2

# alphas of the reference's parallel line search (controller.py:592)
DEFAULT_ALPHAS = (1.0, 0.8, 0.5, 0.2, 0.1)


class Problem(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("T", C.c_int32), ("nx", C.c_int32), ("nu", C.c_int32),
        ("dtype", C.c_int32),
        ("C_batched", C.c_int32), ("Cf_batched", C.c_int32),
        ("xref_batched", C.c_int32), ("uref_batched", C.c_int32),
        ("ls_cost_shared", C.c_int32),
        ("dyn_id", C.c_int32), ("jac_mode", C.c_int32),
        ("dyn_params", C.c_double * 8),
        ("dyn_A", C.c_void_p), ("dyn_B", C.c_void_p),
        ("dt", C.c_double), ("reg_eps", C.c_double), ("fd_eps", C.c_double),
        ("max_iter", C.c_int32), ("n_alpha", C.c_int32),
        ("alphas", C.c_double * MAX_ALPHA),
        ("has_umin", C.c_int32), ("has_umax", C.c_int32),
        ("umin", C.c_double * MAX_NU), ("umax", C.c_double * MAX_NU),
        ("C_diag", C.c_int32), ("compat_exact", C.c_int32),
    ]


class ForwardIO(C.Structure):
    _fields_ = [(name, C.c_void_p) for name in (
        "x0", "C", "c", "Cf", "cf", "C_ls", "c_ls", "Cf_ls", "cf_ls", "xref", "uref", "Uinit",
        "replay_iters", "replay_alpha", "replay_flags",
        "X", "U", "A", "Bm", "tight", "iters", "alpha_trace", "flags", "cost", "cost_trace")]


class BackwardIO(C.Structure):
    _fields_ = [(name, C.c_void_p) for name in (
        "X", "U", "C", "Cf", "xref", "uref", "tight", "A", "Bm", "gX", "gU",
        "grad_C", "grad_c", "grad_Cf", "grad_cf", "grad_x0", "grad_xref", "grad_uref", "dX", "dU")]


class EnvIO(C.Structure):
    """tacd_env_io (include/tacd_ilqr.h)."""
    _fields_ = [("state", C.c_void_p), ("action", C.c_void_p), ("reset_state", C.c_void_p), ("steps", C.c_void_p),
                ("obs", C.c_void_p), ("next_state", C.c_void_p), ("reward", C.c_void_p), ("terminated", C.c_void_p),
                ("truncated", C.c_void_p), ("max_steps", C.c_int32), ("clip_action", C.c_int32),
                ("reward_on_next", C.c_int32), ("reserved_", C.c_int32),
                ("term_limit", C.c_double * MAX_NX), ("w_x", C.c_double * MAX_NX), ("w_u", C.c_double * MAX_NU)]


def make_problem(*, B, T, nx, nu, dtype, dyn_id, dt, reg_eps=1e-6, max_iter=40,
                 C_batched=0, Cf_batched=0, xref_batched=1, uref_batched=1, ls_cost_shared=0,
                 jac_mode=JAC_ANALYTIC, fd_eps=1e-4, dyn_params=(), dyn_A=None, dyn_B=None,
                 alphas=DEFAULT_ALPHAS, umin=None, umax=None, C_diag=0, compat_exact=0) -> Problem:
    """Fill a ``tacd_problem``.  ``dyn_A``/``dyn_B`` are raw addresses (int) or None."""
    if nx > MAX_NX or nu > MAX_NU:
        raise ValueError(f"nx<={MAX_NX}, nu<={MAX_NU} supported, got nx={nx}, nu={nu}")
    if not (1 <= len(alphas) <= MAX_ALPHA):
        raise ValueError("1..8 line-search step sizes supported")
    p = Problem()
    p.B, p.T, p.nx, p.nu, p.dtype = int(B), int(T), int(nx), int(nu), int(dtype)
    p.C_batched, p.Cf_batched = int(C_batched), int(Cf_batched)
    p.xref_batched, p.uref_batched = int(xref_batched), int(uref_batched)
    p.ls_cost_shared = int(ls_cost_shared)
    p.C_diag = int(C_diag)
    p.compat_exact = int(compat_exact)
    p.dyn_id, p.jac_mode = int(dyn_id), int(jac_mode)
    for i, v in enumerate(dyn_params):
        p.dyn_params[i] = float(v)
    p.dyn_A = dyn_A
    p.dyn_B = dyn_B
    p.dt, p.reg_eps, p.fd_eps = float(dt), float(reg_eps), float(fd_eps)
    p.max_iter, p.n_alpha = int(max_iter), len(alphas)
    for i, a in enumerate(alphas):
        p.alphas[i] = float(a)
    p.has_umin = int(umin is not None)
    p.has_umax = int(umax is not None)
    if umin is not None:
        for i, v in enumerate(umin):
            p.umin[i] = float(v)
    if umax is not None:
        for i, v in enumerate(umax):
            p.umax[i] = float(v)
    return p
