"""Tensor-level wrappers of the C ABI (include/tacd_ilqr.h).

Each function takes CUDA tensors, allocates outputs / workspace with torch, and enqueues the
kernels on ``torch.cuda.current_stream()``.  Nothing here synchronises the host, and nothing
falls back to PyTorch math: a non-CUDA tensor or a missing library is an error.
"""
from __future__ import annotations

import ctypes as C
import functools
from dataclasses import dataclass, field
from typing import Optional, Sequence

import torch

from . import _abi, _lib

_DT = {torch.float32: _abi.F32, torch.float64: _abi.F64}


@dataclass
class Spec:
    """Problem family (everything of tacd_problem that does not depend on the tensors of one call)."""
    nx: int
    nu: int
    T: int
    dt: float
    dyn_id: int
    dyn_params: Sequence[float] = ()
    dyn_A: Optional[torch.Tensor] = None   # LTI, device tensors in the solve dtype
    dyn_B: Optional[torch.Tensor] = None
    reg_eps: float = 1e-6
    max_iter: int = 40
    jac_mode: int = _abi.JAC_ANALYTIC
    fd_eps: float = 1e-4
    alphas: Sequence[float] = _abi.DEFAULT_ALPHAS
    u_min: Optional[Sequence[float]] = None
    u_max: Optional[Sequence[float]] = None
    _keep: list = field(default_factory=list, repr=False)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.TacdError("tacdmpc_b200 runs on CUDA tensors only (there is no CPU path); got a "
                                 f"{t.device} tensor")


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _problem(spec: Spec, B: int, dtype, **layout) -> _abi.Problem:
    if dtype not in _DT:
        raise TypeError(f"solve dtype must be float32 or float64, got {dtype}")
    return _abi.make_problem(
        B=B, T=spec.T, nx=spec.nx, nu=spec.nu, dtype=_DT[dtype], dyn_id=spec.dyn_id, dt=spec.dt,
        reg_eps=spec.reg_eps, max_iter=spec.max_iter, jac_mode=spec.jac_mode, fd_eps=spec.fd_eps,
        dyn_params=spec.dyn_params, dyn_A=_ptr(spec.dyn_A), dyn_B=_ptr(spec.dyn_B), alphas=spec.alphas,
        umin=spec.u_min, umax=spec.u_max, **layout)


def _prep(t, dtype, device):
    return t.detach().to(device=device, dtype=dtype).contiguous()


def _layout(B, Cm, Cf, xref, uref, C_diag=False):
    # compact diagonal costs (tacd_problem.C_diag) are always per-element: C[B,T,n], C_final[B,n]
    return dict(C_batched=int(Cm.ndim == (3 if C_diag else 4)), Cf_batched=int(Cf.ndim == (2 if C_diag else 3)),
                xref_batched=int(xref.shape[0] == B), uref_batched=int(uref.shape[0] == B), C_diag=int(C_diag))


def ilqr_forward(spec: Spec, x0, Cm, c, Cf, cf, xref, uref, Uinit, *, ls_cost=None, replay=None,
                 want_AB=False, want_trace=True, want_cost_trace=False, C_diag=False):
    """solve_step (controller.py:275-315) on the GPU.  Returns a dict of device tensors:
    X, U, tight (uint8), iters (int32), flags (uint8), cost, [alpha_trace, A, Bm, cost_trace].
    ``C_diag``: ``Cm [B,T,n]`` / ``Cf [B,n]`` (and ``ls_cost``) hold diagonals (the ActorMPC layout)."""
    _need_cuda(x0, Cm, c, Cf, cf, xref, uref, Uinit)
    dev, dt = x0.device, x0.dtype
    x0, Cm, c, Cf, cf, xref, uref, Uinit = (_prep(t, dt, dev) for t in (x0, Cm, c, Cf, cf, xref, uref, Uinit))
    B, T, nx, nu = x0.shape[0], spec.T, spec.nx, spec.nu
    if Uinit.shape != (B, T, nu):
        raise RuntimeError(f"U_init must be [B={B}, T={T}, nu={nu}], got {tuple(Uinit.shape)}")
    if xref.ndim != 3 or xref.shape[1:] != (T + 1, nx) or xref.shape[0] not in (1, B):
        raise RuntimeError(f"x_ref must be [B or 1, T+1={T + 1}, nx={nx}], got {tuple(xref.shape)}")
    if uref.ndim != 3 or uref.shape[1:] != (T, nu) or uref.shape[0] not in (1, B):
        raise RuntimeError(f"u_ref must be [B or 1, T={T}, nu={nu}], got {tuple(uref.shape)}")
    if C_diag and (Cm.shape != (B, T, nx + nu) or Cf.shape != (B, nx + nu)):
        raise RuntimeError(f"C_diag costs must be C[B,T,n], C_final[B,n]; got {tuple(Cm.shape)}, {tuple(Cf.shape)}")
    p = _problem(spec, B, dt, ls_cost_shared=int(ls_cost is not None), **_layout(B, Cm, Cf, xref, uref, C_diag))
    io = _abi.ForwardIO()
    keep = [x0, Cm, c, Cf, cf, xref, uref, Uinit]
    out = dict(X=torch.empty(B, T + 1, nx, device=dev, dtype=dt), U=torch.empty(B, T, nu, device=dev, dtype=dt),
               tight=torch.empty(B, T, nu, device=dev, dtype=torch.uint8),
               iters=torch.empty(B, device=dev, dtype=torch.int32), flags=torch.empty(B, device=dev, dtype=torch.uint8),
               cost=torch.empty(B, device=dev, dtype=dt))
    if want_trace:
        out["alpha_trace"] = torch.empty(B, max(p.max_iter, 1), device=dev, dtype=torch.uint8)
    if want_AB:
        out["A"] = torch.empty(B, T, nx, nx, device=dev, dtype=dt)
        out["Bm"] = torch.empty(B, T, nx, nu, device=dev, dtype=dt)
    if want_cost_trace:
        out["cost_trace"] = torch.full((B, max(p.max_iter, 1), p.n_alpha + 2), float("nan"), device=dev, dtype=dt)
    io.x0, io.C, io.c, io.Cf, io.cf = _ptr(x0), _ptr(Cm), _ptr(c), _ptr(Cf), _ptr(cf)
    io.xref, io.uref, io.Uinit = _ptr(xref), _ptr(uref), _ptr(Uinit)
    if ls_cost is not None:
        ls = [_prep(t, dt, dev) for t in ls_cost]
        keep += ls
        io.C_ls, io.c_ls, io.Cf_ls, io.cf_ls = (_ptr(t) for t in ls)
    if replay is not None:
        r = [replay["iters"].to(device=dev, dtype=torch.int32).contiguous(),
             replay["alpha_trace"].to(device=dev, dtype=torch.uint8).contiguous(),
             replay["flags"].to(device=dev, dtype=torch.uint8).contiguous()]
        keep += r
        io.replay_iters, io.replay_alpha, io.replay_flags = (_ptr(t) for t in r)
    for k, v in out.items():
        setattr(io, k, _ptr(v))
    L = _lib.lib()
    ws_bytes = int(L.tacd_forward_workspace_bytes(C.byref(p)))
    ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
    with torch.cuda.device(dev):
        _lib.check(L.tacd_ilqr_forward(C.byref(p), C.byref(io), C.c_void_p(ws.data_ptr()), ws_bytes, _stream()),
                   "tacd_ilqr_forward")
    # inputs / workspace were allocated on this stream: the caching allocator re-uses their blocks in
    # stream order, so they may be dropped by Python while the kernel is still queued
    del keep, ws
    return out


_GRAD_NAMES = ("grad_C", "grad_c", "grad_Cf", "grad_cf", "grad_x0", "grad_xref", "grad_uref", "dX", "dU")


def ilqr_backward(spec: Spec, X, U, Cm, xref, uref, tight, gX, gU, *, A=None, Bm=None, Cf_batched=False,
                  want=("grad_C", "grad_c", "grad_Cf", "grad_cf", "grad_x0", "grad_xref", "grad_uref"), C_diag=False,
                  exact=False, Cf=None, out=None):
    """ILQRSolve.backward (controller.py:63-115) on the GPU.  Outputs for shared (un-batched) inputs are
    already summed over the batch.  ``C_diag``: ``Cm [B,T,n]`` holds diagonals; ``grad_C [B,T,n]`` and
    ``grad_Cf [B,n]`` come back as diagonals too.  ``exact`` (needs ``Cf``): exact implicit differentiation instead
    of the reference's (``tacd_problem.compat_exact``).  ``out``: optional ``{name: tensor}`` of caller-owned result buffers
    (e.g. views into ONE flat bucket that is then all-reduced across ranks without a gather copy)."""
    _need_cuda(X, U, Cm, xref, uref, tight, gX, gU, A, Bm)
    dev, dt = X.device, X.dtype
    X, U, Cm, xref, uref, gX, gU = (_prep(t, dt, dev) for t in (X, U, Cm, xref, uref, gX, gU))
    tight = tight.to(device=dev, dtype=torch.uint8).contiguous()
    B, T, nx, nu = X.shape[0], spec.T, spec.nx, spec.nu
    n = nx + nu
    Cb, xb, ub = Cm.ndim == (3 if C_diag else 4), xref.shape[0] == B, uref.shape[0] == B
    if C_diag:
        Cf_batched = True
    if exact:
        if Cf is None:
            raise RuntimeError("exact backward needs the terminal cost Cf")
        Cf = _prep(Cf, dt, dev)
        Cf_batched = Cf.ndim == (2 if C_diag else 3)
    p = _problem(spec, B, dt, C_batched=int(Cb), Cf_batched=int(Cf_batched), xref_batched=int(xb), uref_batched=int(ub),
                 C_diag=int(C_diag), compat_exact=int(exact))
    io = _abi.BackwardIO()
    if exact:
        io.Cf = _ptr(Cf)
    keep = [X, U, Cm, xref, uref, tight, gX, gU, Cf]
    io.X, io.U, io.C, io.xref, io.uref, io.tight, io.gX, io.gU = (_ptr(t) for t in keep[:8])
    if A is not None:
        A, Bm = _prep(A, dt, dev), _prep(Bm, dt, dev)
        keep += [A, Bm]
        io.A, io.Bm = _ptr(A), _ptr(Bm)
    lead = lambda b: (B,) if b else ()
    shapes = dict(grad_C=lead(Cb) + ((T, n) if C_diag else (T, n, n)), grad_c=lead(Cb) + (T, n),
                  grad_Cf=lead(Cf_batched) + ((n,) if C_diag else (n, n)),
                  grad_cf=lead(Cf_batched) + (n,), grad_x0=(B, nx), grad_xref=((B,) if xb else (1,)) + (T + 1, nx),
                  grad_uref=((B,) if ub else (1,)) + (T, nu), dX=(B, T + 1, nx), dU=(B, T, nu))
    given = out or {}
    out = {}
    for k in want:
        t = given.get(k)
        if t is None:
            t = torch.empty(shapes[k], device=dev, dtype=dt)
        elif t.numel() != torch.Size(shapes[k]).numel() or t.dtype != dt or t.device != dev or not t.is_contiguous():
            raise RuntimeError(f"out[{k!r}] must be a contiguous {dt} tensor of {torch.Size(shapes[k]).numel()} elements on {dev}")
        out[k] = t
    for k, v in out.items():
        setattr(io, k, _ptr(v))
    L = _lib.lib()
    ws_bytes = int(L.tacd_backward_workspace_bytes(C.byref(p)))
    ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
    with torch.cuda.device(dev):
        _lib.check(L.tacd_ilqr_backward(C.byref(p), C.byref(io), C.c_void_p(ws.data_ptr()), ws_bytes, _stream()),
                   "tacd_ilqr_backward")
    del keep, ws
    return out


def _vp(*ts):
    return [C.c_void_p(_ptr(t)) for t in ts]


def rollout(spec: Spec, x0, U):
    _need_cuda(x0, U)
    dev, dt = x0.device, x0.dtype
    x0, U = _prep(x0, dt, dev), _prep(U, dt, dev)
    B = x0.shape[0]
    X = torch.empty(B, spec.T + 1, spec.nx, device=dev, dtype=dt)
    p = _problem(spec, B, dt)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().tacd_rollout(C.byref(p), *_vp(x0, U, X), _stream()), "tacd_rollout")
    return X


def linearize(spec: Spec, X, U):
    _need_cuda(X, U)
    dev, dt = X.device, X.dtype
    X, U = _prep(X, dt, dev), _prep(U, dt, dev)
    B, T, nx, nu = X.shape[0], spec.T, spec.nx, spec.nu
    A = torch.empty(B, T, nx, nx, device=dev, dtype=dt)
    Bm = torch.empty(B, T, nx, nu, device=dev, dtype=dt)
    p = _problem(spec, B, dt)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().tacd_linearize(C.byref(p), *_vp(X, U, A, Bm), _stream()), "tacd_linearize")
    return A, Bm


def objective(spec: Spec, X, U, Cm, c, Cf, cf, xref, uref):
    _need_cuda(X, U, Cm, c, Cf, cf, xref, uref)
    dev, dt = X.device, X.dtype
    X, U, Cm, c, Cf, cf, xref, uref = (_prep(t, dt, dev) for t in (X, U, Cm, c, Cf, cf, xref, uref))
    B = X.shape[0]
    cost = torch.empty(B, device=dev, dtype=dt)
    p = _problem(spec, B, dt, **_layout(B, Cm, Cf, xref, uref))
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().tacd_objective(C.byref(p), *_vp(X, U, Cm, c, Cf, cf, xref, uref, cost), _stream()),
                   "tacd_objective")
    return cost


def riccati(spec: Spec, A, Bm, lx, lu, lxx, lxu, luu, lxN, lxxN):
    _need_cuda(A, Bm, lx, lu, lxx, lxu, luu, lxN, lxxN)
    dev, dt = A.device, A.dtype
    ts = [_prep(t, dt, dev) for t in (A, Bm, lx, lu, lxx, lxu, luu, lxN, lxxN)]
    B, T, nx, nu = A.shape[0], spec.T, spec.nx, spec.nu
    K = torch.empty(B, T, nu, nx, device=dev, dtype=dt)
    k = torch.empty(B, T, nu, device=dev, dtype=dt)
    flags = torch.empty(B, device=dev, dtype=torch.uint8)
    p = _problem(spec, B, dt)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().tacd_riccati(C.byref(p), *_vp(*ts, K, k, flags), _stream()), "tacd_riccati")
    return K, k, flags


def linesearch(spec: Spec, x0, X, U, K, k, Cm, c, Cf, cf, xref, uref):
    _need_cuda(x0, X, U, K, k, Cm, c, Cf, cf, xref, uref)
    dev, dt = x0.device, x0.dtype
    ts = [_prep(t, dt, dev) for t in (x0, X, U, K, k, Cm, c, Cf, cf, xref, uref)]
    B, T, nx, nu = x0.shape[0], spec.T, spec.nx, spec.nu
    p = _problem(spec, B, dt, **_layout(B, ts[5], ts[7], ts[9], ts[10]))
    na = p.n_alpha
    Xc = torch.empty(B, na, T + 1, nx, device=dev, dtype=dt)
    Uc = torch.empty(B, na, T, nu, device=dev, dtype=dt)
    costs = torch.empty(B, na, device=dev, dtype=dt)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().tacd_linesearch(C.byref(p), *_vp(*ts, Xc, Uc, costs), _stream()), "tacd_linesearch")
    return Xc, Uc, costs


def pnqp(H, q, lo, hi):
    """Independent box QPs: H[N,m,m], q/lo/hi[N,m] -> x[N,m], iters[N] (utils.py:66-124)."""
    _need_cuda(H, q, lo, hi)
    dev, dt = H.device, H.dtype
    H, q, lo, hi = (_prep(t, dt, dev) for t in (H, q, lo, hi))
    N, m = q.shape
    x = torch.empty(N, m, device=dev, dtype=dt)
    iters = torch.empty(N, device=dev, dtype=torch.int32)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().tacd_pnqp(_DT[dt], N, m, *_vp(H, q, lo, hi, x, iters), _stream()), "tacd_pnqp")
    return x, iters


def closed_loop_advance(spec: Spec, k: int, n_sim: int, x, U_opt, Xs, Us, U_warm=None, xref_full=None, uref_full=None,
                        xref_win=None, uref_win=None):
    """One receding-horizon step after the solve of simulation step ``k`` (``tacd_closed_loop_advance``): applies
    ``U_opt[:,0]`` to ``x`` in place, records ``Xs[:,k+1]``, ``Us[:,k]``, writes the shifted warm start and the
    reference windows of the next solve.  All tensors are caller-owned, contiguous, on the device."""
    _need_cuda(x, U_opt, Xs, Us, U_warm, xref_full, uref_full, xref_win, uref_win)
    dt = x.dtype
    B = x.shape[0]
    for t in (x, U_opt, Xs, Us, U_warm, xref_full, uref_full, xref_win, uref_win):
        if t is not None and (t.dtype != dt or not t.is_contiguous()):
            raise RuntimeError("closed_loop_advance: tensors must be contiguous and of one dtype")
    p = _problem(spec, B, dt, xref_batched=int(xref_full is not None and xref_full.shape[0] == B),
                 uref_batched=int(uref_full is not None and uref_full.shape[0] == B))
    ref_len = int(xref_full.shape[1]) if xref_full is not None else 0
    if xref_full is not None and uref_full is not None and int(uref_full.shape[1]) != ref_len - 1:
        raise RuntimeError(f"closed_loop_advance: uref_full must hold ref_len - 1 = {ref_len - 1} steps (the kernel strides it by "
                           f"that), got {int(uref_full.shape[1])}")
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().tacd_closed_loop_advance(C.byref(p), int(k), int(n_sim), *_vp(x, U_opt, Xs, Us, U_warm, xref_full, uref_full),
                                                       ref_len, *_vp(xref_win, uref_win), _stream()), "tacd_closed_loop_advance")


def al_update(X, xref, x_min, x_max, lam_lo, lam_hi, rho: float, rho_next: float, update_lambda: bool = True):
    """Augmented-Lagrangian multiplier update + quadratic penalty model for state box constraints
    (``tacd_al_update``).  ``lam_lo``/``lam_hi`` [B,T+1,nx] are updated in place; returns ``(dq, dp, viol)``:
    additions to the diagonal / linear state cost of every step (t = T: terminal cost) and the largest violation
    per element."""
    _need_cuda(X, xref, x_min, x_max, lam_lo, lam_hi)
    dt, dev = X.dtype, X.device
    B, T1, nx = X.shape
    X, xref = _prep(X, dt, dev), _prep(xref, dt, dev)
    x_min, x_max = _prep(x_min, dt, dev).expand(nx).contiguous(), _prep(x_max, dt, dev).expand(nx).contiguous()
    if lam_lo.dtype != dt or lam_hi.dtype != dt or not lam_lo.is_contiguous() or not lam_hi.is_contiguous():
        raise RuntimeError("multipliers must be contiguous tensors of the solve dtype")
    dq, dp = torch.empty_like(X), torch.empty_like(X)
    viol = torch.empty(B, device=dev, dtype=dt)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().tacd_al_update(_DT[dt], B, T1 - 1, nx, *_vp(X, xref), int(xref.shape[0] == B), *_vp(x_min, x_max, lam_lo, lam_hi),
                                             float(rho), float(rho_next), int(update_lambda), *_vp(dq, dp, viol), _stream()), "tacd_al_update")
    return dq, dp, viol


# ---- PPO glue ----------------------------------------------------------------------------------
@functools.lru_cache(maxsize=64)
def _env_vecs(term_limit, w_x, w_u):
    """The fixed-size host arrays of tacd_env_io, built once per environment definition."""
    def arr(n, vals, fill):
        return (C.c_double * n)(*[float(vals[i]) if vals is not None and i < len(vals) else fill for i in range(n)])
    return arr(_abi.MAX_NX, term_limit, float("inf")), arr(_abi.MAX_NX, w_x, 0.0), arr(_abi.MAX_NU, w_u, 0.0)


def env_step(spec: Spec, state, action, *, steps=None, reset_state=None, max_steps: int = 0, term_limit=None,
             w_x=None, w_u=None, clip_action: bool = False, reward_on_next: bool = True, next_state=None):
    """One step of ``B`` environments driven by the dynamics plugin of ``spec`` (``tacd_env_step``;
    ACMPC/parallel_env.py:34-54 without the Python loop).  ``steps`` (int32 [B]) is advanced in place and zeroed
    where an episode ended and ``reset_state`` supplied the next start.  Returns
    ``(obs, next_state, reward, terminated, truncated)``; ``next_state`` may be passed in (it may be ``state``)."""
    _need_cuda(state, action, steps, reset_state)
    dt, dev = state.dtype, state.device
    B, nx = state.shape
    state, action = _prep(state, dt, dev), _prep(action, dt, dev).reshape(B, spec.nu)
    if reset_state is not None:
        reset_state = _prep(reset_state, dt, dev)
    if steps is not None and (steps.dtype != torch.int32 or not steps.is_contiguous()):
        raise RuntimeError("steps must be a contiguous int32 tensor")
    obs = torch.empty(B, nx, device=dev, dtype=dt)
    if next_state is None:
        next_state = torch.empty(B, nx, device=dev, dtype=dt)
    reward = torch.empty(B, device=dev, dtype=dt)
    term = torch.empty(B, device=dev, dtype=torch.uint8)
    trunc = torch.empty(B, device=dev, dtype=torch.uint8)
    io = _abi.EnvIO()
    io.state, io.action, io.reset_state, io.steps = _ptr(state), _ptr(action), _ptr(reset_state), _ptr(steps)
    io.obs, io.next_state, io.reward, io.terminated, io.truncated = _ptr(obs), _ptr(next_state), _ptr(reward), _ptr(term), _ptr(trunc)
    io.max_steps, io.clip_action, io.reward_on_next = int(max_steps), int(clip_action), int(reward_on_next)
    io.term_limit, io.w_x, io.w_u = _env_vecs(None if term_limit is None else tuple(term_limit), None if w_x is None else tuple(w_x),
                                              None if w_u is None else tuple(w_u))
    p = _problem(spec, B, dt)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().tacd_env_step(C.byref(p), C.byref(io), _stream()), "tacd_env_step")
    return obs, next_state, reward, term.bool(), trunc.bool()


def gae(rewards, values, gamma: float = 0.99, lam: float = 1.0):
    """``(returns, advantages)`` of ACMPC/training_loop.py:30-41 for ``[N, T]`` rewards / values (``tacd_gae``)."""
    _need_cuda(rewards, values)
    dt, dev = rewards.dtype, rewards.device
    if rewards.ndim != 2 or values.shape != rewards.shape:
        raise RuntimeError(f"rewards, values must be [N, T]; got {tuple(rewards.shape)}, {tuple(values.shape)}")
    rewards, values = _prep(rewards, dt, dev), _prep(values, dt, dev)
    N, T = rewards.shape
    ret, adv = torch.empty_like(rewards), torch.empty_like(rewards)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().tacd_gae(_DT[dt], N, T, *_vp(rewards, values), float(gamma), float(lam), *_vp(ret, adv), _stream()), "tacd_gae")
    return ret, adv


def discounted_targets(rewards, bootstrap=None, gamma: float = 0.99):
    """``target_t = r_t + gamma target_{t+1}``, ``target_H = bootstrap`` for ``[N, H]`` rewards (``tacd_discounted_targets``;
    the value-expansion targets of ACMPC/training_loop.py:162-170)."""
    _need_cuda(rewards, bootstrap)
    dt, dev = rewards.dtype, rewards.device
    if rewards.ndim != 2 or (bootstrap is not None and bootstrap.shape != rewards.shape[:1]):
        raise RuntimeError("rewards must be [N, H] and bootstrap [N]")
    rewards = _prep(rewards, dt, dev)
    bootstrap = None if bootstrap is None else _prep(bootstrap, dt, dev)
    N, H = rewards.shape
    out = torch.empty_like(rewards)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().tacd_discounted_targets(_DT[dt], N, H, *_vp(rewards, bootstrap), float(gamma), *_vp(out), _stream()),
                   "tacd_discounted_targets")
    return out


def fp32_probe(kind: int = 0, iters: int = 4096, blocks: int = 0, device=None, reps: int = 5):
    """Measured FP32 throughput of this GPU in TFLOP/s (``tacd_fp32_probe``): ``kind`` 0 = scalar FFMA, 1 = packed FFMA2.
    Best of ``reps`` CUDA-event timings after a warm-up launch.  Returns ``(tflops, ms)``."""
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if blocks <= 0:
        blocks = torch.cuda.get_device_properties(device).multi_processor_count * 8
    seed = torch.tensor([0.999999, 1e-6, 1.0], device=device, dtype=torch.float32)
    out = torch.empty(blocks * 256, device=device, dtype=torch.float32)
    flops = C.c_double(0.0)
    best = float("inf")
    with torch.cuda.device(device):
        for r in range(reps + 1):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _lib.check(_lib.lib().tacd_fp32_probe(int(kind), int(iters), int(blocks), C.c_void_p(seed.data_ptr()), C.c_void_p(out.data_ptr()),
                                                  C.byref(flops), _stream()), "tacd_fp32_probe")
            b.record()
            b.synchronize()
            if r > 0:
                best = min(best, a.elapsed_time(b))
    return flops.value / (best * 1e-3) / 1e12, best
