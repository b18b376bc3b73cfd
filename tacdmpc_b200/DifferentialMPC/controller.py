"""``DifferentiableMPCController`` / ``ILQRSolve`` with the reference's interface
(DifferentialMPC/controller.py), executing on hand-written sm_100a kernels.

What stays the same: constructor keyword arguments (live and dead ones), ``forward(x0,
U_init=None) -> (X*, U*)``, ``solve_step``, ``reset``, the autograd function's argument order and its
9-tuple of gradients, the side effects on ``cost_module`` and on the ``*_last`` attributes, the three
``grad_method`` modes and ``ValueError`` when ``analytic`` comes without ``f_dyn_jac``.

What is different underneath:
  * the whole iLQR solve (all iterations, line search, accept/reject, final linearisation, active-set
    mask) is ONE kernel launch with per-element early exit and no host synchronisation
    (``tacd_ilqr_forward``), when ``f_dyn`` is one of :mod:`tacdmpc_b200.dynamics`;
  * the backward pass is one launch of the KKT kernel for the whole batch (``tacd_ilqr_backward``)
    instead of a Python loop over batch elements;
  * an arbitrary Python ``f_dyn`` takes the *generic path*: dynamics and Jacobians stay in torch, the
    Riccati recursion, the cost evaluations and the backward run in CUDA.

Reference quirks Q1-Q12 are reproduced under ``compat="reference"`` (default).
"""
from __future__ import annotations

import enum
from typing import Callable, Optional, Tuple

import torch
from torch import Tensor

from .. import _abi, _lib, ops
from ..dynamics import NativeDynamics
from .cost import DiagQuadCost, GeneralQuadCost
from .AugmentedLagrangianManager import ConstrainedQuadCost
from .utils import batched_jacobian, jacobian_finite_diff_batched


class GradMethod(enum.Enum):
    """How A_t = df/dx, B_t = df/du are obtained (controller.py:119-129)."""
    ANALYTIC = "analytic"
    AUTO_DIFF = "auto_diff"
    FINITE_DIFF = "finite_diff"


_JAC = {GradMethod.ANALYTIC: _abi.JAC_ANALYTIC, GradMethod.AUTO_DIFF: _abi.JAC_AUTO_DIFF,
        GradMethod.FINITE_DIFF: _abi.JAC_FINITE_DIFF}


def _solve_dtype(x0: Tensor) -> torch.dtype:
    # AMP: half / bfloat16 inputs are solved in fp32 (the reference would run its Riccati in reduced
    # precision under autocast; SURVEY §5 "AMP" documents this deliberate deviation)
    return torch.float64 if x0.dtype == torch.float64 else torch.float32


class ILQRSolve(torch.autograd.Function):
    """forward(x0, C, c, C_final, c_final, x_ref, u_ref, controller, U_init) -> (X*, U*)  (controller.py:25-115)"""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x0, C, c, C_final, c_final, x_ref, u_ref, controller, U_init):
        cm = controller.cost_module
        # side effects the callers rely on (controller.py:37-41)
        cm.C, cm.c, cm.C_final, cm.c_final = C, c, C_final, c_final
        cm.set_reference(x_ref, u_ref)

        X, U = controller.solve_step(x0, U_init)
        if (not controller.converged) and controller.detach_unconverged:
            X, U = X.detach(), U.detach()

        ctx.controller = controller
        ctx.spec = controller._last_spec
        ctx.Cf_batched = controller._last_Cf.ndim == 3   # what the kernel was given (_costs squeezes a batch of 1)
        ctx.exact = controller.compat == "exact"
        ctx.Cf = controller._last_Cf if ctx.exact else None
        ctx.shapes = tuple(t.shape for t in (x0, C, c, C_final, c_final, x_ref, u_ref))
        ctx.in_dtypes = tuple(t.dtype for t in (x0, C, c, C_final, c_final, x_ref, u_ref))
        saved = [X, U, controller._last_C, controller._last_xref, controller._last_uref, controller._last_tight_u8]
        if controller._last_AB is not None:
            saved += list(controller._last_AB)
        ctx.save_for_backward(*saved)
        return X, U

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_X, grad_U):
        saved = ctx.saved_tensors
        X, U, Cm, xref, uref, tight = saved[:6]
        A, Bm = (saved[6], saved[7]) if len(saved) > 6 else (None, None)
        need = ctx.needs_input_grad
        names = ("grad_x0", "grad_C", "grad_c", "grad_Cf", "grad_cf", "grad_xref", "grad_uref")
        want = tuple(nm for nm, nd in zip(names, need[:7]) if nd)
        out = ops.ilqr_backward(ctx.spec, X, U, Cm, xref, uref, tight,
                                grad_X.to(X.dtype), grad_U.to(U.dtype), A=A, Bm=Bm,
                                Cf_batched=ctx.Cf_batched, want=want, exact=ctx.exact, Cf=ctx.Cf) if want else {}
        grads = []
        for nm, shape, dt in zip(names, ctx.shapes, ctx.in_dtypes):
            g = out.get(nm)
            if g is not None:
                g = g.reshape(shape).to(dt)
            grads.append(g)
        return (*grads, None, None)


class ILQRSolveDiag(torch.autograd.Function):
    """ILQRSolve for a ``DiagQuadCost``: forward(x0, q, p, q_final, p_final, x_ref, u_ref, controller, U_init).
    Same solve and same implicit differentiation (controller.py:25-115); the cost enters and its gradient
    leaves as diagonals, never expanded to ``[B,T,n,n]``."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x0, q, p, q_final, p_final, x_ref, u_ref, controller, U_init):
        cm = controller.cost_module
        cm.set_diag(q, p, q_final, p_final)
        cm.set_reference(x_ref, u_ref)
        X, U = controller.solve_step(x0, U_init)
        if (not controller.converged) and controller.detach_unconverged:
            X, U = X.detach(), U.detach()
        ctx.spec = controller._last_spec
        ctx.diag = controller._last_diag
        ctx.Cf_batched = controller._last_Cf.ndim == (2 if ctx.diag else 3)
        ctx.exact = controller.compat == "exact"
        ctx.Cf = controller._last_Cf if ctx.exact else None
        ctx.shapes = tuple(t.shape for t in (x0, q, p, q_final, p_final, x_ref, u_ref))
        ctx.in_dtypes = tuple(t.dtype for t in (x0, q, p, q_final, p_final, x_ref, u_ref))
        saved = [X, U, controller._last_C, controller._last_xref, controller._last_uref, controller._last_tight_u8]
        if controller._last_AB is not None:
            saved += list(controller._last_AB)
        ctx.save_for_backward(*saved)
        return X, U

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_X, grad_U):
        saved = ctx.saved_tensors
        X, U, Cm, xref, uref, tight = saved[:6]
        A, Bm = (saved[6], saved[7]) if len(saved) > 6 else (None, None)
        need = ctx.needs_input_grad
        names = ("grad_x0", "grad_C", "grad_c", "grad_Cf", "grad_cf", "grad_xref", "grad_uref")
        want = tuple(nm for nm, nd in zip(names, need[:7]) if nd)
        out = ops.ilqr_backward(ctx.spec, X, U, Cm, xref, uref, tight, grad_X.to(X.dtype), grad_U.to(U.dtype),
                                A=A, Bm=Bm, Cf_batched=ctx.Cf_batched, want=want, C_diag=ctx.diag, exact=ctx.exact, Cf=ctx.Cf) if want else {}
        grads = []
        for nm, shape, dt in zip(names, ctx.shapes, ctx.in_dtypes):
            g = out.get(nm)
            if g is not None:
                if not ctx.diag and nm in ("grad_C", "grad_Cf"):   # dense fallback path: keep the diagonal
                    g = torch.diagonal(g, dim1=-2, dim2=-1)
                g = g.reshape(shape).to(dt)
            grads.append(g)
        return (*grads, None, None)


class DifferentiableMPCController(torch.nn.Module):
    """Batched differentiable iLQR-MPC (controller.py:132-824)."""

    def __init__(
            self,
            f_dyn: Callable,
            total_time: float,
            step_size: float,
            horizon: int,
            cost_module: torch.nn.Module,
            u_min: Optional[torch.Tensor] = None,
            u_max: Optional[torch.Tensor] = None,
            reg_eps: float = 1e-6,
            device: str = "cuda:0",
            N_sim: Optional[int] = None,
            grad_method: GradMethod | str = GradMethod.AUTO_DIFF,
            f_dyn_jac: Optional[Callable] = None,
            fd_eps: float = 1e-4,
            max_iter: int = 40,
            tol_x: float = 1e-6,
            tol_u: float = 1e-6,
            exit_unconverged: bool = False,
            detach_unconverged: bool = True,
            converge_tol: float = 1e-6,
            delta_u: Optional[float] = None,
            best_cost_eps: float = 1e-6,
            not_improved_lim: int = 10,
            verbose: int = 0,
            compat: str = "reference",
    ):
        super().__init__()
        self.device = torch.device(device)
        self.f_dyn = f_dyn
        self.total_time = total_time
        self.dt = step_size
        self.horizon = horizon
        self.cost_module = cost_module
        self.nx, self.nu = cost_module.nx, cost_module.nu
        self.u_min = u_min.to(self.device) if u_min is not None else None
        self.u_max = u_max.to(self.device) if u_max is not None else None
        self._bounds_cache = {}
        self.reg_eps = reg_eps
        self.N_sim = N_sim if N_sim is not None else horizon
        if isinstance(grad_method, str):
            grad_method = GradMethod(grad_method.lower())
        self.grad_method = grad_method
        self.f_dyn_jac = f_dyn_jac
        self.fd_eps = fd_eps
        if self.grad_method is GradMethod.ANALYTIC and self.f_dyn_jac is None and not isinstance(f_dyn, NativeDynamics):
            raise ValueError("grad_method='analytic' needs f_dyn_jac(x, u, dt) -> (A, B)")
        # accepted and ignored exactly as the reference does (dead on its live path)
        self.delta_u = delta_u
        self.best_cost_eps = best_cost_eps
        self.not_improved_lim = not_improved_lim
        self.verbose = verbose
        self.max_iter = int(max_iter)
        self.tol_u, self.tol_x = float(tol_u), float(tol_x)
        self.exit_unconverged = exit_unconverged
        self.detach_unconverged = bool(detach_unconverged)
        self.converge_tol = converge_tol
        self.converged = True
        if compat not in ("reference", "exact"):
            raise ValueError("compat must be 'reference' or 'exact'")
        self.compat = compat
        # multi-GPU: the sharding helper installs the GLOBAL element 0's cost here (quirk Q4)
        self.ls_cost_override = None
        self.force_dense_cost = False   # DiagQuadCost: expand to dense [B,T,n,n] before the kernels (A/B tests)
        self._last_diag = False
        self.U_prev = None
        self.reset()
        self.converged = True

    # ------------------------------------------------------------------ state of the last solve
    def reset(self) -> None:
        """Drop everything remembered from the last solve (controller.py:792-806)."""
        if getattr(self, "verbose", 0) > 0:
            print("Resetting MPC controller internal state.")
        self.U_last = self.X_last = self.H_last = self.lmb_last = self.tight_mask_last = None
        self.iters_last = self.flags_last = self.cost_last = self.alpha_trace_last = None
        self._F_last = None
        self._last_spec = self._last_C = self._last_xref = self._last_uref = self._last_tight_u8 = self._last_AB = None
        self._last_Cf = None
        self.converged = None

    @property
    def F_last(self):
        """(A, B) at the last solution.  The fused kernel does not write them unless asked; they are
        produced on first access by ``tacd_linearize``."""
        if self._F_last is None and self.X_last is not None and self._last_spec is not None:
            self._F_last = ops.linearize(self._last_spec, self.X_last, self.U_last)
        return self._F_last

    @F_last.setter
    def F_last(self, v):
        self._F_last = v

    # ------------------------------------------------------------------ problem description
    def _is_native(self) -> bool:
        return isinstance(self.f_dyn, NativeDynamics) and type(self.cost_module) in (GeneralQuadCost, DiagQuadCost, ConstrainedQuadCost)

    def _diag_eligible(self) -> bool:
        """Compact diagonal costs go to the kernels as they are when the fused five-lanes-per-element forward
        kernel takes the problem (shipped non-LTI plugin or small LTI, analytic / auto-diff Jacobians)."""
        return (isinstance(self.cost_module, DiagQuadCost) and isinstance(self.f_dyn, NativeDynamics)
                and self.nx + self.nu <= 5 and self.nx <= 4 and self.grad_method != GradMethod.FINITE_DIFF
                and not self.force_dense_cost)

    def _spec(self, dtype, device) -> ops.Spec:
        f = self.f_dyn
        native = isinstance(f, NativeDynamics)
        dynA = dynB = None
        if native:
            if (f.nx, f.nu) != (self.nx, self.nu):
                raise ValueError(f"dynamics are nx={f.nx}, nu={f.nu}; cost module says nx={self.nx}, nu={self.nu}")
            dynA, dynB = f.device_matrices(device, dtype)

        def bounds(b):
            # The kernels take the box as launch constants (tacd_problem.umin/umax).  The tensors live on the
            # device (controller.py:187-188 keeps them there), so reading them is a device->host copy + sync:
            # done once per tensor version, not per solve (a per-call sync would stall the stream and break
            # CUDA-graph capture).
            if b is None:
                return None
            key = (id(b), b._version)
            hit = self._bounds_cache.get(key)
            if hit is None:
                hit = [float(v) for v in b.detach().to("cpu", torch.float64).flatten().expand(self.nu).tolist()]
                if len(self._bounds_cache) > 8:
                    self._bounds_cache.clear()
                self._bounds_cache[key] = hit
            return hit

        return ops.Spec(nx=self.nx, nu=self.nu, T=self.horizon, dt=float(self.dt),
                        dyn_id=f.dyn_id if native else _abi.DYN_EXTERNAL, dyn_params=tuple(getattr(f, "params", ())),
                        dyn_A=dynA, dyn_B=dynB, reg_eps=float(self.reg_eps), max_iter=self.max_iter,
                        jac_mode=_JAC[self.grad_method], fd_eps=1e-4,  # quirk Q9: self.fd_eps is never read
                        u_min=bounds(self.u_min), u_max=bounds(self.u_max))

    def _ls_override(self, dtype, device, diag: bool):
        """The installed line-search cost (sharding.broadcast_line_search_cost) in the layout of the path that is about
        to run: compact diagonals ([T,n], [T,n], [n], [n]) for the C_diag kernels, dense ([T,n,n], [T,n], [n,n], [n])
        otherwise.  An override of the other layout is converted (a compact one is expanded; a dense one is reduced to
        its diagonal, which is exact for the diagonal costs that reach the compact path); anything else raises."""
        C, c, Cf, cf = (t.detach().to(device=device, dtype=dtype) for t in self.ls_cost_override)
        T, n = self.horizon, self.nx + self.nu
        if C.shape == (T, n) and Cf.shape == (n,):
            if not diag:
                C, Cf = torch.diag_embed(C), torch.diag_embed(Cf)
        elif C.shape == (T, n, n) and Cf.shape == (n, n):
            if diag:
                C, Cf = torch.diagonal(C, dim1=-2, dim2=-1), torch.diagonal(Cf, dim1=-2, dim2=-1)
        else:
            raise RuntimeError(f"ls_cost_override must be ([T,n,n],[T,n],[n,n],[n]) or compact ([T,n],[T,n],[n],[n]); got "
                               f"{tuple(C.shape)}, {tuple(c.shape)}, {tuple(Cf.shape)}, {tuple(cf.shape)}")
        if c.shape != (T, n) or cf.shape != (n,):
            raise RuntimeError(f"ls_cost_override linear terms must be [T,n], [n]; got {tuple(c.shape)}, {tuple(cf.shape)}")
        return tuple(t.contiguous() for t in (C, c, Cf, cf))

    def _costs(self, B, dtype, device, dense_only: bool = False):
        cm = self.cost_module
        if self._diag_eligible() and cm.q.shape[0] == B and not dense_only:
            q, p, qf, pf = (t.detach().to(device=device, dtype=dtype).contiguous() for t in (cm.q, cm.p, cm.q_final, cm.p_final))
            xref = cm.x_ref.detach().to(device=device, dtype=dtype).contiguous()
            uref = cm.u_ref.detach().to(device=device, dtype=dtype).contiguous()
            ls = None
            if self.compat == "reference":   # quirk Q4, compact as well
                ls = self._ls_override(dtype, device, True) if self.ls_cost_override is not None else (q[0], p[0], qf[0], pf[0])
            self._last_diag = True
            return q, p, qf, pf, xref, uref, ls
        self._last_diag = False
        C, c, Cf, cf = (t.detach().to(device=device, dtype=dtype).contiguous() for t in (cm.C, cm.c, cm.C_final, cm.c_final))
        if C.ndim == 4 and C.shape[0] == 1 and B > 1:
            C, c = C[0], c[0]
        if Cf.ndim == 3 and Cf.shape[0] == 1 and B > 1:
            Cf, cf = Cf[0], cf[0]
        xref = cm.x_ref.detach().to(device=device, dtype=dtype).contiguous()
        uref = cm.u_ref.detach().to(device=device, dtype=dtype).contiguous()
        # quirk Q4: with per-element costs the reference scores every element's line-search candidates
        # with batch element 0's cost (cost.py:40-62 un-batched branch, called from controller.py:616-618)
        ls = None
        if self.compat == "reference" and (C.ndim == 4 or Cf.ndim == 3):
            if self.ls_cost_override is not None:
                ls = self._ls_override(dtype, device, False)
            else:
                ls = (C[0] if C.ndim == 4 else C, c[0] if c.ndim == 3 else c, Cf[0] if Cf.ndim == 3 else Cf,
                      cf[0] if cf.ndim == 2 else cf)
        return C, c, Cf, cf, xref, uref, ls

    # ------------------------------------------------------------------ public API
    def forward(self, x0: Tensor, U_init: Optional[Tensor] = None, *, x_ref_full: Optional[Tensor] = None,
                u_ref_full: Optional[Tensor] = None, warm_start: bool = True) -> Tuple[Tensor, Tensor]:
        """``mpc(x0, U_init)``: one solve, ``(X*[B,T+1,nx], U*[B,T,nu])`` (controller.py:266-273).

        ``mpc(x0, x_ref_full=..., u_ref_full=...)``: the receding-horizon closed loop of the reference's demos and
        of tests/test_batch_tracking.py:76-84 / tests/test_forward.py:92-95 — ``N_sim`` steps, each solving from the
        current state against the reference window ``[k, k+T]`` and applying the first input; returns
        ``(Xs[B,N_sim+1,nx], Us[B,N_sim,nu])``.  See ``simulate``."""
        if x_ref_full is not None or u_ref_full is not None:
            return self.simulate(x0, x_ref_full, u_ref_full, U_init=U_init, warm_start=warm_start)
        if isinstance(self.cost_module, ConstrainedQuadCost):
            return self.solve_constrained(x0, U_init)
        B = x0.shape[0] if x0.ndim > 1 else 1
        if U_init is None:
            U_init = torch.zeros(B, self.horizon, self.nu, device=x0.device, dtype=x0.dtype)
        cm = self.cost_module
        if isinstance(cm, DiagQuadCost):
            return ILQRSolveDiag.apply(x0, cm.q, cm.p, cm.q_final, cm.p_final, cm.x_ref, cm.u_ref, self, U_init)
        return ILQRSolve.apply(x0, cm.C, cm.c, cm.C_final, cm.c_final, cm.x_ref, cm.u_ref, self, U_init)

    def simulate(self, x0: Tensor, x_ref_full: Tensor, u_ref_full: Optional[Tensor] = None, *, U_init: Optional[Tensor] = None,
                 warm_start: bool = True, N_sim: Optional[int] = None) -> Tuple[Tensor, Tensor]:
        """Closed-loop receding-horizon simulation.

        Per step k < N_sim: references = ``x_ref_full[:, k:k+T+1]``, ``u_ref_full[:, k:k+T]``; solve from the current
        state; ``u_k = U*[:,0]``; ``x <- f(x, u_k, dt)``; the next solve starts from the previous plan shifted by
        one step with the last input zeroed (``warm_start=True``, examples/03_demo_unicycle_tracking.py:139-140) or
        from zeros (``False``, tests/test_batch_tracking.py).  ``x_ref_full`` is ``[B or 1, >= N_sim+T+1, nx]``.

        With a shipped dynamics plugin and no gradient required the whole loop stays on the device: per step the
        solve kernels plus ONE ``tacd_closed_loop_advance`` launch, no host synchronisation, no per-step tensor
        allocation on the reference path.  When a gradient is required (or ``f_dyn`` is a Python callable) the same
        loop runs through autograd, one ``ILQRSolve`` node per step."""
        if x0.ndim != 2:
            raise RuntimeError("x0 must be [B, nx] (quirk Q12)")
        N = int(N_sim if N_sim is not None else self.N_sim)
        T, nx, nu, B = self.horizon, self.nx, self.nu, x0.shape[0]
        cm = self.cost_module
        if x_ref_full is None:
            raise RuntimeError("simulate needs x_ref_full [B or 1, N_sim + T + 1, nx]")
        if x_ref_full.ndim == 2:
            x_ref_full = x_ref_full.unsqueeze(0)
        if u_ref_full is None:
            u_ref_full = torch.zeros(x_ref_full.shape[0], x_ref_full.shape[1] - 1, nu, device=x_ref_full.device, dtype=x_ref_full.dtype)
        if u_ref_full.ndim == 2:
            u_ref_full = u_ref_full.unsqueeze(0)
        if x_ref_full.shape[1] < N + T + 1 or u_ref_full.shape[1] < N + T:
            raise RuntimeError(f"reference trajectories must cover N_sim + T + 1 = {N + T + 1} states and {N + T} inputs")
        params = [x0, x_ref_full, u_ref_full] + ([cm.q, cm.p, cm.q_final, cm.p_final] if isinstance(cm, DiagQuadCost)
                                                 else [cm.C, cm.c, cm.C_final, cm.c_final])
        needs_grad = torch.is_grad_enabled() and any(isinstance(t, Tensor) and t.requires_grad for t in params)
        # (a ConstrainedQuadCost runs the augmented-Lagrangian outer loop per step: forward() dispatches it below; the
        # device loop would solve the unconstrained problem)
        if not needs_grad and self._is_native() and x0.is_cuda and not isinstance(cm, ConstrainedQuadCost):
            return self._simulate_device(x0, x_ref_full, u_ref_full, U_init, warm_start, N)
        # autograd / generic path: the loop of the reference's demos, step by step
        x = x0
        warm = U_init if U_init is not None else torch.zeros(B, T, nu, device=x0.device, dtype=x0.dtype)
        Xs, Us = [x0], []
        for k in range(N):
            cm.set_reference(x_ref_full[:, k:k + T + 1], u_ref_full[:, k:k + T])
            _, U = self.forward(x, warm)
            u = U[:, 0]
            x = self.f_dyn(x, u, self.dt)
            Xs.append(x)
            Us.append(u)
            if warm_start:
                warm = torch.cat((U[:, 1:].detach(), torch.zeros(B, 1, nu, device=U.device, dtype=U.dtype)), dim=1)
            else:
                warm = torch.zeros(B, T, nu, device=x0.device, dtype=x0.dtype)
        return torch.stack(Xs, dim=1), torch.stack(Us, dim=1)

    def _simulate_device(self, x0, x_ref_full, u_ref_full, U_init, warm_start, N):
        dtype, device = _solve_dtype(x0), x0.device
        T, nx, nu, B = self.horizon, self.nx, self.nu, x0.shape[0]
        spec = self._spec(dtype, device)
        prep = lambda t: t.detach().to(device=device, dtype=dtype).contiguous()
        # tacd_closed_loop_advance strides the input reference by ref_len - 1: hand it exactly that many steps (the Python loop
        # only ever reads [k, k+T) windows, so any u_ref_full of at least N_sim + T steps means the same thing on both paths)
        L = min(x_ref_full.shape[1], u_ref_full.shape[1] + 1)
        xf, uf = prep(x_ref_full[:, :L]), prep(u_ref_full[:, :L - 1])
        xw, uw = xf[:, :T + 1].contiguous(), uf[:, :T].contiguous()
        self.cost_module.set_reference(xw, uw)
        C, c, Cf, cf, _, _, ls = self._costs(B, dtype, device)
        x = prep(x0).clone()
        warm = prep(U_init).clone() if U_init is not None else torch.zeros(B, T, nu, device=device, dtype=dtype)
        Xs = torch.empty(B, N + 1, nx, device=device, dtype=dtype)
        Us = torch.empty(B, N, nu, device=device, dtype=dtype)
        out = None
        for k in range(N):
            out = ops.ilqr_forward(spec, x, C, c, Cf, cf, xw, uw, warm, ls_cost=ls, want_trace=False, C_diag=self._last_diag)
            ops.closed_loop_advance(spec, k, N, x, out["U"], Xs, Us, warm if warm_start else None, xf, uf, xw, uw)
            if not warm_start and k == 0 and U_init is not None:
                warm.zero_()   # cold start: U_init seeds the first solve only, as in the step-by-step loop of simulate()
        if out is not None:
            self.X_last, self.U_last = out["X"].to(x0.dtype), out["U"].to(x0.dtype)
            self.tight_mask_last = out["tight"].bool()
            self.iters_last, self.flags_last, self.cost_last = out["iters"], out["flags"], out["cost"]
        self.converged = True
        return Xs.to(x0.dtype), Us.to(x0.dtype)

    def solve_constrained(self, x0: Tensor, U_init: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
        """State-box-constrained solve by an augmented Lagrangian (``ConstrainedQuadCost``; schedule defined in
        ``AugmentedLagrangianManager.py``): ``n_outer`` warm-started iLQR solves on per-element costs that carry the
        quadratic model of the active penalties, one ``tacd_al_update`` launch between them.  Forward only (no
        gradient flows through the multipliers); the multipliers and the violation stay on ``cost_module.al_manager``."""
        cm = self.cost_module
        al = cm.al_manager
        if x0.ndim != 2:
            raise RuntimeError("x0 must be [B, nx] (quirk Q12)")
        if not (self._is_native() and x0.is_cuda):
            raise RuntimeError("state constraints need a shipped dynamics plugin on a CUDA device")
        dtype, device = _solve_dtype(x0), x0.device
        B, T, nx, nu, n = x0.shape[0], self.horizon, self.nx, self.nu, self.nx + self.nu
        spec = self._spec(dtype, device)
        prep = lambda t: t.detach().to(device=device, dtype=dtype).contiguous()
        C, c, Cf, cf = prep(cm.C), prep(cm.c), prep(cm.C_final), prep(cm.c_final)
        xref, uref = prep(cm.x_ref), prep(cm.u_ref)
        # per-element copies the penalty model is added to (the penalty is diagonal in the state block)
        Cb = (C if C.ndim == 4 else C.unsqueeze(0).expand(B, T, n, n)).clone()
        cb = (c if c.ndim == 3 else c.unsqueeze(0).expand(B, T, n)).clone()
        Cfb = (Cf if Cf.ndim == 3 else Cf.unsqueeze(0).expand(B, n, n)).clone()
        cfb = (cf if cf.ndim == 2 else cf.unsqueeze(0).expand(B, n)).clone()
        C0d = torch.diagonal(Cb, dim1=-2, dim2=-1)[..., :nx].clone()      # [B,T,nx] base diagonals
        c0 = cb[..., :nx].clone()
        Cf0d = torch.diagonal(Cfb, dim1=-2, dim2=-1)[..., :nx].clone()
        cf0 = cfb[..., :nx].clone()
        al.reset(B, T, nx, device, dtype)
        x0d = prep(x0)
        warm = prep(U_init) if U_init is not None else torch.zeros(B, T, nu, device=device, dtype=dtype)
        out = None
        for outer in range(max(1, int(al.n_outer))):
            # the reference's Q4 quirk has no meaning here: every element is scored with its own cost
            out = ops.ilqr_forward(spec, x0d, Cb, cb, Cfb, cfb, xref, uref, warm, ls_cost=None, want_trace=False)
            rho_next = min(al.rho * al.rho_growth, al.rho_max) if outer > 0 else al.rho
            dq, dp, viol = ops.al_update(out["X"], xref, cm.x_min, cm.x_max, al.lam_lo, al.lam_hi, al.rho, rho_next, update_lambda=True)
            al.rho, al.violation, al.outer_iters = rho_next, viol, outer + 1
            torch.diagonal(Cb, dim1=-2, dim2=-1)[..., :nx].copy_(C0d + dq[:, :T])
            cb[..., :nx].copy_(c0 + dp[:, :T])
            torch.diagonal(Cfb, dim1=-2, dim2=-1)[..., :nx].copy_(Cf0d + dq[:, T])
            cfb[..., :nx].copy_(cf0 + dp[:, T])
            warm = out["U"]
            if al.tol > 0 and outer > 0 and float(viol.max()) <= al.tol:
                break
        X, U = out["X"].to(x0.dtype), out["U"].to(x0.dtype)
        self.converged = True
        self.X_last, self.U_last = X, U
        self.tight_mask_last = out["tight"].bool()
        self.iters_last, self.flags_last, self.cost_last = out["iters"], out["flags"], out["cost"]
        return X, U

    def solve_step(self, x0: Tensor, U_init: Tensor) -> Tuple[Tensor, Tensor]:
        """One full iLQR solve from (x0, U_init) (controller.py:275-315)."""
        if x0.ndim != 2:
            raise RuntimeError("x0 must be [B, nx] (quirk Q12)")
        if not x0.is_cuda:
            raise RuntimeError("tacdmpc_b200 solves on a CUDA device only; x0 is on " + str(x0.device))
        dtype, device = _solve_dtype(x0), x0.device
        B = x0.shape[0]
        spec = self._spec(dtype, device)
        C, c, Cf, cf, xref, uref, ls = self._costs(B, dtype, device)
        x0d = x0.detach().to(dtype)
        U0 = U_init.detach().to(device=device, dtype=dtype)
        if self._is_native():
            try:
                out = ops.ilqr_forward(spec, x0d, C, c, Cf, cf, xref, uref, U0, ls_cost=ls, want_trace=True,
                                       C_diag=self._last_diag)
            except _lib.TacdUnsupported:
                # compact diagonals are taken by the five-lanes-per-element kernels only (shipped (nx, nu, dyn) instances, horizon
                # within shared memory); anything else solves on the dense expansion, as GeneralQuadCost would
                if not self._last_diag:
                    raise
                C, c, Cf, cf, xref, uref, ls = self._costs(B, dtype, device, dense_only=True)
                out = ops.ilqr_forward(spec, x0d, C, c, Cf, cf, xref, uref, U0, ls_cost=ls, want_trace=True)
            AB = None
        else:
            out, AB = self._solve_generic(spec, x0d, C, c, Cf, cf, xref, uref, U0, ls)
        X, U = out["X"].to(x0.dtype), out["U"].to(x0.dtype)
        self.converged = True
        self.X_last, self.U_last = X, U
        self.tight_mask_last = out["tight"].bool()
        self.iters_last, self.flags_last, self.cost_last = out["iters"], out["flags"], out["cost"]
        self.alpha_trace_last = out.get("alpha_trace")
        nx = self.nx
        if self._last_diag:
            self.H_last = None   # dense Hessian blocks are not materialised on the compact path (cost_module.C is)
        else:
            Cb = C if C.ndim == 4 else C.unsqueeze(0).expand(B, *C.shape)
            self.H_last = (Cb[..., :nx, :nx], Cb[..., nx:, nx:], Cb[..., :nx, nx:])
        self._F_last = AB
        self._last_spec, self._last_C, self._last_xref, self._last_uref = spec, C, xref, uref
        self._last_Cf = Cf
        self._last_tight_u8, self._last_AB = out["tight"], AB
        return X, U

    # ------------------------------------------------------------------ stage wrappers (reference names)
    def rollout_trajectory(self, x0: Tensor, U: Tensor) -> Tensor:
        """X[B,T+1,nx] from x0, U without clamping (controller.py:361-398)."""
        if x0.ndim == 1:
            x0 = x0.unsqueeze(0)
        if U.ndim == 2:
            U = U.unsqueeze(0)
        if isinstance(self.f_dyn, NativeDynamics) and x0.is_cuda:
            return ops.rollout(self._spec(_solve_dtype(x0), x0.device), x0, U).to(x0.dtype)
        X = [x0]
        for t in range(U.shape[1]):
            X.append(self.f_dyn(X[-1], U[:, t], self.dt))
        return torch.stack(X, dim=1)

    def linearize_dynamics(self, X: Tensor, U: Tensor):
        """A[B,T,nx,nx], B[B,T,nx,nu] along (X[:, :-1], U) (controller.py:448-471)."""
        Bsz, T, nx, nu = X.shape[0], U.shape[1], self.nx, self.nu
        if isinstance(self.f_dyn, NativeDynamics) and X.is_cuda:
            A, Bm = ops.linearize(self._spec(_solve_dtype(X), X.device), X, U)
            return A.to(X.dtype), Bm.to(X.dtype)
        x_flat, u_flat = X[:, :-1].reshape(-1, nx), U.reshape(-1, nu)
        if self.grad_method is GradMethod.ANALYTIC:
            A, Bm = self.f_dyn_jac(x_flat, u_flat, self.dt)
        elif self.grad_method is GradMethod.AUTO_DIFF:
            A, Bm = batched_jacobian(lambda x, u: self.f_dyn(x, u, self.dt), x_flat, u_flat)
        else:
            # the reference's FD branch raises at HEAD (it shadows B, controller.py:468-471); this is the
            # stage it meant to call, with the eps it hard-codes (quirk Q9)
            A, Bm = jacobian_finite_diff_batched(self.f_dyn, x_flat, u_flat, dt=self.dt)
        return A.reshape(Bsz, T, nx, nx), Bm.reshape(Bsz, T, nx, nu)

    def compute_cost(self, X: Tensor, U: Tensor) -> Tensor:
        return self.cost_module.objective(X, U)

    def _compute_tight_mask(self, U: Tensor, atol: float = 1e-7) -> Tensor:
        """isclose(U, bound, rtol=1e-5, atol) for either bound (controller.py:636-642, quirk Q7)."""
        mask = torch.zeros_like(U, dtype=torch.bool)
        if self.u_min is not None:
            mask |= torch.isclose(U, self.u_min.to(U.dtype).expand_as(U), atol=atol)
        if self.u_max is not None:
            mask |= torch.isclose(U, self.u_max.to(U.dtype).expand_as(U), atol=atol)
        return mask

    # ------------------------------------------------------------------ generic path (Python f_dyn)
    def _solve_generic(self, spec, x0, C, c, Cf, cf, xref, uref, U, ls):
        """iLQR with a user-supplied Python ``f_dyn``: rollouts / Jacobians in torch, Riccati recursion
        (``tacd_riccati``) and all cost evaluations (``tacd_objective``) in CUDA.  Like the reference
        this loop synchronises once per iteration (``improved.any()``)."""
        B, T, nx, nu = x0.shape[0], self.horizon, self.nx, self.nu
        dt_ = x0.dtype
        alphas = torch.tensor(_abi.DEFAULT_ALPHAS, dtype=dt_, device=x0.device)
        na = alphas.numel()
        Cb = C if C.ndim == 4 else C.unsqueeze(0).expand(B, *C.shape)
        cb = c if c.ndim == 3 else c.unsqueeze(0).expand(B, *c.shape)
        Cfb = Cf if Cf.ndim == 3 else Cf.unsqueeze(0).expand(B, *Cf.shape)
        cfb = cf if cf.ndim == 2 else cf.unsqueeze(0).expand(B, *cf.shape)
        xr = xref if xref.shape[0] == B else xref.expand(B, -1, -1)
        ur = uref if uref.shape[0] == B else uref.expand(B, -1, -1)
        lsC = ls if ls is not None else (C, c, Cf, cf)
        umin = None if self.u_min is None else self.u_min.to(dt_)
        umax = None if self.u_max is None else self.u_max.to(dt_)
        U = U.clone()
        X = self.rollout_trajectory(x0, U)
        best = ops.objective(spec, X, U, C, c, Cf, cf, xref, uref)
        iters = torch.zeros(B, dtype=torch.int32, device=x0.device)
        live = torch.ones(B, dtype=torch.bool, device=x0.device)
        flags = torch.zeros(B, dtype=torch.uint8, device=x0.device)
        trace = torch.full((B, max(self.max_iter, 1)), 255, dtype=torch.uint8, device=x0.device)
        for it in range(self.max_iter):
            tau = torch.cat((X[:, :T] - xr[:, :T], U - ur), dim=-1)
            l = (Cb @ tau.unsqueeze(-1)).squeeze(-1) + cb
            dxN = X[:, -1] - xr[:, -1]
            lxN = (Cfb[:, :nx, :nx] @ dxN.unsqueeze(-1)).squeeze(-1) + cfb[:, :nx]
            A, Bm = self.linearize_dynamics(X, U)
            K, k, f = ops.riccati(spec, A, Bm, l[..., :nx], l[..., nx:], Cb[..., :nx, :nx], Cb[..., :nx, nx:],
                                  Cb[..., nx:, nx:], lxN, Cfb[:, :nx, :nx])
            flags |= f
            # evaluate_alphas (controller.py:589-620), all candidates of all elements at once
            xa = x0.unsqueeze(1).expand(B, na, nx)
            Xc, Uc = [xa], []
            for t in range(T):
                du = alphas.view(1, na, 1) * k[:, t].unsqueeze(1) + ((xa - X[:, t].unsqueeze(1)) @ K[:, t].mT)
                ua = U[:, t].unsqueeze(1) + du
                if umin is not None:
                    ua = torch.max(ua, umin)
                if umax is not None:
                    ua = torch.min(ua, umax)
                xa = self.f_dyn(xa, ua, self.dt)
                Uc.append(ua)
                Xc.append(xa)
            Xc, Uc = torch.stack(Xc, dim=2), torch.stack(Uc, dim=2)            # [B,na,T+1,nx], [B,na,T,nu]
            rep = lambda t_, shared: t_ if shared else t_.repeat_interleave(na, dim=0)
            costs = ops.objective(spec, Xc.reshape(B * na, T + 1, nx), Uc.reshape(B * na, T, nu),
                                  rep(lsC[0], lsC[0].ndim == 3), rep(lsC[1], lsC[1].ndim == 2),
                                  rep(lsC[2], lsC[2].ndim == 2), rep(lsC[3], lsC[3].ndim == 1),
                                  xr.repeat_interleave(na, dim=0), ur.repeat_interleave(na, dim=0)).view(B, na)
            ai = torch.argmin(costs, dim=1)
            Xn = Xc[torch.arange(B, device=x0.device), ai]
            Un = Uc[torch.arange(B, device=x0.device), ai]
            newc = ops.objective(spec, Xn, Un, C, c, Cf, cf, xref, uref)
            improved = (newc < best) & live
            trace[:, it] = torch.where(live, ai.to(torch.uint8), trace[:, it])
            iters += live.to(torch.int32)
            if not bool(improved.any()):
                break
            live = improved
            best = torch.where(improved, newc, best)
            X = torch.where(improved.view(B, 1, 1), Xn, X)
            U = torch.where(improved.view(B, 1, 1), Un, U)
        else:
            flags |= (live.to(torch.uint8) * _abi.FLAG_MAXITER)
        A, Bm = self.linearize_dynamics(X, U)
        tight = self._compute_tight_mask(U).to(torch.uint8)
        out = dict(X=X, U=U, tight=tight, iters=iters, flags=flags, cost=best, alpha_trace=trace)
        return out, (A.contiguous(), Bm.contiguous())
