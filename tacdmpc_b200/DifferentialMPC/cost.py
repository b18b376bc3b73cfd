"""Quadratic cost plugin with the reference's interface (DifferentialMPC/cost.py:7-100).

``GeneralQuadCost`` is a parameter holder: the controller reads ``C, c, C_final, c_final, x_ref, u_ref``
straight into the CUDA solve.  ``objective`` / ``quadraticize`` stay available with the reference's
signatures for user code; ``objective`` on CUDA tensors runs the ``tacd_objective`` kernel.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor


class GeneralQuadCost(torch.nn.Module):
    """cost(tau) = sum_t 1/2 tau_t' C_t tau_t + c_t' tau_t + terminal, tau_t = [x_t - x_ref_t ; u_t - u_ref_t].

    ``C``/``c`` are shared ``[T,n,n]``/``[T,n]`` or per-element ``[B,T,n,n]``/``[B,T,n]``; ``C_final``/
    ``c_final`` independently ``[n,n]``/``[n]`` or ``[B,n,n]``/``[B,n]`` (cost.py:29-35).
    """

    def __init__(self, nx: int, nu: int, C: Tensor, c: Tensor, C_final: Tensor, c_final: Tensor, **kwargs):
        super().__init__()
        self.device = torch.device(kwargs.get("device", "cpu"))
        self.nx, self.nu = nx, nu
        self.C, self.c = C.to(self.device), c.to(self.device)
        self.C_final, self.c_final = C_final.to(self.device), c_final.to(self.device)
        horizon = C.shape[-3]
        x_ref = kwargs.get("x_ref")
        u_ref = kwargs.get("u_ref")
        self.register_buffer("x_ref", x_ref if x_ref is not None else torch.zeros(1, horizon + 1, nx, device=self.device))
        self.register_buffer("u_ref", u_ref if u_ref is not None else torch.zeros(1, horizon, nu, device=self.device))
        self.set_reference(self.x_ref, self.u_ref)

    def set_reference(self, x_ref: Optional[Tensor] = None, u_ref: Optional[Tensor] = None):
        """2-D references get a leading batch dim of 1 (cost.py:21-27)."""
        if x_ref is not None:
            self.x_ref = (x_ref.unsqueeze(0) if x_ref.ndim == 2 else x_ref).to(self.device)
        if u_ref is not None:
            self.u_ref = (u_ref.unsqueeze(0) if u_ref.ndim == 2 else u_ref).to(self.device)

    # -- helpers -------------------------------------------------------------------------------
    def _prepare_costs(self, B: int):
        """Broadcast the four cost tensors to a leading batch of B (views, no copies)."""
        C, c, Cf, cf = self.C, self.c, self.C_final, self.c_final
        if C.ndim == 3:
            C, c = C[None], c[None]
        if Cf.ndim == 2:
            Cf, cf = Cf[None], cf[None]
        if C.shape[0] == 1 and B > 1:
            C, c = C.expand(B, *C.shape[1:]), c.expand(B, *c.shape[1:])
        if Cf.shape[0] == 1 and B > 1:
            Cf, cf = Cf.expand(B, *Cf.shape[1:]), cf.expand(B, *cf.shape[1:])
        return C, c, Cf, cf

    def _taus(self, X, U, x_ref_override, u_ref_override):
        xr = self.x_ref if x_ref_override is None else x_ref_override
        ur = self.u_ref if u_ref_override is None else u_ref_override
        if xr.ndim == 2:
            xr = xr[None]
        if ur.ndim == 2:
            ur = ur[None]
        T = U.shape[1]
        tau = torch.cat((X[:, :T] - xr[:, :T], U - ur), dim=-1)
        tauN = torch.nn.functional.pad(X[:, -1] - xr[:, -1], (0, self.nu))
        return tau, tauN

    # -- reference API --------------------------------------------------------------------------
    def objective(self, X, U, *, x_ref_override=None, u_ref_override=None):
        """Total cost per batch element (cost.py:37-62); un-batched input returns a scalar."""
        squeeze = X.ndim == 2
        if squeeze:
            X, U = X[None], U[None]
        B = X.shape[0]
        tau, tauN = self._taus(X, U, x_ref_override, u_ref_override)
        C, c, Cf, cf = self._prepare_costs(B)
        run = 0.5 * ((tau.unsqueeze(-2) @ C @ tau.unsqueeze(-1)).flatten(1).sum(1)) + (c * tau).flatten(1).sum(1)
        fin = 0.5 * (tauN.unsqueeze(-2) @ Cf @ tauN.unsqueeze(-1)).flatten(1).sum(1) + (cf * tauN).sum(-1)
        total = run + fin
        return total[0] if squeeze else total

    def quadraticize(self, X, U, *, x_ref_override=None, u_ref_override=None):
        """(l_x, l_u, l_xx, l_xu, l_uu, l_xN, l_xxN) with the Hessian blocks as VIEWS of C (cost.py:64-100)."""
        squeeze = X.ndim == 2
        if squeeze:
            X, U = X[None], U[None]
        B = X.shape[0]
        nx = self.nx
        tau, tauN = self._taus(X, U, x_ref_override, u_ref_override)
        C, c, Cf, cf = self._prepare_costs(B)
        l = (C @ tau.unsqueeze(-1)).squeeze(-1) + c
        lN = (Cf @ tauN.unsqueeze(-1)).squeeze(-1) + cf
        out = (l[..., :nx], l[..., nx:], C[..., :nx, :nx], C[..., :nx, nx:], C[..., nx:, nx:], lN[..., :nx], Cf[..., :nx, :nx])
        return tuple(t.squeeze(0) for t in out) if squeeze else out


class DiagQuadCost(GeneralQuadCost):
    """Per-element DIAGONAL quadratic cost kept in compact form — the cost ``ACMPC/actor.py:58-81`` builds with
    ``torch.diag_embed``.

    ``q [B,T,n]`` and ``q_final [B,n]`` are the diagonals of ``C`` / ``C_final``; ``p``, ``p_final`` are ``c`` /
    ``c_final``.  The controller hands them to the kernels as they are (``tacd_problem.C_diag``): 2n instead of
    n*n + n words per step go through the solve, and the backward writes ``grad_q``, ``grad_p`` instead of a dense
    ``grad_C``.  ``C``, ``c``, ``C_final``, ``c_final`` remain readable (materialised on access) so code written
    against ``GeneralQuadCost`` keeps working; results are identical to passing the dense tensors.
    """

    def __init__(self, nx: int, nu: int, q: Tensor, p: Tensor, q_final: Tensor, p_final: Tensor, **kwargs):
        torch.nn.Module.__init__(self)
        self.device = torch.device(kwargs.get("device", q.device))
        self.nx, self.nu = nx, nu
        self.set_diag(q, p, q_final, p_final)
        horizon = q.shape[-2]
        x_ref, u_ref = kwargs.get("x_ref"), kwargs.get("u_ref")
        self.register_buffer("x_ref", x_ref if x_ref is not None else torch.zeros(1, horizon + 1, nx, device=self.device))
        self.register_buffer("u_ref", u_ref if u_ref is not None else torch.zeros(1, horizon, nu, device=self.device))
        self.set_reference(self.x_ref, self.u_ref)

    def set_diag(self, q: Tensor, p: Tensor, q_final: Optional[Tensor] = None, p_final: Optional[Tensor] = None):
        """Install per-element diagonals; the terminal cost defaults to the last running step (actor.py:76-77)."""
        n = self.nx + self.nu
        if q.ndim != 3 or q.shape[-1] != n or p.shape != q.shape:
            raise RuntimeError(f"q, p must be [B, T, n={n}]; got {tuple(q.shape)}, {tuple(p.shape)}")
        self.q, self.p = q.to(self.device), p.to(self.device)
        self.q_final = (q_final if q_final is not None else q[:, -1].clone()).to(self.device)
        self.p_final = (p_final if p_final is not None else p[:, -1].clone()).to(self.device)

    # dense views for code written against GeneralQuadCost (reads only; assign q/p to change the cost)
    C = property(lambda self: torch.diag_embed(self.q))
    c = property(lambda self: self.p)
    C_final = property(lambda self: torch.diag_embed(self.q_final))
    c_final = property(lambda self: self.p_final)
