"""State box constraints by an augmented Lagrangian.

The reference's ``DifferentialMPC/AugmentedLagrangianManager.py`` defines ``ConstrainedQuadCost(al_manager, x_max,
x_min, *cost_args)`` but cannot be imported (the ``AugmentedLagrangianManager`` class it needs does not exist, and
``quadraticize`` disagrees with its base class), so there is NO reference behaviour to reproduce.  What the file
states is kept — the constraint ``g = [x - x_max ; x_min - x] <= 0`` on every state of the horizon
(``utils.py:31-35``), a multiplier ``lambda`` and a penalty ``rho`` that add ``(lambda + rho relu(g))' J_g`` and
``rho J_g' J_g`` to the quadratic model (``:45-61``) — and the missing schedule is DEFINED here:

* outer loop, ``n_outer`` times: solve the iLQR problem with the quadratic model of the penalty of the currently
  active constraints added to the cost (``J_g = +-I``, so the model is diagonal: it only touches ``C_t[i,i]`` and
  ``c_t[i]``, and the inner solve is the ordinary fused kernel on a per-element cost), warm-started with the previous
  plan; then ``lambda <- max(0, lambda + rho g(x))``, ``rho <- min(rho_growth * rho, rho_max)``
  (Powell-Hestenes-Rockafellar; one ``tacd_al_update`` launch).
* a constraint is in the model of a solve while its multiplier is positive after that update.

Results are checked against a dense QP solution of a state-constrained linear-quadratic problem
(``tests/test_gpu_al.py``): there is no oracle, parity is *unpinned* by construction.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch
from torch import Tensor

from .cost import GeneralQuadCost


@dataclass
class AugmentedLagrangianManager:
    """Penalty schedule and the multipliers of the last solve (``lam_lo``, ``lam_hi``: ``[B, T+1, nx]``)."""
    rho_init: float = 10.0
    rho_growth: float = 3.0
    rho_max: float = 1e5
    n_outer: int = 10
    tol: float = 1e-4            # stop early when the largest violation of the batch is below tol (one host sync per outer step)
    lam_lo: Optional[Tensor] = None
    lam_hi: Optional[Tensor] = None
    rho: float = 0.0
    violation: Optional[Tensor] = None   # [B] largest constraint violation after the last solve
    outer_iters: int = 0

    def reset(self, B: int, T: int, nx: int, device, dtype):
        self.lam_lo = torch.zeros(B, T + 1, nx, device=device, dtype=dtype)
        self.lam_hi = torch.zeros(B, T + 1, nx, device=device, dtype=dtype)
        self.rho = float(self.rho_init)
        self.violation = None
        self.outer_iters = 0


class ConstrainedQuadCost(GeneralQuadCost):
    """``GeneralQuadCost`` + box constraints ``x_min <= x_t <= x_max`` on all T+1 states, argument order of the
    reference's class (``AugmentedLagrangianManager.py:6-16``).  Use ``+-inf`` for an unbounded side."""

    def __init__(self, al_manager: AugmentedLagrangianManager, x_max: Tensor, x_min: Tensor, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.al_manager = al_manager
        self.x_max = torch.as_tensor(x_max).to(self.device)
        self.x_min = torch.as_tensor(x_min).to(self.device)

    def constraint_values(self, X: Tensor) -> Tensor:
        """g(X) = [X - x_max ; x_min - X] (utils.py:31-35), shape [..., 2 nx]."""
        return torch.cat((X - self.x_max.to(X.dtype), self.x_min.to(X.dtype) - X), dim=-1)
