"""Utilities exported by the reference package (DifferentialMPC/utils.py).

``pnqp`` on CUDA tensors runs the ``tacd_pnqp`` kernel (one thread per problem); the Jacobian helpers
differentiate arbitrary Python callables and therefore stay in torch.func.
"""
from __future__ import annotations

import torch
from torch.func import jacrev, vmap

from .. import ops


def bmv(A, x):
    return (A @ x.unsqueeze(-1)).squeeze(-1)


def bdot(a, b):
    return (a * b).sum(-1)


def bquad(x, H):
    return (x.unsqueeze(-2) @ H @ x.unsqueeze(-1)).squeeze(-1).squeeze(-1)


def bger(a, b):
    return a.unsqueeze(-1) * b.unsqueeze(-2)


def eclamp(x, lo, hi):
    lo = torch.as_tensor(lo, dtype=x.dtype, device=x.device)
    hi = torch.as_tensor(hi, dtype=x.dtype, device=x.device)
    return torch.max(torch.min(x, hi), lo)


def get_state_constraints(state, x_max, x_min):
    return torch.cat([state - x_max, x_min - state], dim=1)


def batched_jacobian(f, x, u):
    """A = df/dx, B = df/du of f(x, u) at every row of (x, u) (utils.py:37-46)."""
    nx = x.shape[-1]

    def joint(z):
        return f(z[..., :nx], z[..., nx:])

    J = vmap(jacrev(joint))(torch.cat([x, u], dim=-1))
    return J[..., :nx], J[..., nx:]


def jacobian_finite_diff_batched(f_dyn, x, u, dt, eps: float = 1e-4):
    """Central differences over the joint (x,u) vector (utils.py:48-64)."""
    nx = x.shape[-1]
    z = torch.cat([x, u], dim=-1)
    cols = []
    for j in range(z.shape[-1]):
        e = torch.zeros(z.shape[-1], dtype=z.dtype, device=z.device)
        e[j] = eps
        zp, zm = z + e, z - e
        cols.append((f_dyn(zp[..., :nx], zp[..., nx:], dt) - f_dyn(zm[..., :nx], zm[..., nx:], dt)) / (2 * eps))
    J = torch.stack(cols, dim=-1)
    return J[..., :nx], J[..., nx:]


def pnqp(H, q, lower, upper, x_init=None, n_iter: int = 20, reg: float = 1e-10, tol: float = 1e-5, gamma: float = 0.1):
    """Projected-Newton box QP, min 1/2 x'Hx + q'x s.t. lower <= x <= upper (utils.py:66-124).

    Every problem of the batch is solved independently (the reference couples the batch through
    ``.all()`` / ``.max()``, but only ever calls it with a batch of one).  The kernel implements the
    reference's default hyper-parameters; other values are rejected rather than silently ignored.
    """
    if x_init is not None or n_iter != 20 or reg != 1e-10 or tol != 1e-5 or gamma != 0.1:
        raise NotImplementedError("tacd_pnqp implements the reference defaults (x_init=None, n_iter=20, reg=1e-10, "
                                  "tol=1e-5, gamma=0.1), which is how the solver calls it")
    lead = q.shape[:-1]
    m = q.shape[-1]
    lo = torch.as_tensor(lower, dtype=q.dtype, device=q.device).expand_as(q)
    hi = torch.as_tensor(upper, dtype=q.dtype, device=q.device).expand_as(q)
    x, iters = ops.pnqp(H.reshape(-1, m, m), q.reshape(-1, m), lo.reshape(-1, m), hi.reshape(-1, m))
    return x.reshape(*lead, m), {"iters": iters.reshape(lead)}
