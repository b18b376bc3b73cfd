"""Dynamics plugins with an embedded sm_100a implementation.

The reference takes any Python callable ``f_dyn(x, u, dt)`` (and optionally
``f_dyn_jac(x, u, dt) -> (A, B)``).  A Python callable cannot enter a fused CUDA kernel, so
the dynamics the reference ships are provided here as callables that ALSO
carry the id/parameters of their device implementation.  Passing one of these objects as
``f_dyn`` selects the fused whole-solve kernels; any other callable takes the generic path
(torch for f_dyn / Jacobians, CUDA for the Riccati recursion, costs and the KKT backward).

Each object is itself a valid reference-style plugin: ``obj(x, u, dt)`` evaluates the
dynamics with torch ops (for simulation / environments) and ``obj.jac(x, u, dt)`` is a
valid ``f_dyn_jac``.
"""
from __future__ import annotations

import math

import torch

from . import _abi


class NativeDynamics:
    """Base class: a reference-compatible f_dyn with a device twin."""
    dyn_id: int = -1
    nx: int = 0
    nu: int = 0
    params: tuple = ()

    def __call__(self, x, u, dt):
        raise NotImplementedError

    def jac(self, x, u, dt):
        raise NotImplementedError

    # LTI only: device matrices in the solve dtype
    def device_matrices(self, device, dtype):
        return None, None


class LTI(NativeDynamics):
    """x+ = A x + B u (examples/01_demo_linear.py:12-26)."""
    dyn_id = _abi.DYN_LTI

    def __init__(self, A, B):
        self.A = torch.as_tensor(A).detach().clone()
        self.B = torch.as_tensor(B).detach().clone()
        if self.A.ndim != 2 or self.A.shape[0] != self.A.shape[1] or self.B.ndim != 2 or self.B.shape[0] != self.A.shape[0]:
            raise ValueError("LTI needs A[nx,nx], B[nx,nu]")
        self.nx, self.nu = int(self.A.shape[0]), int(self.B.shape[1])
        self._cache = {}

    def device_matrices(self, device, dtype):
        key = (str(device), dtype)
        if key not in self._cache:
            self._cache[key] = (self.A.to(device=device, dtype=dtype).contiguous(),
                                self.B.to(device=device, dtype=dtype).contiguous())
        return self._cache[key]

    def __call__(self, x, u, dt):
        A, B = self.device_matrices(x.device, x.dtype)
        return x @ A.mT + u @ B.mT

    def jac(self, x, u, dt):
        A, B = self.device_matrices(x.device, x.dtype)
        N = x.shape[0]
        return A.unsqueeze(0).expand(N, -1, -1), B.unsqueeze(0).expand(N, -1, -1)


class DoubleIntegrator(LTI):
    """Explicit-Euler double integrator (examples/01_demo_linear.py:12-16); ``semi_implicit`` gives the
    variant of examples/01_demo_linear_AC.py:61-72 (velocity updated first)."""

    def __init__(self, dt: float, semi_implicit: bool = False):
        A = torch.tensor([[1.0, dt], [0.0, 1.0]], dtype=torch.float64)
        B = torch.tensor([[dt * dt if semi_implicit else 0.0], [dt]], dtype=torch.float64)
        super().__init__(A, B)
        self.dt = dt


class CartPole(NativeDynamics):
    """Explicit-Euler gym cart-pole (examples/02_demo_cartpole_regulation.py:29-80); defaults are
    CartPole-v1's constants, which is what ``CartPoleParams.from_gym()`` reads."""
    dyn_id = _abi.DYN_CARTPOLE
    nx, nu = 4, 1

    def __init__(self, m_c: float = 1.0, m_p: float = 0.1, l: float = 0.5, g: float = 9.8):
        self.params = (float(m_c), float(m_p), float(l), float(g))

    def __call__(self, x, u, dt):
        m_c, m_p, l, g = self.params
        vel, th, om = x[..., 1], x[..., 2], x[..., 3]
        f = u[..., 0]
        s, c = torch.sin(th), torch.cos(th)
        M, mpl = m_c + m_p, m_p * l
        temp = (f + mpl * om * om * s) / M
        thdd = (g * s - c * temp) / (l * (4.0 / 3.0 - m_p * c * c / M))
        vdd = temp - mpl * thdd * c / M
        return x + torch.stack((vel, vdd, om, thdd), dim=-1) * dt

    def jac(self, x, u, dt):
        m_c, m_p, l, g = self.params
        th, om, f = x[:, 2], x[:, 3], u[:, 0]
        s, c = torch.sin(th), torch.cos(th)
        M, mpl = m_c + m_p, m_p * l
        temp = (f + mpl * om * om * s) / M
        num = g * s - c * temp
        den = l * (4.0 / 3.0 - m_p * c * c / M)
        thdd = num / den
        t_f, t_th, t_om = 1.0 / M, mpl * om * om * c / M, 2 * mpl * om * s / M
        n_th = g * c + s * temp - c * t_th
        d_th = l * (2 * m_p * c * s / M)
        a_th = (n_th * den - num * d_th) / (den * den)
        a_om = -c * t_om / den
        a_f = -c * t_f / den
        r = mpl / M
        v_th = t_th - r * (a_th * c - thdd * s)
        v_om = t_om - r * a_om * c
        v_f = t_f - r * a_f * c
        N = x.shape[0]
        A = torch.eye(4, dtype=x.dtype, device=x.device).repeat(N, 1, 1)
        A[:, 0, 1] = dt
        A[:, 1, 2] = v_th * dt
        A[:, 1, 3] = v_om * dt
        A[:, 2, 3] = dt
        A[:, 3, 2] = a_th * dt
        A[:, 3, 3] = 1.0 + a_om * dt
        B = torch.zeros(N, 4, 1, dtype=x.dtype, device=x.device)
        B[:, 1, 0] = v_f * dt
        B[:, 3, 0] = a_f * dt
        return A, B


class Unicycle(NativeDynamics):
    """Explicit-Euler unicycle (examples/03_demo_unicycle_tracking.py:8-49); ``wrapped`` wraps theta to
    [-pi, pi) as examples/03_demo_unicycle_AC.py:107-125 does."""
    nx, nu = 3, 2

    def __init__(self, wrapped: bool = False):
        self.wrapped = bool(wrapped)
        self.dyn_id = _abi.DYN_UNICYCLE_WRAPPED if wrapped else _abi.DYN_UNICYCLE

    def __call__(self, x, u, dt):
        th, v, om = x[..., 2], u[..., 0], u[..., 1]
        tn = th + om * dt
        if self.wrapped:
            tn = torch.remainder(tn + math.pi, 2 * math.pi) - math.pi
        return torch.stack((x[..., 0] + v * torch.cos(th) * dt, x[..., 1] + v * torch.sin(th) * dt, tn), dim=-1)

    def jac(self, x, u, dt):
        th, v = x[:, 2], u[:, 0]
        s, c = torch.sin(th), torch.cos(th)
        N = x.shape[0]
        A = torch.eye(3, dtype=x.dtype, device=x.device).repeat(N, 1, 1)
        A[:, 0, 2] = -v * s * dt
        A[:, 1, 2] = v * c * dt
        B = torch.zeros(N, 3, 2, dtype=x.dtype, device=x.device)
        B[:, 0, 0] = c * dt
        B[:, 1, 0] = s * dt
        B[:, 2, 1] = dt
        return A, B
