"""numpy restatement of the EXACT implicit differentiation of a converged box-constrained iLQR solve
(TEST INFRASTRUCTURE for compat="exact"; the reference has no such mode — its backward carries
quirks Q1-Q3, Q6 — so this is pinned by finite differences of the forward solve, tests/test_oracle_exact.py).

Amos et al., "Differentiable MPC" (2018), eq. 8-9 and Module 1/2: the gradient of a loss w.r.t. the LQR data is
obtained from one more LQR solve on the KKT system of the last iLQR iteration with the upstream gradient as linear
term, zero initial state, and the active inputs frozen.
"""
import numpy as np


def exact_backward(nx, nu, X, U, C, c, Cf, cf, xref, uref, tight, gX, gU, A, Bm, reg_eps=0.0):
    """All arrays carry a leading batch dim; C [B,T,n,n], Cf [B,n,n] (broadcast them before the call).
    Returns dict(grad_C, grad_c, grad_Cf, grad_cf, grad_x0, grad_xref, grad_uref, dX, dU) per element."""
    B, T = U.shape[0], U.shape[1]
    n = nx + nu
    out = dict(grad_C=np.zeros((B, T, n, n)), grad_c=np.zeros((B, T, n)), grad_Cf=np.zeros((B, n, n)), grad_cf=np.zeros((B, n)),
               grad_x0=np.zeros((B, nx)), grad_xref=np.zeros((B, T + 1, nx)), grad_uref=np.zeros((B, T, nu)),
               dX=np.zeros((B, T + 1, nx)), dU=np.zeros((B, T, nu)))
    for b in range(B):
        V = Cf[b, :nx, :nx].astype(np.float64)
        v = -gX[b, T].astype(np.float64)
        Ks, ks = [None] * T, [None] * T
        for t in range(T - 1, -1, -1):
            At, Bt, Ct = A[b, t].astype(np.float64), Bm[b, t].astype(np.float64), C[b, t].astype(np.float64)
            Qxx = Ct[:nx, :nx] + At.T @ V @ At
            Qxu = Ct[:nx, nx:] + At.T @ V @ Bt
            Quu = Ct[nx:, nx:] + Bt.T @ V @ Bt
            qx = -gX[b, t] + At.T @ v
            qu = -gU[b, t] + Bt.T @ v
            free = ~tight[b, t].astype(bool)
            K, k = np.zeros((nu, nx)), np.zeros(nu)
            if free.any():
                Qff = Quu[np.ix_(free, free)] + 1e-8 * reg_eps * np.eye(int(free.sum()))
                K[free] = -np.linalg.solve(Qff, Qxu.T[free])
                k[free] = -np.linalg.solve(Qff, qu[free])
            V = Qxx + K.T @ Quu @ K + K.T @ Qxu.T + Qxu @ K
            v = qx + K.T @ Quu @ k + K.T @ qu + Qxu @ k
            Ks[t], ks[t] = K, k
        out["grad_x0"][b] = -v
        dx = np.zeros(nx)
        for t in range(T):
            du = Ks[t] @ dx + ks[t]
            dtau = np.concatenate([dx, du])
            tau = np.concatenate([X[b, t] - xref[b, t], U[b, t] - uref[b, t]]).astype(np.float64)
            out["grad_C"][b, t] = -0.5 * (np.outer(dtau, tau) + np.outer(tau, dtau))
            out["grad_c"][b, t] = -dtau
            g = C[b, t].astype(np.float64) @ dtau
            out["grad_xref"][b, t], out["grad_uref"][b, t] = g[:nx], g[nx:]
            out["dX"][b, t], out["dU"][b, t] = dx, du
            dx = A[b, t] @ dx + Bm[b, t] @ du
        tauN = (X[b, T] - xref[b, T]).astype(np.float64)
        out["grad_Cf"][b, :nx, :nx] = -0.5 * (np.outer(dx, tauN) + np.outer(tauN, dx))
        out["grad_cf"][b, :nx] = -dx
        out["grad_xref"][b, T] = Cf[b, :nx, :nx] @ dx
        out["dX"][b, T] = dx
    return out
