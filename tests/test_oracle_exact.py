"""CPU: the numpy restatement of the EXACT implicit differentiation (tests/exact_ref.py, the checker of
compat="exact") is pinned by central finite differences of the oracle's forward solve on a linear-quadratic
problem, where one Newton step is the exact optimum and every gradient is well defined."""
import numpy as np

from exact_ref import exact_backward


def test_exact_backward_matches_finite_differences_of_the_forward_solve(oracle):
    rng = np.random.default_rng(0)
    nx, nu, T, B = 3, 2, 8, 3
    n = nx + nu
    A = np.eye(nx) + 0.2 * rng.standard_normal((nx, nx))
    Bm = 0.5 * rng.standard_normal((nx, nu))
    M = 0.05 * rng.standard_normal((n, n))
    C1 = np.diag(rng.uniform(0.5, 2.0, n)) + M @ M.T
    d = dict(C=np.tile(C1, (T, 1, 1)) * (1 + 0.05 * np.arange(T))[:, None, None], c=0.1 * rng.standard_normal((T, n)),
             Cf=np.zeros((n, n)), cf=np.zeros(n), x0=0.3 * rng.standard_normal((B, nx)),
             xref=0.1 * rng.standard_normal((B, T + 1, nx)), uref=0.05 * rng.standard_normal((B, T, nu)))
    d["Cf"][:nx, :nx] = 3 * C1[:nx, :nx]
    d["cf"][:nx] = 0.1 * rng.standard_normal(nx)
    spec = dict(nx=nx, nu=nu, T=T, dt=0.1, dyn="lti", A=A, B=Bm, reg_eps=1e-8, max_iter=1)

    def fwd(dd):
        return oracle.forward(spec, dd["x0"], dd["C"], dd["c"], dd["Cf"], dd["cf"], dd["xref"], dd["uref"], np.zeros((B, T, nu)),
                              want_AB=True, want_trace=False)

    f = fwd(d)
    wX, wU = rng.standard_normal(f["X"].shape), rng.standard_normal(f["U"].shape)
    loss = lambda ff: float((wX * ff["X"]).sum() + (wU * ff["U"]).sum())
    Cb, Cfb = np.broadcast_to(d["C"], (B,) + d["C"].shape), np.broadcast_to(d["Cf"], (B,) + d["Cf"].shape)
    g = exact_backward(nx, nu, f["X"], f["U"], Cb, d["c"], Cfb, d["cf"], d["xref"], d["uref"], f["tight"], wX, wU, f["A"], f["Bm"],
                       spec["reg_eps"])

    def fd(key, idx, sym=None, eps=1e-3):
        dp, dm = {k: v.copy() for k, v in d.items()}, {k: v.copy() for k, v in d.items()}
        for i in (idx, sym):
            if i is not None:
                dp[key][i] += eps
                dm[key][i] -= eps
        return (loss(fwd(dp)) - loss(fwd(dm))) / (2 * eps)

    checks = [("x0", (1, 0), None, g["grad_x0"][1, 0]), ("x0", (0, nx - 1), None, g["grad_x0"][0, nx - 1]),
              ("xref", (1, 3, 0), None, g["grad_xref"][1, 3, 0]), ("xref", (0, T, 1), None, g["grad_xref"][0, T, 1]),
              ("uref", (2, 2, 0), None, g["grad_uref"][2, 2, 0]), ("c", (4, 1), None, g["grad_c"][:, 4, 1].sum()),
              ("c", (0, nx), None, g["grad_c"][:, 0, nx].sum()), ("C", (3, 0, 0), None, g["grad_C"][:, 3, 0, 0].sum()),
              ("C", (5, 1, nx), (5, nx, 1), g["grad_C"][:, 5, 1, nx].sum() + g["grad_C"][:, 5, nx, 1].sum()),
              ("Cf", (1, 1), None, g["grad_Cf"][:, 1, 1].sum()), ("cf", (0,), None, g["grad_cf"][:, 0].sum())]
    for key, idx, sym, an in checks:
        num = fd(key, idx, sym)
        assert abs(an - num) <= 1e-6 * max(1.0, abs(num)), (key, idx, an, num)
