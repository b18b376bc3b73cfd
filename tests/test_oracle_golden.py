"""Pins the CPU oracle (oracle/) against outputs of the unmodified reference
(tests/golden/*.npz, produced by oracle/gen_golden.py).  CPU only."""
import numpy as np
import pytest

import parity as P


@pytest.mark.parametrize("name", P.SOLVE_CASES)
def test_forward_replay_matches_reference(oracle, name):
    g = P.load(name)
    spec = P.spec_of(g)
    tol = P.case_tol(g)
    out = oracle.forward(spec, g["x0"], g["C"], g["c"], g["Cf"], g["cf"], g["xref"], g["uref"], g["Uinit"],
                         ls_cost=P.ls_cost_of(g),
                         replay=dict(iters=g["iters"], alpha_trace=g["alpha_trace"], flags=g["flags"]))
    assert np.array_equal(out["iters"], g["iters"])
    assert np.array_equal(out["alpha_trace"], g["alpha_trace"])
    ex, eu = P.rel_err(out["X"], g["X"]), P.rel_err(out["U"], g["U"])
    tx, tu = P.elem_tol(g, "X"), P.elem_tol(g, "U")
    rep = P.reproducible(g)
    assert (ex[rep] <= tx[rep]).all() and (eu[rep] <= tu[rep]).all(), (ex, tx, eu, tu)
    well = (tu <= tol) & (tx <= tol)   # elements where the reference itself is reproducible to tol
    print(f"{name}: {well.sum()}/{len(well)} elements held to {tol:g}; max err X {ex[well].max() if well.any() else 0:.2e} "
          f"U {eu[well].max() if well.any() else 0:.2e}")
    assert (P.rel_err(out["A"], g["A"])[rep] <= np.maximum(tx, tu)[rep]).all()
    assert (P.rel_err(out["Bm"], g["Bm"])[rep] <= np.maximum(tx, tu)[rep]).all()
    assert np.array_equal(out["tight"][well], g["tight"][well])  # active-set mask: bit-exact
    assert (P.rel_err(out["cost"][:, None], g["cost"][:, None])[rep] <= 10 * np.maximum(tx, tu)[rep]).all()
    print(f"{name}: {rep.sum()}/{len(rep)} elements reproducible in the reference itself")


@pytest.mark.parametrize("name", P.SOLVE_CASES)
def test_forward_lockstep_decisions(oracle, name):
    g = P.load(name)
    spec = P.spec_of(g)
    dt = g["x0"].dtype
    out = oracle.forward(spec, g["x0"], g["C"], g["c"], g["Cf"], g["cf"], g["xref"], g["uref"], g["Uinit"],
                         ls_cost=P.ls_cost_of(g))
    same, below, bad = P.lockstep(g, out["iters"], out["alpha_trace"], out["flags"], P.MARGIN[dt])
    assert not bad, f"decisions differ above the noise margin: {bad[:5]}"
    B = g["x0"].shape[0]
    print(f"{name}: {same}/{B} traces identical, {below} departed below margin")
    # end-to-end: final objective agrees wherever the reference itself is reproducible
    tu = np.maximum(P.elem_tol(g, "U"), P.elem_tol(g, "X"))
    well = tu <= P.case_tol(g)
    scale = np.maximum(np.abs(g["cost"].astype(np.float64)), 1.0)
    cerr = np.abs(out["cost"].astype(np.float64) - g["cost"]) / scale
    assert (cerr[well] <= 200 * P.case_tol(g)).all(), cerr
    assert np.isfinite(out["cost"][P.reproducible(g)]).all()
    # elements whose trace is identical match to tolerance
    idx = [b for b in range(B) if out["iters"][b] == g["iters"][b]
           and np.array_equal(out["alpha_trace"][b], g["alpha_trace"][b])]
    idx = [b for b in idx if P.reproducible(g)[b]]
    if idx:
        assert (P.rel_err(out["U"][idx], g["U"][idx]) <= tu[idx]).all()
        w = [b for b in idx if well[b]]
        assert np.array_equal(out["tight"][w], g["tight"][w])


@pytest.mark.parametrize("name", P.SOLVE_CASES)
def test_first_iteration_stages(oracle, name):
    """Riccati gains and line-search candidates on identical inputs (single-shot, well conditioned)."""
    g = P.load(name)
    spec = P.spec_of(g)
    dt = g["x0"].dtype
    tol = P.case_tol(g)
    n = g["it0_A"].shape[0]
    A, Bm = oracle.linearize(spec, g["it0_X"], g["it0_U"])
    assert P.rel_err(A, g["it0_A"]).max() <= tol and P.rel_err(Bm, g["it0_Bm"]).max() <= tol
    K, k, _ = oracle.riccati(spec, g["it0_A"], g["it0_Bm"], g["it0_lx"], g["it0_lu"], g["it0_lxx"], g["it0_lxu"],
                             g["it0_luu"], g["it0_lxN"], g["it0_lxxN"])
    # fp32: the recursion's own rounding error (reference fp32 vs reference fp64) widens the bar
    fK = 8 * P.rel_err(g["it0_K"], g["it0_K64"]) if "it0_K64" in g else 0.0
    fk = 8 * P.rel_err(g["it0_k"], g["it0_k64"]) if "it0_k64" in g else 0.0
    okK = np.asarray(fK) <= 8 * P.REPRO_LIMIT   # the reference's own fp32 gains still carry digits
    okk = np.asarray(fk) <= 8 * P.REPRO_LIMIT
    assert (np.where(okK, P.rel_err(K, g["it0_K"]), 0) <= 20 * tol + fK).all(), P.rel_err(K, g["it0_K"]).max()
    assert (np.where(okk, P.rel_err(k, g["it0_k"]), 0) <= 20 * tol + fk).all(), P.rel_err(k, g["it0_k"]).max()
    ls = P.ls_cost_of(g)
    Cm, c, Cf, cf = ls if ls is not None else (g["C"], g["c"], g["Cf"], g["cf"])
    Xc, Uc, costs = oracle.linesearch(spec, g["x0"][:n], g["it0_X"], g["it0_U"], g["it0_K"], g["it0_k"],
                                      Cm, c, Cf, cf, g["xref"][:n], g["uref"][:n])
    slack = P.STAGE_SLACK.get(name, 1.0)
    assert P.rel_err(Xc, g["it0_Xc"]).max() <= slack * tol
    assert P.rel_err(Uc, g["it0_Uc"]).max() <= slack * tol
    assert P.rel_err(costs, g["it0_costs"]).max() <= 10 * slack * tol


@pytest.mark.parametrize("name", P.SOLVE_CASES)
@pytest.mark.parametrize("loss", ["Usum", "U0sum", "weighted"])
def test_backward_matches_reference(oracle, name, loss):
    """Implicit-differentiation backward on the reference's saved tensors (deterministic)."""
    g = P.load(name)
    spec = P.spec_of(g)
    dt = g["x0"].dtype
    tol = P.case_tol(g)
    X, U = g["X"], g["U"]
    if loss == "Usum":
        gX, gU = np.zeros_like(X), np.ones_like(U)
    elif loss == "U0sum":
        gX, gU = np.zeros_like(X), np.zeros_like(U)
        gU[:, 0] = 1
    else:
        gX, gU = g["wX"], g["wU"]
    out = oracle.backward(spec, X, U, g["C"], g["xref"], g["uref"], g["tight"], gX, gU,
                          Cf_batched=int(g["Cf"].ndim == 3))
    for nm in ("C", "c", "Cf", "cf", "x0", "xref", "uref"):
        ref = g[f"grad_{loss}_{nm}"]
        got = out[f"grad_{nm}"].reshape(ref.shape)
        scale = max(np.abs(ref).max(), 1.0)
        err = np.abs(got.astype(np.float64) - ref).max() / scale
        floor = 0.0
        if f"grad64_{loss}_{nm}" in g:   # fp32: how far the reference's own fp32 gradient is from fp64
            floor = 4 * np.abs(g[f"grad64_{loss}_{nm}"] - ref).max() / scale
        if floor > 4 * P.REPRO_LIMIT:
            continue   # the reference's own fp32 gradient has lost its digits against fp64: nothing to pin
        assert err <= 50 * tol + floor, (nm, err, floor)


@pytest.mark.parametrize("tag", ["f32", "f64"])
@pytest.mark.parametrize("m", [1, 2, 3, 4])
def test_pnqp_known_answers(oracle, tag, m):
    g = P.load("pnqp_kat")
    H, q, lo, hi = (g[f"{tag}_m{m}_{k}"] for k in ("H", "q", "lo", "hi"))
    x, iters = oracle.pnqp(H, q, lo, hi)
    tol = 1e-5 if tag == "f32" else 1e-8
    assert np.abs(x - g[f"{tag}_m{m}_x"]).max() <= tol * max(1.0, np.abs(g[f"{tag}_m{m}_x"]).max())
    assert (iters == g[f"{tag}_m{m}_iters"]).mean() >= 0.8


@pytest.mark.parametrize("tag", ["f32", "f64"])
@pytest.mark.parametrize("dyn", ["cartpole", "unicycle"])
def test_finite_difference_jacobians(oracle, tag, dyn):
    g = P.load("fd_jacobian_kat")
    x, u = g[f"{dyn}_{tag}_x"], g[f"{dyn}_{tag}_u"]
    spec = dict(nx=x.shape[1], nu=u.shape[1], T=1, dt=0.05 if dyn == "cartpole" else 0.1, dyn=dyn,
                dyn_params=(1.0, 0.1, 0.5, 9.8), jac="finite_diff")
    X = np.concatenate([x[:, None], x[:, None]], axis=1)
    A, Bm = oracle.linearize(spec, X, u[:, None])
    tol = 2e-3 if tag == "f32" else 1e-9   # fp32 central differences with eps=1e-4 are noise-limited
    assert np.abs(A[:, 0] - g[f"{dyn}_{tag}_A"]).max() <= tol
    assert np.abs(Bm[:, 0] - g[f"{dyn}_{tag}_B"]).max() <= tol


def test_config1_double_integrator_demo_known_answer(oracle):
    """SURVEY §8d config 1 / §8c soft KAT (examples/01_demo_linear.py as the script runs it): the oracle driven in
    closed loop for 150 cold-started solves of 50 agents reproduces the reference's recorded states and its
    position MSE of 0.1296."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "config1_di_closed_loop.npz"))
    B, DT, H, N_SIM, nx, nu = 50, 0.05, 20, 150, 2, 1
    C = np.zeros((H, 3, 3))
    C[:, 0, 0], C[:, 1, 1], C[:, 2, 2] = 100.0, 1.0, 0.01
    c, Cf, cf = np.zeros((H, 3)), C[0] * 10, np.zeros(3)
    ref_len = N_SIM + H + 1
    t_full = np.arange(ref_len) * DT
    pos = g["amp"] * np.sin(g["om"] * t_full[None])
    vel = g["amp"] * g["om"] * np.cos(g["om"] * t_full[None])
    xref = np.stack([pos, vel], axis=2)
    A, Bm = np.array([[1.0, DT], [0.0, 1.0]]), np.array([[0.0], [DT]])
    spec = dict(nx=nx, nu=nu, T=H, dt=DT, dyn="lti", A=A, B=Bm, reg_eps=1e-6, max_iter=40, u_min=[-50.0], u_max=[50.0])
    x = np.zeros((B, nx))
    Xs = [x]
    for k in range(N_SIM):
        out = oracle.forward(spec, x, C, c, Cf, cf, np.ascontiguousarray(xref[:, k:k + H + 1]), np.zeros((B, H, nu)),
                             np.zeros((B, H, nu)), want_AB=False)
        x = x @ A.T + out["U"][:, 0] @ Bm.T
        Xs.append(x)
    Xs = np.stack(Xs, 1)
    mse = np.mean((Xs[:, :, 0] - xref[:, :N_SIM + 1, 0]) ** 2)
    assert abs(mse - 0.1296) < 5e-5 and abs(mse - float(g["mse_analytic"])) < 1e-9
    assert np.abs(Xs - g["Xs"]).max() < 1e-6
