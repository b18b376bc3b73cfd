"""-m gpu parity tests proper: the CUDA path (through the C ABI, tacdmpc_b200.ops) against
 (1) the golden vectors recorded from the unmodified reference, and
 (2) the CPU oracle on the same inputs.
Protocol and tolerances: tests/parity.py."""
import numpy as np
import pytest
import torch

import parity as P

pytestmark = pytest.mark.gpu


def _refresh_knobs(request):
    """The library caches its TACD_* knobs; re-read them now and again when monkeypatch has restored the environment."""
    from tacdmpc_b200 import _lib
    _lib.lib().tacd_refresh_env()
    request.addfinalizer(lambda: _lib.lib().tacd_refresh_env())

# Multipliers on the north-star tolerance (1e-5 fp32 / 1e-9 fp64).  Each is at most 4x what the kernels achieve on the worst
# golden case (the ledger tests/conftest.py prints; DESIGN.md 5 has the table).
COST_K = 1.0          # replay: final objective, in units of the element's value tolerance (achieved 0.19; round 1 allowed 10)
FREE_COST_K = 1.0     # free-running final objective, in units of the north-star tolerance (achieved 0.19; round 1 allowed 200)
GRAD_K = 2.0          # gradients beyond the reference's own fp32-vs-fp64 floor (achieved <= 0.49 on 16 of 18 cases; round 1: 50)
# the two cases whose one-sample floor underestimates the spread of fp32 answers (achieved 7.0 and 35.4)
GRAD_K_CASE = {"unicycle_loose_f32": 28.0, "lti16_f32": 50.0}
PNQP_ITERS_FRAC = 0.85   # achieved 0.896 (m = 4): the stop test max|dx| < 1e-5 is decided by rounding on a few problems


@pytest.fixture(scope="module")
def G():
    import gpu_common
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return gpu_common


@pytest.fixture(params=["group", "thread"])
def fwd_kernel(request, monkeypatch):
    """The shipped small plugins have two whole-solve forward kernels: five lanes per element (default,
    csrc/ilqr_group.cuh) and one thread per element (csrc/ilqr_small.cuh; finite differences, > 5 alphas).
    TACD_FWD_KERNEL=thread forces the second; both are held to the same parity bar."""
    if request.param == "thread":
        monkeypatch.setenv("TACD_FWD_KERNEL", "thread")
    else:
        monkeypatch.delenv("TACD_FWD_KERNEL", raising=False)
    _refresh_knobs(request)
    return request.param


def _skip_dup(name, fwd_kernel):
    if fwd_kernel == "thread" and name not in P.SMALL_CASES:
        pytest.skip("generic CTA-per-element path: no second kernel to select")


@pytest.mark.parametrize("name", P.SOLVE_CASES)
def test_forward_replay_matches_reference(G, name, fwd_kernel):
    _skip_dup(name, fwd_kernel)
    g = P.load(name)
    spec = P.spec_of(g)
    tol = P.case_tol(g, values=True)
    out = G.forward(spec, g, replay=True)
    assert np.array_equal(out["iters"], g["iters"])
    assert np.array_equal(out["alpha_trace"][:, :g["alpha_trace"].shape[1]], g["alpha_trace"])
    ex, eu = P.rel_err(out["X"], g["X"]), P.rel_err(out["U"], g["U"])
    tx, tu = P.elem_tol(g, "X"), P.elem_tol(g, "U")
    rep = P.reproducible(g)
    assert (ex[rep] <= tx[rep]).all() and (eu[rep] <= tu[rep]).all(), (ex, tx, eu, tu)
    well = (tu <= tol) & (tx <= tol)
    print(f"{name}: {well.sum()}/{len(well)} elements held to {tol:g}; max err X {ex[well].max() if well.any() else 0:.2e} "
          f"U {eu[well].max() if well.any() else 0:.2e}")
    tag = f"{name}[{fwd_kernel}]"
    P.record(tag, "replay elems held to north-star tol", float(well.mean()), note=f"({int(well.sum())}/{len(well)} at {tol:g})")
    if well.any():
        P.record(tag, "replay max relerr X (held elems)", float(ex[well].max()), tol)
        P.record(tag, "replay max relerr U (held elems)", float(eu[well].max()), tol)
    if rep.any():
        P.record(tag, "replay max relerr/elem_tol (all reproducible)", float(np.maximum(ex / tx, eu / tu)[rep].max()), 1.0)
    assert (P.rel_err(out["A"], g["A"])[rep] <= np.maximum(tx, tu)[rep]).all()
    assert (P.rel_err(out["Bm"], g["Bm"])[rep] <= np.maximum(tx, tu)[rep]).all()
    assert np.array_equal(out["tight"][well], g["tight"][well])
    ce = P.rel_err(out["cost"][:, None], g["cost"][:, None])
    if rep.any():
        P.record(tag, "replay cost relerr/elem_tol", float((ce / np.maximum(tx, tu))[rep].max()), COST_K)
    assert (ce[rep] <= COST_K * np.maximum(tx, tu)[rep]).all()
    print(f"{name}: {rep.sum()}/{len(rep)} elements reproducible in the reference itself")


@pytest.mark.parametrize("name", P.SOLVE_CASES)
def test_forward_lockstep_decisions(G, name, fwd_kernel):
    _skip_dup(name, fwd_kernel)
    g = P.load(name)
    spec = P.spec_of(g)
    dt = g["x0"].dtype
    out = G.forward(spec, g)
    same, below, bad = P.lockstep(g, out["iters"], out["alpha_trace"], out["flags"], P.MARGIN[dt])
    assert not bad, f"decisions differ above the noise margin: {bad[:5]}"
    B = g["x0"].shape[0]
    print(f"{name}: {same}/{B} traces identical, {below} departed below margin")
    tu = np.maximum(P.elem_tol(g, "U"), P.elem_tol(g, "X"))
    well = tu <= P.case_tol(g)
    scale = np.maximum(np.abs(g["cost"].astype(np.float64)), 1.0)
    cerr = np.abs(out["cost"].astype(np.float64) - g["cost"]) / scale
    if well.any():
        P.record(f"{name}[{fwd_kernel}]", "free-run cost relerr / tol", float(cerr[well].max() / P.case_tol(g)), FREE_COST_K)
    P.record(f"{name}[{fwd_kernel}]", "free-run identical trace frac", same / B, note="(recorded, not a bar)")
    assert (cerr[well] <= FREE_COST_K * P.case_tol(g)).all(), cerr
    assert np.isfinite(out["cost"][P.reproducible(g)]).all()
    idx = [b for b in range(B) if out["iters"][b] == g["iters"][b]
           and np.array_equal(out["alpha_trace"][b, :g["alpha_trace"].shape[1]], g["alpha_trace"][b])]
    idx = [b for b in idx if P.reproducible(g)[b]]
    if idx:
        assert (P.rel_err(out["U"][idx], g["U"][idx]) <= tu[idx]).all()
        w = [b for b in idx if well[b]]
        assert np.array_equal(out["tight"][w], g["tight"][w])


@pytest.mark.parametrize("name", P.SOLVE_CASES)
def test_forward_matches_oracle_in_replay(G, oracle, name, fwd_kernel):
    """CUDA vs CPU oracle on the same inputs and the same decision sequence."""
    _skip_dup(name, fwd_kernel)
    g = P.load(name)
    spec = P.spec_of(g)
    rp = dict(iters=g["iters"], alpha_trace=g["alpha_trace"], flags=g["flags"])
    ref = oracle.forward(spec, g["x0"], g["C"], g["c"], g["Cf"], g["cf"], g["xref"], g["uref"], g["Uinit"],
                         ls_cost=P.ls_cost_of(g), replay=rp)
    out = G.forward(spec, g, replay=True, cost_trace=True)
    tx, tu = P.elem_tol(g, "X"), P.elem_tol(g, "U")
    rep = P.reproducible(g)
    assert (P.rel_err(out["X"], ref["X"])[rep] <= tx[rep]).all()
    assert (P.rel_err(out["U"], ref["U"])[rep] <= tu[rep]).all()
    assert np.array_equal(out["iters"], ref["iters"])
    # the non-PD flag of an element the reference itself cannot reproduce is decided by rounding
    assert np.array_equal(out["flags"][rep], ref["flags"][rep])


@pytest.mark.parametrize("name", P.SOLVE_CASES)
def test_first_iteration_stages(G, name):
    from tacdmpc_b200 import ops
    g = P.load(name)
    spec = P.spec_of(g)
    dt = g["x0"].dtype
    tol = P.case_tol(g)
    s = G.gpu_spec(spec, dt)
    n = g["it0_A"].shape[0]
    d = G.dev
    A, Bm = ops.linearize(s, d(g["it0_X"]), d(g["it0_U"]))
    assert P.rel_err(A.cpu().numpy(), g["it0_A"]).max() <= tol and P.rel_err(Bm.cpu().numpy(), g["it0_Bm"]).max() <= tol
    X = ops.rollout(s, d(g["x0"][:n]), d(g["it0_U"]))
    assert P.rel_err(X.cpu().numpy(), g["it0_X"]).max() <= P.STAGE_SLACK.get(name, 1.0) * tol
    K, k, _ = ops.riccati(s, *(d(g["it0_" + nm]) for nm in ("A", "Bm", "lx", "lu", "lxx", "lxu", "luu", "lxN", "lxxN")))
    K, k = K.cpu().numpy(), k.cpu().numpy()
    fK = 8 * P.rel_err(g["it0_K"], g["it0_K64"]) if "it0_K64" in g else 0.0
    fk = 8 * P.rel_err(g["it0_k"], g["it0_k64"]) if "it0_k64" in g else 0.0
    okK = np.asarray(fK) <= 8 * P.REPRO_LIMIT   # the reference's own fp32 gains still carry digits
    okk = np.asarray(fk) <= 8 * P.REPRO_LIMIT
    assert (np.where(okK, P.rel_err(K, g["it0_K"]), 0) <= 20 * tol + fK).all(), P.rel_err(K, g["it0_K"]).max()
    assert (np.where(okk, P.rel_err(k, g["it0_k"]), 0) <= 20 * tol + fk).all(), P.rel_err(k, g["it0_k"]).max()
    ls = P.ls_cost_of(g)
    Cm, c, Cf, cf = ls if ls is not None else (g["C"], g["c"], g["Cf"], g["cf"])
    Xc, Uc, costs = ops.linesearch(s, d(g["x0"][:n]), d(g["it0_X"]), d(g["it0_U"]), d(g["it0_K"]), d(g["it0_k"]),
                                   d(Cm), d(c), d(Cf), d(cf), d(g["xref"][:n]), d(g["uref"][:n]))
    slack = P.STAGE_SLACK.get(name, 1.0)
    assert P.rel_err(Xc.cpu().numpy(), g["it0_Xc"]).max() <= slack * tol
    assert P.rel_err(Uc.cpu().numpy(), g["it0_Uc"]).max() <= slack * tol
    assert P.rel_err(costs.cpu().numpy(), g["it0_costs"]).max() <= 10 * slack * tol
    cost = ops.objective(s, d(g["X"]), d(g["U"]), d(g["C"]), d(g["c"]), d(g["Cf"]), d(g["cf"]), d(g["xref"]), d(g["uref"]))
    assert P.rel_err(cost.cpu().numpy()[:, None], g["cost"][:, None]).max() <= 10 * tol


@pytest.fixture(params=["group", "thread"])
def bwd_kernel(request, monkeypatch):
    """Two backward kernels for the shipped small plugins: one thread per element (default,
    csrc/ilqr_small.cuh) and five lanes per element (csrc/ilqr_group_bwd.cuh, opt-in with
    TACD_BWD_KERNEL=group: measured slower at large batches).  Both are held to the same bar."""
    if request.param == "group":
        monkeypatch.setenv("TACD_BWD_KERNEL", "group")
    else:
        monkeypatch.delenv("TACD_BWD_KERNEL", raising=False)
    _refresh_knobs(request)
    return request.param


@pytest.mark.parametrize("name", P.SOLVE_CASES)
@pytest.mark.parametrize("loss", ["Usum", "U0sum", "weighted"])
def test_backward_matches_reference(G, oracle, name, loss, bwd_kernel):
    if bwd_kernel == "group" and name not in P.SMALL_CASES:
        pytest.skip("generic CTA-per-element path: no second kernel to select")
    from tacdmpc_b200 import ops
    g = P.load(name)
    spec = P.spec_of(g)
    dt = g["x0"].dtype
    tol = P.case_tol(g)
    s = G.gpu_spec(spec, dt)
    d = G.dev
    X, U = g["X"], g["U"]
    if loss == "Usum":
        gX, gU = np.zeros_like(X), np.ones_like(U)
    elif loss == "U0sum":
        gX, gU = np.zeros_like(X), np.zeros_like(U)
        gU[:, 0] = 1
    else:
        gX, gU = g["wX"], g["wU"]
    want = ("grad_C", "grad_c", "grad_Cf", "grad_cf", "grad_x0", "grad_xref", "grad_uref", "dX", "dU")
    out = ops.ilqr_backward(s, d(X), d(U), d(g["C"]), d(g["xref"]), d(g["uref"]), d(g["tight"]), d(gX), d(gU),
                            Cf_batched=g["Cf"].ndim == 3, want=want)
    torch.cuda.synchronize()
    out = {k: v.cpu().numpy() for k, v in out.items()}
    ref = oracle.backward(spec, X, U, g["C"], g["xref"], g["uref"], g["tight"], gX, gU, Cf_batched=int(g["Cf"].ndim == 3))
    # fp32: if ANY of the reference's own gradients has lost its digits against fp64, the KKT recursion on
    # these saved tensors is numerically meaningless (open-loop-unstable lti8_f32): nothing to pin
    floors = [4 * np.abs(g[f"grad64_{loss}_{nm}"] - g[f"grad_{loss}_{nm}"]).max() / max(np.abs(g[f"grad_{loss}_{nm}"]).max(), 1.0)
              for nm in ("C", "c", "Cf", "cf", "xref", "uref") if f"grad64_{loss}_{nm}" in g]
    if floors and max(floors) > 4 * P.REPRO_LIMIT:
        pytest.skip(f"reference fp32 gradients are {max(floors) / 4:.1e} away from its fp64 gradients")
    for nm in ("C", "c", "Cf", "cf", "x0", "xref", "uref"):
        gold = g[f"grad_{loss}_{nm}"]
        got = out[f"grad_{nm}"].reshape(gold.shape)
        scale = max(np.abs(gold).max(), 1.0)
        err = np.abs(got.astype(np.float64) - gold).max() / scale
        floor = 0.0
        if f"grad64_{loss}_{nm}" in g:
            floor = 4 * np.abs(g[f"grad64_{loss}_{nm}"] - gold).max() / scale
        if floor > 4 * P.REPRO_LIMIT:
            continue   # the reference's own fp32 gradient has lost its digits against fp64: nothing to pin
        tag = f"{name}[{bwd_kernel}]"
        GK = GRAD_K_CASE.get(name, GRAD_K)
        P.record(tag, f"bwd {loss} grad_{nm} vs reference: (err - floor)/tol", max(err - floor, 0.0) / tol, GK)
        assert err <= GK * tol + floor, (nm, err, floor)
        # and against the oracle
        eo = np.abs(got.astype(np.float64) - ref[f"grad_{nm}"].reshape(gold.shape)).max() / scale
        P.record(tag, f"bwd {loss} grad_{nm} vs oracle: (err - floor)/tol", max(eo - floor, 0.0) / tol, GK)
        assert eo <= GK * tol + floor, ("oracle", nm, eo)
    assert np.all(out["grad_x0"] == 0)  # quirk Q3


@pytest.mark.parametrize("name", ["cartpole_shared_f32", "unicycle_tight_f64", "lti8_T20_f32"])
def test_backward_external_jacobians(G, name):
    """TACD_DYN_EXTERNAL (A, B supplied as tensors: the generic-callable path) gives the same gradients."""
    from tacdmpc_b200 import _abi, ops
    g = P.load(name)
    spec = P.spec_of(g)
    dt = g["x0"].dtype
    s = G.gpu_spec(spec, dt)
    d = G.dev
    args = (d(g["X"]), d(g["U"]), d(g["C"]), d(g["xref"]), d(g["uref"]), d(g["tight"]), d(g["wX"]), d(g["wU"]))
    a = ops.ilqr_backward(s, *args, Cf_batched=g["Cf"].ndim == 3)
    s.dyn_id = _abi.DYN_EXTERNAL
    b = ops.ilqr_backward(s, *args, A=d(g["A"]), Bm=d(g["Bm"]), Cf_batched=g["Cf"].ndim == 3)
    for k in a:
        x, y = a[k].cpu().numpy().astype(np.float64), b[k].cpu().numpy().astype(np.float64)
        assert np.abs(x - y).max() <= 50 * P.case_tol(g) * max(1.0, np.abs(x).max()), k


@pytest.mark.parametrize("tag", ["f32", "f64"])
@pytest.mark.parametrize("m", [1, 2, 3, 4])
def test_pnqp_known_answers(G, tag, m):
    from tacdmpc_b200 import ops
    g = P.load("pnqp_kat")
    H, q, lo, hi = (G.dev(g[f"{tag}_m{m}_{k}"]) for k in ("H", "q", "lo", "hi"))
    x, iters = ops.pnqp(H, q, lo, hi)
    tol = 1e-5 if tag == "f32" else 1e-8
    ref = g[f"{tag}_m{m}_x"]
    assert np.abs(x.cpu().numpy() - ref).max() <= tol * max(1.0, np.abs(ref).max())
    frac = float((iters.cpu().numpy() == g[f"{tag}_m{m}_iters"]).mean())
    P.record(f"pnqp_{tag}_m{m}", "x max abs err / tol", float(np.abs(x.cpu().numpy() - ref).max() / (tol * max(1.0, np.abs(ref).max()))), 1.0)
    P.record(f"pnqp_{tag}_m{m}", "iteration counts equal frac", frac, PNQP_ITERS_FRAC)
    assert frac >= PNQP_ITERS_FRAC


@pytest.mark.parametrize("tag", ["f32", "f64"])
@pytest.mark.parametrize("dyn", ["cartpole", "unicycle"])
def test_finite_difference_jacobians(G, tag, dyn):
    from tacdmpc_b200 import ops
    g = P.load("fd_jacobian_kat")
    x, u = g[f"{dyn}_{tag}_x"], g[f"{dyn}_{tag}_u"]
    spec = dict(nx=x.shape[1], nu=u.shape[1], T=1, dt=0.05 if dyn == "cartpole" else 0.1, dyn=dyn,
                dyn_params=(1.0, 0.1, 0.5, 9.8), jac="finite_diff")
    s = G.gpu_spec(spec, x.dtype)
    X = np.concatenate([x[:, None], x[:, None]], axis=1)
    A, Bm = ops.linearize(s, G.dev(X), G.dev(u[:, None]))
    tol = 2e-3 if tag == "f32" else 1e-9
    assert np.abs(A.cpu().numpy()[:, 0] - g[f"{dyn}_{tag}_A"]).max() <= tol
    assert np.abs(Bm.cpu().numpy()[:, 0] - g[f"{dyn}_{tag}_B"]).max() <= tol
