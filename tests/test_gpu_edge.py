"""-m gpu edge cases of the whole-solve kernels (both forward kernels, both dtypes): ragged horizons (T not a
multiple of the five-step Jacobian blocks, T = 1), batches that do not fill a lane group or a warp, fewer
line-search candidates than lanes, max_iter 0 / 1, time-varying and per-element costs, non-finite inputs.  Each
solve is checked against the CPU oracle driven along the kernel's own decision trace (so the check is independent
of rounding-level decision flips), and the backward against the oracle's backward on the kernel's outputs."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _refresh_knobs(request):
    """The library caches its TACD_* knobs; re-read them now and again when monkeypatch has restored the environment."""
    from tacdmpc_b200 import _lib
    _lib.lib().tacd_refresh_env()
    request.addfinalizer(lambda: _lib.lib().tacd_refresh_env())


@pytest.fixture(scope="module")
def G():
    import gpu_common
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return gpu_common


def _problem(dyn, T, B, dtype, seed, per_element=False, n_alpha=5, max_iter=40, bounds=True):
    rng = np.random.default_rng(seed)
    nx, nu = {"cartpole": (4, 1), "unicycle": (3, 2), "unicycle_wrapped": (3, 2), "lti": (2, 1)}[dyn]
    n = nx + nu
    w = rng.uniform(0.5, 2.0, n) * np.array(([10, 1, 100, 1, 0.1] if dyn == "cartpole" else [20, 20, 5, 0.1, 0.1] if nx == 3 else [100, 1, 0.01])[:n])
    C = np.tile(np.diag(w), (T, 1, 1)) * (1 + 0.03 * np.arange(T))[:, None, None]
    M = 0.02 * rng.standard_normal((n, n))
    C = C + (M @ M.T)[None]
    c = 0.05 * rng.standard_normal((T, n))
    Cf = np.zeros((n, n))
    Cf[:nx, :nx] = 5 * C[0, :nx, :nx]
    cf = np.zeros(n)
    cf[:nx] = 0.05 * rng.standard_normal(nx)
    if per_element:
        s = rng.uniform(0.5, 1.5, (B, 1, 1, 1))
        C, c = C[None] * s, np.tile(c, (B, 1, 1))
        Cf, cf = Cf[None] * s[:, 0], np.tile(cf, (B, 1))
    x0 = 0.4 * rng.standard_normal((B, nx))
    if dyn == "cartpole":
        x0[:, 2] += 0.2
    xref = 0.1 * rng.standard_normal((B, T + 1, nx))
    uref = 0.05 * rng.standard_normal((B, T, nu))
    spec = dict(nx=nx, nu=nu, T=T, dt=0.05 if dyn != "unicycle" else 0.1, dyn=dyn, dyn_params=(1.0, 0.1, 0.5, 9.8), reg_eps=1e-6,
                max_iter=max_iter, alphas=(1.0, 0.8, 0.5, 0.2, 0.1)[:n_alpha])
    if dyn == "lti":
        spec.update(A=np.array([[1.0, 0.05], [0.0, 1.0]]), B=np.array([[0.0], [0.05]]))
    if bounds:
        b = [3.0] if nu == 1 else [1.5, 1.0]
        spec.update(u_min=[-v for v in b], u_max=b)
    d = dict(x0=x0, C=C, c=c, Cf=Cf, cf=cf, xref=xref, uref=uref, Uinit=0.01 * rng.standard_normal((B, T, nu)))
    return spec, {k: v.astype(dtype) for k, v in d.items()}


CASES = [
    # dyn, T, B, dtype, per_element, n_alpha, max_iter
    ("cartpole", 1, 7, np.float64, False, 5, 40),
    ("cartpole", 2, 1, np.float32, False, 5, 40),
    ("cartpole", 7, 33, np.float32, False, 5, 40),
    ("cartpole", 13, 200, np.float64, True, 5, 40),
    ("cartpole", 20, 5, np.float32, True, 3, 40),
    ("cartpole", 20, 64, np.float64, False, 1, 40),
    ("cartpole", 20, 64, np.float32, False, 5, 1),
    ("cartpole", 20, 9, np.float64, False, 5, 0),
    ("unicycle", 11, 50, np.float64, True, 5, 40),
    ("unicycle", 30, 37, np.float32, False, 4, 40),
    ("unicycle_wrapped", 6, 20, np.float64, False, 5, 40),
    ("lti", 9, 45, np.float64, True, 5, 40),
    ("lti", 20, 3, np.float32, False, 2, 40),
]


@pytest.mark.parametrize("kernel", ["group", "thread"])
@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}-T{c[1]}-B{c[2]}-{np.dtype(c[3]).name}-pe{int(c[4])}-a{c[5]}-it{c[6]}")
def test_solve_and_backward_against_oracle(G, oracle, monkeypatch, request, case, kernel):
    from tacdmpc_b200 import ops
    dyn, T, B, dtype, per_element, n_alpha, max_iter = case
    if kernel == "thread":
        monkeypatch.setenv("TACD_FWD_KERNEL", "thread")
    else:
        monkeypatch.delenv("TACD_FWD_KERNEL", raising=False)
    _refresh_knobs(request)
    spec, d = _problem(dyn, T, B, dtype, seed=T * 131 + B, per_element=per_element, n_alpha=n_alpha, max_iter=max_iter)
    s = G.gpu_spec(spec, dtype)
    s.alphas = spec["alphas"]
    dv = G.dev
    ls = (d["C"][0], d["c"][0], d["Cf"][0], d["cf"][0]) if per_element else None
    out = ops.ilqr_forward(s, dv(d["x0"]), dv(d["C"]), dv(d["c"]), dv(d["Cf"]), dv(d["cf"]), dv(d["xref"]), dv(d["uref"]), dv(d["Uinit"]),
                           ls_cost=None if ls is None else tuple(dv(a) for a in ls), want_AB=True)
    torch.cuda.synchronize()
    o = {k: v.cpu().numpy() for k, v in out.items()}
    assert (o["iters"] >= min(1, max_iter)).all() and (o["iters"] <= max(max_iter, 0)).all()
    rp = dict(iters=o["iters"], alpha_trace=o["alpha_trace"], flags=o["flags"])
    ref = oracle.forward(spec, d["x0"], d["C"], d["c"], d["Cf"], d["cf"], d["xref"], d["uref"], d["Uinit"], ls_cost=ls, replay=rp)
    # fp32: the oracle is not FMA-contracted and the solves run up to 40 iterations over up to 30 clamped steps, so
    # this shape-sweeping test uses 1e-3 (the golden-vector tests hold the well-conditioned cases to 1e-5)
    tol = 1e-3 if dtype == np.float32 else 1e-9
    for k in ("X", "U", "A", "Bm"):
        scale = max(1.0, np.abs(ref[k]).max())
        assert np.abs(o[k].astype(np.float64) - ref[k]).max() / scale <= tol, (k, np.abs(o[k] - ref[k]).max())
    assert np.abs(o["cost"].astype(np.float64) - ref["cost"]).max() <= 10 * tol * max(1.0, np.abs(ref["cost"]).max())
    # backward on the kernel's own outputs (dense per-element / shared layouts, both through the C ABI)
    rng = np.random.default_rng(1)
    gX, gU = rng.standard_normal(o["X"].shape).astype(dtype), rng.standard_normal(o["U"].shape).astype(dtype)
    want = ("grad_C", "grad_c", "grad_Cf", "grad_cf", "grad_x0", "grad_xref", "grad_uref")
    bw = ops.ilqr_backward(s, out["X"], out["U"], dv(d["C"]), dv(d["xref"]), dv(d["uref"]), out["tight"], dv(gX), dv(gU),
                           Cf_batched=per_element, want=want)
    rb = oracle.backward(spec, o["X"], o["U"], d["C"], d["xref"], d["uref"], o["tight"], gX, gU, Cf_batched=int(per_element))
    for nm in want:
        got, exp = bw[nm].cpu().numpy().astype(np.float64), rb[nm]
        scale = max(1.0, np.abs(exp).max())
        assert np.abs(got - exp.reshape(got.shape)).max() / scale <= 50 * tol, (nm, np.abs(got - exp.reshape(got.shape)).max() / scale)


def test_non_finite_inputs_terminate(G):
    """NaN / inf states must not hang the work queue or poison other elements."""
    from tacdmpc_b200 import ops
    spec, d = _problem("cartpole", 20, 64, np.float32, seed=3)
    d["x0"][5, 2] = np.nan
    d["x0"][9, 0] = np.inf
    s = G.gpu_spec(spec, np.float32)
    dv = G.dev
    out = ops.ilqr_forward(s, *(dv(d[k]) for k in ("x0", "C", "c", "Cf", "cf", "xref", "uref", "Uinit")))
    torch.cuda.synchronize()
    it = out["iters"].cpu().numpy()
    U = out["U"].cpu().numpy()
    good = np.ones(64, bool)
    good[[5, 9]] = False
    assert np.isfinite(U[good]).all() and (it[good] >= 2).all()
    assert (it[~good] <= 40).all() and not np.isfinite(out["cost"].cpu().numpy()[~good]).any()


# ---- the run-time-dimension kernels (csrc/ilqr_generic.cu): warp-per-element (nx <= 16) and CTA-per-element modes,
# compile-time-size instances ((8,4), (16,8), (32,8)) and the run-time fallback (5,2), (20,3), ragged batches that do
# not fill a four-element CTA, with and without the input box (box QP in the KKT pass)
GEN_CASES = [
    # nx, nu, T, B, dtype, per_element, bounds
    (5, 2, 7, 1, np.float64, False, True),
    (5, 2, 12, 7, np.float32, True, False),
    (8, 4, 10, 5, np.float64, True, True),
    (8, 4, 25, 130, np.float32, False, False),
    (16, 8, 6, 9, np.float64, False, True),
    (16, 4, 8, 3, np.float32, True, False),
    (20, 3, 5, 6, np.float64, True, True),
    (32, 8, 4, 5, np.float64, False, False),
    (32, 8, 6, 3, np.float32, True, True),
]


def _lti_problem(nx, nu, T, B, dtype, seed, per_element, bounds):
    rng = np.random.default_rng(seed)
    n = nx + nu
    A = np.eye(nx) * 0.9 + 0.3 * rng.standard_normal((nx, nx)) / np.sqrt(nx)
    A *= 0.95 / np.abs(np.linalg.eigvals(A)).max()   # spectral radius 0.95: an unstable A over T = 25 is ill-conditioned in
    #                                                  fp32 to the point where the fp32 and fp64 oracles part by 3-90 % (measured)
    Bm = 0.3 * rng.standard_normal((nx, nu))
    w = rng.uniform(0.5, 1.5, n)
    w[nx:] *= 0.2
    M = 0.05 * rng.standard_normal((n, n))
    C = np.tile(np.diag(w) + M @ M.T, (T, 1, 1)) * (1 + 0.02 * np.arange(T))[:, None, None]
    c = 0.1 * rng.standard_normal((T, n))
    Cf = np.zeros((n, n))
    Cf[:nx, :nx] = 3 * C[0, :nx, :nx]
    cf = np.zeros(n)
    cf[:nx] = 0.1 * rng.standard_normal(nx)
    if per_element:
        s = rng.uniform(0.5, 1.5, (B, 1, 1, 1))
        C, c = C[None] * s, np.tile(c, (B, 1, 1))
        Cf, cf = Cf[None] * s[:, 0], np.tile(cf, (B, 1))
    spec = dict(nx=nx, nu=nu, T=T, dt=0.05, dyn="lti", A=A, B=Bm, reg_eps=1e-6, max_iter=40)
    if bounds:
        spec.update(u_min=[-0.4] * nu, u_max=[0.4] * nu)
    d = dict(x0=rng.standard_normal((B, nx)), C=C, c=c, Cf=Cf, cf=cf, xref=0.1 * rng.standard_normal((B, T + 1, nx)),
             uref=0.05 * rng.standard_normal((B, T, nu)), Uinit=0.01 * rng.standard_normal((B, T, nu)))
    return spec, {k: v.astype(dtype) for k, v in d.items()}


@pytest.mark.parametrize("case", GEN_CASES, ids=lambda c: f"nx{c[0]}-nu{c[1]}-T{c[2]}-B{c[3]}-{np.dtype(c[4]).name}-pe{int(c[5])}-box{int(c[6])}")
def test_generic_kernels_against_oracle(G, oracle, case):
    from tacdmpc_b200 import ops
    nx, nu, T, B, dtype, per_element, bounds = case
    spec, d = _lti_problem(nx, nu, T, B, dtype, seed=nx * 1000 + T * 10 + B, per_element=per_element, bounds=bounds)
    s = G.gpu_spec(spec, dtype)
    dv = G.dev
    ls = (d["C"][0], d["c"][0], d["Cf"][0], d["cf"][0]) if per_element else None
    out = ops.ilqr_forward(s, dv(d["x0"]), dv(d["C"]), dv(d["c"]), dv(d["Cf"]), dv(d["cf"]), dv(d["xref"]), dv(d["uref"]), dv(d["Uinit"]),
                           ls_cost=None if ls is None else tuple(dv(a) for a in ls))
    torch.cuda.synchronize()
    o = {k: v.cpu().numpy() for k, v in out.items()}
    rp = dict(iters=o["iters"], alpha_trace=o["alpha_trace"], flags=o["flags"])
    ref = oracle.forward(spec, d["x0"], d["C"], d["c"], d["Cf"], d["cf"], d["xref"], d["uref"], d["Uinit"], ls_cost=ls, replay=rp)
    tol = 2e-4 if dtype == np.float32 else 1e-9
    for k in ("X", "U"):
        scale = max(1.0, np.abs(ref[k]).max())
        assert np.abs(o[k].astype(np.float64) - ref[k]).max() / scale <= tol, (k, np.abs(o[k] - ref[k]).max())
    assert np.array_equal(o["tight"].astype(bool), ref["tight"].astype(bool))
    if bounds:
        assert o["tight"].any()                       # the box is exercised
    rng = np.random.default_rng(2)
    gX, gU = rng.standard_normal(o["X"].shape).astype(dtype), rng.standard_normal(o["U"].shape).astype(dtype)
    want = ("grad_C", "grad_c", "grad_Cf", "grad_cf", "grad_x0", "grad_xref", "grad_uref")
    bw = ops.ilqr_backward(s, out["X"], out["U"], dv(d["C"]), dv(d["xref"]), dv(d["uref"]), out["tight"], dv(gX), dv(gU),
                           Cf_batched=per_element, want=want)
    rb = oracle.backward(spec, o["X"], o["U"], d["C"], d["xref"], d["uref"], o["tight"], gX, gU, Cf_batched=int(per_element))
    for nm in want:
        got, exp = bw[nm].cpu().numpy().astype(np.float64), rb[nm]
        scale = max(1.0, np.abs(exp).max())
        assert np.abs(got - exp.reshape(got.shape)).max() / scale <= 50 * tol, (nm, np.abs(got - exp.reshape(got.shape)).max() / scale)
