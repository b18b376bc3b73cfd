"""-m gpu: every BASELINE.json configuration at ITS size, forward AND backward,
against the CPU oracle on a random sample of the batch (VERDICT r1 next #1a).

Forward: the full batch is solved free-running on the GPU; the sampled elements are re-solved by the oracle (1) driven along the
kernel's own decision trace (REPLAY: X*, U* per element, active-set mask and non-PD flag bit-exact where the values agree) and
(2) free-running (END-TO-END: final objective; the fraction of identical decision traces is recorded — it is NOT 100 %, iLQR
termination compares costs that differ by rounding).  Backward: the full batch on the GPU with a random weighting
of X* and U* (and PPO's U*[:,0].sum()); the oracle's backward on the sample from the SAME saved tensors; shared-cost gradients
against the float64 sum of the oracle's per-element gradients over the FULL batch.

Bars are per quantity and per configuration, set from the achieved errors the ledger prints (tests/parity.py::record):
at most 4x what the kernels achieve, never looser than the reference's own fp32 noise allows.
"""
import numpy as np
import pytest
import torch

import parity as P

pytestmark = pytest.mark.gpu

NORTH_STAR = {np.dtype(np.float32): 1e-5, np.dtype(np.float64): 1e-9}   # relative tolerances of BASELINE.json


@pytest.fixture(scope="module", autouse=True)
def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def _workloads():
    from tacdmpc_b200 import workloads as W
    return {
        "2a_cartpole_shared_B65536": (lambda: W.cartpole(65536, "shared"), 4096),
        "2b_cartpole_per_element_B65536": (lambda: W.cartpole(65536, "per_element"), 2048),
        "2b_cartpole_diag_B65536": (lambda: W.cartpole(65536, "diag"), 2048),
        "3i_unicycle_loose_B16384": (lambda: W.unicycle(16384, tight=False), 2048),
        "3ii_unicycle_tight_B16384": (lambda: W.unicycle(16384, tight=True), 2048),
        "4_ppo_actor_diag_B4096": (lambda: W.ppo_actor(4096, diag=True), 1024),
        "4_ppo_actor_diag_B512": (lambda: W.ppo_actor(512, diag=True), 512),
        "5_lti8x4_B4096": (lambda: W.lti(8, 4), 256),
        "5_lti16x4_B4096": (lambda: W.lti(16, 4), 256),
        "5_lti16x8_B4096": (lambda: W.lti(16, 8), 256),
        "5_lti32x8_B4096": (lambda: W.lti(32, 8), 256),
        "5_lti32x8_rho0.9_B4096": (lambda: W.lti(32, 8, rho=0.9), 64),
        "5_lti8x4_rho0.9_B4096": (lambda: W.lti(8, 4, rho=0.9), 256),
        # the same shapes in fp64: configs 3 and 5 as SURVEY states them are not reproducible in fp32 by the reference algorithm
        # itself (ledger: "oracle-reproducible frac"), so the kernels are also pinned at size where the arithmetic carries digits
        "3i_unicycle_loose_B16384_f64": (lambda: W.unicycle(16384, tight=False), 1024),
        "3ii_unicycle_tight_B16384_f64": (lambda: W.unicycle(16384, tight=True), 1024),
        "5_lti8x4_B4096_f64": (lambda: W.lti(8, 4), 256),
        "5_lti16x8_B4096_f64": (lambda: W.lti(16, 8), 128),
        "5_lti32x8_B4096_f64": (lambda: W.lti(32, 8), 64),
    }


# ---- bars (at most 4x the achieved value in gpurun_out/parity_achieved.json of the round-2 build; table in DESIGN.md §5) ---------
# fwd      max relative error of X*, U* in replay over the STRICT elements (those whose own noise floor is below 1e-5 / 8)
# fwd_ok   required fraction of the oracle-reproducible elements within their tolerance max(1e-5, 8 x floor)
# cost     free-running final objective, 99th percentile of the relative difference
# grad     max relative error of the per-element gradients over the strict elements (floor below grad / 4)
# grad_ok  required fraction of the reproducible elements within max(grad, 4 x floor)
# A bar of None = recorded, not asserted: on the chaotic fp32 configurations the maximum over a sample is a rare-event statistic
# (one pnqp / clamp decision flipping moves an element by O(1e-3)); there the fraction within tolerance is what is held.
DEFAULT_BAR = dict(fwd=1e-5, fwd_ok=1.0, cost=1e-5, grad=1e-5, grad_ok=1.0)
# grad_ok = 0.6: the restated reference compiled with FMA contraction agrees with the same code compiled without it on 72 % of the
# reproducible elements of config 3(i) under this very criterion (a one-sample fp32-vs-fp64 floor underestimates the spread of fp32
# answers; measured in round 2) — the CUDA kernels reach 70-93 %.
_CHAOTIC32 = dict(fwd=None, fwd_ok=0.95, cost=None, grad=None, grad_ok=0.6)
_F64 = dict(fwd=1e-9, fwd_ok=1.0, cost=1e-9, grad=1e-9, grad_ok=1.0)
BARS = {
    # achieved (round 2, B200): fwd 2.9e-6 / 2.1e-6 / 2.1e-6, grad 1.1e-6 / 2.0e-6, cost 4.9e-7 / 5.9e-7
    "2a_cartpole_shared_B65536": dict(cost=2e-6, grad=5e-6),
    "2b_cartpole_per_element_B65536": dict(cost=2.4e-6, grad=8e-6),
    "2b_cartpole_diag_B65536": dict(cost=2.4e-6, grad=8e-6),
    # achieved: fwd 6.1e-7, grad 2.5e-6, cost 1.4e-6
    "4_ppo_actor_diag_B4096": dict(fwd=2.5e-6, cost=5.6e-6),
    "4_ppo_actor_diag_B512": dict(fwd=2.5e-6, cost=5.6e-6),
    # config 3 in fp32: 15-29 % of the sample is reproducible by the restated reference itself (median floor 7e-3 / 2e-2)
    "3i_unicycle_loose_B16384": _CHAOTIC32,
    "3ii_unicycle_tight_B16384": _CHAOTIC32,
    # config 5 as SURVEY states it (open-loop unstable over T = 50): the fp32 gradients of the restated reference are O(1) away
    # from its fp64 ones (0 % reproducible) and free-running decisions are noise; forward values in replay are held
    "5_lti8x4_B4096": _CHAOTIC32, "5_lti16x4_B4096": _CHAOTIC32, "5_lti16x8_B4096": _CHAOTIC32, "5_lti32x8_B4096": _CHAOTIC32,
    # the well-conditioned twins (spectral radius 0.9): achieved fwd 2.6e-7, grad 8.8e-7, cost 1.4e-5 / 4.2e-6 (40 iterations)
    "5_lti8x4_rho0.9_B4096": dict(fwd=1e-6, grad=3.5e-6, cost=5e-5),
    "5_lti32x8_rho0.9_B4096": dict(cost=1.6e-5),
    # fp64 at the same sizes.  achieved: unicycle fwd 3.5e-10 / 2.0e-9, grad 7.8e-10; the unstable LTI loses 7+ digits even in fp64
    "3i_unicycle_loose_B16384_f64": dict(_F64, fwd=1.9e-9, fwd_ok=0.99, grad=3e-9),
    "3ii_unicycle_tight_B16384_f64": dict(_F64, fwd=8e-9, grad=2.2e-9),
    "5_lti8x4_B4096_f64": dict(_F64, fwd=None, fwd_ok=0.95, grad=None, grad_ok=0.0, cost=None),
    "5_lti16x8_B4096_f64": dict(_F64, fwd=None, fwd_ok=0.95, grad=None, grad_ok=0.0, cost=None),
    "5_lti32x8_B4096_f64": dict(_F64, fwd=None, fwd_ok=0.95, grad=None, grad_ok=0.0, cost=None),
}


def _bar(case, k):
    b = BARS.get(case, {})
    return b[k] if k in b else DEFAULT_BAR[k]


def _dense(w, inp):
    """The oracle restates the reference, which only knows dense costs."""
    if not w["C_diag"]:
        return inp["C"], inp["Cf"], w["ls"]
    ls = w["ls"]
    return (torch.diag_embed(inp["C"]), torch.diag_embed(inp["Cf"]),
            (torch.diag_embed(ls[0]), ls[1], torch.diag_embed(ls[2]), ls[3]))


def _np(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module", params=list(_workloads()))
def solved(request, oracle):
    from tacdmpc_b200 import ops
    from tacdmpc_b200 import workloads as W
    case = request.param
    make, ns = _workloads()[case]
    w = make()
    dt = torch.float64 if case.endswith("_f64") else torch.float32
    w["inp"] = {k: v.to(dt) for k, v in w["inp"].items()}
    w["ls"] = None if w["ls"] is None else tuple(t.to(dt) for t in w["ls"])
    spec, inp = w["spec"], w["inp"]
    s = W.gpu_spec(spec, "cuda", dt)
    d = {k: v.cuda() for k, v in inp.items()}
    ls = None if w["ls"] is None else tuple(t.cuda() for t in w["ls"])
    fw = ops.ilqr_forward(s, d["x0"], d["C"], d["c"], d["Cf"], d["cf"], d["xref"], d["uref"], d["Uinit"], ls_cost=ls,
                          want_trace=True, C_diag=w["C_diag"])
    torch.cuda.synchronize()
    B = inp["x0"].shape[0]
    idx = torch.randperm(B, generator=torch.Generator().manual_seed(11))[:ns].sort().values
    return dict(case=case, w=w, s=s, d=d, fw=fw, idx=idx, B=B)


def _sample_inputs(S):
    w, idx, B = S["w"], S["idx"], S["B"]
    inp = w["inp"]
    Cd, Cfd, lsd = _dense(w, inp)
    h = dict(inp, C=Cd, Cf=Cfd)
    batched_ndim = dict(x0=2, C=4, c=3, Cf=3, cf=2, xref=3, uref=3, Uinit=3)
    sub = {k: _np(v[idx] if (v.ndim == batched_ndim[k] and v.shape[0] == B) else v) for k, v in h.items()}
    return sub, None if lsd is None else tuple(_np(t) for t in lsd)


def _oracle_fwd(oracle, spec, sub, ls, rp, dtype=None, x0=None, C=None):
    c = (lambda a: a.astype(dtype)) if dtype is not None else (lambda a: a)
    return oracle.forward(spec, c(sub["x0"] if x0 is None else x0), c(sub["C"] if C is None else C), c(sub["c"]), c(sub["Cf"]), c(sub["cf"]), c(sub["xref"]),
                          c(sub["uref"]), c(sub["Uinit"]), ls_cost=None if ls is None else tuple(c(a) for a in ls), replay=rp,
                          want_AB=False, want_trace=False)


def test_forward_sample_against_oracle(solved, oracle):
    S = solved
    case, w, fw, idx = S["case"], S["w"], S["fw"], S["idx"]
    spec = w["spec"]
    sub, ls = _sample_inputs(S)
    gi = idx.cuda()
    rp = dict(iters=_np(fw["iters"][gi]), alpha_trace=_np(fw["alpha_trace"][gi]), flags=_np(fw["flags"][gi]))
    ref = _oracle_fwd(oracle, spec, sub, ls, rp)
    X, U = _np(fw["X"][gi]), _np(fw["U"][gi])
    assert np.isfinite(X).all() and np.isfinite(U).all()
    e = np.maximum(P.rel_err(X, ref["X"]), P.rel_err(U, ref["U"]))
    # The noise floor of each sampled element, measured the way oracle/gen_golden.py measures the reference's (tests/parity.py::
    # elem_tol): (1) how far the ORACLE's own answer moves when x0 and C move by one ulp with the decisions held fixed, (2) how far its
    # fp32 answer is from the same decision sequence carried out in fp64.  Where the floor exceeds the north-star tolerance (the
    # clamped line search of the unicycle is chaotic in fp32: median floor 3e-2 on config 3(ii); the open-loop-unstable LTI of
    # config 5 amplifies rounding over T = 50) the element's bar is 8x its floor; above REPRO_LIMIT the restated reference cannot
    # reproduce its own answer and only finiteness is required.
    noise = np.zeros(len(e))
    rng = np.random.default_rng(99)
    for _ in range(3):
        sgn = np.where(rng.random(sub["x0"].shape) < 0.5, -1.0, 1.0).astype(sub["x0"].dtype)
        # (the running cost moves by one ulp too: for the LTI plugin the map x0 -> solution is affine and well conditioned,
        # the gains are what the unstable Riccati recursion of config 5 loses)
        sgc = np.where(rng.random(sub["C"].shape) < 0.5, -1.0, 1.0).astype(sub["C"].dtype)
        pert = _oracle_fwd(oracle, spec, sub, ls, rp, x0=np.nextafter(sub["x0"], sub["x0"] + sgn), C=np.nextafter(sub["C"], sub["C"] + sgc))
        noise = np.maximum(noise, np.maximum(P.rel_err(pert["X"], ref["X"]), P.rel_err(pert["U"], ref["U"])))
    ns_tol = NORTH_STAR[sub["x0"].dtype]
    if sub["x0"].dtype == np.float32:
        r64 = _oracle_fwd(oracle, spec, sub, ls, rp, dtype=np.float64)
        with np.errstate(invalid="ignore"):
            f64 = np.maximum(P.rel_err(ref["X"], r64["X"]), P.rel_err(ref["U"], r64["U"]))
        noise = np.where(np.isfinite(f64), np.maximum(noise, f64), np.inf)
    tol = np.maximum(ns_tol, 8.0 * noise)
    rep = noise <= P.REPRO_LIMIT
    strict = tol <= ns_tol                # elements the oracle itself reproduces to the north-star tolerance
    it = rp["iters"]
    P.record(case, "mean iters (sample)", it.mean())
    P.record(case, "nonPD fraction", float((rp["flags"] & 1).mean()))
    P.record(case, "oracle noise floor, median", float(np.median(noise)))
    P.record(case, "oracle-reproducible frac (floor <= 1e-3)", float(rep.mean()))
    P.record(case, "strict frac (floor <= tol/8)", float(strict.mean()))
    P.record(case, "fwd frac of sample within north-star tol", float((e <= ns_tol).mean()))
    checks = []
    if strict.any():
        P.record(case, "fwd replay p99 relerr, strict elems", float(np.quantile(e[strict], 0.99)))
        b = _bar(case, "fwd")
        P.record(case, "fwd replay max relerr, strict elems", float(e[strict].max()), b)
        checks.append(b is None or e[strict].max() <= b)
    if rep.any():
        okf = float((e <= tol)[rep].mean())
        P.record(case, "fwd frac of reproducible within tol", okf, _bar(case, "fwd_ok"))
        checks.append(okf >= _bar(case, "fwd_ok"))
    ok = (e <= ns_tol) & rep
    # bit-exact integer outputs where the values agree: active-set mask and flags (iteration counts are the replayed ones)
    tg = _np(fw["tight"][gi]).astype(bool)
    tight_eq = float((tg[ok] == ref["tight"][ok].astype(bool)).all(axis=(1, 2)).mean()) if ok.any() else 1.0
    P.record(case, "tight mask equal (elems <= 1e-5)", tight_eq, 1.0)
    assert tight_eq == 1.0
    fl_eq = float(((rp["flags"] & 1)[ok] == (ref["flags"] & 1)[ok]).mean()) if ok.any() else 1.0
    P.record(case, "nonPD flag equal (elems <= 1e-5)", fl_eq)
    assert fl_eq >= 0.99
    # free-running oracle: its own decisions
    free = _oracle_fwd(oracle, spec, sub, ls, None)
    same = (free["iters"] == it) & (free["alpha_trace"][:, :rp["alpha_trace"].shape[1]] == rp["alpha_trace"]).all(axis=1)
    P.record(case, "free-run identical trace frac", float(same.mean()), note="(recorded, not a bar: DESIGN.md 5)")
    P.record(case, "free-run |mean iters diff|", abs(float(free["iters"].mean()) - float(it.mean())))
    J = _np(fw["cost"][gi]).astype(np.float64)
    fin = np.isfinite(free["cost"]) & np.isfinite(J) & rep
    if fin.any():
        ce = np.abs(J - free["cost"])[fin] / np.maximum(1.0, np.abs(free["cost"][fin]))
        P.record(case, "free-run cost max relerr", float(ce.max()))
        b = _bar(case, "cost")
        P.record(case, "free-run cost p99 relerr", float(np.quantile(ce, 0.99)), b)
        checks.append(b is None or np.quantile(ce, 0.99) <= b)
    assert all(checks), (case, checks)


@pytest.mark.parametrize("loss", ["weighted", "U0sum"])
def test_backward_sample_against_oracle(solved, oracle, loss):
    from tacdmpc_b200 import ops
    S = solved
    case, w, fw, idx, d, s, B = S["case"], S["w"], S["fw"], S["idx"], S["d"], S["s"], S["B"]
    spec = w["spec"]
    T, nx, nu = spec["T"], spec["nx"], spec["nu"]
    g = torch.Generator().manual_seed(5)
    if loss == "weighted":
        gX, gU = torch.randn(B, T + 1, nx, generator=g), torch.randn(B, T, nu, generator=g)
    else:
        gX, gU = torch.zeros(B, T + 1, nx), torch.zeros(B, T, nu)
        gU[:, 0] = 1
    gX, gU = gX.to(d["x0"].dtype), gU.to(d["x0"].dtype)
    want = ("grad_C", "grad_c", "grad_Cf", "grad_cf", "grad_x0", "grad_xref", "grad_uref", "dX", "dU")
    out = ops.ilqr_backward(s, fw["X"], fw["U"], d["C"], d["xref"], d["uref"], fw["tight"], gX.cuda(), gU.cuda(),
                            Cf_batched=w["Cf_batched"], want=want, C_diag=w["C_diag"])
    torch.cuda.synchronize()
    sub, _ = _sample_inputs(S)
    gi = idx.cuda()
    Xs, Us, tg = _np(fw["X"][gi]), _np(fw["U"][gi]), _np(fw["tight"][gi])
    args = (Xs, Us, sub["C"], sub["xref"], sub["uref"])
    ref = oracle.backward(spec, *args, tg, _np(gX[idx]), _np(gU[idx]), Cf_batched=int(w["Cf_batched"]), want=want)
    # the same backward carried out in fp64 on the same saved tensors: how far the restated reference's own fp32 gradients are from
    # exact arithmetic (the floor tests/test_gpu_parity.py takes from the golden grad64_* vectors)
    r64 = oracle.backward(spec, *(a.astype(np.float64) for a in args), tg, _np(gX[idx]).astype(np.float64), _np(gU[idx]).astype(np.float64),
                          Cf_batched=int(w["Cf_batched"]), want=want)
    per_elem = ["grad_x0", "grad_xref", "grad_uref", "dX", "dU"]
    batched_C = d["C"].ndim == (3 if w["C_diag"] else 4)
    if batched_C:
        per_elem += ["grad_C", "grad_c"]
    if w["Cf_batched"]:
        per_elem += ["grad_Cf", "grad_cf"]
    ns = len(idx)
    err, floor = np.zeros(ns), np.zeros(ns)
    for k in per_elem:
        got = _np(out[k][gi]).astype(np.float64).reshape(ns, -1)
        r, r6 = ref[k].astype(np.float64), r64[k]
        if w["C_diag"] and k in ("grad_C", "grad_Cf"):
            r, r6 = np.diagonal(r, axis1=-2, axis2=-1), np.diagonal(r6, axis1=-2, axis2=-1)
        r, r6 = r.reshape(ns, -1), r6.reshape(ns, -1)
        with np.errstate(invalid="ignore", over="ignore"):
            scale = np.maximum(1.0, np.abs(r6).max(axis=1))
            ek = np.abs(got - r).max(axis=1) / scale
            fk = np.abs(r - r6).max(axis=1) / scale
        err = np.maximum(err, np.where(np.isfinite(ek), ek, np.inf))
        floor = np.maximum(floor, np.where(np.isfinite(fk), fk, np.inf))
    f64case = Xs.dtype == np.float64
    bar = _bar(case, "grad")
    hold_max = bar is not None
    bar = bar if hold_max else NORTH_STAR[Xs.dtype]
    if f64case:
        floor[:] = 0.0          # (the fp64 run IS the exact-arithmetic stand-in; its bar is 1e-9)
    rep = floor <= P.REPRO_LIMIT
    strict = 4.0 * floor <= bar
    tol = np.maximum(bar, 4.0 * floor)
    P.record(case, f"bwd {loss} oracle fp32-vs-fp64 floor, median", float(np.median(floor)))
    P.record(case, f"bwd {loss} reproducible frac (floor <= 1e-3)", float(rep.mean()))
    P.record(case, f"bwd {loss} strict frac", float(strict.mean()))
    checks = []
    if strict.any():
        P.record(case, f"bwd {loss} p99 relerr, strict elems", float(np.quantile(err[strict], 0.99)))
        P.record(case, f"bwd {loss} max relerr, strict elems", float(err[strict].max()), bar if hold_max else None)
        checks.append(not hold_max or err[strict].max() <= bar)
    if rep.any():
        okf = float((err <= tol)[rep].mean())
        P.record(case, f"bwd {loss} frac of reproducible within tol", okf, _bar(case, "grad_ok"))
        checks.append(okf >= _bar(case, "grad_ok"))
    assert all(checks), (case, loss, checks)
    assert float(out["grad_x0"].abs().max()) == 0.0   # quirk Q3
    # shared-cost gradients: float64 sum over the FULL batch of the oracle's per-element gradients
    if not batched_C and B * T * (nx + nu) ** 2 * 4 <= 256 << 20:
        n = nx + nu
        Ce = np.broadcast_to(_np(d["C"])[None], (B, T, n, n)).copy()
        fa = (_np(fw["X"]), _np(fw["U"]), Ce, _np(d["xref"]), _np(d["uref"]))
        full = oracle.backward(spec, *fa, _np(fw["tight"]), _np(gX), _np(gU), Cf_batched=0, want=("grad_C", "grad_c"))
        f64 = oracle.backward(spec, *(a.astype(np.float64) for a in fa), _np(fw["tight"]), _np(gX).astype(np.float64),
                              _np(gU).astype(np.float64), Cf_batched=0, want=("grad_C", "grad_c"))
        for k in ("grad_C", "grad_c"):
            tot, t64 = full[k].astype(np.float64).sum(0), f64[k].sum(0)
            scale = max(1.0, np.abs(t64).max())
            fl = np.abs(tot - t64).max() / scale
            err_s = np.abs(_np(out[k]).astype(np.float64) - tot).max() / scale
            P.record(case, f"bwd {loss} shared {k}: oracle f32-vs-f64 batch-sum floor", fl)
            P.held(case, f"bwd {loss} shared {k} vs oracle batch sum", err_s, max(bar, 4 * fl))
