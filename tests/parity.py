"""Shared parity protocol (SURVEY §7 'Hard parts' 1).

iLQR termination in the reference is decided by a strict `new_cost < best_cost`
with no tolerance (controller.py:300-301), i.e. by floating-point noise once the
solve has converged.  Parity is therefore checked in three well-posed ways:

 (a) REPLAY: the implementation under test is driven with the reference's
     recorded decisions (argmin index + improved flag per iteration); X*, U*,
     Jacobians, masks and iteration counts must then match to the stated
     tolerance / bit-exactly.
 (b) LOCK-STEP: free-running, every decision must equal the reference's
     whenever the reference's own decision margin exceeds `margin`; the
     fraction of elements whose trace departs below the margin is reported.
 (c) END-TO-END: final objective agreement for all elements.
"""
from __future__ import annotations

import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# relative tolerances stated by BASELINE.json north_star
TOL = {np.dtype(np.float32): 1e-5, np.dtype(np.float64): 1e-9}
# decision margin below which the reference's own choice is noise
MARGIN = {np.dtype(np.float32): 2e-5, np.dtype(np.float64): 1e-11}

SOLVE_CASES = [
    "di_f64", "di_f32", "cartpole_shared_f32", "cartpole_shared_f64", "cartpole_actor_f32", "cartpole_actor_f64",
    "cartpole_ad_f64", "unicycle_tight_f32", "unicycle_tight_f64", "unicycle_loose_f64", "unicycle_loose_f32",
    "unicycle_wrapped_f64", "lti8_f32", "lti8_T20_f32", "lti8_f64", "lti8_T50_f64", "lti8_mixed_f64", "lti16_f32", "lti32_f32",
    "gradcheck_f64",
]
# lti8_f32 rolls an open-loop-unstable LTI system out for T=50 steps in fp32; the reference's own fp32
# answer is 5.2 (relative) away from its fp64 answer on this case (gen_golden.py prints it), so one
# 50-step fp32 rollout is only reproducible to ~1e-4 (U, through the feedback gains).
STAGE_SLACK = {"lti8_f32": 30.0}
SMALL_CASES = [c for c in SOLVE_CASES if not c.startswith("lti")] 


# Conditioning-limited cases.  lti8_T50_*: open-loop-unstable LTI rolled out for T=50 steps; the Riccati
# recursion amplifies rounding by ~1e9 (in fp64 two correct implementations agree to ~1e-6 on K and ~3e-9
# on U; in fp32 nothing survives, see REPRO_LIMIT below).  The tolerance of such a case is scaled.
CASE_SLACK = {"lti8_T50_f64": 1e3}
# ... but X*, U* of that case are reproduced far better than its gains: achieved 1.05e-9 relative in replay, so the VALUES are
# held to 4e-9 (round 1 held them to 1e-6 with the blanket slack above, which now only covers gains / gradients / margins)
VALUE_SLACK = {"lti8_T50_f64": 4.0}


def load(name):
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))
    g["__name__"] = name
    return g


def case_tol(g, values=False):
    slack = (VALUE_SLACK if values else CASE_SLACK).get(str(g.get("__name__", "")), 1.0)
    return TOL[g["x0"].dtype] * slack


def spec_of(g):
    dyn = str(g["meta_dyn"])
    s = dict(nx=int(g["meta_nx"]), nu=int(g["meta_nu"]), T=int(g["meta_T"]), dt=float(g["meta_dt"]),
             reg_eps=float(g["meta_reg_eps"]), max_iter=int(g["meta_max_iter"]), dyn=dyn,
             jac={"analytic": "analytic", "auto_diff": "auto_diff", "finite_diff": "finite_diff"}[str(g["meta_grad_method"])])
    if dyn == "lti":
        s["A"], s["B"] = g["meta_lti_A"], g["meta_lti_B"]
    if "meta_dyn_params" in g:
        s["dyn_params"] = tuple(float(v) for v in g["meta_dyn_params"])
    if "u_min" in g:
        s["u_min"], s["u_max"] = g["u_min"], g["u_max"]
    return s


def ls_cost_of(g):
    """Quirk Q4: with per-element costs the reference scores line-search candidates with element 0's cost."""
    if g["C"].ndim == 4 or g["Cf"].ndim == 3:
        C0 = g["C"][0] if g["C"].ndim == 4 else g["C"]
        c0 = g["c"][0] if g["c"].ndim == 3 else g["c"]
        Cf0 = g["Cf"][0] if g["Cf"].ndim == 3 else g["Cf"]
        cf0 = g["cf"][0] if g["cf"].ndim == 2 else g["cf"]
        return (C0, c0, Cf0, cf0)
    return None


def rel_err(a, b):
    """max |a-b| per batch element, relative to that element's max |b| (floored at 1)."""
    a = np.asarray(a, np.float64).reshape(a.shape[0], -1)
    b = np.asarray(b, np.float64).reshape(b.shape[0], -1)
    scale = np.maximum(np.abs(b).max(axis=1), 1.0)
    return (np.abs(a - b).max(axis=1) / scale)


def elem_tol(g, what="U", k=8.0):
    """Per-element tolerance: the north-star tolerance, widened where the REFERENCE ITSELF is noisier than
    that — measured by gen_golden.py as (1) how far the reference's answer moves when x0 moves by one ulp
    with decisions held fixed (noise_*), and (2) for fp32 cases, how far the reference's fp32 answer is from
    the same decision sequence carried out in fp64 (X64/U64)."""
    tol = case_tol(g, values=True)
    floor = np.asarray(g["noise_" + what], np.float64)
    if what + "64" in g:
        floor = np.maximum(floor, rel_err(g[what], g[what + "64"]))
    return np.maximum(tol, k * floor)


# Above this measured noise floor the reference cannot reproduce ITS OWN answer under a one-ulp change of
# x0 (chaotic clamped line search: unicycle_tight_f32) or its fp32 recursion has lost all digits against
# fp64 (open-loop-unstable LTI over T=50: lti8_f32).  Such elements are excluded from value comparisons
# (finiteness of the objective is still required); every test prints how many elements it held.
REPRO_LIMIT = 1e-3


def reproducible(g):
    """Boolean mask of batch elements whose reference answer is reproducible (noise floor <= REPRO_LIMIT)."""
    floor = np.maximum(np.asarray(g["noise_U"], np.float64), np.asarray(g["noise_X"], np.float64))
    if "U64" in g:
        floor = np.maximum(floor, np.maximum(rel_err(g["U"], g["U64"]), rel_err(g["X"], g["X64"])))
    return floor <= REPRO_LIMIT


def lockstep(g, iters, alpha_trace, flags, margin):
    """Compare a free-running decision trace with the reference's.

    Returns (n_identical, n_departed_below_margin, violations) where violations lists
    (element, iteration, kind, margin) for departures ABOVE the margin."""
    B = g["x0"].shape[0]
    cand, newc, best = g["tr_cand"].astype(np.float64), g["tr_newc"].astype(np.float64), g["tr_best"].astype(np.float64)
    max_iter = int(g["meta_max_iter"])
    identical, departed, bad = 0, 0, []
    # where the reference's own trajectory is noisier than the tolerance, its cost margins are too
    floor = np.maximum(elem_tol(g, "U", k=16.0), elem_tol(g, "X", k=16.0))
    base_margin = margin
    rep = reproducible(g)
    for b in range(B):
        margin = max(base_margin * CASE_SLACK.get(str(g.get("__name__", "")), 1.0), floor[b] if floor[b] > case_tol(g) else 0.0)
        if not rep[b]:
            margin = float("inf")
        n_ref = int(g["iters"][b])
        capped_ref = bool(g["flags"][b] & 2)
        n_imp_ref = n_ref if capped_ref else n_ref - 1
        n_got = int(iters[b])
        n_imp_got = n_got if (flags[b] & 2) else n_got - 1
        ok = True
        for j in range(min(n_ref, n_got)):
            a_ref, a_got = int(g["alpha_trace"][b, j]), int(alpha_trace[b, j])
            scale = max(abs(best[b, j]), 1e-30)
            if a_ref != a_got:
                m = abs(cand[b, j, a_got] - cand[b, j, a_ref]) / scale
                if not (m <= margin):
                    bad.append((b, j, "alpha", m))
                ok = False
                break
            imp_ref, imp_got = j < n_imp_ref, j < n_imp_got
            if imp_ref != imp_got:
                m = abs(newc[b, j] - best[b, j]) / scale
                if not (m <= margin):
                    bad.append((b, j, "improved", m))
                ok = False
                break
        if ok and n_ref != n_got:
            bad.append((b, min(n_ref, n_got), "length", float("inf")))
            ok = False
        if ok:
            identical += 1
        else:
            departed += 1
    return identical, departed - len({v[0] for v in bad}), bad


# ---- achieved-error ledger ------------------------------------------------------------------------------------
# Every -m gpu parity test records what the CUDA path ACHIEVED next to the bar it was held to; tests/conftest.py
# prints the ledger in pytest's terminal summary (visible under -q, so the driver's record shows it) and writes it to
# gpurun_out/parity_achieved.json.  The bars in the tests are set from this ledger: at most 4x the achieved error.
ACHIEVED = []


def record(case, what, achieved, bar=None, note=""):
    ACHIEVED.append(dict(case=str(case), what=str(what), achieved=float(achieved), bar=None if bar is None else float(bar), note=str(note)))


def held(case, what, achieved, bar, note=""):
    """Record and assert achieved <= bar."""
    record(case, what, achieved, bar, note)
    assert achieved <= bar, f"{case}: {what} = {achieved:.3e} exceeds the bar {bar:.3e} {note}"
