"""-m gpu: the cost-mode variants of the five-lane forward kernel (csrc/ilqr_group.cuh::group_cost_mode).  A plain shared
running cost is classified ON THE DEVICE — 0 general, 1 time-invariant, 2 time-invariant and diagonal — and the solve runs in
the variant compiled for that mode (all three are enqueued, two exit).  The variants only skip loads of values that are equal
and products with exact zeros, so every output must be BIT-IDENTICAL to the general kernel's (TACD_NO_COST_MODES=1), for every
shipped plugin and both dtypes; production uses them on multi-wave launches only, TACD_FORCE_COST_MODES=1 turns them on here."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _costs(kind, T, n, nx, rng):
    w = rng.uniform(0.5, 2.0, n) * np.array([10, 1, 100, 1, 0.1][:n])
    C = np.tile(np.diag(w), (T, 1, 1))
    c = np.tile(0.05 * rng.standard_normal(n), (T, 1))
    if kind == "dense_const":                       # mode 1: time-invariant, off-diagonal entries present
        M = 0.1 * rng.standard_normal((n, n))
        C = C + (M @ M.T)[None]
    elif kind == "time_varying":                    # mode 0
        C = C * (1 + 0.03 * np.arange(T))[:, None, None]
    elif kind == "linear_time_varying":             # mode 0: only c_t depends on t
        c = 0.05 * rng.standard_normal((T, n))
    elif kind == "negative_zero":                   # mode 2: -0.0 off-diagonals are exact zeros too
        C = C.copy()
        C[:, 0, 1] = -0.0
    Cf = np.zeros((n, n))
    Cf[:nx, :nx] = 5 * C[0, :nx, :nx]
    cf = np.zeros(n)
    return C, c, Cf, cf


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("dyn", ["cartpole", "unicycle", "unicycle_wrapped", "lti21", "lti11"])
@pytest.mark.parametrize("kind", ["diag_const", "dense_const", "time_varying", "linear_time_varying", "negative_zero"])
def test_cost_mode_variants_are_bit_identical_to_the_general_kernel(monkeypatch, dyn, kind, dtype):
    import gpu_common as G
    from tacdmpc_b200 import _lib, ops
    import zlib
    rng = np.random.default_rng(zlib.crc32(f"{dyn}/{kind}".encode()))   # (str hashes are salted per process)
    nx, nu = {"cartpole": (4, 1), "unicycle": (3, 2), "unicycle_wrapped": (3, 2), "lti21": (2, 1), "lti11": (1, 1)}[dyn]
    n, T, B = nx + nu, 12, 77
    C, c, Cf, cf = _costs(kind, T, n, nx, rng)
    spec = dict(nx=nx, nu=nu, T=T, dt=0.05, dyn="lti" if dyn.startswith("lti") else dyn, dyn_params=(1.0, 0.1, 0.5, 9.8), reg_eps=1e-6,
                max_iter=12, u_min=[-3.0] * nu, u_max=[3.0] * nu)
    if dyn == "lti21":
        spec.update(A=np.array([[1.0, 0.05], [0.0, 1.0]]), B=np.array([[0.0], [0.05]]))
    if dyn == "lti11":
        spec.update(A=np.array([[1.02]]), B=np.array([[0.05]]))
    d = dict(x0=0.4 * rng.standard_normal((B, nx)), C=C, c=c, Cf=Cf, cf=cf, xref=0.1 * rng.standard_normal((B, T + 1, nx)),
             uref=0.05 * rng.standard_normal((B, T, nu)), Uinit=0.01 * rng.standard_normal((B, T, nu)))
    d = {k: v.astype(dtype) for k, v in d.items()}
    s = G.gpu_spec(spec, dtype)
    t = {k: G.dev(v) for k, v in d.items()}

    def run(env):
        for k in ("TACD_NO_COST_MODES", "TACD_FORCE_COST_MODES"):
            monkeypatch.delenv(k, raising=False)
        monkeypatch.setenv(env, "1")
        _lib.lib().tacd_refresh_env()
        n0 = _lib.lib().tacd_kernel_launches()
        out = ops.ilqr_forward(s, t["x0"], t["C"], t["c"], t["Cf"], t["cf"], t["xref"], t["uref"], t["Uinit"], want_trace=True)
        torch.cuda.synchronize()
        return {k: v.clone() for k, v in out.items() if torch.is_tensor(v)}, _lib.lib().tacd_kernel_launches() - n0

    try:
        ref, n_ref = run("TACD_NO_COST_MODES")
        got, n_got = run("TACD_FORCE_COST_MODES")
    finally:
        for k in ("TACD_NO_COST_MODES", "TACD_FORCE_COST_MODES"):
            monkeypatch.delenv(k, raising=False)
        _lib.lib().tacd_refresh_env()
    assert n_got == n_ref + 3, "the forced run launches the classifier and the two extra variants"
    assert int(ref["iters"].max()) >= 2
    for k in ("X", "U", "iters", "flags", "alpha_trace", "tight", "cost"):
        assert torch.equal(ref[k], got[k]), k
