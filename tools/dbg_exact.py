import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import parity as P, gpu_common as G
from exact_ref import exact_backward
from tacdmpc_b200 import ops
for name in sys.argv[1:]:
    g = P.load(name); spec = P.spec_of(g); s = G.gpu_spec(spec, g["x0"].dtype); d = G.dev
    B, nx, nu = g["x0"].shape[0], spec["nx"], spec["nu"]
    X, U, gX, gU = g["X"], g["U"], g["wX"], g["wU"]
    want = ("grad_C", "grad_c", "grad_Cf", "grad_cf", "grad_x0", "grad_xref", "grad_uref", "dX", "dU")
    out = ops.ilqr_backward(s, d(X), d(U), d(g["C"]), d(g["xref"]), d(g["uref"]), d(g["tight"]), d(gX), d(gU), Cf_batched=g["Cf"].ndim == 3, want=want, exact=True, Cf=d(g["Cf"]))
    Cb = g["C"] if g["C"].ndim == 4 else np.broadcast_to(g["C"], (B,) + g["C"].shape)
    Cfb = g["Cf"] if g["Cf"].ndim == 3 else np.broadcast_to(g["Cf"], (B,) + g["Cf"].shape)
    xr = np.broadcast_to(g["xref"], (B,) + g["xref"].shape[1:]); ur = np.broadcast_to(g["uref"], (B,) + g["uref"].shape[1:])
    ref = exact_backward(nx, nu, X, U, Cb, g["c"], Cfb, g["cf"], xr, ur, g["tight"], gX, gU, g["A"], g["Bm"], spec.get("reg_eps", 1e-6))
    print(name, "reg", spec.get("reg_eps"), "tight", g["tight"].mean(), "B", B)
    for nm in ("dU", "dX", "grad_x0", "grad_c"):
        got = out[nm].cpu().numpy().astype(np.float64); exp = ref[nm]
        if got.shape != exp.shape: exp = exp.sum(0).reshape(got.shape)
        e = np.abs(got - exp)
        print("  ", nm, "max abs err", e.max(), "scale", np.abs(exp).max(), "argmax", np.unravel_index(e.argmax(), e.shape))
    got = out["dU"].cpu().numpy(); e = np.abs(got - ref["dU"])[..., 0]
    print("   dU err per (b,t):\n", np.array2string(e[:2], precision=2))
    print("   tight[:2]\n", g["tight"][:2, :, 0])
