"""diagnostic: LTI warp kernels vs generic kernels on identical inputs (backward = LU path only; forward with max_iter = 1)"""
import os, sys
import torch
sys.path.insert(0, ".")
from tacdmpc_b200 import _lib, ops, workloads as WL
dev = torch.device("cuda")

def setk(generic):
    if generic: os.environ["TACD_LTI_KERNEL"] = "0"
    else: os.environ.pop("TACD_LTI_KERNEL", None)
    _lib.lib().tacd_refresh_env()

for dtype in (torch.float64, torch.float32):
  for nx, nu, rho in ((8, 4, None), (32, 8, None), (8, 4, 0.9), (32, 8, 0.9)):
    w = WL.lti(nx, nu, B=256, T=50, seed=3, rho=rho)
    w["spec"]["max_iter"] = 1
    spec = WL.gpu_spec(w["spec"], dev, dtype)
    d = {k: v.to(device=dev, dtype=dtype) for k, v in w["inp"].items()}
    ls = tuple(t.to(device=dev, dtype=dtype) for t in w["ls"])
    outs = {}
    for g in (True, False):
        setk(g)
        o = ops.ilqr_forward(spec, d["x0"], d["C"], d["c"], d["Cf"], d["cf"], d["xref"], d["uref"], d["Uinit"], ls_cost=ls, want_trace=True, want_cost_trace=True)
        outs[g] = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in o.items()}
    a, b = outs[False], outs[True]
    def rel(x, y):
        sc = y.abs().flatten(1).amax(dim=1).clamp_min(1e-30)
        e = (x - y).abs().flatten(1).amax(dim=1) / sc
        return float(e.median()), float(e.max())
    print(f"{str(dtype)[6:]} nx={nx} rho={rho}: fwd(max_iter=1) iters lti {a['iters'].float().mean():.2f} gen {b['iters'].float().mean():.2f} "
          f"nonpd lti {(a['flags'] & 1).float().mean():.2f} gen {(b['flags'] & 1).float().mean():.2f}  alpha equal {(a['alpha_trace'] == b['alpha_trace']).float().mean():.3f}  "
          f"X relerr med/max {rel(a['X'], b['X'])}  cost {rel(a['cost'][:, None], b['cost'][:, None])}")
    if "cost_trace" in a and a["cost_trace"] is not None:
        ca, cb = a["cost_trace"][:, 0], b["cost_trace"][:, 0]
        print("     cost_trace[0] (5 candidates, newc, best) elem 0: lti", [f"{v:.6g}" for v in ca[0].tolist()], " gen", [f"{v:.6g}" for v in cb[0].tolist()])
        print("     rel diff of candidate costs med/max", rel(ca[:, :5], cb[:, :5]))
    # backward on the generic kernel's trajectory with both kernels
    X, U, tight = b["X"], b["U"], b["tight"]
    gX = torch.zeros_like(X); gU = torch.zeros_like(U); gU[:, 0] = 1
    gr = {}
    for g in (True, False):
        setk(g)
        bw = ops.ilqr_backward(spec, X, U, d["C"], d["xref"], d["uref"], tight, gX, gU, Cf_batched=True)
        gr[g] = {k: v.clone() for k, v in bw.items() if v is not None}
    for k in ("grad_C", "grad_c", "grad_xref", "grad_uref"):
        print(f"     bwd {k}: relerr med/max {rel(gr[False][k], gr[True][k])}")
setk(False)
