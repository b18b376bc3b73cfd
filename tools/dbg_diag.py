import sys, torch
sys.path.insert(0, '.')
import bench
from tacdmpc_b200 import ops
from tacdmpc_b200.dynamics import CartPole
f = CartPole()
T, N, NX, NU = bench.T, bench.N, bench.NX, bench.NU
for B in (64, 1024, 20000, 65536):
    for tv in (False, True):
        h = bench.make_inputs(B, "per_element", 5, torch)
        if tv:
            h["C"] = h["C"] * (1 + 0.1 * torch.arange(T).view(1, T, 1, 1))
        d = {k: v.cuda() for k, v in h.items()}
        spec = ops.Spec(nx=NX, nu=NU, T=T, dt=bench.DT, dyn_id=f.dyn_id, dyn_params=f.params, u_min=[-20.0], u_max=[20.0])
        ls = (d["C"][0], d["c"][0], d["Cf"][0], d["cf"][0])
        a = ops.ilqr_forward(spec, d["x0"], d["C"], d["c"], d["Cf"], d["cf"], d["xref"], d["uref"], d["Uinit"], ls_cost=ls)
        dg = lambda t: torch.diagonal(t, dim1=-2, dim2=-1).contiguous()
        lsd = (dg(ls[0]), ls[1], dg(ls[2]), ls[3])
        b = ops.ilqr_forward(spec, d["x0"], dg(d["C"]), d["c"], dg(d["Cf"]), d["cf"], d["xref"], d["uref"], d["Uinit"], ls_cost=lsd, C_diag=True)
        torch.cuda.synchronize()
        print(B, "time-varying" if tv else "const", "iters", float(a["iters"].float().mean()), float(b["iters"].float().mean()),
              "equal U:", bool(torch.equal(a["U"], b["U"])), "cost diff", float((a["cost"] - b["cost"]).abs().max()))
