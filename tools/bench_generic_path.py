"""The GENERIC path: a Python callable as f_dyn — rollouts, Jacobians (analytic callable / vmap(jacrev))
and the line-search rollouts stay in torch with one host synchronisation per iLQR iteration (controller.py::_solve_generic); the
Riccati recursion, every cost evaluation and the backward are CUDA (tacd_riccati, tacd_objective, tacd_ilqr_backward with
TACD_DYN_EXTERNAL).  Times one forward + backward of the cart-pole regulation problem of bench.py, next to the fused path on the
same inputs (developer tool).

    python tools/bench_generic_path.py
"""
import sys
import time

import torch

sys.path.insert(0, ".")
from tacdmpc_b200.DifferentialMPC import DifferentiableMPCController, GeneralQuadCost  # noqa: E402
from tacdmpc_b200.dynamics import CartPole  # noqa: E402
from tacdmpc_b200 import workloads as WL  # noqa: E402

dev = torch.device("cuda")
native = CartPole()


def py_dyn(x, u, dt):   # an opaque Python callable with the plugin's arithmetic: takes the generic path
    return native(x, u, dt)


for B in (256, 4096):
    w = WL.cartpole(B, "shared")
    d = {k: v.to(dev) for k, v in w["inp"].items()}
    for name, f, jac, gm in (("fused (device plugin)", native, native.jac, "analytic"),
                             ("generic, analytic jacobian callable", py_dyn, native.jac, "analytic"),
                             ("generic, auto_diff (vmap(jacrev))", py_dyn, None, "auto_diff")):
        C = d["C"].clone().requires_grad_(True)
        cm = GeneralQuadCost(4, 1, C, d["c"], d["Cf"], d["cf"], device="cuda", x_ref=d["xref"], u_ref=d["uref"])
        mpc = DifferentiableMPCController(f, 1.0, 0.05, 20, cm, u_min=torch.tensor([-20.0]), u_max=torch.tensor([20.0]), device="cuda",
                                          grad_method=gm, f_dyn_jac=jac)

        def step():
            X, U = mpc(d["x0"], d["Uinit"])
            U[:, 0].sum().backward()
            return X

        for _ in range(2):
            step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 5
        for _ in range(n):
            step()
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / n * 1e3
        print(f"B={B:5d} {name:40s} {ms:9.2f} ms per forward+backward  {B / ms * 1e3:12.4g} solves/s", flush=True)
