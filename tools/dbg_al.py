import sys, numpy as np, torch
sys.path.insert(0, '.')
from tacdmpc_b200 import ops
from tacdmpc_b200.DifferentialMPC import AugmentedLagrangianManager, ConstrainedQuadCost, DifferentiableMPCController
from tacdmpc_b200.dynamics import DoubleIntegrator
nx, nu, T, dt = 2, 1, 20, 0.1
n = nx + nu
f = DoubleIntegrator(dt)
dd = dict(device="cuda", dtype=torch.float64)
C = torch.zeros(T, n, n, **dd); C[:, 0, 0] = 10; C[:, 1, 1] = 1; C[:, 2, 2] = 0.1
Cf = torch.zeros(n, n, **dd); Cf[0, 0] = 100; Cf[1, 1] = 10
xmax = torch.tensor([float("inf"), 1.0]); xmin = -xmax
x0 = torch.tensor([[2.0, -0.8]], **dd)
for no in range(1, 9):
    al = AugmentedLagrangianManager(rho_init=10.0, rho_growth=4.0, rho_max=1e7, n_outer=no, tol=0.0)
    cm = ConstrainedQuadCost(al, xmax, xmin, nx, nu, C, torch.zeros(T, n, **dd), Cf, torch.zeros(n, **dd), device="cuda")
    mpc = DifferentiableMPCController(f, T * dt, dt, T, cm, grad_method="analytic", f_dyn_jac=f.jac, device="cuda", reg_eps=1e-9)
    X, U = mpc(x0)
    print(no, "viol", float(al.violation.max()), "rho", al.rho, "iters", int(mpc.iters_last[0]), "v:", np.round(X[0, :8, 1].cpu().numpy(), 3), "lam_lo", np.round(al.lam_lo[0, :6, 1].cpu().numpy(), 2))
