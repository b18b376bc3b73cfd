import sys, torch
sys.path.insert(0, '.')
from tacdmpc_b200.DifferentialMPC import AugmentedLagrangianManager, ConstrainedQuadCost, DifferentiableMPCController
from tacdmpc_b200.dynamics import CartPole
B, T, dt, nx, nu = 4096, 20, 0.05, 4, 1
n = nx + nu
C = torch.zeros(T, n, n, device="cuda"); C[:, :nx, :nx] = torch.diag(torch.tensor([10.0, 1.0, 100.0, 1.0])); C[:, nx:, nx:] = 0.1
inf = float("inf"); f = CartPole()
torch.manual_seed(0)
for ang, lim in ((0.25, 0.4), (0.2, 0.4), (0.2, 0.5), (0.15, 0.3)):
    x0 = torch.tensor([0.0, 0.0, ang, 0.0], device="cuda") + 0.03 * torch.randn(B, nx, device="cuda")
    xmax = torch.tensor([lim, inf, inf, inf])
    for no in (1, 4, 8, 12):
        al = AugmentedLagrangianManager(rho_init=50.0, rho_growth=3.0, n_outer=no, tol=0.0)
        cm = ConstrainedQuadCost(al, xmax, -xmax, nx, nu, C, torch.zeros(T, n, device="cuda"), C[0] * 10, torch.zeros(n, device="cuda"), device="cuda")
        mpc = DifferentiableMPCController(f, T * dt, dt, T, cm, u_min=torch.tensor([-20.0]), u_max=torch.tensor([20.0]), grad_method="analytic", f_dyn_jac=f.jac, device="cuda")
        X, U = mpc(x0)
        v = al.violation
        print(f"ang {ang} lim {lim} n_outer {no}: max|x| {float(X[:,:,0].abs().max()):.4f} viol max {float(v.max()):.4f} mean {float(v.mean()):.5f} frac>1e-3 {float((v>1e-3).float().mean()):.3f} iters {float(mpc.iters_last.float().mean()):.1f} |u|max {float(U.abs().max()):.1f}")
