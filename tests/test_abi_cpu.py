"""CPU-only checks of the boundary: the C-ABI library loads and exports every symbol include/tacd_ilqr.h
declares (no compute calls without a GPU), struct layouts agree, argument validation works."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "tacd_ilqr.h")).read()
    return sorted(set(re.findall(r"\b(tacd_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from tacdmpc_b200 import _lib, build
    build.build_extension()
    L = _lib.lib()
    names = _declared()
    assert len(names) >= 12
    for nm in names:
        assert hasattr(L, nm), nm
    assert set(names) == set(_lib.EXPORTS)
    assert L.tacd_abi_version() == 4


def test_oracle_exports_same_entry_points(oracle):
    L = oracle.lib()
    for nm in ("ilqr_forward", "ilqr_backward", "rollout", "linearize", "objective", "riccati", "linesearch", "pnqp"):
        assert hasattr(L, "tacd_oracle_" + nm)


def test_struct_layout_matches_header():
    """sizeof(tacd_problem) etc. as compiled by gcc == the ctypes mirror."""
    import subprocess, tempfile
    from tacdmpc_b200 import _abi
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "tacd_ilqr.h"
int main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(tacd_problem), sizeof(tacd_forward_io), sizeof(tacd_backward_io),
 offsetof(tacd_problem, dyn_params), offsetof(tacd_problem, alphas), offsetof(tacd_problem, umax), offsetof(tacd_problem, C_diag),
 offsetof(tacd_problem, compat_exact), sizeof(tacd_env_io), offsetof(tacd_env_io, max_steps), offsetof(tacd_env_io, w_u));return 0;}'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(prog)
        subprocess.run(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")], check=True)
        out = subprocess.run([os.path.join(d, "t")], capture_output=True, text=True, check=True).stdout.split()
    got = [int(v) for v in out]
    exp = [C.sizeof(_abi.Problem), C.sizeof(_abi.ForwardIO), C.sizeof(_abi.BackwardIO), _abi.Problem.dyn_params.offset,
           _abi.Problem.alphas.offset, _abi.Problem.umax.offset, _abi.Problem.C_diag.offset, _abi.Problem.compat_exact.offset,
           C.sizeof(_abi.EnvIO), _abi.EnvIO.max_steps.offset, _abi.EnvIO.w_u.offset]
    assert got == exp


def test_argument_validation_without_gpu():
    from tacdmpc_b200 import _abi, _lib
    L = _lib.lib()
    p = _abi.make_problem(B=4, T=0, nx=4, nu=1, dtype=_abi.F32, dyn_id=_abi.DYN_CARTPOLE, dt=0.05)
    io = _abi.ForwardIO()
    rc = L.tacd_ilqr_forward(C.byref(p), C.byref(io), None, 0, None)
    assert rc == _abi.EINVAL and b"bad sizes" in L.tacd_last_error()
    p = _abi.make_problem(B=4, T=5, nx=3, nu=1, dtype=_abi.F32, dyn_id=_abi.DYN_CARTPOLE, dt=0.05)
    assert L.tacd_ilqr_forward(C.byref(p), C.byref(io), None, 0, None) == _abi.EINVAL
    with pytest.raises(ValueError):
        _lib.check(_abi.EINVAL, "x")


def test_cpu_tensors_are_rejected_loudly():
    import torch
    from tacdmpc_b200 import _lib
    from tacdmpc_b200.DifferentialMPC import DifferentiableMPCController, GeneralQuadCost
    from tacdmpc_b200.dynamics import CartPole
    T, n = 5, 5
    cm = GeneralQuadCost(4, 1, torch.eye(n).repeat(T, 1, 1), torch.zeros(T, n), torch.eye(n), torch.zeros(n), device="cpu",
                         x_ref=torch.zeros(2, T + 1, 4), u_ref=torch.zeros(2, T, 1))
    mpc = DifferentiableMPCController(CartPole(), T * 0.05, 0.05, T, cm, device="cpu", grad_method="analytic")
    with pytest.raises(RuntimeError, match="CUDA"):
        mpc(torch.zeros(2, 4))


def test_ctor_errors_match_reference():
    import torch
    from tacdmpc_b200.DifferentialMPC import DifferentiableMPCController, GeneralQuadCost, GradMethod
    T, n = 3, 3
    cm = GeneralQuadCost(2, 1, torch.zeros(T, n, n), torch.zeros(T, n), torch.zeros(n, n), torch.zeros(n))
    with pytest.raises(ValueError):
        DifferentiableMPCController(lambda x, u, dt: x, 1.0, 0.1, T, cm, device="cpu", grad_method="analytic")
    c = DifferentiableMPCController(lambda x, u, dt: x, 1.0, 0.1, T, cm, device="cpu", grad_method="AUTO_DIFF")
    assert c.grad_method is GradMethod.AUTO_DIFF and c.N_sim == T and c.converged is True
    assert cm.x_ref.shape == (1, T + 1, 2) and cm.u_ref.shape == (1, T, 1)
