"""Drop-in for the reference package ``DifferentialMPC`` (DifferentialMPC/__init__.py:1-16): same names,
same constructor arguments, same side effects — the solve itself runs in hand-written sm_100a kernels.

    import tacdmpc_b200.DifferentialMPC as DifferentialMPC
"""
from .cost import DiagQuadCost, GeneralQuadCost
from .AugmentedLagrangianManager import AugmentedLagrangianManager, ConstrainedQuadCost
from .controller import DifferentiableMPCController, GradMethod, ILQRSolve, ILQRSolveDiag
from .utils import pnqp, batched_jacobian, jacobian_finite_diff_batched

__all__ = [
    "GeneralQuadCost",
    "DiagQuadCost",
    "ConstrainedQuadCost",
    "AugmentedLagrangianManager",
    "ILQRSolve",
    "ILQRSolveDiag",
    "DifferentiableMPCController",
    "GradMethod",
    "pnqp",
    "batched_jacobian",
    "jacobian_finite_diff_batched",
]
