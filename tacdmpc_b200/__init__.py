"""tacdmpc_b200 — B200-native batched differentiable iLQR (drop-in for TACDMPC's DifferentialMPC).

Importing the package is cheap (no torch import, no CUDA library load); the
CUDA library is loaded on first use by :mod:`tacdmpc_b200._lib` and its absence
is a hard error — there is no CPU fallback.
"""
__version__ = "0.1.0"
