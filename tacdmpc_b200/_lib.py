"""ctypes binding of the CUDA library (tacdmpc_b200/libtacd_ilqr.so, C ABI in include/tacd_ilqr.h).

There is NO fallback: if the library is missing or fails to load, every entry point
raises.  (The CPU oracle under oracle/ is test infrastructure and is never imported here.)
"""
from __future__ import annotations

import ctypes as C
import os
import threading

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
# TACD_LIB=path loads an experiment build of the same ABI instead (python -m tacdmpc_b200.build -DX=.. --out=path); A/B runs only
LIB_PATH = os.environ.get("TACD_LIB") or os.path.join(_HERE, "libtacd_ilqr.so")
_lock = threading.Lock()
_lib = None

_VP = C.c_void_p
_SIGS = {
    "tacd_abi_version": (C.c_int, []),
    "tacd_last_error": (C.c_char_p, []),
    "tacd_kernel_launches": (C.c_ulonglong, []),
    "tacd_refresh_env": (None, []),
    "tacd_forward_workspace_bytes": (C.c_size_t, [C.POINTER(_abi.Problem)]),
    "tacd_backward_workspace_bytes": (C.c_size_t, [C.POINTER(_abi.Problem)]),
    "tacd_ilqr_forward": (C.c_int, [C.POINTER(_abi.Problem), C.POINTER(_abi.ForwardIO), _VP, C.c_size_t, _VP]),
    "tacd_ilqr_backward": (C.c_int, [C.POINTER(_abi.Problem), C.POINTER(_abi.BackwardIO), _VP, C.c_size_t, _VP]),
    "tacd_rollout": (C.c_int, [C.POINTER(_abi.Problem)] + [_VP] * 4),
    "tacd_linearize": (C.c_int, [C.POINTER(_abi.Problem)] + [_VP] * 5),
    "tacd_objective": (C.c_int, [C.POINTER(_abi.Problem)] + [_VP] * 10),
    "tacd_riccati": (C.c_int, [C.POINTER(_abi.Problem)] + [_VP] * 13),
    "tacd_linesearch": (C.c_int, [C.POINTER(_abi.Problem)] + [_VP] * 15),
    "tacd_pnqp": (C.c_int, [C.c_int32, C.c_int32, C.c_int32] + [_VP] * 7),
    "tacd_al_update": (C.c_int, [C.c_int32] * 4 + [_VP, _VP, C.c_int32] + [_VP] * 4 + [C.c_double, C.c_double, C.c_int32] + [_VP] * 4),
    "tacd_env_step": (C.c_int, [C.POINTER(_abi.Problem), C.POINTER(_abi.EnvIO), _VP]),
    "tacd_gae": (C.c_int, [C.c_int32] * 3 + [_VP, _VP, C.c_double, C.c_double, _VP, _VP, _VP]),
    "tacd_discounted_targets": (C.c_int, [C.c_int32] * 3 + [_VP, _VP, C.c_double, _VP, _VP]),
    "tacd_fp32_probe": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, _VP, _VP, C.POINTER(C.c_double), _VP]),
    "tacd_probe_lti_gains": (C.c_int, [C.c_int32] * 4 + [C.c_double] + [_VP] * 6),
    "tacd_probe_fvf": (C.c_int, [C.c_int32] * 5 + [_VP] * 5),
    "tacd_closed_loop_advance": (C.c_int, [C.POINTER(_abi.Problem), C.c_int32, C.c_int32] + [_VP] * 7 + [C.c_int32] + [_VP] * 3),
}
EXPORTS = tuple(_SIGS)


class TacdError(RuntimeError):
    pass


class TacdUnsupported(TacdError):
    """TACD_EUNSUPPORTED: a valid request this build has no kernel for (callers may re-route, e.g. dense instead of
    compact-diagonal costs)."""


def lib():
    """Load (once) and return the CUDA library; raises if it is not built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise TacdError(
                        f"{LIB_PATH} is missing: build it with `python -m tacdmpc_b200.build` "
                        "(tacdmpc_b200 has no CPU or PyTorch fallback)")
                L = C.CDLL(LIB_PATH)
                for name, (res, args) in _SIGS.items():
                    fn = getattr(L, name)
                    fn.restype = res
                    fn.argtypes = args
                if L.tacd_abi_version() != _abi.ABI_VERSION:
                    raise TacdError("libtacd_ilqr.so ABI version mismatch; rebuild")
                _lib = L
    return _lib


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    msg = lib().tacd_last_error().decode("utf-8", "replace")
    kind = {_abi.EINVAL: "invalid argument", _abi.EUNSUPPORTED: "unsupported", _abi.ECUDA: "CUDA error",
            _abi.EWORKSPACE: "workspace too small"}.get(rc, f"error {rc}")
    if rc == _abi.EINVAL:
        raise ValueError(f"{what}: {kind}: {msg}")
    if rc == _abi.EUNSUPPORTED:
        raise TacdUnsupported(f"{what}: {kind}: {msg}")
    raise TacdError(f"{what}: {kind}: {msg}")
