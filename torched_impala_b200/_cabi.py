"""ctypes binding of include/impala_b200.h (the C-ABI boundary of the hot path).

There is no CPU fallback anywhere in this package: if `libimpala_b200.so` is
missing, or a call returns non-zero, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libimpala_b200.so")

MODE_REFERENCE = 0
MODE_PAPER = 1
MODES = {"reference": MODE_REFERENCE, "paper": MODE_PAPER}

_ERRORS = {-1: "IMPALA_ERR_BAD_ARG", -2: "IMPALA_ERR_UNSUPPORTED_SHAPE",
           -3: "IMPALA_ERR_WORKSPACE_TOO_SMALL"}

_p = C.c_void_p
_i = C.c_int
_i64 = C.c_int64
_f = C.c_float

# name -> (restype, argtypes); must list every symbol include/impala_b200.h declares
SIGNATURES = {
    "impala_abi_version": (_i, []),
    "impala_compiled_sm": (_i, []),
    "impala_param_layout": (_i, [_i, _i, _i, C.POINTER(_i64), C.POINTER(_i64)]),
    "impala_batch_layout": (_i, [_i, _i, _i, _i, C.POINTER(_i64), C.POINTER(_i64)]),
    "impala_ingest": (_i, [_p, _p, _i64, _p]),
    "impala_ingest_shard": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "impala_mlp_forward": (_i, [_p, _p, _p, _i, _i, _i, _i, _p]),
    "impala_launch_count": (C.c_longlong, []),
    "impala_mlp_forward_pair": (_i, [_p] * 5 + [_i] * 6 + [_p]),
    "impala_mlp_backward_workspace": (_i64, [_i, _i, _i, _i]),
    "impala_mlp_backward": (_i, [_p, _p, _p, _p, _p, _i64, _i, _i, _i, _i, _p]),
    "impala_mlp_backward_pair": (_i, [_p] * 8 + [_i64, _p, _i64] + [_i] * 6 + [_p]),
    "impala_peer_alloc": (_i, [_i64, _p, _p]),
    "impala_peer_open": (_i, [_p, _p]),
    "impala_peer_close": (_i, [_p]),
    "impala_peer_free": (_i, [_p]),
    "impala_peer_push": (_i, [_p, _i64, _p, _p, _i64, _i64, _i, _i, _p]),
    "impala_mlp_backward_pair_push_supported": (_i, [_i] * 6),
    "impala_mlp_backward_pair_push": (_i, [_p] * 6 + [_i64, _p, _i64] + [_i] * 6 + [_p, _i, _p, _p, _i64, _i64, _i, _i, _p]),
    "impala_gather_clip_adam": (_i, [_p] * 4 + [_i64, _i64, _i, _i, _p, _p, _p, _i64, _i64] + [_f] * 5
                                + [_p, _p, C.c_double, _p]),
    "impala_vtrace": (_i, [_p] * 9 + [_i, _i, _i, _f, _f, _f, _i, _p]),
    "impala_vtrace_loss_workspace": (_i64, [_i, _i, _i]),
    "impala_vtrace_loss": (_i, [_p] * 13 + [_i64, _i, _i, _i] + [_f] * 7 + [_i, _p]),
    "impala_clip_adam": (_i, [_p, _p, _p, _p, _p, _i64, _i64, _f, _f, _f, _f, _f, _p, _p]),
    "impala_policy_terms": (_i, [_p, _p, _p, _p, _i, _i, _p]),
    "impala_policy_terms_backward": (_i, [_p, _p, _p, _p, _p, _i, _i, _p]),
    "impala_reduce": (_i, [_p, _p, _i64, _i, _p, _p]),
}

_lib = None


class ImpalaCudaError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImpalaCudaError(
                f"{LIB_PATH} is missing - build it with `python -m torched_impala_b200.build` "
                "(there is no CPU fallback for the learner hot path)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    if rc < 0:
        raise ImpalaCudaError(f"{what}: {_ERRORS.get(rc, rc)}")
    raise ImpalaCudaError(f"{what}: CUDA error {rc}")


def param_layout(O: int, H: int, N2: int):
    offs = (_i64 * 4)()
    total = _i64()
    check(lib().impala_param_layout(O, H, N2, offs, C.byref(total)), "impala_param_layout")
    return list(offs), total.value


def batch_layout(T: int, B: int, O: int, A: int):
    offs = (_i64 * 6)()
    total = _i64()
    check(lib().impala_batch_layout(T, B, O, A, offs, C.byref(total)), "impala_batch_layout")
    return list(offs), total.value
