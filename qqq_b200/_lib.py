"""ctypes binding of libqqq_b200.so (the C ABI declared in include/qqq_b200.h).

There is NO fallback: if the shared library is missing or a symbol is absent, importing/using the ops
raises.  Build it with `python -m qqq_b200.build` (or `__graft_entry__.build()`).
"""
from __future__ import annotations

import ctypes
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libqqq_b200.so"

# Every symbol include/qqq_b200.h declares.
SYMBOLS = (
    "qqq_gemm_sm100a",
    "qqq_gemm_bias_sm100a",
    "qqq_gemm_reduce_sm100a",
    "qqq_gemm_acc_sm100a",
    "qqq_gemm_scatter_sm100a",
    "qqq_tp_reduce_quant_sm100a",
    "qqq_act_quant_sm100a",
    "qqq_act_quant_strided_sm100a",
    "qqq_b200_version",
    "qqq_b200_last_error",
    "qqq_b200_launch_count",
    "qqq_b200_plan",
)

_lib = None


class QQQLibraryError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("QQQ_B200_LIB", LIB_PATH))
    if not path.exists():
        raise QQQLibraryError(
            f"{path} not found: the CUDA library is not built. Run `python -m qqq_b200.build`. "
            "qqq_b200 has no CPU/PyTorch fallback."
        )
    lib = ctypes.CDLL(str(path))
    for s in SYMBOLS:
        if not hasattr(lib, s):
            raise QQQLibraryError(f"{path} does not export `{s}` (stale build?)")
    vp, ci = ctypes.c_void_p, ctypes.c_int
    lib.qqq_gemm_sm100a.argtypes = [vp, vp, vp, vp, vp, vp, vp, ci, ci, ci, vp, ci, ci, vp, ci, ci, ci, ci]
    lib.qqq_gemm_sm100a.restype = ci
    lib.qqq_gemm_bias_sm100a.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, ci, ci, ci, vp, ci, ci, vp, ci, ci]
    lib.qqq_gemm_bias_sm100a.restype = ci
    lib.qqq_gemm_reduce_sm100a.argtypes = list(lib.qqq_gemm_sm100a.argtypes)
    lib.qqq_gemm_reduce_sm100a.restype = ci
    lib.qqq_gemm_acc_sm100a.argtypes = [vp, vp, vp, vp, vp, ci, ci, ci, vp, ci, ci, vp, ci, ci]
    lib.qqq_gemm_acc_sm100a.restype = ci
    lib.qqq_gemm_scatter_sm100a.argtypes = [vp, vp, vp, vp, vp, vp, vp, ci, ci, ci, vp, ci, ci, vp, ci, ci, ci, ci, ci]
    lib.qqq_gemm_scatter_sm100a.restype = ci
    lib.qqq_tp_reduce_quant_sm100a.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, ci, vp]
    lib.qqq_tp_reduce_quant_sm100a.restype = ci
    lib.qqq_act_quant_sm100a.argtypes = [vp, vp, vp, ci, ci, ci, vp]
    lib.qqq_act_quant_sm100a.restype = ci
    lib.qqq_act_quant_strided_sm100a.argtypes = [vp, ctypes.c_longlong, vp, vp, ci, ci, ci, vp]
    lib.qqq_act_quant_strided_sm100a.restype = ci
    lib.qqq_b200_version.argtypes = []
    lib.qqq_b200_version.restype = ci
    lib.qqq_b200_last_error.argtypes = []
    lib.qqq_b200_last_error.restype = ctypes.c_char_p
    lib.qqq_b200_launch_count.argtypes = []
    lib.qqq_b200_launch_count.restype = ctypes.c_longlong
    lib.qqq_b200_plan.argtypes = [ci, ci, ci, ci, ci, ci, ctypes.POINTER(ci)]
    lib.qqq_b200_plan.restype = ci
    _lib = lib
    return lib


def last_error() -> str:
    msg = load().qqq_b200_last_error()
    return msg.decode() if msg else ""
