"""qqq_b200 — B200-native (sm_100a, tcgen05) W4A8 GEMM behind QQQ's `qqq_gemm()` / `QuantLinear` interface.

Public surface (mirrors the reference hot path only):
    qqq_gemm(A, B, C, D, s1, s2, s3, workspace, thread_k, thread_n, sms, max_par)   <- QQQ._CUDA.qqq_gemm
    mul(...), QuantLinear (alias QQQLinear)                                          <- QQQ.gptq.qlinear
    model.{make_quant, pack_model, build_quantized_model, fuse_qkv_gate_up, ...}     <- QQQ.gptq.apply_gptq / QQQ.gptq.models
"""
from .ops import dynamic_quant, launch_count, qqq_gemm  # noqa: F401
from . import model  # noqa: F401
from .qlinear import (QQQLinear, QuantizedActivation, QuantLinear, merge_quant_linears, mul, pack_int4_weights,  # noqa: F401
                      set_act_quant_cache)

__all__ = ["qqq_gemm", "dynamic_quant", "mul", "QuantLinear", "QQQLinear", "QuantizedActivation", "pack_int4_weights", "merge_quant_linears", "launch_count", "set_act_quant_cache"]
__version__ = "0.1.0"
