"""Python face of the drop-in boundary: `qqq_gemm(...)` with the reference extension's exact signature
(csrc/qqq_gemm.h:23-36, exported as QQQ._CUDA.qqq_gemm at csrc/pybind.cpp:3-5) and `dynamic_quant`
(QQQ/gptq/qlinear/qlinear_marlin.py:265-268), both over the C ABI in include/qqq_b200.h.

PyTorch is used only for device memory, dtypes and the current stream.
"""
from __future__ import annotations

import torch

from . import _lib

_ERR_PROB_SHAPE = 1
_ERR_KERN_SHAPE = 2


def _ptr(t: torch.Tensor) -> int:
    return t.data_ptr()


def check_gemm_args(A, B, C, D, s1, s2, s3, workspace, max_par):
    """The reference's argument checks (csrc/qqq_gemm.cu:1060-1075, same messages) plus the layout / dtype / device
    checks it omits (it would read garbage or fault instead).  Returns (prob_m, prob_n, prob_k, groupsize) exactly as
    the reference derives them: N from C.size(1) (:1063), groupsize from s3 (:1065)."""
    prob_m = A.size(0)
    prob_n = C.size(1)
    prob_k = A.size(1)
    groupsize = -1 if s3.numel() == 0 else prob_k // s3.size(0)
    if groupsize != -1 and groupsize * s3.size(0) != prob_k:
        raise RuntimeError(f"k={prob_k} not compatible with {s3.size(0)} groups.")
    if workspace.numel() < prob_n // 128 * max_par:
        raise RuntimeError(f"workspace must be of size at least {prob_n // 128 * max_par}.")
    if s1.dtype != torch.float32:
        raise RuntimeError(f"s1 dtype must be float32, but got {s1.dtype}.")
    if s2.dtype != torch.float32:
        raise RuntimeError(f"s2 dtype must be float32, but got {s2.dtype}.")
    if s3.dtype != torch.float16:
        raise RuntimeError(f"s3 dtype must be float16, but got {s3.dtype}.")
    for name, t, dt in (("A", A, torch.int8), ("B", B, torch.int32), ("C", C, torch.int32), ("D", D, torch.float16),
                        ("workspace", workspace, torch.int32)):
        if t.dtype != dt:
            raise RuntimeError(f"{name} dtype must be {dt}, but got {t.dtype}.")
        if not t.is_contiguous():
            raise RuntimeError(f"{name} must be contiguous.")
        if t.device != A.device:
            raise RuntimeError(f"{name} is on {t.device}, expected {A.device}.")
    if C.size(0) < 64 * max_par and prob_m > 0:
        raise RuntimeError(f"C must have at least {64 * max_par} rows.")
    return prob_m, prob_n, prob_k, groupsize


def install_as_qqq_cuda() -> None:
    """Make `from QQQ._CUDA import qqq_gemm` (QQQ/gptq/qlinear/qlinear_marlin.py:22) resolve to this library, so the
    reference's own `QuantLinear` / `mul` run on the sm_100a kernel unchanged (option B)."""
    import sys
    import types

    mod = types.ModuleType("QQQ._CUDA")
    mod.__doc__ = "qqq_b200 shim for the reference's compiled extension (csrc/pybind.cpp:3-5)"

    def _positional(A, B, C, D, s1, s2, s3, workspace, thread_k, thread_n, sms, max_par, /):
        # pybind's m.def has no py::arg names: all twelve arguments are required and positional
        return qqq_gemm(A, B, C, D, s1, s2, s3, workspace, thread_k, thread_n, sms, max_par)

    mod.qqq_gemm = _positional
    sys.modules["QQQ._CUDA"] = mod
    if "QQQ" in sys.modules:
        setattr(sys.modules["QQQ"], "_CUDA", mod)


def qqq_gemm(A, B, C, D, s1, s2, s3, workspace, thread_k=-1, thread_n=-1, sms=-1, max_par=8):
    """INT8 x INT4 -> FP16 GEMM.  Same 12 arguments, meaning and errors as the reference `qqq_gemm`
    (csrc/qqq_gemm.cu:1048-1106):

    A int8 [M,K]; B int32 [K/16, 2N] (reference packing); C int32 [>=64*max_par, N] scratch;
    D fp16 [M,N] out; s1 fp32 [M,1]; s2 fp32 [1,N]; s3 fp16 [K/g, N] or empty; workspace int32
    [>= N/128*max_par] zeros.  Raises RuntimeError on the same conditions the reference raises AT_ERROR.
    N is taken from C.size(1) exactly like the reference (:1063).
    """
    prob_m, prob_n, prob_k, groupsize = check_gemm_args(A, B, C, D, s1, s2, s3, workspace, max_par)
    if not A.is_cuda:
        raise RuntimeError("qqq_gemm: tensors must be CUDA tensors (qqq_b200 has no CPU path).")
    dev = A.get_device()
    stream = torch.cuda.current_stream(dev).cuda_stream
    lib = _lib.load()
    err = lib.qqq_gemm_sm100a(
        _ptr(A), _ptr(B), _ptr(C), _ptr(D), _ptr(s1), _ptr(s2), _ptr(s3) if s3.numel() else None,
        prob_m, prob_n, prob_k, _ptr(workspace), groupsize, dev, stream, thread_k, thread_n, sms, max_par,
    )
    if err == _ERR_PROB_SHAPE:
        raise RuntimeError(
            f"Problem (m={prob_m}, n={prob_n}, k={prob_k}) not compatible with thread_k={thread_k}, thread_n={thread_n}."
        )
    if err == _ERR_KERN_SHAPE:
        raise RuntimeError(
            f"No kernel implementation for thread_k={thread_k}, thread_n={thread_n}, groupsize={groupsize}."
        )
    if err != 0:
        raise RuntimeError(f"qqq_gemm_sm100a failed (rc={err}): {_lib.last_error()}")


def qqq_gemm_bias(A, B, C, D, s1, s2, s3, workspace, bias, max_par=16, sms=-1):
    """`qqq_gemm` with QuantLinear.forward's bias add folded into the epilogue (qqq_gemm_bias_sm100a): same checks, same bits
    as `qqq_gemm(...)` followed by the reference's eager `D + bias`."""
    prob_m, prob_n, prob_k, groupsize = check_gemm_args(A, B, C, D, s1, s2, s3, workspace, max_par)
    if not A.is_cuda:
        raise RuntimeError("qqq_gemm_bias: tensors must be CUDA tensors (qqq_b200 has no CPU path).")
    if bias.dtype != torch.float16 or bias.numel() != prob_n or not bias.is_contiguous() or bias.device != A.device:
        raise RuntimeError("qqq_gemm_bias: bias must be a contiguous float16 [n] tensor on the device of A.")
    dev = A.get_device()
    err = _lib.load().qqq_gemm_bias_sm100a(
        _ptr(A), _ptr(B), _ptr(C), _ptr(D), _ptr(s1), _ptr(s2), _ptr(s3) if s3.numel() else None, _ptr(bias),
        prob_m, prob_n, prob_k, _ptr(workspace), groupsize, dev, torch.cuda.current_stream(dev).cuda_stream, sms, max_par,
    )
    if err != 0:
        raise RuntimeError(f"qqq_gemm_bias_sm100a failed (rc={err}): {_lib.last_error()}")


def qqq_gemm_reduce(A, B, C, d_multicast_ptr: int, s1, s2, s3, workspace, prob_n: int, max_par=16, sms=-1):
    """Row-shard GEMM whose epilogue ADDS its fp16 output into the multicast address `d_multicast_ptr` (an fp16 [M, prob_n]
    buffer replicated on every rank of the tensor-parallel group) instead of storing it: qqq_gemm_reduce_sm100a in
    include/qqq_b200.h.  Zeroing the replicas and the cross-rank barrier are the caller's (tp.AllReduceWorkspace)."""
    prob_m, prob_k = A.size(0), A.size(1)
    groupsize = -1 if s3.numel() == 0 else prob_k // s3.size(0)
    if not A.is_cuda:
        raise RuntimeError("qqq_gemm_reduce: tensors must be CUDA tensors (qqq_b200 has no CPU path).")
    if not d_multicast_ptr or d_multicast_ptr % 16:
        raise RuntimeError("qqq_gemm_reduce: a 16-byte aligned multicast address is required (no NVLS multicast support?).")
    if workspace.numel() < prob_n // 128 * max_par or C.numel() < 64 * max_par * prob_n:
        raise RuntimeError("qqq_gemm_reduce: workspace / C too small.")
    dev = A.get_device()
    err = _lib.load().qqq_gemm_reduce_sm100a(
        _ptr(A), _ptr(B), _ptr(C), d_multicast_ptr, _ptr(s1), _ptr(s2), _ptr(s3) if s3.numel() else None,
        prob_m, prob_n, prob_k, _ptr(workspace), groupsize, dev, torch.cuda.current_stream(dev).cuda_stream, -1, -1, sms,
        max_par,
    )
    if err != 0:
        raise RuntimeError(f"qqq_gemm_reduce_sm100a failed (rc={err}): {_lib.last_error()}")


def _ptr_array(ptrs):
    import ctypes

    return (ctypes.c_void_p * len(ptrs))(*[int(v) for v in ptrs])


def qqq_gemm_scatter(A, B, C, peer_partials, s1, s2, s3, workspace, prob_n: int, tp_rank: int, tp_world: int, tp_rows: int,
                     max_par=16, sms=-1):
    """Row-shard GEMM whose epilogue stores output row m into slot `tp_rank` of the partial-sum buffer of the rank that owns
    the row (m // tp_rows): qqq_gemm_scatter_sm100a in include/qqq_b200.h.  `peer_partials`: tp_world device addresses
    (peer-mapped) of the fp16 [tp_world][tp_rows][prob_n] buffers.  Followed on every rank by `tp_reduce_quant`."""
    prob_m, prob_k = A.size(0), A.size(1)
    groupsize = -1 if s3.numel() == 0 else prob_k // s3.size(0)
    if not A.is_cuda:
        raise RuntimeError("qqq_gemm_scatter: tensors must be CUDA tensors (qqq_b200 has no CPU path).")
    if len(peer_partials) != tp_world:
        raise RuntimeError("qqq_gemm_scatter: one partial-sum buffer address per rank is required.")
    if workspace.numel() < prob_n // 128 * max_par or C.numel() < 64 * max_par * prob_n:
        raise RuntimeError("qqq_gemm_scatter: workspace / C too small.")
    dev = A.get_device()
    err = _lib.load().qqq_gemm_scatter_sm100a(
        _ptr(A), _ptr(B), _ptr(C), _ptr_array(peer_partials), _ptr(s1), _ptr(s2), _ptr(s3) if s3.numel() else None,
        prob_m, prob_n, prob_k, _ptr(workspace), groupsize, dev, torch.cuda.current_stream(dev).cuda_stream, sms, max_par,
        tp_rank, tp_world, tp_rows,
    )
    if err != 0:
        raise RuntimeError(f"qqq_gemm_scatter_sm100a failed (rc={err}): {_lib.last_error()}")


def tp_reduce_quant(partials_ptr: int, a8_dst, a8_mc: int, s1_dst, s1_mc: int, h_out, bias, flags_ptr: int, peer_flags,
                    tp_rank: int, tp_world: int, tp_rows: int, prob_m: int, prob_n: int, dev: int):
    """Second half of the fused row-parallel exchange (qqq_tp_reduce_quant_sm100a): sum the partial-sum slots of this rank's
    rows, quantise per token, deliver int8 rows + scales to every rank.  Addresses are plain ints (symmetric memory)."""
    err = _lib.load().qqq_tp_reduce_quant_sm100a(
        partials_ptr, _ptr_array(a8_dst), a8_mc or None, _ptr_array(s1_dst), s1_mc or None,
        _ptr(h_out) if h_out is not None else None, _ptr(bias) if bias is not None else None, flags_ptr,
        _ptr_array(peer_flags), tp_rank, tp_world, tp_rows, prob_m, prob_n, dev, torch.cuda.current_stream(dev).cuda_stream,
    )
    if err != 0:
        raise RuntimeError(f"qqq_tp_reduce_quant_sm100a failed (rc={err}): {_lib.last_error()}")


def dynamic_quant(x: torch.Tensor):
    """Per-token int8 quantisation, bit-identical to the reference's 5 eager ops
    (qlinear_marlin.py:265-268), as ONE kernel.  x fp16 [M,K] -> (int8 [M,K], fp32 [M,1])."""
    if x.dtype != torch.float16 or not x.is_cuda:
        raise RuntimeError("dynamic_quant expects a CUDA fp16 tensor (qqq_b200 has no CPU path).")
    M, K = x.shape
    # a column slice of a wider row-major matrix (output of a merged GEMM) is quantised in place
    strided_ok = x.stride(1) == 1 and x.stride(0) >= K and x.stride(0) % 8 == 0 and x.data_ptr() % 16 == 0
    if not strided_ok:
        x = x.contiguous()
    q = torch.empty((M, K), dtype=torch.int8, device=x.device)
    s = torch.empty((M, 1), dtype=torch.float32, device=x.device)
    if M == 0:
        return q, s
    dev = x.get_device()
    err = _lib.load().qqq_act_quant_strided_sm100a(_ptr(x), x.stride(0), _ptr(q), _ptr(s), M, K, dev,
                                                   torch.cuda.current_stream(dev).cuda_stream)
    if err != 0:
        raise RuntimeError(f"qqq_act_quant_sm100a failed (rc={err}): {_lib.last_error()}")
    return q, s


def launch_count() -> int:
    return int(_lib.load().qqq_b200_launch_count())


def qqq_gemm_acc(A, B, C, D32, s3, workspace, max_par=16, sms=-1):
    """The W4A8 GEMM without its epilogue scales: D32 int32 [M, N] receives the exact integer accumulators
    (qqq_gemm_acc_sm100a in include/qqq_b200.h) — building block of the bit-exact tensor-parallel mode."""
    prob_m, prob_k, prob_n = A.size(0), A.size(1), C.size(1)
    groupsize = -1 if s3.numel() == 0 else prob_k // s3.size(0)
    if not A.is_cuda:
        raise RuntimeError("qqq_gemm_acc: tensors must be CUDA tensors (qqq_b200 has no CPU path).")
    if D32.dtype != torch.int32 or tuple(D32.shape) != (prob_m, prob_n) or not D32.is_contiguous():
        raise RuntimeError("qqq_gemm_acc: D32 must be a contiguous int32 [M, N] tensor.")
    if workspace.numel() < prob_n // 128 * max_par or C.size(0) < 64 * max_par:
        raise RuntimeError("qqq_gemm_acc: workspace / C too small.")
    dev = A.get_device()
    err = _lib.load().qqq_gemm_acc_sm100a(
        _ptr(A), _ptr(B), _ptr(C), _ptr(D32), _ptr(s3) if s3.numel() else None, prob_m, prob_n, prob_k, _ptr(workspace),
        groupsize, dev, torch.cuda.current_stream(dev).cuda_stream, sms, max_par,
    )
    if err != 0:
        raise RuntimeError(f"qqq_gemm_acc_sm100a failed (rc={err}): {_lib.last_error()}")
