"""CPU tests of the drop-in boundary: the C-ABI library builds, loads without a GPU driver, exports every
symbol include/qqq_b200.h declares, and rejects bad shapes with the reference's return codes before touching
the device."""
import ctypes
import os
import re
import subprocess

import pytest

from qqq_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "qqq_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qqq_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported():
    lib = _lib.load()
    decl = _declared_symbols()
    assert set(decl) == set(_lib.SYMBOLS)
    for s in decl:
        assert hasattr(lib, s), s
    out = subprocess.check_output(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], text=True)
    for s in decl:
        assert re.search(rf"\bT {s}\b", out), f"{s} is not an exported text symbol"


def test_no_torch_or_libcuda_link_dependency():
    out = subprocess.check_output(["ldd", str(_lib.LIB_PATH)], text=True)
    assert "libtorch" not in out and "libc10" not in out and "libcuda.so" not in out


def test_version_and_counters():
    lib = _lib.load()
    assert lib.qqq_b200_version() >= 100
    assert lib.qqq_b200_launch_count() >= 0


@pytest.mark.parametrize("n,k,tk,tn,gs,rc", [
    (100, 128, -1, -1, -1, 1),   # N not a multiple of 64
    (128, 100, -1, -1, -1, 1),   # K not a multiple of 64
    (128, 128, 96, 128, -1, 1),  # thread_k must be 64 or 128
    (128, 256, -1, -1, 100, 1),  # K % groupsize
    (128, 256, -1, -1, 64, 2),   # groupsize without a kernel (reference: ERR_KERN_SHAPE)
])
def test_shape_errors_before_any_cuda_call(n, k, tk, tn, gs, rc):
    lib = _lib.load()
    for entry in (lib.qqq_gemm_sm100a, lib.qqq_gemm_reduce_sm100a):  # the reducing variant shares every check
        got = entry(None, None, None, None, None, None, None, 4, n, k, None, gs, 0, None, tk, tn, -1, 16)
        assert got == rc
        assert len(_lib.last_error()) > 0


def test_empty_problem_returns_ok_without_device():
    lib = _lib.load()
    assert lib.qqq_gemm_sm100a(None, None, None, None, None, None, None, 0, 128, 128, None, -1, 0, None, -1, -1, -1, 16) == 0
    assert lib.qqq_gemm_reduce_sm100a(None, None, None, None, None, None, None, 0, 128, 128, None, -1, 0, None, -1, -1, -1, 16) == 0
    assert lib.qqq_gemm_acc_sm100a(None, None, None, None, None, 0, 128, 128, None, -1, 0, None, -1, 16) == 0
    assert lib.qqq_gemm_acc_sm100a(None, None, None, None, None, 4, 100, 128, None, -1, 0, None, -1, 16) == 1
    assert lib.qqq_act_quant_sm100a(None, None, None, 0, 128, 0, None) == 0
    assert lib.qqq_act_quant_sm100a(None, None, None, 4, 100, 0, None) == 1


def test_missing_library_is_a_loud_error(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setenv("QQQ_B200_LIB", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.QQQLibraryError, match="no CPU/PyTorch fallback"):
        _lib.load()
    monkeypatch.delenv("QQQ_B200_LIB")
    monkeypatch.setattr(_lib, "_lib", None)
    _lib.load()
