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


C_CLIENT = r"""
/* A host written in plain C99 binding the boundary exactly as include/qqq_b200.h declares it. */
#include <stdio.h>
#include <string.h>
#include "qqq_b200.h"

int main(void) {
  int plan[20];
  if (qqq_b200_version() < 100) return 10;
  /* configs[4], M = 1024: pure host planning, no GPU needed */
  if (qqq_b200_plan(1024, 21760, 8192, -1, 148, 16, plan) != QQQ_OK) return 11;
  if (plan[0] < 1 || plan[0] > 148 || plan[1] % 16 != 0 || plan[3] != 170 || plan[4] != 64) return 12;
  /* the reference's return codes (csrc/qqq_gemm.cu:947-948) come back before any CUDA call */
  if (qqq_gemm_sm100a(NULL, NULL, NULL, NULL, NULL, NULL, NULL, 4, 100, 128, NULL, -1, 0, NULL, -1, -1, -1, 16) !=
      QQQ_ERR_PROB_SHAPE) return 13;
  if (strlen(qqq_b200_last_error()) == 0) return 14;
  if (qqq_gemm_sm100a(NULL, NULL, NULL, NULL, NULL, NULL, NULL, 4, 128, 256, NULL, 64, 0, NULL, -1, -1, -1, 16) !=
      QQQ_ERR_KERN_SHAPE) return 15;
  if (qqq_gemm_bias_sm100a(NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, 0, 128, 128, NULL, -1, 0, NULL, -1, 16) != QQQ_OK)
    return 16; /* empty problem: silent no-op like the reference (:1002-1003) */
  if (qqq_act_quant_strided_sm100a(NULL, 128, NULL, NULL, 0, 128, 0, NULL) != QQQ_OK) return 17;
  printf("grid=%d n_tok=%d launches=%lld\n", plan[0], plan[1], qqq_b200_launch_count());
  return 0;
}
"""


def test_plain_c_host_compiles_links_and_calls(tmp_path):
    """The header is C (not C++-only) and a C host can bind every kind of entry point: gcc -std=c99 -pedantic -Werror."""
    _lib.load()
    src = tmp_path / "client.c"
    src.write_text(C_CLIENT)
    exe = tmp_path / "client"
    libdir = os.path.dirname(str(_lib.LIB_PATH))
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           str(src), "-o", str(exe), "-L", libdir, "-lqqq_b200", f"-Wl,-rpath,{libdir}"])
    out = subprocess.check_output([str(exe)], text=True)
    assert out.startswith("grid=") and "launches=" in out
