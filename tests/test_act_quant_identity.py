"""CPU: the division-free quotient of the activation-quant kernel (qqq_b200/csrc/act_quant.cu) is exact.

The reference computes `x / quant_scale` as an IEEE fp32 division per element (qlinear_marlin.py:267).  The kernel
uses r = RN(1/s) once per row and q = fma(fma(-s, RN(x*r), x), r, RN(x*r)); oracle/div_identity.c enumerates all
2.0e9 (finite fp16 x, positive finite fp16 s) pairs and compares with the hardware division bit for bit."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")
def test_reciprocal_newton_quotient_is_exact_for_all_fp16_pairs(tmp_path):
    flags = open("/proc/cpuinfo").read()
    if " fma" not in flags:
        pytest.skip("host CPU has no FMA instruction")
    exe = tmp_path / "div_identity"
    subprocess.check_call(["gcc", "-O2", "-mfma", "-ffp-contract=off", "-fopenmp", "-o", str(exe),
                           os.path.join(ROOT, "oracle", "div_identity.c"), "-lm"])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "fp32 quotient mismatches=0" in out.stdout and "int8 mismatches=0" in out.stdout, out.stdout
