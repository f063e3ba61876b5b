"""Build integration/_CUDA*.so — the pybind11 `QQQ._CUDA` replacement (integration/qqq_cuda_b200.cpp) — in-tree with g++
against this interpreter's torch headers, linked to ../qqq_b200/libqqq_b200.so by relative rpath.

    python integration/build_ext.py [--force]      ->  integration/_CUDA.so

Copy (or symlink) the result to <QQQ checkout>/QQQ/_CUDA.so together with libqqq_b200.so (or install qqq_b200 and adjust the
rpath): `from QQQ._CUDA import qqq_gemm` then resolves to the B200 kernel with no change to the reference's Python."""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent
SRC = HERE / "qqq_cuda_b200.cpp"
OUT = HERE / "_CUDA.so"


def build(force: bool = False) -> Path:
    sys.path.insert(0, str(ROOT))
    from qqq_b200 import build as qbuild

    lib = qbuild.build()
    if OUT.exists() and not force and OUT.stat().st_mtime > max(SRC.stat().st_mtime, (ROOT / "include/qqq_b200.h").stat().st_mtime):
        return OUT
    import torch
    from torch.utils import cpp_extension as ce

    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    inc = [f"-I{p}" for p in ce.include_paths()] + [f"-I{cuda_home}/include", f"-I{sysconfig.get_paths()['include']}"]
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = (["g++", "-O2", "-std=c++17", "-fPIC", "-shared", str(SRC), "-o", str(OUT), "-DTORCH_EXTENSION_NAME=_CUDA",
            "-DTORCH_API_INCLUDE_EXTENSION_H", f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
           + inc + [f"-L{tlib}", "-ltorch", "-ltorch_cpu", "-ltorch_python", "-lc10", "-lc10_cuda", "-ltorch_cuda",
                    f"-L{lib.parent}", "-lqqq_b200", f"-Wl,-rpath,{tlib}", "-Wl,-rpath,$ORIGIN/../qqq_b200"])
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
