"""Build libqqq_b200.so in-tree with nvcc for sm_100a.   python -m qqq_b200.build [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.  Static cudart, no libcuda link
dependency (the driver API is reached through cudaGetDriverEntryPoint), no torch dependency.
"""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OUT = HERE / "libqqq_b200.so"
SOURCES = ["qqq_c_api.cu", "qqq_gemm_sm100.cu", "act_quant.cu", "tp_reduce_quant.cu"]
HEADERS = ["qqq_common.cuh", "quant_common.cuh", "qqq_gemm_sm100.h", "../../include/qqq_b200.h"]


def nvcc_path() -> str:
    home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    return os.path.join(home, "bin", "nvcc")


def needs_build() -> bool:
    if not OUT.exists():
        return True
    t = OUT.stat().st_mtime
    deps = [CSRC / s for s in SOURCES] + [(CSRC / h).resolve() for h in HEADERS] + [Path(__file__)]
    return any(d.stat().st_mtime > t for d in deps if d.exists())


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return OUT
    objs = []
    flags = [
        "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
        "--compiler-options", "-fPIC", "-Xptxas", "-v" if verbose else "-warn-spills",
    ]
    procs = []
    for s in SOURCES:
        o = CSRC / (s + ".o")
        cmd = [nvcc_path(), "-c", str(CSRC / s), "-o", str(o)] + flags
        if verbose:
            print(" ".join(cmd), flush=True)
        procs.append((subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True), s))
        objs.append(str(o))
    for pr, s in procs:
        out, _ = pr.communicate()
        if verbose or pr.returncode != 0:
            print(out)
        if pr.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}")
    link = [nvcc_path(), "-shared", "-o", str(OUT)] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
    subprocess.check_call(link)
    return OUT


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(p)
