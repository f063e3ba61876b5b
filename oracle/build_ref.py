"""Build recipe for oracle/_ref: the UNMODIFIED reference extension, compiled where it lies.

TEST INFRASTRUCTURE ONLY.  Nothing under qqq_b200/ may import this.

Compiles /root/reference/csrc/{pybind.cpp,qqq_gemm.cu} (reference `setup.py:26-31` builds the same two
files as `QQQ._CUDA`) for sm_100a into oracle/_ref/qqq_ref_cuda.so.  No reference source is copied into
this repository: nvcc/g++ read the files from /root/reference directly and only the objects/.so land in
oracle/_ref/ (git-ignored, but shipped to the GPU box by gpurun).

The reference kernel is legacy mma.sync/cp.async code (csrc/qqq_gemm.cu:106-117); it compiles for
sm_100a unmodified and serves as (a) the GPU oracle that pins parity and (b) the on-box speed baseline.

Usage:  python oracle/build_ref.py            (no-op when the .so is newer than the sources)
"""
import os
import subprocess
import sys
import sysconfig
from pathlib import Path

HERE = Path(__file__).resolve().parent
OUT = HERE / "_ref"
REF = Path(os.environ.get("QQQ_REFERENCE_DIR", "/root/reference"))
NAME = "qqq_ref_cuda"


def so_path() -> Path:
    return OUT / f"{NAME}.so"


def build(force: bool = False, verbose: bool = True) -> Path | None:
    srcs = [REF / "csrc" / "pybind.cpp", REF / "csrc" / "qqq_gemm.cu"]
    so = so_path()
    if not all(s.exists() for s in srcs):
        # GPU box: /root/reference is absent; use the prebuilt .so if it travelled with the snapshot.
        return so if so.exists() else None
    if so.exists() and not force and all(so.stat().st_mtime > s.stat().st_mtime for s in srcs):
        return so
    import torch
    from torch.utils import cpp_extension as ce

    OUT.mkdir(parents=True, exist_ok=True)
    inc = []
    for p in ce.include_paths("cuda"):
        inc += ["-isystem", p]
    inc += ["-isystem", sysconfig.get_paths()["include"]]
    common = [
        "-DTORCH_EXTENSION_NAME=" + NAME,
        "-DTORCH_API_INCLUDE_EXTENSION_H",
        "-D_GLIBCXX_USE_CXX11_ABI=" + str(int(torch._C._GLIBCXX_USE_CXX11_ABI)),
        "-std=c++17",
        "-O3",
    ]
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    nvcc = os.path.join(cuda_home, "bin", "nvcc")
    o_cu = OUT / "qqq_gemm.o"
    o_cpp = OUT / "pybind.o"
    cmds = [
        [nvcc, "-c", str(srcs[1]), "-o", str(o_cu), "-gencode", "arch=compute_100a,code=sm_100a",
         "--compiler-options", "-fPIC", "--expt-relaxed-constexpr", "-w"] + common + inc,
        ["g++", "-c", str(srcs[0]), "-o", str(o_cpp), "-fPIC", "-w"] + common + inc,
    ]
    for c in cmds:
        if verbose:
            print("[build_ref]", " ".join(c[:6]), "...", flush=True)
        subprocess.check_call(c)
    libdirs = ce.library_paths("cuda")
    link = ["g++", "-shared", str(o_cpp), str(o_cu), "-o", str(so)]
    for d in libdirs:
        link += ["-L" + d, "-Wl,-rpath," + d]
    link += ["-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart"]
    subprocess.check_call(link)
    if verbose:
        print("[build_ref] built", so, flush=True)
    return so


def load():
    """Import the reference extension (needs torch imported first). Returns module or None."""
    so = build(verbose=False)
    if so is None or not so.exists():
        return None
    import importlib.util
    import torch  # noqa: F401  (libtorch symbols)

    spec = importlib.util.spec_from_file_location(NAME, str(so))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    p = build(force="--force" in sys.argv)
    print(p)
