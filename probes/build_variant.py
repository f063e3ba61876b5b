"""Build a variant of libqqq_b200.so with extra -D flags (dev tooling).
usage: python probes/build_variant.py OUT.so [-DNAME[=V] ...]     e.g.  probes/libqqq_b200_trace.so -DQQQ_TRACE -DQQQ_TRACE_CTA=5"""
import os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "qqq_b200", "csrc")
out = os.path.abspath(sys.argv[1])
defs = sys.argv[2:]
nvcc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "--compiler-options", "-fPIC"] + defs
with tempfile.TemporaryDirectory() as td:
    objs, procs = [], []
    for s in ("qqq_c_api.cu", "qqq_gemm_sm100.cu", "act_quant.cu", "tp_reduce_quant.cu"):
        o = os.path.join(td, s + ".o")
        procs.append(subprocess.Popen([nvcc, "-c", os.path.join(CSRC, s), "-o", o] + flags))
        objs.append(o)
    assert all(p.wait() == 0 for p in procs)
    subprocess.check_call([nvcc, "-shared", "-o", out] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"])
print(out)
