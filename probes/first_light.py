"""First-light check for the GEMM kernel on the GPU box (dev tooling): a few parity cases with diagnostics."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import qqq_oracle as O
from gpu_util import run_gemm, bits

for (M, K, N, gs) in [(16, 128, 128, -1), (16, 256, 128, -1), (1, 1024, 256, -1), (16, 128, 128, 128), (40, 2048, 384, 128), (300, 512, 256, -1), (20, 4096, 4096, -1)]:
    p = O.make_problem(M, K, N, gs, seed=1)
    t0 = time.time()
    try:
        D, C, ws = run_gemm(p, N)
    except Exception as e:
        print("CASE", M, K, N, gs, "EXC", e); continue
    Dref = O.qqq_gemm_oracle(p["A8"], p["B"], p["s1"], p["s2"], p["s3"])
    bad = bits(D) != bits(Dref)
    print(f"CASE M={M} K={K} N={N} g={gs}: mismatches {int(bad.sum())}/{bad.size}  C_clean={int(C.abs().sum())==0} ws_clean={int(ws.abs().sum())==0} [{time.time()-t0:.2f}s]", flush=True)
    if bad.any():
        idx = np.argwhere(bad)
        print("  first bad (m,n):", idx[:8].tolist())
        print("  cols with errors (mod 128):", sorted(set((idx[:,1] % 128).tolist()))[:40])
        print("  rows with errors:", sorted(set(idx[:,0].tolist()))[:40])
        m, n = idx[0]
        print("  got", D[m, n], "ref", Dref[m, n], "row0 got", D[m, :6], "ref", Dref[m, :6])
