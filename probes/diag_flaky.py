"""Dev: hunt a flaky mismatch.  usage: python probes/diag_flaky.py M K N gs [iters]  (env knobs apply: QQQ_B200_SPLIT, ...)"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import qqq_b200
from oracle import qqq_oracle as O, build_ref
M, K, N, gs = (int(v) for v in sys.argv[1:5])
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 100
dev = "cuda:0"
p = O.make_problem(M, K, N, gs, seed=1000 + M)
want = torch.from_numpy(O.qqq_gemm_oracle(p["A8"], p["B"], p["s1"], p["s2"], p["s3"]).view(np.int16)).to(dev)
t = {k: torch.from_numpy(np.ascontiguousarray(p[k])).to(dev) for k in ("A8", "B", "s1", "s2", "s3")}
plan = (__import__("ctypes").c_int * 20)()
from qqq_b200 import _lib
_lib.load().qqq_b200_plan(M, N, K, gs, 148, 16, plan)
print("plan grid,n_tok,m_tiles,n_tiles,k_blocks,ksub,k_units,a_tiles,a_units,a_upc,b_tiles,b_tpc,nsw,nst,G,smem,pair:", list(plan)[:17])
try:
    ref = build_ref.load()
except Exception as e:
    ref = None; print("no ref kernel", e)
big = torch.empty(64 << 20, dtype=torch.int32, device=dev)
for name, fn in (("ours", qqq_b200.qqq_gemm), ("ref", ref.qqq_gemm if ref else None)):
    if fn is None: continue
    bad_runs = 0
    for it in range(iters):
        C = torch.randint(-2**31, 2**31 - 1, (16 * 64, N), dtype=torch.int32, device=dev)
        ws = torch.zeros(N // 128 * 16 + 16, dtype=torch.int32, device=dev)
        D = torch.full((M, N), float("nan"), dtype=torch.float16, device=dev)
        if it % 2: big.fill_(it)  # dirty L2 every other run
        fn(t["A8"], t["B"], C, D, t["s1"], t["s2"], t["s3"], ws, -1, -1, -1, 16)
        torch.cuda.synchronize()
        neq = (D.view(torch.int16) != want)
        n = int(neq.sum())
        if n or int(ws.abs().sum()):
            bad_runs += 1
            if bad_runs <= 4:
                idx = neq.nonzero()
                r, c = idx[:, 0], idx[:, 1]
                print(f"  {name} it={it}: {n} bad; rows {int(r.min())}..{int(r.max())} cols {int(c.min())}..{int(c.max())}; "
                      f"distinct 32-col blocks {sorted(set((c // 32).tolist()))[:8]} distinct 16-row blocks {sorted(set((r // 16).tolist()))[:12]} ws={int(ws.abs().sum())}")
                got = D[r[0], c[0]].item(); w = want.view(torch.float16)[r[0], c[0]].item()
                print(f"     first: D[{int(r[0])},{int(c[0])}] = {got} want {w}")
    print(f"{name}: {bad_runs} of {iters} runs differ from the oracle", flush=True)
