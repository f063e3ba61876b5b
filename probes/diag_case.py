"""Diagnose a parity failure: run one big random case several times against torch._int_mm (dev tooling)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import qqq_oracle as O
import qqq_b200

M = int(sys.argv[1]); K = int(sys.argv[2]); N = int(sys.argv[3]); reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
dev = "cuda:0"
rng = np.random.default_rng(M)
w = rng.integers(-8, 8, size=(K, N))
B = torch.from_numpy(O.pack_B(w, False)).to(dev)
W8 = torch.from_numpy(O.w8_per_channel(w & 0xF).astype(np.int8)).to(dev)
s2_nat = (rng.random(N).astype(np.float32) + 0.5) * 1e-3
s2 = torch.from_numpy(O.permute_s_channel(s2_nat)).to(dev)
g = torch.Generator(device="cpu").manual_seed(M + 17)
A8 = torch.randint(-128, 128, (M, K), dtype=torch.int8, generator=g).to(dev)
s1 = ((torch.rand((M, 1), generator=g) + 0.5) * 1e-2).to(dev)
acc = torch._int_mm(A8, W8)
ref = ((acc.float() * torch.from_numpy(s2_nat).to(dev)[None, :]) * s1).half()
C = torch.zeros((1024, N), dtype=torch.int32, device=dev); ws = torch.zeros(N // 128 * 16, dtype=torch.int32, device=dev)
s3 = torch.zeros(0, dtype=torch.float16, device=dev)
for r in range(reps):
    D = torch.full((M, N), float("nan"), dtype=torch.float16, device=dev)
    qqq_b200.qqq_gemm(A8, B, C, D, s1, s2, s3, ws, -1, -1, -1, 16)
    torch.cuda.synchronize()
    bad = (D.view(torch.int16) != ref.view(torch.int16))
    nb = int(bad.sum())
    print(f"rep {r}: M={M} K={K} N={N} mismatches {nb}/{bad.numel()} nan={int(torch.isnan(D).sum())}", flush=True)
    if nb:
        idx = bad.nonzero()
        rows = idx[:, 0].unique(); cols = idx[:, 1].unique()
        print("  rows:", rows[:20].tolist(), "... n=", len(rows), " row tiles(256):", (rows // 256).unique().tolist()[:20])
        print("  cols:", cols[:20].tolist(), "... n=", len(cols), " col tiles(128):", (cols // 128).unique().tolist()[:20])
        m, n = idx[0].tolist()
        print("  first:", (m, n), "got", float(D[m, n]), "ref", float(ref[m, n]), "acc", int(acc[m, n]))
        # is the wrong value a different row's scale? check ratio
        print("  got/ref ratios sample:", [(float(D[i, j]) / (float(ref[i, j]) + 1e-9)) for i, j in idx[:6].tolist()])
