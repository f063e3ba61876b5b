"""Run the GEMM a few times at one (M, groupsize) of the BASELINE sweep shape (dev tooling for ncu)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import qqq_b200

M = int(sys.argv[1]); gs = int(sys.argv[2]); iters = int(sys.argv[3]) if len(sys.argv) > 3 else 6
K = int(sys.argv[4]) if len(sys.argv) > 4 else 8192
N = int(sys.argv[5]) if len(sys.argv) > 5 else 21760
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
Bs = [torch.randint(-2**31, 2**31 - 1, (K // 16, 2 * N), dtype=torch.int32, device=dev, generator=g) for _ in range(3)]
s2 = torch.rand(1, N, device=dev) * 1e-3 + 5e-4
s3 = (torch.rand(K // 128, N, device=dev) * 8 + 4).half() if gs == 128 else torch.zeros(0, dtype=torch.float16, device=dev)
C = torch.zeros(16 * 64, N, dtype=torch.int32, device=dev)
ws = torch.zeros(max(N // 128 * 16, 16), dtype=torch.int32, device=dev)
A = torch.randint(-127, 128, (M, K), dtype=torch.int8, device=dev)
s1 = torch.rand(M, 1, device=dev) * 1e-2 + 1e-3
D = torch.empty(M, N, dtype=torch.float16, device=dev)
for i in range(iters):
    qqq_b200.qqq_gemm(A, Bs[i % 3], C, D, s1, s2, s3, ws, -1, -1, -1, 16)
torch.cuda.synchronize()
print("ok", M, gs, K, N)
