"""Time the reference kernel (oracle/_ref) and fp16 cuBLAS on the BASELINE sweep (dev tooling).
Weights rotate through enough copies to exceed L2 (126 MB) so small-M numbers are HBM numbers."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import build_ref  # noqa: E402


def time_fn(fn, n_iter=20, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n_iter):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n_iter * 1e3  # us


def main():
    mod = build_ref.load()
    dev = "cuda:0"
    K, N = 8192, 21760
    ncopy = 4
    g = torch.Generator(device=dev).manual_seed(0)
    Bs = [torch.randint(-2**31, 2**31 - 1, (K // 16, 2 * N), dtype=torch.int32, device=dev, generator=g) for _ in range(ncopy)]
    Wh = [torch.randn(K, N, dtype=torch.float16, device=dev) * 0.02 for _ in range(ncopy)]
    s2 = torch.rand(1, N, device=dev) * 1e-2 + 5e-3
    s3g = (torch.rand(K // 128, N, device=dev) * 8 + 4).half()
    s3e = torch.zeros(0, dtype=torch.float16, device=dev)
    C = torch.zeros(16 * 64, N, dtype=torch.int32, device=dev)
    ws = torch.zeros(N // 128 * 16 + 64, dtype=torch.int32, device=dev)
    print(f"{'M':>5} {'fp16 us':>9} {'fp16 TF':>8} | {'ref pc us':>9} {'TOP/s':>7} {'x fp16':>6} | {'ref g128':>9} {'TOP/s':>7} {'x fp16':>6}")
    for M in (1, 16, 128, 1024, 4096):
        A = torch.randint(-127, 128, (M, K), dtype=torch.int8, device=dev)
        Ah = torch.randn(M, K, dtype=torch.float16, device=dev)
        s1 = torch.rand(M, 1, device=dev) * 1e-2 + 1e-3
        D = torch.empty(M, N, dtype=torch.float16, device=dev)
        Dh = torch.empty(M, N, dtype=torch.float16, device=dev)
        t_h = time_fn(lambda i: torch.matmul(Ah, Wh[i % ncopy], out=Dh))
        t_pc = time_fn(lambda i: mod.qqq_gemm(A, Bs[i % ncopy], C, D, s1, s2, s3e, ws, -1, -1, -1, 16))
        t_g = time_fn(lambda i: mod.qqq_gemm(A, Bs[i % ncopy], C, D, s1, s2, s3g, ws, -1, -1, -1, 16))
        fl = 2.0 * M * K * N
        print(f"{M:>5} {t_h:9.1f} {fl / t_h * 1e-6:8.1f} | {t_pc:9.1f} {fl / t_pc * 1e-6:7.1f} {t_h / t_pc:6.2f} | "
              f"{t_g:9.1f} {fl / t_g * 1e-6:7.1f} {t_h / t_g:6.2f}", flush=True)


if __name__ == "__main__":
    main()
