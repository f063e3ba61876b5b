"""Time qqq_b200.qqq_gemm on a list of shapes (dev tooling). Weights rotate through >L2 worth of copies."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import qqq_b200

dev = "cuda:0"
HBM, INT8 = 6570.0, 4428.0


def time_fn(fn, n_iter=20, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n_iter):
        fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n_iter * 1e3


def run(M, K, N, gs, graph=True):
    g = torch.Generator(device=dev).manual_seed(0)
    ncopy = max(2, int(300e6 // (K * N // 2)) + 1)
    ncopy = min(ncopy, 24)
    Bs = [torch.randint(-2**31, 2**31 - 1, (K // 16, 2 * N), dtype=torch.int32, device=dev, generator=g) for _ in range(ncopy)]
    s2 = torch.rand(1, N, device=dev) * 1e-3 + 5e-4
    s3 = (torch.rand(K // 128, N, device=dev) * 8 + 4).half() if gs == 128 else torch.zeros(0, dtype=torch.float16, device=dev)
    C = torch.zeros(16 * 64, N, dtype=torch.int32, device=dev)
    ws = torch.zeros(max(N // 128 * 16, 16), dtype=torch.int32, device=dev)
    A = torch.randint(-127, 128, (M, K), dtype=torch.int8, device=dev)
    s1 = torch.rand(M, 1, device=dev) * 1e-2 + 1e-3
    D = torch.empty(M, N, dtype=torch.float16, device=dev)
    n_it = 20 if M * K * N < 4e11 else 6
    if graph:
        # one graph replays ncopy launches (rotating weights) so host launch cost is excluded
        for i in range(2):
            qqq_b200.qqq_gemm(A, Bs[i % ncopy], C, D, s1, s2, s3, ws, -1, -1, -1, 16)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for i in range(ncopy):
                qqq_b200.qqq_gemm(A, Bs[i], C, D, s1, s2, s3, ws, -1, -1, -1, 16)
        reps = max(1, n_it // ncopy)
        t = time_fn(lambda i: gr.replay(), reps, 2) / ncopy
    else:
        t = time_fn(lambda i: qqq_b200.qqq_gemm(A, Bs[i % ncopy], C, D, s1, s2, s3, ws, -1, -1, -1, 16), n_it)
    fl = 2.0 * M * K * N
    by = M * K + K * N / 2 + 2 * M * N + 4 * M + 4 * N + (2 * (K // 128) * N if gs == 128 else 0)
    print(f"M={M:5d} K={K:5d} N={N:5d} g={gs:4d}: {t:8.2f} us  {fl / t * 1e-6:7.1f} TOP/s ({fl / t * 1e-6 / INT8 * 100:4.1f}% int8)  "
          f"{by / t * 1e-3:7.1f} GB/s ({by / t * 1e-3 / HBM * 100:4.1f}% hbm)", flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which == "one":
        run(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]))
        sys.exit(0)
    if which in ("all", "sweep"):
        for gs in (-1, 128):
            for M in (1, 16, 64, 128, 256, 1024, 4096):
                run(M, 8192, 21760, gs)
    if which in ("all", "llama"):
        for (K, N) in ((4096, 4096), (4096, 11008), (11008, 4096)):
            for M in (1, 16, 1024):
                run(M, K, N, -1)
        for (K, N) in ((4096, 4096), (4096, 14336), (14336, 4096), (4096, 1024)):
            run(32, K, N, 128)
