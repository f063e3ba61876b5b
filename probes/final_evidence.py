"""Round-2 closing evidence (dev tooling): what the decode step (configs[2]) and the activation quantisation really cost.

    python probes/final_evidence.py time     # CUDA-event timings, graph replays, inputs rotated beyond L2 and L2-warm
    ncu --set full -k regex:'act_quant|qqq_gemm' -s 7 -c 7 ... python probes/final_evidence.py ncu
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import qqq_b200  # noqa: E402
from qqq_b200 import ops  # noqa: E402

dev = "cuda:0"
HBM = 6551.0
AQ = [(1024, 4096), (1024, 11008), (32, 4096)]
DEC = [(32, 4096, 14336), (32, 4096, 1024), (32, 14336, 4096), (32, 4096, 4096)]  # Llama-3-8B decode, g128


def gemm_args(M, K, N, ncopy, gen):
    Bs = [torch.randint(-2**31, 2**31 - 1, (K // 16, 2 * N), dtype=torch.int32, device=dev, generator=gen) for _ in range(ncopy)]
    s3s = [(torch.rand(K // 128, N, device=dev) * 8 + 4).half() for _ in range(ncopy)]
    s2 = torch.rand(1, N, device=dev) * 1e-3 + 5e-4
    C = torch.zeros(16 * 64, N, dtype=torch.int32, device=dev)
    ws = torch.zeros(max(N // 128 * 16, 16), dtype=torch.int32, device=dev)
    A = torch.randint(-127, 128, (M, K), dtype=torch.int8, device=dev)
    s1 = torch.rand(M, 1, device=dev) * 1e-2 + 1e-3
    D = torch.empty(M, N, dtype=torch.float16, device=dev)
    return A, Bs, s3s, C, D, s1, s2, ws


def replay_us(launch_all, n):
    launch_all()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        launch_all()
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5 / n * 1e3


def main_time():
    gen = torch.Generator(device=dev).manual_seed(0)
    print(torch.cuda.get_device_name(0), flush=True)
    for (M, K) in AQ:
        by = 3 * M * K + 4 * M
        ncold = max(2, int(300e6 // (2 * M * K)) + 1)
        xs = [torch.randn(M, K, device=dev, dtype=torch.float16) for _ in range(ncold)]
        cold = replay_us(lambda: [ops.dynamic_quant(x) for x in xs], ncold)
        warm = replay_us(lambda: [ops.dynamic_quant(xs[0]) for _ in range(16)], 16)
        print(f"act_quant M={M:5d} K={K:5d}: {cold:6.2f} us rotating inputs ({by / cold * 1e-3:6.0f} GB/s = {by / cold * 1e-3 / HBM:4.2f} of HBM)"
              f"   {warm:6.2f} us same input (L2-resident, back-to-back launches)", flush=True)
    for (M, K, N) in DEC:
        by = M * K + K * N / 2 + 2 * M * N + 4 * M + 4 * N + 2 * (K // 128) * N
        ncopy = min(24, max(2, int(300e6 // (K * N // 2)) + 1))
        A, Bs, s3s, C, D, s1, s2, ws = gemm_args(M, K, N, ncopy, gen)
        t = replay_us(lambda: [qqq_b200.qqq_gemm(A, Bs[i], C, D, s1, s2, s3s[i], ws, -1, -1, -1, 16) for i in range(ncopy)], ncopy)
        x = torch.randn(M, K, device=dev, dtype=torch.float16)

        def both():
            for i in range(ncopy):
                q, s = ops.dynamic_quant(x)
                qqq_b200.qqq_gemm(q, Bs[i], C, D, s, s2, s3s[i], ws, -1, -1, -1, 16)

        t2 = replay_us(both, ncopy)
        print(f"decode g128 M={M} K={K:5d} N={N:5d}: GEMM {t:6.2f} us ({by / t * 1e-3:6.0f} GB/s = {by / t * 1e-3 / HBM:4.2f} of HBM)"
              f"   act-quant + GEMM {t2:6.2f} us", flush=True)


def main_ncu():
    gen = torch.Generator(device=dev).manual_seed(0)
    xs = [torch.randn(M, K, device=dev, dtype=torch.float16) for (M, K) in AQ]
    gs = [gemm_args(M, K, N, 1, gen) for (M, K, N) in DEC]
    for _ in range(2):  # pass 0 is skipped by ncu (-s 7), pass 1 is captured (-c 7)
        for x in xs:
            ops.dynamic_quant(x)
        for (A, Bs, s3s, C, D, s1, s2, ws) in gs:
            qqq_b200.qqq_gemm(A, Bs[0], C, D, s1, s2, s3s[0], ws, -1, -1, -1, 16)
        torch.cuda.synchronize()


if __name__ == "__main__":
    (main_ncu if sys.argv[1:] == ["ncu"] else main_time)()
