"""Generate tests/golden/kernel_*.npz by running the UNMODIFIED reference CUDA kernel (oracle/_ref, built by
oracle/build_ref.py from /root/reference/csrc) on a B200, on seeded inputs from oracle.make_problem().

Run on the GPU box:   python tests/golden/gen_kernel_golden.py  [--out gpurun_out/golden]
It also bit-compares the CPU oracle (oracle/qqq_oracle.py) with the reference kernel on every case and on a
few larger shapes that are not stored, printing a PIN line per case.  The stored fixtures are what
tests/test_oracle_golden.py (CPU) and tests/test_gemm_parity.py (GPU) check against.
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import build_ref, qqq_oracle as O  # noqa: E402

# (M, K, N, group_size, seed, store)
CASES = [
    (1, 128, 64, -1, 11, True),
    (17, 256, 256, -1, 12, True),
    (33, 512, 128, 128, 13, True),
    (64, 384, 192, 128, 14, True),
    (7, 1024, 320, -1, 15, True),
    (130, 256, 256, 128, 16, True),
    (16, 4096, 4096, -1, 21, False),
    (16, 4096, 4096, 128, 22, False),
    (1000, 2048, 1024, -1, 23, False),
    (1024, 2048, 1024, 128, 24, False),
    (1, 8192, 2176, 128, 25, False),
]


def run_ref(mod, p, M, K, N, max_par=16):
    dev = "cuda:0"
    A = torch.from_numpy(p["A8"]).to(dev)
    B = torch.from_numpy(p["B"]).to(dev)
    s1 = torch.from_numpy(p["s1"]).to(dev)
    s2 = torch.from_numpy(p["s2"]).to(dev)
    s3 = torch.from_numpy(p["s3"]).to(dev)
    C = torch.zeros((max_par * 64, N), dtype=torch.int32, device=dev)
    D = torch.empty((M, N), dtype=torch.float16, device=dev)
    ws = torch.zeros(N // 128 * max_par + 16, dtype=torch.int32, device=dev)
    mod.qqq_gemm(A, B, C, D, s1, s2, s3, ws, -1, -1, -1, max_par)
    torch.cuda.synchronize()
    assert int(ws.abs().sum().item()) == 0, "reference left workspace non-zero"
    return D.cpu().numpy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "golden"))
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    mod = build_ref.load()
    if mod is None:
        print("oracle/_ref is missing: run `python oracle/build_ref.py` where /root/reference exists")
        return 2
    print("device:", torch.cuda.get_device_name(0))
    rc = 0
    for (M, K, N, gs, seed, store) in CASES:
        t0 = time.time()
        p = O.make_problem(M, K, N, gs, seed)
        D_ref = run_ref(mod, p, M, K, N)
        D_or = O.qqq_gemm_oracle(p["A8"], p["B"], p["s1"], p["s2"], p["s3"])
        same = np.array_equal(D_ref.view(np.uint16), D_or.view(np.uint16))
        nbad = int((D_ref.view(np.uint16) != D_or.view(np.uint16)).sum())
        print(f"PIN M={M} K={K} N={N} g={gs} seed={seed}: oracle==reference_kernel bitwise: {same} "
              f"(mismatch {nbad}/{D_ref.size}) max|D|={np.abs(D_ref.astype(np.float32)).max():.3f} "
              f"[{time.time() - t0:.1f}s]", flush=True)
        if not same:
            rc = 1
            d = np.abs(D_ref.astype(np.float32) - D_or.astype(np.float32))
            print("   max abs diff", d.max(), "at", np.unravel_index(d.argmax(), d.shape))
        if store:
            name = os.path.join(args.out, f"kernel_M{M}_K{K}_N{N}_g{gs if gs > 0 else 'pc'}.npz")
            np.savez_compressed(name, M=M, K=K, N=N, group_size=gs, seed=seed, A8=p["A8"], B=p["B"], s1=p["s1"],
                                s2=p["s2"], s3=p["s3"], D=D_ref)
    return rc


if __name__ == "__main__":
    sys.exit(main())
