"""GPU: qqq_gemm_acc (the GEMM without epilogue scales, raw int32 accumulators — building block of the bit-exact
tensor-parallel mode) against the oracle's integer model, incl. shapes whose tiles are split along K (fix-up path).
Named test_zzz_*: this kernel variant has not run on hardware yet (written after round 1's GPU budget was spent), so it
goes last — a fault in it must not take the CUDA context away from the other GPU tests."""
import numpy as np
import pytest
import torch

from oracle import qqq_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,K,N,gs", [(5, 512, 256, -1), (70, 1024, 384, 128), (16, 4096, 4096, -1), (300, 2048, 1024, 128),
                                      (257, 768, 640, -1)])
@pytest.mark.parametrize("sms", [-1, 37])
def test_raw_accumulators_are_exact(M, K, N, gs, sms):
    from qqq_b200 import ops

    dev = "cuda:0"
    p = O.make_problem(M, K, N, gs, seed=77 + M)
    W8 = O.weights_int8(p["B"], p["s3"] if gs != -1 else None)
    want = p["A8"].astype(np.int64) @ W8.astype(np.int64)
    A = torch.from_numpy(p["A8"]).to(dev)
    B = torch.from_numpy(p["B"]).to(dev)
    s3 = torch.from_numpy(p["s3"]).to(dev)
    C = torch.zeros(16 * 64, N, dtype=torch.int32, device=dev)
    ws = torch.zeros(max(N // 128 * 16, 16), dtype=torch.int32, device=dev)
    D = torch.full((M, N), -(2**31), dtype=torch.int32, device=dev)
    ops.qqq_gemm_acc(A, B, C, D, s3, ws, 16, sms)
    torch.cuda.synchronize()
    assert np.array_equal(D.cpu().numpy().astype(np.int64), want)
    assert int(ws.abs().sum()) == 0
