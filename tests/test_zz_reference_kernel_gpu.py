"""GPU: this library against the UNMODIFIED reference CUDA kernel (oracle/_ref/qqq_ref_cuda.so, built by
oracle/build_ref.py from /root/reference/csrc where it lies; the prebuilt .so travels to the GPU box), live on the same
device tensors — the strongest parity statement available: identical packed-int4 / int8 inputs in, identical fp16 bits
out.  Shapes go beyond the stored kernel_*.npz fixtures (ragged M, K not a multiple of 128, both modes, M above the
reference's 64-row blocks).  Skipped when the reference build did not travel."""
import numpy as np
import pytest
import torch

from oracle import build_ref
from oracle import qqq_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref_mod():
    try:
        mod = build_ref.load()
    except Exception as e:  # the checker's plumbing, not the product
        pytest.skip(f"reference kernel extension unavailable: {e!r}")
    if mod is None:
        pytest.skip("oracle/_ref/qqq_ref_cuda.so not present (built where /root/reference exists)")
    return mod


@pytest.mark.parametrize("M,K,N,gs", [(1, 4096, 4096, -1), (3, 256, 128, 128), (16, 2048, 1024, -1), (48, 1024, 512, 128),
                                      (65, 192, 256, -1), (100, 640, 384, 128), (128, 4096, 1024, -1), (257, 1024, 1152, 128),
                                      (512, 2048, 2048, -1), (1000, 512, 256, 128), (1024, 4096, 4096, -1),
                                      (1024, 4096, 1024, 128)])
def test_same_bits_as_the_reference_kernel(ref_mod, M, K, N, gs):
    import qqq_b200

    dev = "cuda:0"
    p = O.make_problem(M, K, N, gs, seed=1000 + M)
    t = {k: torch.from_numpy(np.ascontiguousarray(p[k])).to(dev) for k in ("A8", "B", "s1", "s2", "s3")}
    max_par = 16

    def run(fn):
        C = torch.zeros((max_par * 64, N), dtype=torch.int32, device=dev)
        ws = torch.zeros(N // 128 * max_par + 16, dtype=torch.int32, device=dev)
        D = torch.full((M, N), float("nan"), dtype=torch.float16, device=dev)
        fn(t["A8"], t["B"], C, D, t["s1"], t["s2"], t["s3"], ws, -1, -1, -1, max_par)
        torch.cuda.synchronize()
        assert int(ws.abs().sum()) == 0  # both return the lock words zeroed
        return D.view(torch.int16)

    ours = run(qqq_b200.qqq_gemm)
    # The product is held to the CPU oracle first (exact integer model, itself pinned on reference-kernel outputs).
    want = torch.from_numpy(O.qqq_gemm_oracle(p["A8"], p["B"], p["s1"], p["s2"], p["s3"]).view(np.int16)).to(dev)
    assert torch.equal(ours, want), f"{int((ours != want).sum())} of {M * N} fp16 values differ from the oracle"
    # Then to the reference kernel itself.  The UNMODIFIED reference build is not race-free on a B200: at
    # (1024, 4096, 1024, g128) about 28 % of its launches return whole 64-row x 256-column blocks that differ from the
    # exact result (its lock-serialised cross-CTA reduce, csrc/qqq_gemm.cu:213-237,606-676; 0 of 350 launches of this
    # library differ — profiles/r02/call_c/diag_*.log).  So the reference gets several launches: the product must equal
    # the bits the reference produces when it produces the exact result, and every reference launch that differs from
    # the product must ALSO differ from the oracle (i.e. be the reference's own error, not ours).
    agree = 0
    for _ in range(6):
        ref = run(ref_mod.qqq_gemm)
        if torch.equal(ref, ours):
            agree += 1
        else:
            assert not torch.equal(ref, want)
    assert agree >= 1, "the reference kernel never reproduced the product's (= the oracle's) bits in 6 launches"
