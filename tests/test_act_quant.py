"""GPU: the fused activation-quant kernel vs the reference's eager formula (qlinear_marlin.py:265-268) evaluated
by torch ON THE GPU (that is the reference's execution), and vs the oracle's CUDA-semantics restatement."""
import numpy as np
import pytest
import torch

from oracle import qqq_oracle as O

pytestmark = pytest.mark.gpu


def reference_dynamic_quant(x):
    # restated from QQQ/gptq/qlinear/qlinear_marlin.py:265-268
    quant_scale = x.abs().max(dim=-1, keepdim=True)[0].div(127.0).to(torch.float32)
    x = (x / quant_scale).round().clamp(-128, 127).to(torch.int8)
    return x, quant_scale


@pytest.mark.parametrize("M,K", [(1, 128), (1, 4096), (7, 8192), (33, 11008), (257, 4096), (16, 21760), (3, 40960), (1024, 1024)])
def test_matches_torch_eager_bit_exact(M, K):
    import qqq_b200

    g = torch.Generator(device="cuda").manual_seed(M * 1000 + K)
    x = (torch.randn((M, K), device="cuda", generator=g) * 3).half()
    x[0, K // 2] = 200.0
    if M > 2:
        x[2] *= 1e-3  # tiny row
    q, s = qqq_b200.dynamic_quant(x)
    q_ref, s_ref = reference_dynamic_quant(x)
    torch.cuda.synchronize()
    assert torch.equal(s, s_ref)
    assert torch.equal(q, q_ref)
    q_or, s_or = O.dynamic_quant(x.cpu().numpy(), cuda_semantics=True)
    assert np.array_equal(s.cpu().numpy(), s_or)
    assert np.array_equal(q.cpu().numpy(), q_or)


def test_many_random_rows_scale_rounding():
    """The fp16(fp32(amax) * fp32(1/127)) rounding matters only near fp16 ties: sweep many amax values."""
    import qqq_b200

    M, K = 4096, 128
    x = torch.zeros((M, K), device="cuda", dtype=torch.float16)
    amax = torch.linspace(0.01, 600.0, M, device="cuda").half()
    x[:, 0] = amax
    x[:, 1] = -amax / 3
    q, s = qqq_b200.dynamic_quant(x)
    q_ref, s_ref = reference_dynamic_quant(x)
    assert torch.equal(s, s_ref) and torch.equal(q, q_ref)


def test_zero_row_does_not_poison():
    import qqq_b200

    x = torch.zeros((4, 256), device="cuda", dtype=torch.float16)
    x[1] = 1.0
    q, s = qqq_b200.dynamic_quant(x)
    assert float(s[0]) == 0.0 and int(q[0].abs().sum()) == 0
    assert int(q[1, 0]) == 127
