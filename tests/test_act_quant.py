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


def test_exact_ties_round_half_even_and_saturation():
    """amax = 127 -> scale exactly 1.0, so k + 0.5 inputs are exact ties of the quotient (torch.round is half-even);
    amax = 254 -> scale 2.0: odd inputs are ties."""
    import qqq_b200

    K = 256
    x = torch.zeros((3, K), device="cuda", dtype=torch.float16)
    ties = torch.arange(-127, 127, device="cuda", dtype=torch.float32) + 0.5
    x[0, : ties.numel()] = ties.half()
    x[0, -1] = 127.0
    x[1, : ties.numel()] = (2 * ties).half()
    x[1, -1] = -254.0
    x[2, :128] = torch.arange(128, device="cuda").half() * 2.0 - 127.0  # plain integers, scale 1
    x[2, -1] = -127.0
    q, s = qqq_b200.dynamic_quant(x)
    q_ref, s_ref = reference_dynamic_quant(x)
    assert torch.equal(s, s_ref) and torch.equal(q, q_ref)
    assert float(s[0]) == 1.0 and int(q[0, 0]) == -126 and int(q[0, 127]) == 0 and int(q[0, 128]) == 2


def test_every_finite_fp16_value_against_eager():
    """All 63488 finite fp16 inputs, under 24 different row scales (rows keep only |x| <= a cap, which sets amax)."""
    import qqq_b200

    bits16 = torch.arange(0, 65536, dtype=torch.int32)
    bits16 = bits16[((bits16 >> 10) & 31) != 31].to(torch.int16)
    allv = bits16.view(torch.float16).cuda()  # 63488 values, a multiple of 8
    caps = [6e-8, 1e-6, 3.1e-5, 6.2e-5, 1e-3, 0.0123, 0.5, 1.0, 1.27, 3.0, 17.0, 127.0, 128.0, 254.0, 300.0, 1000.0,
            1016.0, 4000.0, 8128.0, 20000.0, 32512.0, 40000.0, 65024.0, 65504.0]
    rows = [torch.where(allv.float().abs() <= c, allv, torch.zeros_like(allv)) for c in caps]
    x = torch.stack(rows).contiguous()
    q, s = qqq_b200.dynamic_quant(x)
    q_ref, s_ref = reference_dynamic_quant(x)
    assert torch.equal(s, s_ref)
    assert torch.equal(q, q_ref)
