"""GPU: the model-level harness on the real kernels — a stock HF decoder whose linears were swapped
for QuantLinear runs through the fused act-quant kernel + tcgen05 GEMM; scenarios mirror tests/test_model_harness.py
(which runs them on the oracle).  Named test_zz_* so that it runs after the kernel parity tests."""
import numpy as np
import pytest
import torch

from oracle import qqq_oracle as O

pytestmark = pytest.mark.gpu

transformers = pytest.importorskip("transformers")

from test_model_harness import _logits, tiny_model  # noqa: E402


def _quantized_on_gpu(model_type, gs):
    import qqq_b200
    from qqq_b200 import model as qmodel

    m = tiny_model(model_type)
    quantizers = qmodel.rtn_quantizers(m, gs)
    fq = _logits(m.to("cuda:0"), torch.arange(24, device="cuda:0").reshape(2, 12) % 128)
    m = m.to("cpu")
    qmodel.pack_model(m, quantizers, 4, gs)
    m = m.to("cuda:0")
    assert len(qmodel.find_layers(m, [qqq_b200.QuantLinear])) == 14
    return m, fq


@pytest.mark.parametrize("model_type,gs", [("llama", -1), ("llama", 128), ("qwen2", 128)])
def test_quantized_model_tracks_the_fake_quant_model_on_gpu(model_type, gs):
    import qqq_b200

    before = qqq_b200.launch_count()
    m, ref = _quantized_on_gpu(model_type, gs)
    ids = torch.arange(24, device="cuda:0").reshape(2, 12) % 128
    got = _logits(m, ids)
    torch.cuda.synchronize()
    assert qqq_b200.launch_count() - before == 2 * 14  # one act-quant + one GEMM per decoder linear: the CUDA path ran
    err = (got - ref).abs().max().item()
    assert err <= 0.05 * ref.abs().max().item() + 0.02, err


def test_every_swapped_linear_is_bit_exact_vs_oracle_inside_the_model():
    """Hook each QuantLinear of the running model and check its output against the oracle on the same input."""
    import qqq_b200
    from qqq_b200 import model as qmodel

    m, _ = _quantized_on_gpu("llama", 128)
    seen = {}

    def hook(mod, args, out):
        seen[mod] = (args[0].detach().cpu().numpy(), out.detach().cpu().numpy())

    qls = qmodel.find_layers(m, [qqq_b200.QuantLinear])
    hs = [q.register_forward_hook(hook) for q in qls.values()]
    _logits(m, torch.arange(10, device="cuda:0").reshape(1, 10))
    for h in hs:
        h.remove()
    assert len(seen) == 14
    for ql, (x, y) in seen.items():
        x2 = x.reshape(-1, x.shape[-1]).astype(np.float16)
        A8, s1 = O.dynamic_quant(x2, cuda_semantics=True)
        ref = O.qqq_gemm_oracle(A8, ql.B.cpu().numpy(), s1, ql.s_channel.cpu().numpy(), ql.s_group.cpu().numpy())
        assert np.array_equal(y.reshape(ref.shape).view(np.uint16), ref.view(np.uint16))


@pytest.mark.parametrize("model_type,gs", [("llama", -1), ("qwen2", 128)])
def test_fused_qkv_gate_up_is_bit_identical_on_gpu(model_type, gs):
    import qqq_b200
    from qqq_b200 import model as qmodel

    m, _ = _quantized_on_gpu(model_type, gs)
    ids = torch.arange(9, device="cuda:0").reshape(1, 9) * 7 % 128
    before = _logits(m, ids)
    assert qmodel.fuse_qkv_gate_up(m) == 4
    l0 = qqq_b200.launch_count()
    after = _logits(m, ids)
    torch.cuda.synchronize()
    assert qqq_b200.launch_count() - l0 == 2 * 8  # qkv, o, gate_up, down per layer
    assert torch.equal(before, after)


def test_act_quant_cache_on_gpu_is_bit_identical_and_saves_launches():
    import qqq_b200

    m, _ = _quantized_on_gpu("llama", -1)
    ids = torch.arange(11, device="cuda:0").reshape(1, 11) * 5 % 128
    ref = _logits(m, ids)
    qqq_b200.set_act_quant_cache(True)
    try:
        l0 = qqq_b200.launch_count()
        got = _logits(m, ids)
        torch.cuda.synchronize()
        assert qqq_b200.launch_count() - l0 == 14 + 8  # 14 GEMMs, 4 activation quants per layer
    finally:
        qqq_b200.set_act_quant_cache(False)
    assert torch.equal(ref, got)


def test_shared_scratch_on_gpu_is_bit_identical():
    from qqq_b200 import model as qmodel

    m, _ = _quantized_on_gpu("llama", 128)
    ids = torch.arange(40, device="cuda:0").reshape(1, 40) * 3 % 128
    ref = _logits(m, ids)
    assert qmodel.share_scratch(m) > 0
    got = _logits(m, ids)
    assert torch.equal(ref, got)
    import qqq_b200

    for q in qmodel.find_layers(m, [qqq_b200.QuantLinear]).values():
        assert int(q.workspace.abs().sum()) == 0
