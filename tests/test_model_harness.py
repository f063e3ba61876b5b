"""Model-level harness: module swap on a stock HF decoder, RTN quantizers in the reference
Quantizer's conventions, checkpoint format, q/k/v + gate/up fusion.

CPU tests: the product has no CPU path, so the two C-ABI entry points `QuantLinear.forward` calls are replaced — in the
test only — by the oracle (`oracle_backend` fixture); what is under test here is the host logic around them.
The same scenarios run on the real kernels in tests/test_zz_model_gpu.py.
"""
import glob
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import qqq_oracle as O

transformers = pytest.importorskip("transformers")

import qqq_b200  # noqa: E402
from qqq_b200 import model as qmodel  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def tiny_config(model_type="llama", layers=2):
    kw = dict(vocab_size=128, hidden_size=256, intermediate_size=512, num_hidden_layers=layers, num_attention_heads=4,
              num_key_value_heads=2, max_position_embeddings=64, tie_word_embeddings=False)
    if model_type == "llama":
        return transformers.LlamaConfig(**kw)
    return transformers.Qwen2Config(**kw)


def tiny_model(model_type="llama", seed=0, dtype=torch.float16):
    torch.manual_seed(seed)
    prev = torch.get_default_dtype()
    torch.set_default_dtype(dtype)  # like from_pretrained(torch_dtype=...): rotary inv_freq stays fp32
    try:
        m = transformers.AutoModelForCausalLM.from_config(tiny_config(model_type))
    finally:
        torch.set_default_dtype(prev)
    if model_type == "qwen2":  # from_config zero-initialises biases: give q/k/v real ones
        for n, p in m.named_parameters():
            if n.endswith("proj.bias"):
                p.data.normal_(0, 0.05)
    return m.eval()


@pytest.fixture
def oracle_backend(monkeypatch):
    """QuantLinear.forward on the CPU oracle (test infrastructure only)."""
    from qqq_b200 import ops

    def dq(x):
        q, s = O.dynamic_quant(x.detach().cpu().numpy(), cuda_semantics=True)
        return torch.from_numpy(q), torch.from_numpy(s)

    def gemm(A, B, C, D, s1, s2, s3, workspace, thread_k=-1, thread_n=-1, sms=-1, max_par=8):
        ref = O.qqq_gemm_oracle(A.numpy(), B.numpy(), s1.numpy(), s2.numpy(), s3.numpy() if s3.numel() else None)
        D.copy_(torch.from_numpy(ref))

    def gemm_bias(A, B, C, D, s1, s2, s3, workspace, bias, max_par=16, sms=-1):
        ops.qqq_gemm(A, B, C, D, s1, s2, s3, workspace, -1, -1, sms, max_par)  # looked up at call time: tests wrap it
        D += bias  # fp16 add on the rounded output, like the kernel's epilogue and the reference's eager op

    monkeypatch.setattr(ops, "dynamic_quant", dq)
    monkeypatch.setattr(ops, "qqq_gemm", gemm)
    monkeypatch.setattr(ops, "qqq_gemm_bias", gemm_bias)


# ---------------------------------------------------------------------------------------------------------------
# RTN quantizer vs the reference Quantizer (golden fixtures from tests/golden/gen_rtn_golden.py)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "rtn_*.npz"))), ids=os.path.basename)
def test_rtn_quantizer_matches_reference_quantizer(path):
    g = np.load(path)
    gs = -1 if path.endswith("gpc.npz") else 128
    W_fq, scale, zero, s_extra = qmodel.rtn_quantize_weight(torch.from_numpy(g["W"]), gs)
    assert np.array_equal(scale.numpy(), g["scale"])
    assert np.array_equal(zero.numpy(), g["zero"])
    assert np.array_equal(W_fq.numpy(), g["Q"])
    if gs != -1:
        assert np.array_equal(s_extra.numpy(), g["s_extra"])
    else:
        assert s_extra is None


def test_rtn_fixture_set_is_present():
    assert len(glob.glob(os.path.join(GOLDEN, "rtn_*.npz"))) >= 4


# ---------------------------------------------------------------------------------------------------------------
# module swap
# ---------------------------------------------------------------------------------------------------------------
def test_find_layers_and_decoder_names():
    m = tiny_model()
    lin = qmodel.find_layers(m)
    assert "lm_head" in lin and len(lin) == 2 * 7 + 1
    names = list(qmodel.decoder_linear_names(m))
    assert len(names) == 14 and "lm_head" not in names
    assert "model.layers.1.mlp.down_proj" in names and "model.layers.0.self_attn.k_proj" in names


@pytest.mark.parametrize("gs", [-1, 128])
def test_make_quant_swaps_by_name_and_keeps_the_rest(gs):
    m = tiny_model()
    names = list(qmodel.decoder_linear_names(m))
    qmodel.make_quant(m, names, 4, gs)
    ql = qmodel.find_layers(m, [qqq_b200.QuantLinear])
    assert sorted(ql) == sorted(names)
    assert type(m.lm_head) is nn.Linear and type(m.model.embed_tokens) is nn.Embedding
    k = ql["model.layers.0.self_attn.k_proj"]
    assert (k.infeatures, k.outfeatures) == (256, 128) and k.bias is None
    assert k.B.shape == (16, 256) and k.s_channel.shape == (1, 128)
    assert k.s_group.shape == ((2, 128) if gs == 128 else (0,))
    # checkpoint keys are the reference's: <name>.B / .s_channel / .s_group (no workspace / reduce_buffer)
    keys = set(qmodel.quantized_state_dict(m))
    assert "model.layers.0.self_attn.k_proj.B" in keys and "model.layers.0.self_attn.k_proj.s_channel" in keys
    assert ("model.layers.0.self_attn.k_proj.s_group" in keys) == (gs == 128)
    assert not any(k.endswith("workspace") or k.endswith("reduce_buffer") or k.endswith("proj.weight") for k in keys)


def test_make_quant_is_a_noop_on_a_quantlinear_and_rejects_other_modules():
    ql = qqq_b200.QuantLinear(4, -1, 128, 64, bias=False)
    qmodel.make_quant(ql, ["x"], 4, -1)
    m = tiny_model()
    with pytest.raises(NotImplementedError):
        qmodel.make_quant(m, ["model.embed_tokens"], 4, -1)


def test_unsupported_architecture_raises_like_the_reference_registry():
    with pytest.raises(NotImplementedError):
        qmodel.get_model_architecture(transformers.GPT2Config())


# ---------------------------------------------------------------------------------------------------------------
# whole-model scenarios on the oracle backend
# ---------------------------------------------------------------------------------------------------------------
def _logits(m, ids):
    with torch.no_grad():
        return m(input_ids=ids).logits.float()


@pytest.mark.parametrize("model_type,gs", [("llama", -1), ("llama", 128), ("qwen2", 128)])
def test_quantized_model_tracks_the_fake_quant_model(oracle_backend, model_type, gs):
    """W4A8 logits vs the fake-quantized fp16 model (same weights, fp16 activations): only the int8 activation
    rounding and the int8 re-quantisation of per-group weights separate them."""
    m = tiny_model(model_type)
    ids = torch.randint(0, 128, (2, 12), generator=torch.Generator().manual_seed(1))
    quantizers = qmodel.rtn_quantizers(m, gs)
    ref = _logits(m, ids)  # weights are fake-quantized in place now
    qmodel.pack_model(m, quantizers, 4, gs)
    assert len(qmodel.find_layers(m, [qqq_b200.QuantLinear])) == 14
    got = _logits(m, ids)
    err = (got - ref).abs().max().item()
    assert err <= 0.05 * ref.abs().max().item() + 0.02, err
    if model_type == "qwen2":
        assert m.model.layers[0].self_attn.q_proj.bias is not None


@pytest.mark.parametrize("model_type,gs", [("llama", -1), ("qwen2", 128)])
def test_fused_qkv_gate_up_is_bit_identical(oracle_backend, model_type, gs):
    m = qmodel.quantize_model_rtn(tiny_model(model_type), gs)
    ids = torch.randint(0, 128, (1, 9), generator=torch.Generator().manual_seed(2))
    before = _logits(m, ids)
    n_before = len(qmodel.find_layers(m, [qqq_b200.QuantLinear]))
    assert qmodel.fuse_qkv_gate_up(m) == 4  # 2 layers x (qkv, gate_up)
    ql = qmodel.find_layers(m, [qqq_b200.QuantLinear])
    assert n_before == 14 and len(ql) == 2 * 4  # qkv, o, gate_up, down per layer
    assert isinstance(m.model.layers[0].self_attn.k_proj, qmodel.FusedProjection)
    after = _logits(m, ids)
    assert torch.equal(before, after)
    # a second forward with another input must not reuse the cached merged output
    ids2 = torch.randint(0, 128, (1, 9), generator=torch.Generator().manual_seed(3))
    assert not torch.equal(_logits(m, ids2), after)
    assert torch.equal(_logits(m, ids), after)


def test_fused_projection_runs_one_gemm_per_group(oracle_backend, monkeypatch):
    from qqq_b200 import ops

    m = qmodel.quantize_model_rtn(tiny_model("llama"), -1)
    qmodel.fuse_qkv_gate_up(m)
    calls = []
    inner = ops.qqq_gemm
    monkeypatch.setattr(ops, "qqq_gemm", lambda *a, **k: (calls.append(a[3].shape[-1]), inner(*a, **k))[1])
    _logits(m, torch.randint(0, 128, (1, 5)))
    assert len(calls) == 2 * 4 and sorted(set(calls)) == [256, 512, 1024]  # qkv 256+128+128, o/down 256, gate_up 1024


@pytest.mark.parametrize("gs", [-1, 128])
def test_checkpoint_round_trip(oracle_backend, gs, tmp_path):
    m = qmodel.quantize_model_rtn(tiny_model("llama"), gs)
    assert m.config.quantization_config == {"group_size": gs, "quant_method": "qqq", "wbits": 4}
    sd = qmodel.quantized_state_dict(m)
    assert all(v.numel() > 0 for v in sd.values())
    torch.save(sd, tmp_path / "model.pt")
    m2 = qmodel.build_quantized_model(tiny_config("llama"), m.config.quantization_config)
    assert len(qmodel.find_layers(m2, [qqq_b200.QuantLinear])) == 14
    qmodel.load_quantized_state_dict(m2, torch.load(tmp_path / "model.pt"))
    ids = torch.randint(0, 128, (1, 7), generator=torch.Generator().manual_seed(5))
    assert torch.equal(_logits(m, ids), _logits(m2.eval(), ids))
    bad = dict(sd)
    bad["model.layers.0.mlp.up_proj.qweight"] = torch.zeros(1)
    with pytest.raises(RuntimeError):
        qmodel.load_quantized_state_dict(m2, bad)
    short = {k: v for k, v in sd.items() if not k.endswith("down_proj.B")}
    with pytest.raises(RuntimeError):
        qmodel.load_quantized_state_dict(m2, short)


def test_product_forward_has_no_cpu_path():
    """Without the oracle fixture the swapped model must refuse to run on the CPU (no silent fallback)."""
    m = qmodel.quantize_model_rtn(tiny_model("llama"), -1)
    with pytest.raises(RuntimeError):
        _logits(m, torch.randint(0, 128, (1, 4)))


def test_act_quant_cache_reuses_the_shared_input_only(oracle_backend, monkeypatch):
    """Opt-in cache: q/k/v (and gate/up) quantise their shared input once; any change of the input (new tensor, in-place
    write, other view) misses; outputs are bit-identical."""
    from qqq_b200 import ops

    m = qmodel.quantize_model_rtn(tiny_model("llama"), 128)
    ids = torch.randint(0, 128, (1, 6), generator=torch.Generator().manual_seed(8))
    ref = _logits(m, ids)
    calls = []
    inner = ops.dynamic_quant
    monkeypatch.setattr(ops, "dynamic_quant", lambda x: (calls.append(1), inner(x))[1])
    _logits(m, ids)
    assert len(calls) == 14
    qqq_b200.set_act_quant_cache(True)
    try:
        calls.clear()
        assert torch.equal(_logits(m, ids), ref)
        assert len(calls) == 2 * 4  # per layer: (q,k,v) once, o, (gate,up) once, down
        ql = m.model.layers[0].self_attn.q_proj
        x = torch.randn(3, 256).half()
        calls.clear()
        y0 = ql(x)
        ql(x)
        assert len(calls) == 1
        x.mul_(2.0)  # in-place write bumps the version: miss
        y1 = ql(x)
        assert len(calls) == 2 and not torch.equal(y0, y1)
        ql(x[1:])  # another view: miss
        ql(x.clone())  # another storage: miss
        assert len(calls) == 4
    finally:
        qqq_b200.set_act_quant_cache(False)
    calls.clear()
    ql(x), ql(x)
    assert len(calls) == 2


def test_share_scratch_keeps_results_and_releases_memory(oracle_backend):
    m = qmodel.quantize_model_rtn(tiny_model("llama"), 128)
    ids = torch.randint(0, 128, (1, 5), generator=torch.Generator().manual_seed(9))
    ref = _logits(m, ids)
    keys = set(qmodel.quantized_state_dict(m))
    freed = qmodel.share_scratch(m)
    qls = list(qmodel.find_layers(m, [qqq_b200.QuantLinear]).values())
    assert freed > 0
    assert len({q.workspace.data_ptr() for q in qls}) == 1 and len({q.reduce_buffer.data_ptr() for q in qls}) == 1
    for q in qls:  # N is still read from C.size(1); the lock array covers the widest layer
        assert q.reduce_buffer.shape == (q.max_par * 64, q.outfeatures) and q.reduce_buffer.is_contiguous()
        assert q.workspace.numel() >= q.outfeatures // 128 * q.max_par
    assert torch.equal(_logits(m, ids), ref)
    assert set(qmodel.quantized_state_dict(m)) == keys  # scratch stays out of checkpoints


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "rtn_*g128.npz"))), ids=os.path.basename)
def test_rtn_quantizers_take_s_extra_from_the_stored_fp16_weight_like_the_reference(path):
    """GPTQ.fasterquant stores Q in the layer's dtype and runs the 8-bit quantizer_extra on that tensor (gptq.py:191-215):
    the scale comes from the fp16-rounded weights, evaluated in fp32 (the Quantizer promotes, quant.py:70-72)."""
    g = np.load(path)

    class Holder(nn.Module):
        def __init__(self, W):
            super().__init__()
            self.layers = nn.ModuleList([nn.Linear(W.shape[1], W.shape[0], bias=False)])
            self.layers[0].weight.data = torch.from_numpy(W).half()

    m = Holder(g["W"])  # an fp16 layer holding W
    q = qmodel.rtn_quantizers(m, 128, names=["layers.0"])["layers.0"]
    s_extra = q[3]
    assert s_extra.dtype == torch.float32
    assert np.array_equal(s_extra.numpy(), g["s_extra_fp16_layer"])


@pytest.mark.parametrize("cache", [False, True])
def test_fused_model_and_quant_cache_under_inference_mode(oracle_backend, cache):
    """tensors created under torch.inference_mode() have no version counter (`x._version` raises); the merged-GEMM
    cache and the activation-quant cache must work there (serving stacks run under inference_mode) and still never serve a
    stale entry."""
    m = qmodel.quantize_model_rtn(tiny_model("llama"), -1)
    ids = torch.randint(0, 128, (1, 6), generator=torch.Generator().manual_seed(11))
    ids2 = torch.randint(0, 128, (1, 6), generator=torch.Generator().manual_seed(12))
    with torch.no_grad():
        want, want2 = m(input_ids=ids, use_cache=False).logits.clone(), m(input_ids=ids2, use_cache=False).logits.clone()
    qmodel.fuse_qkv_gate_up(m)
    qqq_b200.set_act_quant_cache(cache)
    try:
        with torch.inference_mode():
            got = m(input_ids=ids, use_cache=False).logits
            got2 = m(input_ids=ids2, use_cache=False).logits
            got3 = m(input_ids=ids, use_cache=False).logits
    finally:
        qqq_b200.set_act_quant_cache(False)
    assert torch.equal(got, want) and torch.equal(got2, want2) and torch.equal(got3, want)


def test_shared_gemm_never_serves_an_entry_left_by_an_interrupted_forward(oracle_backend):
    """a forward interrupted between q_proj and v_proj leaves the merged output cached; the next forward starts
    at slot 0 again, which always recomputes — also when the new input sits at the same address with the same version."""
    ql = [qqq_b200.QuantLinear(4, -1, 128, n, bias=False) for n in (64, 64, 128)]
    for i, q in enumerate(ql):
        g = torch.Generator().manual_seed(i)
        q.B = torch.randint(-2**31, 2**31 - 1, q.B.shape, dtype=torch.int32, generator=g)
        q.s_channel = torch.full_like(q.s_channel, 1e-3)
    shared = qmodel._SharedGemm(qqq_b200.merge_quant_linears(ql))
    x = torch.randn(3, 128).half()
    first = shared.slice(x, 0).clone()  # "q_proj" only: the forward is interrupted here
    x.data.copy_(torch.randn(3, 128).half())  # same storage, same address; `.data` does not bump the version
    again = shared.slice(x, 0)
    assert not torch.equal(first, again)
    assert torch.equal(again, ql[0](x)) and torch.equal(shared.slice(x, 1), ql[1](x)) and torch.equal(shared.slice(x, 2), ql[2](x))
    assert shared._out is None and shared._ref is None  # every consumer served: references dropped
