"""CPU tests of the host-side mirror of the reference operator: QuantLinear ctor/buffers/pack()/state dict
against the reference's own pack() outputs (tests/golden/pack_*.npz)."""
import glob
import os

import numpy as np
import pytest
import torch

import qqq_b200
from qqq_b200 import QuantLinear, pack_int4_weights
from oracle import qqq_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PACK = sorted(glob.glob(os.path.join(GOLDEN, "pack_*.npz")))


def _module_from_golden(g):
    K, N, gs = int(g["K"]), int(g["N"]), int(g["group_size"])
    lin = torch.nn.Linear(K, N, bias=True).half()
    lin.weight.data = torch.from_numpy(g["weight_fq"]).clone()
    lin.bias.data = torch.from_numpy(g["bias"]).clone()
    ql = QuantLinear(4, gs, K, N, bias=True)
    s_extra = torch.from_numpy(g["s_extra"]) if "s_extra" in g.files else None
    ql.pack(lin, torch.from_numpy(g["scales"]), s_extra)
    return ql


@pytest.mark.parametrize("path", PACK, ids=os.path.basename)
def test_pack_bit_identical_to_reference(path):
    g = np.load(path)
    ql = _module_from_golden(g)
    assert np.array_equal(ql.B.numpy(), g["B"])
    assert np.array_equal(ql.s_channel.numpy(), g["s_channel"])
    assert np.array_equal(ql.s_group.numpy().view(np.uint16), g["s_group"].view(np.uint16))
    assert np.array_equal(ql.bias.numpy().view(np.uint16), g["packed_bias"].view(np.uint16))
    assert tuple(ql.workspace.shape) == tuple(g["workspace_shape"])
    assert tuple(ql.reduce_buffer.shape) == tuple(g["reduce_buffer_shape"])
    assert ql.B.dtype == torch.int32 and ql.s_channel.dtype == torch.float32 and ql.s_group.dtype == torch.float16
    assert int(ql.workspace.abs().sum()) == 0 and int(ql.reduce_buffer.abs().sum()) == 0
    # the reference instance's permutation attributes (qlinear_marlin.py:139,147-176), stored with the fixture
    assert np.array_equal(ql._perm.numpy(), g["perm"]) and ql._scale_perm == g["scale_perm"].tolist()
    assert ql._scale_perm_single == g["scale_perm_single"].tolist()


@pytest.mark.gpu
@pytest.mark.parametrize("path", PACK, ids=os.path.basename)
def test_pack_on_the_device_is_bit_identical_to_reference(path):
    """`pack()` runs on CUDA tensors (the reference packs on the CPU with numpy, qlinear_marlin.py:181-262)
    and must produce the reference's packed buffers bit for bit; the packed module then runs where it was packed."""
    g = np.load(path)
    dev = torch.device("cuda:0")
    K, N, gs = int(g["K"]), int(g["N"]), int(g["group_size"])
    lin = torch.nn.Linear(K, N, bias=True).half().to(dev)
    lin.weight.data = torch.from_numpy(g["weight_fq"]).to(dev)
    lin.bias.data = torch.from_numpy(g["bias"]).to(dev)
    ql = QuantLinear(4, gs, K, N, bias=True).to(dev)
    s_extra = torch.from_numpy(g["s_extra"]).to(dev) if "s_extra" in g.files else None
    ql.pack(lin, torch.from_numpy(g["scales"]).to(dev), s_extra)
    assert ql.B.is_cuda and ql.s_channel.is_cuda
    assert np.array_equal(ql.B.cpu().numpy(), g["B"])
    assert np.array_equal(ql.s_channel.cpu().numpy(), g["s_channel"])
    assert np.array_equal(ql.s_group.cpu().numpy().view(np.uint16), g["s_group"].view(np.uint16))
    assert np.array_equal(ql.bias.cpu().numpy().view(np.uint16), g["packed_bias"].view(np.uint16))
    # ... and the module packed on the device computes what the oracle computes from the reference's packed buffers
    x = torch.randn(9, K, device=dev).half()
    y = ql(x)
    A8, s1 = O.dynamic_quant(x.cpu().numpy(), cuda_semantics=True)
    s3 = g["s_group"] if g["s_group"].size else None
    want = torch.from_numpy(O.qqq_gemm_oracle(A8, g["B"], s1, g["s_channel"], s3)) + torch.from_numpy(g["packed_bias"])
    assert np.array_equal(y.cpu().numpy().view(np.uint16), want.numpy().view(np.uint16))


@pytest.mark.parametrize("pg", [False, True])
def test_vectorised_packer_equals_oracle_packer(pg):
    rng = np.random.default_rng(7)
    K, N = 160, 192
    w = rng.integers(0, 16, (K, N)) if pg else rng.integers(-8, 8, (K, N))
    assert np.array_equal(pack_int4_weights(torch.from_numpy(w), pg).numpy(), O.pack_B(w, pg))


def test_state_dict_keys_and_roundtrip():
    g = np.load(PACK[0])
    ql = _module_from_golden(g)
    sd = ql.state_dict()
    assert set(sd.keys()) == {"B", "s_channel", "s_group", "bias"}  # workspace / reduce_buffer are non-persistent
    K, N, gs = int(g["K"]), int(g["N"]), int(g["group_size"])
    ql2 = QuantLinear(4, gs, K, N, bias=True)
    ql2.load_state_dict(sd)
    assert torch.equal(ql2.B, ql.B) and torch.equal(ql2.s_channel, ql.s_channel)


def test_apply_pins_scale_dtypes():
    ql = QuantLinear(4, 128, 256, 128, bias=False)
    ql.half()
    assert ql.s_channel.dtype == torch.float32 and ql.s_group.dtype == torch.float16
    ql.float()
    assert ql.s_channel.dtype == torch.float32 and ql.s_group.dtype == torch.float16


def test_ctor_validation_matches_reference():
    with pytest.raises(ValueError, match="Not supported `infeatures`"):
        QuantLinear(4, -1, 100, 128, bias=False)
    with pytest.raises(NotImplementedError, match="Only 4 bits"):
        QuantLinear(8, -1, 128, 128, bias=False)
    with pytest.raises(ValueError, match="Only group_size -1 and 128"):
        QuantLinear(4, 64, 256, 128, bias=False)
    with pytest.raises(NotImplementedError, match="does not support train"):
        QuantLinear(4, -1, 128, 128, bias=False, trainable=True)
    ql = QuantLinear(4, 256, 256, 128, bias=False)  # group_size == infeatures is the per-channel format
    assert ql.s_group.numel() == 0 and ql.maxq == 7
    assert QuantLinear(4, 128, 256, 128, bias=False).maxq == 15
    assert qqq_b200.QQQLinear is QuantLinear


def test_forward_without_gpu_fails_loudly():
    """No CPU fallback: the product path raises instead of silently computing on the host."""
    ql = QuantLinear(4, -1, 128, 128, bias=False)
    with pytest.raises(RuntimeError, match="no CPU path"):
        ql(torch.zeros(2, 128, dtype=torch.float16))


def test_merge_quant_linears_is_pure_concatenation():
    from qqq_b200 import merge_quant_linears

    mods = []
    for path in [p for p in PACK if "K256_N256" in p]:
        mods.append(_module_from_golden(np.load(path)))
    pc = [m for m in mods if not m.per_group]
    a = pc[0]
    b = QuantLinear(4, -1, 256, 128, bias=False)
    b.B.copy_(torch.randint(-2**31, 2**31 - 1, b.B.shape, dtype=torch.int32))
    b.s_channel.copy_(torch.rand(1, 128))
    m = merge_quant_linears([a, b])
    assert m.outfeatures == 384 and m.B.shape == (16, 768)
    assert torch.equal(m.B[:, :512], a.B) and torch.equal(m.B[:, 512:], b.B)
    assert torch.equal(m.s_channel[:, :256], a.s_channel) and torch.equal(m.s_channel[:, 256:], b.s_channel)
    # the merged packed tensor decodes to the side-by-side weight matrices
    W = O.weights_int8(m.B.numpy(), None)
    assert np.array_equal(W[:, :256], O.weights_int8(a.B.numpy(), None))
    assert np.array_equal(W[:, 256:], O.weights_int8(b.B.numpy(), None))
    assert m.bias is not None and torch.equal(m.bias[:256], a.bias) and int(m.bias[256:].abs().sum()) == 0
