"""Drop-in at the reference's own call site: the REFERENCE's `QuantLinear` (QQQ/gptq/qlinear/qlinear_marlin.py, loaded
from /root/reference where it lies, never copied) imports `qqq_gemm` from `QQQ._CUDA`; `qqq_b200.ops.install_as_qqq_cuda()`
puts this library there.  The reference module's pack() and forward() then run unchanged: its forward hands our
`qqq_gemm` the twelve positional arguments with the reference's own buffers (B, reduce_buffer, s_channel, s_group,
workspace), which must pass our argument checks and give the oracle's result.

CPU: the launch itself needs a GPU, so the C-ABI call is replaced by the oracle; the argument checks are the product's.
Skipped where the reference tree is absent (the GPU box)."""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest
import torch

from oracle import qqq_oracle as O

REF = os.environ.get("QQQ_REFERENCE_DIR", "/root/reference")
REF_FILE = os.path.join(REF, "QQQ/gptq/qlinear/qlinear_marlin.py")
pytestmark = pytest.mark.skipif(not os.path.exists(REF_FILE), reason="reference tree not present")


@pytest.fixture
def reference_module(monkeypatch):
    from qqq_b200 import ops

    calls = []

    def launch_on_oracle(A, B, C, D, s1, s2, s3, workspace, thread_k=-1, thread_n=-1, sms=-1, max_par=8):
        m, n, k, gs = ops.check_gemm_args(A, B, C, D, s1, s2, s3, workspace, max_par)  # the product's checks
        calls.append((m, n, k, gs, thread_k, thread_n, sms, max_par))
        D.copy_(torch.from_numpy(O.qqq_gemm_oracle(A.numpy(), B.numpy(), s1.numpy(), s2.numpy(),
                                                   s3.numpy() if s3.numel() else None)))

    monkeypatch.setattr(ops, "qqq_gemm", launch_on_oracle)
    monkeypatch.setitem(sys.modules, "QQQ", types.ModuleType("QQQ"))
    monkeypatch.delitem(sys.modules, "QQQ._CUDA", raising=False)
    ops.install_as_qqq_cuda()
    monkeypatch.setattr(torch.cuda, "get_device_capability", lambda *a, **k: (10, 0))
    spec = importlib.util.spec_from_file_location("ref_qlinear_marlin_dropin", REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    yield mod, calls
    sys.modules.pop("QQQ._CUDA", None)


@pytest.mark.parametrize("K,N,gs,M", [(256, 256, -1, 5), (512, 128, 128, 33), (128, 64, -1, 1)])
def test_reference_quantlinear_runs_on_our_qqq_gemm(reference_module, K, N, gs, M):
    mod, calls = reference_module
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from gen_pack_golden import fake_quant_problem

    Wfq, scales, s_extra = fake_quant_problem(K, N, gs, seed=K + N)
    lin = torch.nn.Linear(K, N, bias=True).half()
    lin.weight.data = Wfq
    lin.bias.data = torch.linspace(-1, 1, N).half()
    ref_ql = mod.QuantLinear(4, gs, K, N, bias=True)
    ref_ql.pack(lin, scales, s_extra)
    x = (torch.randn(M, K, generator=torch.Generator().manual_seed(3))).half()
    y = ref_ql(x.reshape(1, M, K))
    assert y.shape == (1, M, N) and len(calls) == 1
    assert calls[0] == (M, N, K, gs, -1, -1, -1, 16)  # mul()'s defaults, max_par from the module
    # expected: oracle on the reference module's own buffers and its own (CPU-evaluated) activation quant
    A8, s1 = ref_ql.dynamic_quant(x)
    want = O.qqq_gemm_oracle(A8.numpy(), ref_ql.B.numpy(), s1.numpy(), ref_ql.s_channel.numpy(),
                             ref_ql.s_group.numpy() if ref_ql.s_group.numel() else None)
    want = (torch.from_numpy(want) + ref_ql.bias).numpy()
    assert np.array_equal(y.reshape(M, N).numpy().view(np.uint16), want.view(np.uint16))
    # and the reference module is interchangeable with ours: same buffers after pack()
    import qqq_b200

    ours = qqq_b200.QuantLinear(4, gs, K, N, bias=True)
    ours.pack(lin, scales, s_extra)
    for name in ("B", "s_channel", "s_group", "bias"):
        assert torch.equal(getattr(ours, name), getattr(ref_ql, name)), name
    assert ours.workspace.shape == ref_ql.workspace.shape and ours.reduce_buffer.shape == ref_ql.reduce_buffer.shape
    sd_ref = {k for k, v in ref_ql.state_dict().items()}
    sd_ours = {k for k, v in ours.state_dict().items()}
    assert sd_ref == sd_ours


def test_shim_is_positional_only_like_pybind(reference_module):
    mod, _ = reference_module
    import QQQ._CUDA as shim  # noqa: N811

    with pytest.raises(TypeError):
        shim.qqq_gemm(A=None)
    with pytest.raises(TypeError):
        shim.qqq_gemm(*([None] * 8))  # the C++ defaults are invisible to Python: all twelve are required


CTOR_CASES = [
    dict(bits=4, group_size=-1, infeatures=256, outfeatures=256, bias=False),
    dict(bits=4, group_size=128, infeatures=512, outfeatures=128, bias=True),
    dict(bits=4, group_size=256, infeatures=256, outfeatures=64, bias=False),   # group_size == infeatures: per-channel format
    dict(bits=4, group_size=-1, infeatures=192, outfeatures=256, bias=False),   # (64, 256) tile rule
    dict(bits=4, group_size=-1, infeatures=100, outfeatures=256, bias=False),   # unsupported shape
    dict(bits=4, group_size=-1, infeatures=256, outfeatures=96, bias=False),    # unsupported shape
    dict(bits=8, group_size=-1, infeatures=256, outfeatures=256, bias=False),   # only 4 bits
    dict(bits=4, group_size=64, infeatures=256, outfeatures=256, bias=False),   # only -1 / 128 / infeatures
    dict(bits=4, group_size=-1, infeatures=256, outfeatures=256, bias=False, trainable=True),
]


@pytest.mark.parametrize("kw", CTOR_CASES, ids=[str(i) for i in range(len(CTOR_CASES))])
def test_constructor_parity_with_the_reference_module(reference_module, kw):
    """Same acceptance rule, same exception types and texts, same buffers (names, shapes, dtypes, persistence), same maxq
    (QQQ/gptq/qlinear/qlinear_marlin.py:51-139)."""
    import qqq_b200

    mod, _ = reference_module

    def build(cls):
        try:
            return cls(**kw), None
        except Exception as e:  # noqa: BLE001
            return None, e

    ref, ref_err = build(mod.QuantLinear)
    ours, our_err = build(qqq_b200.QuantLinear)
    if ref_err is not None:
        assert our_err is not None, f"reference raises {ref_err!r}, ours accepts"
        assert type(our_err) is type(ref_err) and str(our_err) == str(ref_err)
        return
    assert our_err is None, f"ours raises {our_err!r}, reference accepts"
    rb, ob = dict(ref.named_buffers()), dict(ours.named_buffers())
    assert set(rb) == set(ob)
    for k in rb:
        assert rb[k].shape == ob[k].shape and rb[k].dtype == ob[k].dtype, k
    assert set(ref.state_dict()) == set(ours.state_dict())
    for attr in ("infeatures", "outfeatures", "group_size", "bits", "maxq", "max_par", "tile"):
        assert getattr(ref, attr) == getattr(ours, attr), attr
    # scale dtypes stay pinned through .half() / .to(dtype) on both (qlinear_marlin.py:141-145)
    assert ours.half().s_channel.dtype == ref.half().s_channel.dtype == torch.float32


def test_pack_requires_s_extra_for_per_group_on_both(reference_module):
    import qqq_b200

    mod, _ = reference_module
    lin = torch.nn.Linear(256, 128, bias=False).half()
    scales = torch.full((128, 2), 0.01)
    for cls in (mod.QuantLinear, qqq_b200.QuantLinear):
        ql = cls(4, 128, 256, 128, bias=False)
        with pytest.raises(AssertionError, match="s_extra is needed"):
            ql.pack(lin, scales)


@pytest.mark.parametrize("gs", [-1, 128])
def test_pack_of_extreme_and_clamped_values_matches_the_reference(reference_module, gs):
    """Weights beyond the grid (clamped by pack), exact half-way cases (round-half-even) and negative zero."""
    import qqq_b200

    mod, _ = reference_module
    K, N = 256, 128
    g = torch.Generator().manual_seed(77)
    scales = (torch.rand(N, 1 if gs == -1 else K // 128, generator=g) * 0.02 + 0.005)
    grid = torch.randint(-12, 20, (N, K), generator=g).float() + torch.tensor([0.0, 0.5, -0.5, 0.0])[torch.randint(0, 4, (N, K), generator=g)]
    s_rep = scales if gs == -1 else scales.repeat_interleave(128, dim=1)
    lin = torch.nn.Linear(K, N, bias=False).half()
    lin.weight.data = (grid * s_rep).half()
    lin.weight.data[0, :4] = torch.tensor([-0.0, 0.0, 65504.0, -65504.0], dtype=torch.half)
    s_extra = None if gs == -1 else (lin.weight.data.float().abs().amax(1).clamp_min(1e-6) / 127.0).reshape(1, N)
    ref = mod.QuantLinear(4, gs, K, N, bias=False)
    ours = qqq_b200.QuantLinear(4, gs, K, N, bias=False)
    ref.pack(lin, scales, s_extra)
    ours.pack(lin, scales, s_extra)
    for name in ("B", "s_channel", "s_group"):
        a, b = getattr(ref, name), getattr(ours, name)
        assert a.shape == b.shape and torch.equal(a.view(torch.int32) if a.dtype == torch.float32 else a.view(torch.int16) if a.dtype == torch.half else a,
                                                    b.view(torch.int32) if b.dtype == torch.float32 else b.view(torch.int16) if b.dtype == torch.half else b), name


@pytest.mark.parametrize("gs", [-1, 128])
def test_permutation_attributes_match_the_reference_module(reference_module, gs):
    """`_get_perms()`, `_perm`, `_scale_perm`, `_scale_perm_single`, `wf`, `thread_config` (qlinear_marlin.py:66,134,139,147-176):
    same values and types as the reference instance (ours come from a closed form, not from the reference's loops)."""
    import qqq_b200

    mod, _ = reference_module
    a, b = mod.QuantLinear(4, gs, 256, 256, False), qqq_b200.QuantLinear(4, gs, 256, 256, False)
    assert torch.equal(a._perm, b._perm) and a._perm.dtype == b._perm.dtype
    assert a._scale_perm == b._scale_perm and a._scale_perm_single == b._scale_perm_single
    assert isinstance(b._scale_perm, list) and isinstance(b._scale_perm_single, list)
    assert a.thread_config == b.thread_config and torch.equal(a.wf, b.wf)
    pa, pb = a._get_perms(), b._get_perms()
    assert torch.equal(pa[0], pb[0]) and pa[1:] == pb[1:]
