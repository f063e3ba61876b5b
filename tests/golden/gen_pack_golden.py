"""Generate tests/golden/pack_*.npz by running the REFERENCE's own QuantLinear.pack()/dynamic_quant()
(QQQ/gptq/qlinear/qlinear_marlin.py:181-268) in this container.

Run once, here (needs /root/reference; the GPU box does not have it):   python tests/golden/gen_pack_golden.py
The reference module imports `QQQ._CUDA.qqq_gemm` (qlinear_marlin.py:22) and asks for a CUDA device
capability (:60); both are stubbed because only the CPU-side pack()/dynamic_quant() are exercised.
The committed .npz files are what `tests/test_oracle_golden.py` and `tests/test_pack.py` check against.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("QQQ_REFERENCE_DIR", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference_qlinear():
    qqq = types.ModuleType("QQQ")
    cuda = types.ModuleType("QQQ._CUDA")
    cuda.qqq_gemm = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("stub"))
    sys.modules["QQQ"] = qqq
    sys.modules["QQQ._CUDA"] = cuda
    torch.cuda.get_device_capability = lambda *a, **k: (10, 0)
    path = os.path.join(REF, "QQQ/gptq/qlinear/qlinear_marlin.py")
    spec = importlib.util.spec_from_file_location("ref_qlinear_marlin", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def fake_quant_problem(K, N, group_size, seed):
    """Fake-quantised fp16 Linear + scales in the conventions the reference's GPTQ driver hands to pack()
    (QQQ/gptq/quant.py:85-93 for the grids, QQQ/gptq/gptq.py:204-217 for s_extra)."""
    g = torch.Generator().manual_seed(seed)
    W = (torch.randn(N, K, generator=g) * 0.02).float()  # nn.Linear weight [out, in]
    if group_size == -1:
        s = (W.abs().amax(dim=1, keepdim=True) / 7.0).clamp_min(1e-8).half().float()  # [N,1]
        q = torch.clamp(torch.round(W / s), -7, 7)
        Wfq = (q * s).half()
        scales = s.half()  # [N, 1]  (pack() transposes)
        return Wfq, scales, None
    G = K // group_size
    Wg = W.reshape(N, G, group_size)
    s_g = (2.0 * Wg.abs().amax(dim=2) / 15.0).clamp_min(1e-8).half().float()  # [N, G]
    q = torch.clamp(torch.round(Wg / s_g[:, :, None]) + 8, 0, 15)
    Wfq = ((q - 8) * s_g[:, :, None]).reshape(N, K)
    s_extra = (Wfq.abs().amax(dim=1) / 127.0).clamp_min(1e-12).float()  # [N]
    return Wfq.half(), s_g.half(), s_extra.reshape(1, N)


def main():
    mod = load_reference_qlinear()
    cases = [(256, 256, -1, 1), (256, 256, 128, 2), (512, 128, 128, 3), (128, 64, -1, 4), (384, 192, 128, 5),
             (1024, 320, -1, 6)]
    for (K, N, gs, seed) in cases:
        Wfq, scales, s_extra = fake_quant_problem(K, N, gs, seed)
        lin = torch.nn.Linear(K, N, bias=True).half()
        lin.weight.data = Wfq.clone()
        lin.bias.data = (torch.randn(N, generator=torch.Generator().manual_seed(seed + 100)) * 0.1).half()
        ql = mod.QuantLinear(4, gs, K, N, bias=True)
        ql.pack(lin, scales, s_extra)
        x = (torch.randn(9, K, generator=torch.Generator().manual_seed(seed + 200)) * 1.5).half()
        x[3, 5] = 31.0
        qa, sa = ql.dynamic_quant(x)
        out = dict(
            K=K, N=N, group_size=gs,
            weight_fq=Wfq.numpy(), scales=scales.numpy(), bias=lin.bias.data.numpy(),
            B=ql.B.numpy(), s_channel=ql.s_channel.numpy(), s_group=ql.s_group.numpy(),
            packed_bias=ql.bias.numpy(),
            x=x.numpy(), quant_A=qa.numpy(), s1=sa.numpy(),
            perm=ql._perm.numpy(), scale_perm=np.array(ql._scale_perm), scale_perm_single=np.array(ql._scale_perm_single),
            workspace_shape=np.array(ql.workspace.shape), reduce_buffer_shape=np.array(ql.reduce_buffer.shape),
        )
        if s_extra is not None:
            out["s_extra"] = s_extra.numpy()
        name = os.path.join(HERE, f"pack_K{K}_N{N}_g{gs if gs > 0 else 'pc'}.npz")
        np.savez_compressed(name, **out)
        print("wrote", name, {k: getattr(v, "shape", v) for k, v in out.items() if k in ("B", "s_channel", "s_group")})


if __name__ == "__main__":
    main()
