"""Generate tests/golden/rtn_*.npz by running the REFERENCE's own `Quantizer` / `quantize`
(QQQ/gptq/quant.py:5-158) the way `GPTQ.fasterquant` drives it (QQQ/gptq/gptq.py:83-217) minus the Hessian
update (round to nearest), in this container.

Run once, here (needs /root/reference; the GPU box does not have it):   python tests/golden/gen_rtn_golden.py
The committed .npz files pin `qqq_b200.model.rtn_quantize_weight` (tests/test_model_harness.py).
"""
import importlib.util
import os

import numpy as np
import torch

REF = os.environ.get("QQQ_REFERENCE_DIR", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference_quant():
    spec = importlib.util.spec_from_file_location("ref_quant", os.path.join(REF, "QQQ/gptq/quant.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def reference_rtn(mod, W, groupsize):
    """Per column block: find_params on the block (gptq.py:140-145), quantize every column with it (:160-168);
    then the 8-bit per-channel `quantizer_extra` on the fake-quantized weight (:204-217)."""
    N, K = W.shape
    qz = mod.Quantizer()
    qz.configure(4, perchannel=True, sym=True, groupsize=groupsize, mse=False)
    Wf = W.clone().float()
    Q = torch.zeros_like(Wf)
    if groupsize == -1:
        qz.find_params(Wf, weight=True)
        scales, zeros = [qz.scale], [qz.zero]
        for i in range(K):
            Q[:, i] = mod.quantize(Wf[:, i].unsqueeze(1), qz.scale, qz.zero, qz.maxq, qz.sym, qz.groupsize).flatten()
    else:
        scales, zeros = [], []
        for g0 in range(0, K, groupsize):
            qz.find_params(Wf[:, g0:g0 + groupsize], weight=True)
            scales.append(qz.scale)
            zeros.append(qz.zero)
            for i in range(g0, g0 + groupsize):
                Q[:, i] = mod.quantize(Wf[:, i].unsqueeze(1), qz.scale, qz.zero, qz.maxq, qz.sym, qz.groupsize).flatten()
    scale, zero = torch.cat(scales, dim=1), torch.cat(zeros, dim=1)
    s_extra = None
    if groupsize != -1:
        qe = mod.Quantizer()
        qe.configure(bits=8, perchannel=True, groupsize=-1, sym=True, mse=False)
        qe.find_params(Q.clone(), weight=True)
        s_extra = qe.scale
    return Q, scale, zero, s_extra


def main():
    mod = load_reference_quant()
    for (N, K, gs, seed) in [(64, 256, -1, 11), (64, 256, 128, 12), (128, 384, 128, 13), (32, 128, -1, 14)]:
        g = torch.Generator().manual_seed(seed)
        W = torch.randn(N, K, generator=g) * 0.02
        W[0] = 0.0                      # all-zero row / groups: the reference substitutes the range [-1, 1]
        W[1] = W[1].abs()               # non-negative row: xmin stays 0 in the per-group grid
        W[2, :7] *= 30.0                # outliers
        Q, scale, zero, s_extra = reference_rtn(mod, W, gs)
        out = dict(W=W.numpy(), Q=Q.numpy(), scale=scale.numpy(), zero=zero.numpy())
        if s_extra is not None:
            out["s_extra"] = s_extra.numpy()
            # the flow of GPTQ.fasterquant (gptq.py:191-215): the layer keeps Q in ITS dtype (fp16) and the 8-bit
            # quantizer_extra runs on that fp16 tensor; find_params promotes to fp32 (quant.py:70-72)
            Q16, _, _, _ = reference_rtn(mod, W.half().float(), gs)  # an fp16 layer: W itself is fp16
            qe = mod.Quantizer()
            qe.configure(bits=8, perchannel=True, groupsize=-1, sym=True, mse=False)
            qe.find_params(Q16.half().clone(), weight=True)
            out["s_extra_fp16_layer"] = qe.scale.float().numpy()
            assert qe.scale.dtype == torch.float32
        name = f"rtn_N{N}_K{K}_g{'pc' if gs == -1 else gs}.npz"
        np.savez_compressed(os.path.join(HERE, name), **out)
        print("wrote", name, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
