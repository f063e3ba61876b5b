"""`QuantLinear` — the quantized operator of the hot path, mirroring the reference module
QQQ/gptq/qlinear/qlinear_marlin.py:48-288 (ctor arguments, buffer names/shapes/dtypes, state-dict keys,
`pack()`, `dynamic_quant()`, `forward()`), and `mul()` (:28-45).

What differs underneath:
  * `pack()` builds the packed tensor from the closed-form nibble map (vectorised torch on any device)
    instead of the reference's 1024-entry gather + numpy loop; output is bit-identical (tests/test_pack.py).
  * `dynamic_quant()` is one CUDA kernel instead of five eager ops; bit-identical.
  * `forward()` calls the sm_100a tcgen05 kernel through the C ABI.  There is no CPU path.

`QQQLinear` is an alias (that name is used by the vLLM integration of QQQ; the reference repo's class is
`QuantLinear`).
"""
from __future__ import annotations

from logging import getLogger
from typing import NamedTuple

import torch
import torch.nn as nn

from . import ops

logger = getLogger(__name__)

# (thread_k, thread_n) divisibility rule of the reference, qlinear_marlin.py:66-77
_THREAD_CONFIG = ((64, 256), (128, 128), (128, 64), (64, 128))


def mul(A, B, C, D, s1, s2, s3, workspace, thread_k=-1, thread_n=-1, sms=-1, max_par=16):
    """INT8xINT4 multiply; positional pass-through to `qqq_gemm` (reference `mul`, qlinear_marlin.py:28-45).

    @A int8 (m,k) row-major;  @B int32 packed weights, see `QuantLinear.pack`;  @C int32 (max_par*64, n)
    reduce buffer;  @D half (m,n) out;  @s1 fp32 (m,1) per-token scales;  @s2 fp32 (1,n) per-channel scales;
    @s3 half (k/groupsize, n) per-group scales, empty for per-channel;  @workspace int32 >= n/128*max_par zeros.
    """
    ops.qqq_gemm(A, B, C, D, s1, s2, s3, workspace, thread_k, thread_n, sms, max_par)


def _scale_perm_positions(device):
    """Index vectors such that permuted[..., p] = natural[..., idx[p]] for the two scale permutations
    (reference: qlinear_marlin.py:170-175)."""
    p64 = torch.arange(64, device=device)
    scale_perm = (p64 // 8) + 8 * (p64 % 8)  # position 8i+j holds channel i+8j
    p32 = torch.arange(32, device=device)
    jj = p32 % 8
    scale_perm_single = 2 * (p32 // 8) + 8 * (jj // 2) + (jj % 2)  # position 8i+jj holds 2i+[0,1,8,9,..][jj]
    return scale_perm, scale_perm_single


def pack_int4_weights(w: torch.Tensor, per_group: bool) -> torch.Tensor:
    """w: integer tensor [K, N] (per-channel: signed values in [-8,7]; per-group: unsigned in [0,15]).
    Returns int32 [K/16, 2N] in the reference layout:

    row kt = k//16; word (nb, lane, j) at column nb*128 + lane*4 + j holds, for r in 0..3 and blk in 0..1,
    element k = 16kt + 4(lane%4) + r, n = 64nb + 16j + lane//4 + 8blk in nibble p where
      per-channel: p = 2r + (1 - blk)          per-group: p = [0,4,1,5,2,6,3,7][4blk + r].
    """
    K, N = w.shape
    assert K % 16 == 0 and N % 64 == 0
    v = (w.to(torch.int64) & 0xF)
    # k = 16kt + 4kq + r ; n = 64nb + 16j + 8blk + c
    v = v.reshape(K // 16, 4, 4, N // 64, 4, 2, 8)  # (kt, kq, r, nb, j, blk, c)
    v = v.permute(0, 3, 6, 1, 4, 5, 2)  # (kt, nb, c, kq, j, blk, r)   lane = 4c + kq
    if per_group:
        shift = torch.tensor([[0, 4, 1, 5], [2, 6, 3, 7]], device=w.device, dtype=torch.int64) * 4  # [blk][r]
    else:
        shift = torch.tensor([[4, 12, 20, 28], [0, 8, 16, 24]], device=w.device, dtype=torch.int64)  # 4*(2r+1-blk)
    word = (v << shift).sum(dim=(-1, -2))  # (kt, nb, c, kq, j)
    word = word.reshape(K // 16, 2 * N)
    word = torch.where(word >= 2**31, word - 2**32, word)
    return word.to(torch.int32)


class QuantizedActivation(NamedTuple):
    """An activation that already went through `dynamic_quant` (reference: qlinear_marlin.py:265-268): what the fused
    tensor-parallel exchange (tp.ScatterRowParallelQuantLinear) hands to the linears of the next block, and what linears
    that share an input can share.  `QuantLinear.forward` accepts it in place of the fp16 tensor."""

    q: torch.Tensor   # int8 [M, K]
    s1: torch.Tensor  # fp32 [M, 1]
    lead: tuple = ()  # leading dims of the original activation (output is reshaped to lead + (N,)); () -> (M,)


class _ActQuantCache:
    """Last (input -> int8, scale) pair of `dynamic_quant`, so that linears that consume the SAME tensor (q/k/v,
    gate/up in the reference's 7-module layer structure) quantise it once.  Opt-in (`set_act_quant_cache(True)`):
    a hit needs the same storage, offset, shape, strides, dtype AND tensor version, and the cached input is kept alive
    so its address cannot be recycled — but a writer that bypasses autograd's version counter (a raw-pointer kernel
    writing into the tensor between two forwards) is invisible to it."""

    enabled = False
    _key = None
    _ref = None
    _val = None

    @classmethod
    def key(cls, x):
        # inference tensors (torch.inference_mode) have no version counter — reading `_version` raises — and may still be
        # written in place inside inference mode: they are never cached (every linear quantises for itself, as the reference)
        if x.is_inference():
            return None
        return (x.untyped_storage().data_ptr(), x.storage_offset(), tuple(x.shape), tuple(x.stride()), x.dtype, x._version,
                x.device)

    @classmethod
    def get(cls, x):
        k = cls.key(x)
        return cls._val if k is not None and cls._key == k else None

    @classmethod
    def put(cls, x, val):
        k = cls.key(x)
        if k is None:
            cls.clear()
        else:
            cls._key, cls._ref, cls._val = k, x, val

    @classmethod
    def clear(cls):
        cls._key = cls._ref = cls._val = None


def set_act_quant_cache(enabled: bool) -> None:
    """Reuse the int8 activations across consecutive QuantLinear.forward calls on the same input tensor (bit-identical
    outputs, 3 of the 7 activation-quant launches of a decoder layer instead of 7).  Off by default."""
    _ActQuantCache.enabled = bool(enabled)
    _ActQuantCache.clear()


class QuantLinear(nn.Module):
    QUANT_TYPE = "marlin"

    def __init__(self, bits, group_size, infeatures, outfeatures, bias, trainable=False, **kwargs):
        super().__init__()
        if torch.version.hip:
            raise ValueError("qqq_b200 targets NVIDIA B200 (sm_100a) only.")
        if torch.cuda.is_available() and torch.cuda.get_device_capability()[0] != 10:
            raise ValueError(
                f"qqq_b200 kernels are built for compute capability 10.x (B200); found "
                f"{torch.cuda.get_device_capability()}."
            )
        if not any(infeatures % tk == 0 and outfeatures % tn == 0 for tk, tn in _THREAD_CONFIG):
            raise ValueError("Not supported `infeatures`: {} and `outfeatures`: {}.".format(infeatures, outfeatures))
        if bits not in [4]:
            raise NotImplementedError("Only 4 bits are supported.")
        if group_size not in [-1, 128] and group_size != infeatures:
            raise ValueError("Only group_size -1 and 128 are supported.")
        if trainable:
            raise NotImplementedError("Marlin does not support train.")

        self.thread_config = list(_THREAD_CONFIG)  # attribute parity with the reference module (qlinear_marlin.py:66)
        self.infeatures = infeatures
        self.outfeatures = outfeatures
        self.group_size = group_size if group_size != -1 else infeatures
        if self.infeatures % self.group_size != 0:
            raise ValueError("`infeatures` must be divisible by `group_size`.")
        self.bits = bits
        self.tile = 16
        self.maxq = 2**self.bits - 1 if self.group_size != self.infeatures else 2 ** (self.bits - 1) - 1
        self.max_par = 16
        self.register_buffer("B", torch.empty((infeatures // 16, outfeatures * 16 // 8), dtype=torch.int32))
        self.register_buffer("s_channel", torch.empty((1, outfeatures), dtype=torch.float32))
        if self.group_size != self.infeatures:
            self.register_buffer(
                "s_group", torch.empty((infeatures // self.group_size, outfeatures), dtype=torch.half)
            )
        else:
            self.register_buffer("s_group", torch.tensor([], dtype=torch.half))
        # lock words (zero in / zero out) and split-K scratch, same contract as the reference (include/qqq_b200.h)
        self.register_buffer("workspace", torch.zeros(outfeatures // 128 * 16, dtype=torch.int32), persistent=False)
        self.register_buffer(
            "reduce_buffer", torch.zeros((self.max_par * 16 * 4, outfeatures), dtype=torch.int), persistent=False
        )
        self.wf = torch.tensor(list(range(0, 32, 4)), dtype=torch.int32).unsqueeze(0)  # qlinear_marlin.py:134 (unused there too)
        if bias:
            self.register_buffer("bias", torch.zeros((outfeatures), dtype=torch.half))
        else:
            self.bias = None
        self._perm, self._scale_perm, self._scale_perm_single = self._get_perms()

    def _get_perms(self):
        """The reference's three permutations (qlinear_marlin.py:147-176), same types (int64 tensor of 1024, two lists),
        in closed form: entry e = 32*lane + 8*j + p of `perm` addresses, inside a 16x64 weight tile flattened as
        16*64 -> [k][n], the element k = 4*(lane%4) + r, n = 16*j + lane//4 + 8*blk, where (blk, r) = divmod(t, 4) and
        t = [4,0,5,1,6,2,7,3][p] (per-channel) or [0,2,4,6,1,3,5,7][p] (per-group) is the nibble order inside a word.
        `pack()` does not go through them (pack_int4_weights builds the words directly); they are exposed for code that
        reads them off the reference module."""
        e = torch.arange(1024, dtype=torch.int64)
        lane, j, p_ = e // 32, (e // 8) % 4, e % 8
        order = [0, 2, 4, 6, 1, 3, 5, 7] if self.per_group else [4, 0, 5, 1, 6, 2, 7, 3]
        t = torch.tensor(order, dtype=torch.int64)[p_]
        perm = 16 * (4 * (lane % 4) + t % 4) + lane // 4 + 8 * (t // 4) + 256 * j
        sp64, sp32 = _scale_perm_positions("cpu")
        return perm, sp64.tolist(), sp32.tolist()

    def _apply(self, fn):
        # keep scale dtypes pinned across .half()/.to(dtype) exactly like qlinear_marlin.py:141-145
        super()._apply(fn)
        self.s_group = self.s_group.to(torch.half)
        self.s_channel = self.s_channel.to(torch.float32)
        return self

    def post_init(self):
        pass

    @property
    def per_group(self) -> bool:
        return self.group_size != self.infeatures

    def pack(self, linear, scales, s_extra=None):
        """Pack a fake-quantized linear layer (reference `pack`, qlinear_marlin.py:181-262).
        @linear: fake-quantized `torch.nn.Linear` (fp16 weights)
        @scales: quantization scales of shape `(outfeatures, groups)` (transposed inside, like the reference)
        @s_extra: per-channel int8 scales of shape `(1, outfeatures)`, required for per-group
        """
        if self.per_group:
            assert s_extra is not None, "s_extra is needed"
        if linear.weight.dtype != torch.half:
            logger.warning(
                f"The dtype of weights is {linear.weight.dtype}, while the W4A8 GEMM's output is torch.half; "
                "results are only correct if they do not overflow torch.half."
            )
        K, N = self.infeatures, self.outfeatures
        dev = linear.weight.device
        w = linear.weight.data.t()  # [K, N]
        s = scales.t()  # [groups, N]
        sp64, sp32 = _scale_perm_positions(dev)
        if self.per_group:
            G = K // self.group_size
            s_rep = s.reshape(G, 1, N).expand(G, self.group_size, N).reshape(K, N)
            q = torch.round(w / s_rep).int() + (self.maxq + 1) // 2
            q = torch.clamp(q, 0, self.maxq)
            s_extra = s_extra.reshape(1, -1).to(dtype=torch.float32)
            s_g = (s.reshape(G, N) / s_extra).to(dtype=torch.half)
            s_group = s_g.reshape(G, N // 64, 64)[:, :, sp64].reshape(G, N)
            s_channel = s_extra.reshape(N // 32, 32)[:, sp32].reshape(1, N)
        else:
            q = torch.clamp(torch.round(w / s).int(), -self.maxq, self.maxq)
            # /16: the kernel leaves the nibble in the high half of the int8 (W8 = 16*w4)
            s_channel = (s / (2 ** (8 - self.bits))).reshape(N // 32, 32)[:, sp32].to(dtype=torch.float32).reshape(1, N)
            s_group = None
        self.B[:, :] = pack_int4_weights(q, self.per_group).to(self.B.device)
        if self.per_group:
            self.s_group[:, :] = s_group.to(self.s_group.device)
            self.s_channel[:, :] = s_channel.to(self.s_channel.device)
        else:
            self.s_group = torch.tensor([], dtype=torch.half, device=self.s_channel.device)
            self.s_channel[:, :] = s_channel.to(self.s_channel.device)
        if linear.bias is not None:
            if self.bias is not None:
                self.bias[:] = linear.bias.data.to(self.bias.device).to(torch.half)
            else:
                self.bias = linear.bias.clone().to(torch.half)

    # activation int8 quantization (reference: qlinear_marlin.py:265-268)
    def dynamic_quant(self, x: torch.Tensor):
        return ops.dynamic_quant(x)

    def forward(self, A):
        if isinstance(A, QuantizedActivation):
            quant_A, s1 = A.q, A.s1
            out_shape = (tuple(A.lead) if A.lead else (quant_A.shape[0],)) + (self.outfeatures,)
        else:
            out_shape = A.shape[:-1] + (self.outfeatures,)
            A = A.reshape(-1, A.shape[-1]).half()
            cached = _ActQuantCache.get(A) if _ActQuantCache.enabled else None
            if cached is not None:
                quant_A, s1 = cached
            else:
                quant_A, s1 = self.dynamic_quant(A)
                if _ActQuantCache.enabled:
                    _ActQuantCache.put(A, (quant_A, s1))
        D = torch.empty(quant_A.shape[0], self.outfeatures, dtype=torch.float16, device=quant_A.device)
        if self.bias is not None and quant_A.shape[0] > 0:
            # the reference adds the bias with an eager op after the GEMM (qlinear_marlin.py:286-288); here it rides in the
            # GEMM's epilogue (same fp16 add on the rounded output, same bits)
            ops.qqq_gemm_bias(quant_A, self.B, self.reduce_buffer, D, s1, self.s_channel, self.s_group, self.workspace,
                              self.bias, self.max_par)
        else:
            mul(quant_A, self.B, self.reduce_buffer, D, s1, self.s_channel, self.s_group, self.workspace,
                max_par=self.max_par)
        return D.reshape(out_shape)


def merge_quant_linears(mods) -> QuantLinear:
    """Merge QuantLinears that consume the SAME input (q/k/v, gate/up) into one module whose output is the
    concatenation of theirs along the feature dim: one activation quant and one GEMM instead of len(mods) of each.

    Pure concatenation of the packed tensors along N — no repack: every 64-channel block of `B` is 128
    consecutive words of a row and the scale permutations act inside 32/64-channel blocks, so column blocks of
    different linears can simply be laid side by side.  Outputs are bit-identical to the separate modules
    (each output column depends only on its own weight column and the shared, identically quantised input).
    """
    mods = list(mods)
    m0 = mods[0]
    assert all(m.infeatures == m0.infeatures and m.group_size == m0.group_size and m.bits == m0.bits for m in mods)
    assert all(m.outfeatures % 64 == 0 for m in mods)
    has_bias = any(m.bias is not None for m in mods)
    gs = m0.group_size if m0.per_group else -1
    out = QuantLinear(m0.bits, gs, m0.infeatures, sum(m.outfeatures for m in mods), bias=has_bias)
    dev = m0.B.device
    out.B = torch.cat([m.B for m in mods], dim=1).contiguous()
    out.s_channel = torch.cat([m.s_channel for m in mods], dim=1).contiguous()
    out.s_group = torch.cat([m.s_group for m in mods], dim=1).contiguous() if m0.per_group else m0.s_group
    if has_bias:
        out.bias = torch.cat([m.bias if m.bias is not None else torch.zeros(m.outfeatures, dtype=torch.half, device=dev)
                              for m in mods]).contiguous()
    out.workspace = torch.zeros(max(out.outfeatures // 128 * 16, 16), dtype=torch.int32, device=dev)
    out.reduce_buffer = torch.zeros((out.max_par * 64, out.outfeatures), dtype=torch.int32, device=dev)
    out.split_sizes = [m.outfeatures for m in mods]
    return out


QQQLinear = QuantLinear

__all__ = ["QuantLinear", "QQQLinear", "QuantizedActivation", "mul", "pack_int4_weights", "merge_quant_linears", "set_act_quant_cache"]
