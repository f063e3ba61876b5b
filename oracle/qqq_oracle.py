"""CPU oracle for the QQQ W4A8 GEMM hot path.   *** TEST INFRASTRUCTURE — NOT A PRODUCT PATH ***

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module.  Nothing under qqq_b200/ imports it; the product fails loudly when its CUDA library is missing.

It restates, in plain numpy, the algorithm of the reference (paths relative to /root/reference):

  * Marlin permutations ............... QQQ/gptq/qlinear/qlinear_marlin.py:147-176
  * weight packing (layout only) ...... QQQ/gptq/qlinear/qlinear_marlin.py:181-262
  * per-token activation quant ........ QQQ/gptq/qlinear/qlinear_marlin.py:265-268
  * per-channel nibble -> int8 ........ csrc/qqq_gemm.cu:146-151 (+ use at :540-542)
  * per-group nibble -> int8 .......... csrc/qqq_gemm.cu:167-210 (+ use at :536-538)
  * int8 x int8 -> int32 .............. csrc/qqq_gemm.cu:106-117 (mma ... satfinite.s32.s8.s8.s32)
  * epilogue order and roundings ...... csrc/qqq_gemm.cu:129-143, 695-700

Parity pinning: the reference ships no tests or golden vectors for this path.  The oracle
is pinned instead against (1) the reference's own `QuantLinear.pack()` run in this container
(tests/golden/gen_pack_golden.py -> tests/golden/pack_*.npz) and (2) the reference CUDA kernel itself,
compiled unmodified for sm_100a (oracle/build_ref.py) and run on a B200
(tests/golden/gen_kernel_golden.py -> tests/golden/kernel_*.npz).
"""
from __future__ import annotations

import numpy as np

TILE = 16  # qlinear_marlin.py:91


# ----------------------------------------------------------------------------------------------------
# Permutations (qlinear_marlin.py:147-176)
# ----------------------------------------------------------------------------------------------------
def get_perms(per_group: bool):
    """Return (perm[1024], scale_perm[64], scale_perm_single[32]) exactly as `_get_perms` builds them."""
    perm = []
    for i in range(32):
        perm1 = []
        col = i // 4
        for block in (0, 1):
            for row in (4 * (i % 4), 4 * (i % 4) + 1, 4 * (i % 4) + 2, 4 * (i % 4) + 3):
                perm1.append(16 * row + col + 8 * block)
        for j in range(4):
            perm.extend(p + 256 * j for p in perm1)
    perm = np.array(perm, dtype=np.int64)
    if per_group:
        interleave = np.array([0, 2, 4, 6, 1, 3, 5, 7])  # :167
    else:
        interleave = np.array([4, 0, 5, 1, 6, 2, 7, 3])  # :165
    perm = perm.reshape(-1, 8)[:, interleave].ravel()
    scale_perm = []
    for i in range(8):
        scale_perm.extend(i + 8 * j for j in range(8))
    scale_perm_single = []
    for i in range(4):
        scale_perm_single.extend(2 * i + j for j in (0, 1, 8, 9, 16, 17, 24, 25))
    return perm, np.array(scale_perm), np.array(scale_perm_single)


# ----------------------------------------------------------------------------------------------------
# Packing (layout part of qlinear_marlin.py:228-248)
# ----------------------------------------------------------------------------------------------------
def pack_B(w: np.ndarray, per_group: bool) -> np.ndarray:
    """w: integer [K, N]; per-channel values in [-8, 7] (two's complement nibble), per-group values in
    [0, 15].  Returns the packed int32 [K/16, 2N] tensor `B`."""
    K, N = w.shape
    assert K % TILE == 0 and N % 64 == 0
    perm, _, _ = get_perms(per_group)
    t = w.reshape(K // TILE, TILE, N // TILE, TILE).transpose(0, 2, 1, 3).reshape(K // TILE, N * TILE)
    res = t.reshape(-1, perm.size)[:, perm].reshape(t.shape).astype(np.int64)
    q = np.zeros((res.shape[0], res.shape[1] // 8), dtype=np.uint32)
    for i in range(8):
        q |= ((res[:, i::8] & 0xF).astype(np.uint32)) << np.uint32(4 * i)
    return q.view(np.int32)


def permute_s_channel(s: np.ndarray) -> np.ndarray:
    """s: [N] per-output-channel fp32 scale in natural order -> `s_channel` [1, N] (perm blocks of 32)."""
    _, _, sps = get_perms(False)
    return s.reshape(-1, 32)[:, sps].reshape(1, -1).astype(np.float32)


def unpermute_s_channel(s_channel: np.ndarray) -> np.ndarray:
    _, _, sps = get_perms(False)
    inv = np.argsort(sps)
    return s_channel.reshape(-1, 32)[:, inv].reshape(-1)


def permute_s_group(s: np.ndarray) -> np.ndarray:
    """s: [G, N] fp16 group scales in natural order -> `s_group` [G, N] (perm blocks of 64)."""
    _, sp, _ = get_perms(True)
    G = s.shape[0]
    return s.reshape(-1, 64)[:, sp].reshape(G, -1).astype(np.float16)


def unpermute_s_group(s_group: np.ndarray) -> np.ndarray:
    _, sp, _ = get_perms(True)
    inv = np.argsort(sp)
    G = s_group.shape[0]
    return s_group.reshape(-1, 64)[:, inv].reshape(G, -1)


def unpack_B(B: np.ndarray, per_group: bool) -> np.ndarray:
    """Closed-form inverse of pack_B (SURVEY.md §8a nibble map).  Returns the raw nibbles [K, N] in [0,15].

    Row kt of B covers k in [16kt, 16kt+16).  Each 64-column block nb is 128 consecutive words; word
    (lane, j) at index nb*128 + lane*4 + j holds k = 16kt + 4(lane%4) + r, n = 64nb + 16j + lane//4 + 8blk
    for r in 0..3, blk in 0..1.  Nibble p (bits 4p..4p+3):
      per-channel: blk = 0 if p odd else 1, r = p // 2    (so q & 0xF0F0F0F0 is blk 0 as int8 = 16*w)
      per-group  : e = [0,2,4,6,1,3,5,7][p], blk = e // 4, r = e % 4
    """
    Bu = np.ascontiguousarray(B).view(np.uint32)
    KT, W = Bu.shape
    N = W // 2
    K = KT * 16
    out = np.zeros((K, N), dtype=np.int32)
    widx = np.arange(W)
    nb, lane, j = widx // 128, (widx % 128) // 4, widx % 4
    for p in range(8):
        nib = ((Bu >> np.uint32(4 * p)) & np.uint32(0xF)).astype(np.int32)  # [KT, W]
        if per_group:
            e = [0, 2, 4, 6, 1, 3, 5, 7][p]
            blk, r = e // 4, e % 4
        else:
            blk, r = (0 if p % 2 else 1), p // 2
        n = 64 * nb + 16 * j + lane // 4 + 8 * blk  # [W]
        kin = 4 * (lane % 4) + r  # [W]
        rows = 16 * np.arange(KT)[:, None] + kin[None, :]  # [KT, W]
        out[rows, np.broadcast_to(n[None, :], rows.shape)] = nib
    return out


# ----------------------------------------------------------------------------------------------------
# Nibble -> int8 weight, as the kernel computes it
# ----------------------------------------------------------------------------------------------------
def w8_per_channel(nib: np.ndarray) -> np.ndarray:
    """csrc/qqq_gemm.cu:146-151: the nibble is placed in the HIGH half of the int8, i.e. W8 = 16 * sext4(nib)
    (pack() divides s_channel by 16 to compensate, qlinear_marlin.py:221-226)."""
    return (((nib.astype(np.int32) << 4) & 0xFF).astype(np.uint8)).view(np.int8).astype(np.int32)


def w8_per_group(nib: np.ndarray, s_group_nat: np.ndarray, group_size: int = 128) -> np.ndarray:
    """csrc/qqq_gemm.cu:167-210.  (v - 8) is formed exactly in fp16, then ONE fp16 FMA computes
    (v-8)*s + 1152 with a single rounding; 1152 = 0x6480 has ulp 1, so the low mantissa byte is
    128 + RNE((v-8)*s); `prmt` extracts that byte and `^ 0x80` turns it into int8."""
    K, N = nib.shape
    s = np.repeat(s_group_nat.astype(np.float16).astype(np.float64), group_size, axis=0)  # [K, N]
    exact = (nib.astype(np.float64) - 8.0) * s + 1152.0  # exact in f64 (15-bit product + small add)
    h = exact.astype(np.float16)  # single RNE rounding == the FMA's rounding
    byte = (h.view(np.uint16) & 0xFF).astype(np.uint8) ^ np.uint8(0x80)
    return byte.view(np.int8).astype(np.int32)


# ----------------------------------------------------------------------------------------------------
# Activation quantisation (qlinear_marlin.py:265-268)
# ----------------------------------------------------------------------------------------------------
def dynamic_quant(x: np.ndarray, cuda_semantics: bool = True):
    """x: fp16 [M, K].  scale = (absmax / 127) rounded to fp16, then cast to fp32; x / scale is an fp32
    division (fp16 / fp32 promotes), round-half-even, clamp, int8.

    `.div(127.0)` on an fp16 tensor: PyTorch's CUDA kernel multiplies by the fp32 reciprocal,
    fp16(fp32(a) * fp32(1/127)) (ATen BinaryDivTrueKernel.cu, the "CPU scalar" fast path), while its CPU kernel
    divides, fp16(fp32(a) / 127).  The reference only ever runs on CUDA, so cuda_semantics=True is the
    contract; cuda_semantics=False reproduces the CPU-generated fixtures (tests/golden/pack_*.npz).  The two
    differ only when the fp32 results straddle an fp16 rounding boundary."""
    x = x.astype(np.float16)
    amax = np.abs(x).max(axis=-1, keepdims=True)
    if cuda_semantics:
        scale = (amax.astype(np.float32) * np.float32(1.0 / 127.0)).astype(np.float16).astype(np.float32)
    else:
        scale = (amax.astype(np.float32) / np.float32(127.0)).astype(np.float16).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        q = np.rint(x.astype(np.float32) / scale)
    q = np.clip(q, -128, 127)
    q = np.where(np.isnan(q), 0, q)  # all-zero row: reference is undefined (NaN -> int8); oracle picks 0
    return q.astype(np.int8), scale


# ----------------------------------------------------------------------------------------------------
# The GEMM (exact integer model) and the dequant-to-fp16 CPU path
# ----------------------------------------------------------------------------------------------------
def weights_int8(B: np.ndarray, s3: np.ndarray | None) -> np.ndarray:
    per_group = s3 is not None and s3.size > 0
    nib = unpack_B(B, per_group)
    if per_group:
        K = nib.shape[0]
        return w8_per_group(nib, unpermute_s_group(s3), K // s3.shape[0])
    return w8_per_channel(nib)


def qqq_gemm_oracle(A8: np.ndarray, B: np.ndarray, s1: np.ndarray, s2: np.ndarray, s3: np.ndarray | None,
                    W8: np.ndarray | None = None) -> np.ndarray:
    """D[m,n] = f16( (f32(sum_k A8[m,k] W8[k,n]) * s2[n]) * s1[m] ), csrc/qqq_gemm.cu:695-700.
    A8 int8 [M,K]; B packed int32 [K/16,2N]; s1 fp32 [M,1]; s2 fp32 [1,N] (permuted); s3 fp16 [G,N]
    (permuted) or empty/None.  Returns fp16 [M,N]."""
    if W8 is None:
        W8 = weights_int8(B, s3)
    acc = A8.astype(np.float64) @ W8.astype(np.float64)  # exact: |acc| <= K*128*128 << 2^53
    acc = np.clip(acc, -(2.0**31), 2.0**31 - 1)  # .satfinite (never reached for K <= 131071)
    acc32 = acc.astype(np.int64).astype(np.float32)  # cvt.rn.f32.s32
    s2n = unpermute_s_channel(np.asarray(s2, dtype=np.float32))
    d = (acc32 * s2n[None, :].astype(np.float32)).astype(np.float32)
    d = (d * np.asarray(s1, dtype=np.float32).reshape(-1, 1)).astype(np.float32)
    with np.errstate(over="ignore"):
        return d.astype(np.float16)  # cvt.rn.f16.f32


def dequant_weights_fp16(B, s2, s3):
    """fp16 weights [K, N] as a dequantising consumer would see them: W8 * s_channel'."""
    W8 = weights_int8(B, s3)
    s2n = unpermute_s_channel(np.asarray(s2, dtype=np.float32))
    return (W8.astype(np.float32) * s2n[None, :]).astype(np.float16)


def dequant_matmul_cpu(A8, s1, W_fp16):
    """The 'dequant-to-fp16 torch.matmul' CPU path BASELINE.json names (config 1): not bit-exact to the
    kernel (fp16 products/accumulation), tolerance-checked only; it is what bench.py times on host cores."""
    import torch

    a = torch.from_numpy(A8.astype(np.float32) * np.asarray(s1, dtype=np.float32).reshape(-1, 1)).to(torch.float16)
    w = torch.from_numpy(W_fp16)
    return torch.matmul(a, w).numpy()


# ----------------------------------------------------------------------------------------------------
# Synthetic problems (SURVEY.md §8d conventions; QQQ/gptq/quant.py:85-93, gptq.py:204-217)
# ----------------------------------------------------------------------------------------------------
def make_problem(M: int, K: int, N: int, group_size: int = -1, seed: int = 0):
    """Seeded synthetic (x fp16, A8, s1, B, s2, s3, w_int, scales...) in the reference's conventions."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((M, K)).astype(np.float32)
    out = rng.random((M, K)) < 1e-3
    x = np.where(out, x * 20.0, x).astype(np.float16)
    A8, s1 = dynamic_quant(x)
    W = (rng.standard_normal((K, N)) * 0.02).astype(np.float32)  # W^T, [K, N]
    if group_size == -1:
        s = np.abs(W).max(axis=0) / 7.0  # [N]
        s = np.maximum(s, 1e-8).astype(np.float16).astype(np.float32)
        w_int = np.clip(np.rint(W / s[None, :]), -7, 7).astype(np.int32)
        B = pack_B(w_int, per_group=False)
        s2 = permute_s_channel((s / 16.0).astype(np.float32))
        s3 = np.zeros((0,), dtype=np.float16)
        return dict(x=x, A8=A8, s1=s1, B=B, s2=s2, s3=s3, w_int=w_int, s_w=s)
    G = K // group_size
    Wg = W.reshape(G, group_size, N)
    s_g = (2.0 * np.abs(Wg).max(axis=1) / 15.0)  # [G, N]
    s_g = np.maximum(s_g, 1e-8).astype(np.float16).astype(np.float32)
    q = np.clip(np.rint(Wg / s_g[:, None, :]) + 8, 0, 15).astype(np.int32)
    w_fq = (q - 8).astype(np.float32) * s_g[:, None, :]
    s_extra = np.abs(w_fq.reshape(K, N)).max(axis=0) / 127.0  # [N]
    s_extra = np.maximum(s_extra, 1e-12).astype(np.float32)
    s_group_nat = (s_g / s_extra[None, :]).astype(np.float16)
    B = pack_B(q.reshape(K, N), per_group=True)
    s2 = permute_s_channel(s_extra)
    s3 = permute_s_group(s_group_nat)
    return dict(x=x, A8=A8, s1=s1, B=B, s2=s2, s3=s3, w_int=q.reshape(K, N), s_g=s_g, s_extra=s_extra)
