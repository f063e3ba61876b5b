"""CPU tests: the oracle against (1) the reference's own pack()/dynamic_quant() outputs and (2) the outputs of
the reference CUDA kernel run on a B200 (tests/golden/*.npz; generating scripts beside them)."""
import glob
import os

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import qqq_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PACK = sorted(glob.glob(os.path.join(GOLDEN, "pack_*.npz")))
KERN = sorted(glob.glob(os.path.join(GOLDEN, "kernel_*.npz")))


def test_fixtures_present():
    assert len(PACK) >= 6 and len(KERN) >= 6


@pytest.mark.parametrize("path", PACK, ids=os.path.basename)
def test_perms_match_reference(path):
    g = np.load(path)
    pg = int(g["group_size"]) != -1
    perm, sp, sps = O.get_perms(pg)
    assert np.array_equal(perm, g["perm"])
    assert np.array_equal(sp, g["scale_perm"])
    assert np.array_equal(sps, g["scale_perm_single"])


def _ints_from_golden(g):
    K, N, gs = int(g["K"]), int(g["N"]), int(g["group_size"])
    W = g["weight_fq"].T  # fp16 [K, N]
    if gs == -1:
        s = g["scales"].T  # [1, N] fp16
        w = np.clip(np.rint((W / s).astype(np.float16).astype(np.float32)), -7, 7).astype(np.int32)
        return w
    G = K // gs
    s = g["scales"].T  # [G, N]
    srep = np.repeat(s, gs, axis=0)
    q = np.clip(np.rint((W / srep).astype(np.float16).astype(np.float32)) + 8, 0, 15).astype(np.int32)
    return q


@pytest.mark.parametrize("path", PACK, ids=os.path.basename)
def test_pack_and_unpack_match_reference_pack(path):
    g = np.load(path)
    pg = int(g["group_size"]) != -1
    w = _ints_from_golden(g)
    assert np.array_equal(O.pack_B(w, pg), g["B"])
    nib = O.unpack_B(g["B"], pg)
    assert np.array_equal(nib, w & 0xF)
    if pg:
        s_nat = (g["scales"].T.astype(np.float32) / g["s_extra"].reshape(1, -1)).astype(np.float16)
        assert np.array_equal(O.permute_s_group(s_nat).view(np.uint16), g["s_group"].view(np.uint16))
        assert np.array_equal(O.permute_s_channel(g["s_extra"].reshape(-1)), g["s_channel"])
        assert np.array_equal(O.unpermute_s_group(g["s_group"]).view(np.uint16), s_nat.view(np.uint16))
    else:
        s = (g["scales"].reshape(-1) / np.float16(16)).astype(np.float16).astype(np.float32)
        assert np.array_equal(O.permute_s_channel(s), g["s_channel"])
        assert np.array_equal(O.unpermute_s_channel(g["s_channel"]), s)


@pytest.mark.parametrize("path", PACK, ids=os.path.basename)
def test_dynamic_quant_matches_reference_cpu(path):
    g = np.load(path)
    q, s = O.dynamic_quant(g["x"], cuda_semantics=False)
    assert np.array_equal(q, g["quant_A"])
    assert np.array_equal(s, g["s1"])


@pytest.mark.parametrize("path", KERN, ids=os.path.basename)
def test_oracle_gemm_bit_exact_vs_reference_kernel(path):
    """D produced by the reference's csrc/qqq_gemm.cu on a B200 == oracle, bit for bit."""
    g = np.load(path)
    D = O.qqq_gemm_oracle(g["A8"], g["B"], g["s1"], g["s2"], g["s3"])
    assert np.array_equal(D.view(np.uint16), g["D"].view(np.uint16))


@pytest.mark.parametrize("path", KERN, ids=os.path.basename)
def test_dequant_fp16_matmul_is_close(path):
    """The dequant-to-fp16 matmul CPU path (BASELINE config 1) agrees with the exact kernel within fp16 error."""
    g = np.load(path)
    Wh = O.dequant_weights_fp16(g["B"], g["s2"], g["s3"])
    D4 = O.dequant_matmul_cpu(g["A8"], g["s1"], Wh).astype(np.float32)
    D = g["D"].astype(np.float32)
    tol = 2e-2 * max(1.0, np.abs(D).max())
    assert np.abs(D4 - D).max() <= tol


def test_problem_generator_is_seeded():
    a = O.make_problem(5, 256, 128, 128, seed=3)
    b = O.make_problem(5, 256, 128, 128, seed=3)
    for k in ("A8", "B", "s1", "s2", "s3"):
        assert np.array_equal(a[k], b[k])


def test_per_group_weights_in_int8_range():
    p = O.make_problem(3, 512, 128, 128, seed=9)
    W8 = O.weights_int8(p["B"], p["s3"])
    assert W8.min() >= -128 and W8.max() <= 127
    # definition check: W8 == RNE((v-8) * s_group) with the fp16 scale
    nib = O.unpack_B(p["B"], True)
    s = np.repeat(O.unpermute_s_group(p["s3"]).astype(np.float64), 128, axis=0)
    assert np.array_equal(W8, np.rint((nib - 8) * s).astype(np.int32))


@settings(max_examples=25, deadline=None)
@given(kt=st.integers(1, 6), nb=st.integers(1, 3), pg=st.booleans(), seed=st.integers(0, 2**31 - 1))
def test_pack_unpack_roundtrip_property(kt, nb, pg, seed):
    K, N = 16 * kt, 64 * nb
    rng = np.random.default_rng(seed)
    w = rng.integers(0, 16, size=(K, N)) if pg else rng.integers(-8, 8, size=(K, N))
    B = O.pack_B(w, pg)
    assert B.shape == (K // 16, 2 * N) and B.dtype == np.int32
    assert np.array_equal(O.unpack_B(B, pg), w & 0xF)


@settings(max_examples=15, deadline=None)
@given(m=st.integers(1, 9), seed=st.integers(0, 2**31 - 1), pg=st.booleans())
def test_gemm_linearity_in_activation_scale(m, seed, pg):
    """Doubling s1 doubles D exactly (power-of-two scaling commutes with every rounding in the epilogue)."""
    p = O.make_problem(m, 256, 64, 128 if pg else -1, seed=seed % 1000)
    D1 = O.qqq_gemm_oracle(p["A8"], p["B"], p["s1"], p["s2"], p["s3"]).astype(np.float32)
    D2 = O.qqq_gemm_oracle(p["A8"], p["B"], p["s1"] * 2, p["s2"], p["s3"]).astype(np.float32)
    ok = np.isfinite(D2) & (np.abs(D1) > 1e-3)
    assert np.array_equal(D2[ok], 2 * D1[ok])


def test_config1_plumbing_cpu():
    """BASELINE config 1: single Linear forward M=1, K=N=4096 per-channel via dequant-to-fp16 matmul on CPU."""
    p = O.make_problem(1, 4096, 4096, -1, seed=1)
    D = O.qqq_gemm_oracle(p["A8"], p["B"], p["s1"], p["s2"], p["s3"]).astype(np.float32)
    Wh = O.dequant_weights_fp16(p["B"], p["s2"], p["s3"])
    D4 = O.dequant_matmul_cpu(p["A8"], p["s1"], Wh).astype(np.float32)
    assert np.abs(D4 - D).max() <= 2e-2 * max(1.0, np.abs(D).max())
    # and the quantised product tracks the un-quantised x @ W within the quantisation error budget
    W = p["w_int"].astype(np.float32) * p["s_w"][None, :]
    ref = p["x"].astype(np.float32) @ W
    rel = np.linalg.norm(D - ref) / np.linalg.norm(ref)  # int8 activation noise with 20x outliers: ~0.15
    assert rel < 0.3
