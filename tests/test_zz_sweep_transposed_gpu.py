"""GPU: the BASELINE sweep shape once transposed (K=21760, N=8192) — 170 k-blocks (not a multiple of the
4- or 2-block pipeline stages, so the last stage of every tile is partly zero-filled by TMA), 64 channel tiles."""
import pytest

from test_gemm_parity import check_vs_int_mm

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,gs", [(16, -1), (16, 128), (1024, -1), (1024, 128)])
def test_transposed_sweep_shape_vs_int_mm(M, gs):
    check_vs_int_mm(M, 21760, 8192, gs)
