"""The pybind11 face of the boundary (integration/qqq_cuda_b200.cpp -> integration/_CUDA.so): the module a QQQ checkout
installs as `QQQ._CUDA` in place of the one built from csrc/pybind.cpp + csrc/qqq_gemm.cu.  CPU: it builds against this
torch, imports without a GPU, keeps pybind's positional-only twelve-argument signature and raises the reference's
RuntimeErrors (csrc/qqq_gemm.cu:1066-1075) before touching the device.  GPU: tests/test_zz_integration_gpu.py."""
import importlib.util
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_ext():
    sys.path.insert(0, os.path.join(ROOT, "integration"))
    import build_ext

    so = build_ext.build()
    spec = importlib.util.spec_from_file_location("_CUDA", str(so))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def ext():
    return load_ext()


def _args(M=4, K=128, N=128, max_par=16):
    return [torch.zeros(M, K, dtype=torch.int8), torch.zeros(K // 16, 2 * N, dtype=torch.int32),
            torch.zeros(64 * max_par, N, dtype=torch.int32), torch.zeros(M, N, dtype=torch.half), torch.zeros(M, 1),
            torch.zeros(1, N), torch.zeros(0, dtype=torch.half), torch.zeros(N // 128 * max_par, dtype=torch.int32),
            -1, -1, -1, max_par]


def test_signature_is_positional_with_twelve_required_arguments(ext):
    assert "qqq_gemm(arg0: torch.Tensor" in ext.qqq_gemm.__doc__ and "arg11" in ext.qqq_gemm.__doc__
    with pytest.raises(TypeError):
        ext.qqq_gemm(*_args()[:8])  # the C++ defaults are invisible from Python, like csrc/pybind.cpp:4
    with pytest.raises(TypeError):
        ext.qqq_gemm(*_args()[:11], max_par=16)


def test_reference_error_texts_before_any_device_work(ext):
    a = _args()
    a[4] = a[4].double()
    with pytest.raises(RuntimeError, match="s1 dtype must be float32"):
        ext.qqq_gemm(*a)
    a = _args()
    a[5] = a[5].half()
    with pytest.raises(RuntimeError, match="s2 dtype must be float32"):
        ext.qqq_gemm(*a)
    a = _args()
    a[6] = torch.zeros(3, 128)  # 3 groups do not divide K = 128; also wrong dtype, but the group check comes first
    with pytest.raises(RuntimeError, match="k=128 not compatible with 3 groups"):
        ext.qqq_gemm(*a)
    a = _args()
    a[7] = a[7][:5]
    with pytest.raises(RuntimeError, match="workspace must be of size at least 16"):
        ext.qqq_gemm(*a)
    with pytest.raises(RuntimeError, match="CUDA tensors"):
        ext.qqq_gemm(*_args())  # no CPU path


def test_links_the_product_library_by_relative_rpath(ext):
    import subprocess

    out = subprocess.check_output(["readelf", "-d", os.path.join(ROOT, "integration", "_CUDA.so")], text=True)
    assert "libqqq_b200.so" in out and "$ORIGIN/../qqq_b200" in out
