// QQQ._CUDA for B200: the torch-extension face of libqqq_b200.so.
//
// Builds the module the reference builds from csrc/pybind.cpp + csrc/qqq_gemm.cu (setup.py:26-31) — same module-level
// function `qqq_gemm` with the same twelve positional arguments (csrc/qqq_gemm.h:23-36, csrc/pybind.cpp:3-5) and the
// same AT_ERROR conditions and texts (csrc/qqq_gemm.cu:1060-1106) — but the body hands raw pointers to the C ABI of
// include/qqq_b200.h instead of launching the Marlin-derived kernel.  A QQQ checkout that installs this module as
// `QQQ/_CUDA*.so` runs QQQ/gptq/qlinear/qlinear_marlin.py unchanged on the sm_100a tcgen05 kernel.
// Built by integration/build_ext.py (g++ only: no device code here).
#include <ATen/cuda/CUDAContext.h>
#include <torch/extension.h>

#include "../include/qqq_b200.h"

void qqq_gemm(const torch::Tensor& A, const torch::Tensor& B, torch::Tensor& C, torch::Tensor& D, const torch::Tensor& s1,
              const torch::Tensor& s2, const torch::Tensor& s3, torch::Tensor& workspace, int thread_k = -1,
              int thread_n = -1, int sms = -1, int max_par = 8) {
  const int prob_m = A.size(0);
  const int prob_n = C.size(1);
  const int prob_k = A.size(1);
  const int groupsize = (s3.numel() == 0) ? -1 : prob_k / s3.size(0);
  if (groupsize != -1 && groupsize * s3.size(0) != prob_k) AT_ERROR("k=", prob_k, " not compatible with ", s3.size(0), " groups.");
  if (workspace.numel() < prob_n / 128 * max_par) AT_ERROR("workspace must be of size at least ", prob_n / 128 * max_par, ".");
  if (s1.scalar_type() != at::kFloat) AT_ERROR("s1 dtype must be float32, but got ", s1.scalar_type(), ".");
  if (s2.scalar_type() != at::kFloat) AT_ERROR("s2 dtype must be float32, but got ", s2.scalar_type(), ".");
  if (s3.scalar_type() != at::kHalf) AT_ERROR("s3 dtype must be float16, but got ", s3.scalar_type(), ".");
  // what the reference leaves unchecked (it would read garbage or fault): layout, dtype, device
  TORCH_CHECK(A.is_cuda() && B.is_cuda() && C.is_cuda() && D.is_cuda() && workspace.is_cuda(), "qqq_gemm: tensors must be CUDA tensors.");
  TORCH_CHECK(A.scalar_type() == at::kChar && B.scalar_type() == at::kInt && C.scalar_type() == at::kInt &&
                  D.scalar_type() == at::kHalf && workspace.scalar_type() == at::kInt,
              "qqq_gemm: A int8, B/C/workspace int32, D float16 expected.");
  TORCH_CHECK(A.is_contiguous() && B.is_contiguous() && C.is_contiguous() && D.is_contiguous(), "qqq_gemm: tensors must be contiguous.");
  const int dev = A.get_device();
  const int err = qqq_gemm_sm100a(A.data_ptr(), B.data_ptr(), C.data_ptr(), D.data_ptr(), s1.data_ptr(), s2.data_ptr(),
                                  s3.numel() ? s3.data_ptr() : nullptr, prob_m, prob_n, prob_k, workspace.data_ptr(),
                                  groupsize, dev, at::cuda::getCurrentCUDAStream(dev).stream(), thread_k, thread_n, sms,
                                  max_par);
  if (err == QQQ_ERR_PROB_SHAPE) {
    AT_ERROR("Problem (m=", prob_m, ", n=", prob_n, ", k=", prob_k, ")", " not compatible with thread_k=", thread_k,
             ", thread_n=", thread_n, ".");
  } else if (err == QQQ_ERR_KERN_SHAPE) {
    AT_ERROR("No kernel implementation for thread_k=", thread_k, ", thread_n=", thread_n, ", groupsize=", groupsize, ".");
  } else if (err != QQQ_OK) {
    AT_ERROR("qqq_gemm_sm100a failed (rc=", err, "): ", qqq_b200_last_error());
  }
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  // no py::arg names, like csrc/pybind.cpp:4 — all twelve arguments are positional and required from Python
  m.def("qqq_gemm", &qqq_gemm, "INT8xINT4 matmul based marlin FP16xINT4 kernel (sm_100a tcgen05 implementation).");
}
