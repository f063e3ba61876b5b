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
  TORCH_CHECK(A.is_contiguous() && B.is_contiguous() && C.is_contiguous() && D.is_contiguous() && s1.is_contiguous() &&
                  s2.is_contiguous() && s3.is_contiguous() && workspace.is_contiguous(),
              "qqq_gemm: tensors must be contiguous.");
  TORCH_CHECK(s1.is_cuda() && s2.is_cuda() && (s3.numel() == 0 || s3.is_cuda()), "qqq_gemm: scales must be CUDA tensors.");
  for (const torch::Tensor* t : std::initializer_list<const torch::Tensor*>{&B, &C, &D, &s1, &s2, &workspace})
    TORCH_CHECK(t->get_device() == A.get_device(), "qqq_gemm: all tensors must be on the device of A.");
  TORCH_CHECK(s3.numel() == 0 || s3.get_device() == A.get_device(), "qqq_gemm: all tensors must be on the device of A.");
  // the planner uses C (64*max_par rows of N int32) as split-K scratch up to that capacity: a smaller C would be overrun
  TORCH_CHECK(prob_m == 0 || C.size(0) >= 64 * max_par, "C must have at least ", 64 * max_par, " rows.");
  TORCH_CHECK(A.dim() == 2 && B.dim() == 2 && C.dim() == 2 && D.dim() == 2, "qqq_gemm: A, B, C, D must be matrices.");
  TORCH_CHECK(B.size(0) * 16 == prob_k && B.size(1) == 2 * (int64_t)prob_n, "qqq_gemm: B must be [k/16, 2n] = [", prob_k / 16,
              ", ", 2 * prob_n, "].");
  TORCH_CHECK(D.size(0) == prob_m && D.size(1) == prob_n, "qqq_gemm: D must be [m, n] = [", prob_m, ", ", prob_n, "].");
  TORCH_CHECK(s1.numel() == prob_m && s2.numel() == prob_n, "qqq_gemm: s1 needs m and s2 needs n elements.");
  TORCH_CHECK(s3.numel() == 0 || s3.size(1) == prob_n, "qqq_gemm: s3 must be [k/groupsize, n].");
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
