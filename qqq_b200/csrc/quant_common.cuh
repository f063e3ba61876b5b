// Device helpers shared by the activation-quantisation kernels (act_quant.cu, tp_reduce_quant.cu): the reference's
// per-token int8 quantisation (QQQ/gptq/qlinear/qlinear_marlin.py:265-268) as it evaluates on a CUDA device.
#pragma once
#include "qqq_common.cuh"

namespace qqq {

__device__ __forceinline__ uint32_t habs2_u32(uint32_t v) { return v & 0x7FFF7FFFu; }

// 4 consecutive halves -> 4 int8 packed little-endian:  int8(clamp(rint(x / s), -128, 127)).
// FAST: s is finite and > 0.  The IEEE quotient RN(x/s) is then obtained without a division per element:
// r = RN(1/s) once per row, q0 = RN(x*r), q = fma(fma(-s, q0, x), r, q0).  For every finite fp16 x and every positive
// finite fp16-valued s this equals RN(x/s) bit for bit except for the sign of a zero result (checked exhaustively,
// oracle/div_identity.c / tests/test_act_quant_identity.py), so the int8 result is identical.  Otherwise (all-zero
// row: s = 0; inf/NaN in the row) the true division is used: x/0 = +-inf saturates, 0/0 = NaN converts to 0 (the
// reference's float->int8 cast of NaN is undefined; CUDA's cvt gives 0).
// cvt.rni (round-half-even, like torch.round) + cvt.pack.sat replace round/clamp/cast.
template <bool FAST>
__device__ __forceinline__ uint32_t quant4(uint32_t lo, uint32_t hi, float s, float r) {
  const __half2 a = *reinterpret_cast<const __half2*>(&lo);
  const __half2 b = *reinterpret_cast<const __half2*>(&hi);
  const float f[4] = {__low2float(a), __high2float(a), __low2float(b), __high2float(b)};
  int qi[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float q;
    if (FAST) {
      const float q0 = __fmul_rn(f[i], r);
      q = __fmaf_rn(__fmaf_rn(-s, q0, f[i]), r, q0);
    } else {
      q = __fdiv_rn(f[i], s);
    }
    qi[i] = __float2int_rn(q);  // NaN -> 0, +-inf -> INT_MAX / INT_MIN
  }
  uint32_t t, out;
  asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(t) : "r"(qi[3]), "r"(qi[2]), "r"(0));
  asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(out) : "r"(qi[1]), "r"(qi[0]), "r"(t));
  return out;
}

__device__ __forceinline__ uint32_t hmax2_u32(uint32_t a, uint32_t b) {
  __half2 c = __hmax2_nan(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
  return *reinterpret_cast<uint32_t*>(&c);
}

// s1 = fp32(fp16(max|x| / 127)): PyTorch's CUDA `div(scalar)` on an fp16 tensor multiplies by the fp32 reciprocal
// (ATen BinaryDivTrueKernel.cu) and rounds to fp16.
__device__ __forceinline__ float token_scale(__half row_max) {
  const float inv127 = (float)(1.0 / 127.0);
  return __half2float(__float2half_rn(__fmul_rn(__half2float(row_max), inv127)));
}

}  // namespace qqq
