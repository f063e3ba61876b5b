// Per-token dynamic int8 activation quantisation for sm_100a — one kernel instead of the reference's five
// eager torch ops (QQQ/gptq/qlinear/qlinear_marlin.py:265-268):
//     quant_scale = x.abs().max(dim=-1, keepdim=True)[0].div(127.0).to(torch.float32)
//     x = (x / quant_scale).round().clamp(-128, 127).to(torch.int8)
// Bit-exact restatement of what those ops compute ON A CUDA DEVICE:
//   * abs/max in fp16 are exact;
//   * `.div(127.0)` on a CUDA fp16 tensor with a Python scalar is evaluated by PyTorch as
//     fp16( fp32(a) * fp32(1.0/127.0) )  (multiply by the reciprocal, ATen BinaryDivTrueKernel.cu);
//   * `x / quant_scale` promotes to fp32 and is an IEEE fp32 division; round() is round-half-even.
// HBM-bound: reads 2*M*K bytes, writes M*K + 4*M.  One 256-thread CTA per token row; the row stays in registers
// between the max pass and the quantise pass (NCH 16-byte chunks per thread, NCH chosen from K so that the
// register footprint — and with it the number of resident CTAs per SM — matches the row length).
#include <type_traits>

#include "quant_common.cuh"

namespace qqq {

constexpr int kQuantThreads = 256;

// NCH > 0: the row has at most NCH*256 chunks and lives in registers.  NCH == 0: any K, second pass re-reads x.
// Rows up to 4096 halves (NCH <= 2) are held to 32 registers so that 8 CTAs fit an SM: 1184 rows per wave, i.e. a
// 1024-token prefill is one wave instead of one full wave plus a 15 % tail.
template <int NCH>
__global__ void __launch_bounds__(kQuantThreads, (NCH >= 1 && NCH <= 2) ? 8 : 1)
act_quant_kernel(const uint4* __restrict__ x, uint2* __restrict__ q, float* __restrict__ s1, int K8 /* K/8 */,
                 int ldx8 /* row stride of x in 16-byte units */) {
  grid_launch_dependents();  // the GEMM that consumes q/s1 may start its prologue and weight prefetch right away
  grid_dependency_wait();    // x is the preceding kernel's output (programmatic dependent launch)
  const int row = blockIdx.x;
  const uint4* xr = x + (size_t)row * ldx8;
  uint2* qr = q + (size_t)row * K8;
  constexpr int NC = NCH > 0 ? NCH : 1;
  uint4 cache[NC];
  uint32_t m = 0;  // running max of |x| for both half2 lanes (fp16 bit patterns)
  if (NCH > 0) {
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      const int i = threadIdx.x + j * kQuantThreads;
      cache[j] = (i < K8) ? __ldg(xr + i) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      m = hmax2_u32(m, habs2_u32(cache[j].x));
      m = hmax2_u32(m, habs2_u32(cache[j].y));
      m = hmax2_u32(m, habs2_u32(cache[j].z));
      m = hmax2_u32(m, habs2_u32(cache[j].w));
    }
  } else {
    for (int i = threadIdx.x; i < K8; i += kQuantThreads) {
      const uint4 v = __ldg(xr + i);
      m = hmax2_u32(m, habs2_u32(v.x));
      m = hmax2_u32(m, habs2_u32(v.y));
      m = hmax2_u32(m, habs2_u32(v.z));
      m = hmax2_u32(m, habs2_u32(v.w));
    }
  }
  // reduce the two lanes, then the warp, then the block (every thread finishes the block reduce itself: one barrier)
  __half2 mh = *reinterpret_cast<__half2*>(&m);
  __half mx = __hmax_nan(__low2half(mh), __high2half(mh));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = __hmax_nan(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  __shared__ __half red[kQuantThreads / 32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  __half t = red[0];
#pragma unroll
  for (int i = 1; i < kQuantThreads / 32; ++i) t = __hmax_nan(t, red[i]);
  const float s = token_scale(t);
  if (threadIdx.x == 0) s1[row] = s;
  const bool fast = s > 0.f && s < __int_as_float(0x7F800000);  // uniform over the CTA
  const float r = __frcp_rn(s);
  auto quant_row = [&](auto fast_tag) {
    constexpr bool F = decltype(fast_tag)::value;
    if (NCH > 0) {
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        const int i = threadIdx.x + j * kQuantThreads;
        if (i < K8)
          qr[i] = make_uint2(quant4<F>(cache[j].x, cache[j].y, s, r), quant4<F>(cache[j].z, cache[j].w, s, r));
      }
    } else {
      for (int i = threadIdx.x; i < K8; i += kQuantThreads) {
        const uint4 v = __ldg(xr + i);
        qr[i] = make_uint2(quant4<F>(v.x, v.y, s, r), quant4<F>(v.z, v.w, s, r));
      }
    }
  };
  if (fast)
    quant_row(std::true_type{});
  else
    quant_row(std::false_type{});
}

template <int NCH>
static cudaError_t launch_one(const uint4* xp, uint2* qp, float* sp, int M, int K8, int ldx8, cudaStream_t stream,
                              bool pdl) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(M);
  cfg.blockDim = dim3(kQuantThreads);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, act_quant_kernel<NCH>, xp, qp, sp, K8, ldx8);
}

cudaError_t launch_act_quant(const void* x, long long ldx, void* q, void* s1, int M, int K, cudaStream_t stream,
                             bool pdl) {
  const uint4* xp = reinterpret_cast<const uint4*>(x);
  uint2* qp = reinterpret_cast<uint2*>(q);
  float* sp = reinterpret_cast<float*>(s1);
  const int K8 = K / 8;
  const int ldx8 = (int)(ldx / 8);
  const int nch = (K8 + kQuantThreads - 1) / kQuantThreads;
  if (nch <= 1) return launch_one<1>(xp, qp, sp, M, K8, ldx8, stream, pdl);
  if (nch <= 2) return launch_one<2>(xp, qp, sp, M, K8, ldx8, stream, pdl);
  if (nch <= 4) return launch_one<4>(xp, qp, sp, M, K8, ldx8, stream, pdl);
  if (nch <= 8) return launch_one<8>(xp, qp, sp, M, K8, ldx8, stream, pdl);
  return launch_one<0>(xp, qp, sp, M, K8, ldx8, stream, pdl);
}

}  // namespace qqq
