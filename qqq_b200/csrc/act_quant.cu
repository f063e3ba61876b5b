// Per-token dynamic int8 activation quantisation for sm_100a — one kernel instead of the reference's five
// eager torch ops (QQQ/gptq/qlinear/qlinear_marlin.py:265-268):
//     quant_scale = x.abs().max(dim=-1, keepdim=True)[0].div(127.0).to(torch.float32)
//     x = (x / quant_scale).round().clamp(-128, 127).to(torch.int8)
// Bit-exact restatement of what those ops compute ON A CUDA DEVICE:
//   * abs/max in fp16 are exact;
//   * `.div(127.0)` on a CUDA fp16 tensor with a Python scalar is evaluated by PyTorch as
//     fp16( fp32(a) * fp32(1.0/127.0) )  (multiply by the reciprocal, ATen BinaryDivTrueKernel.cu);
//   * `x / quant_scale` promotes to fp32 and is an IEEE fp32 division; round() is round-half-even.
// HBM-bound: reads 2*M*K bytes, writes M*K + 4*M.  One CTA per row; the row is held in registers
// between the max pass and the quantise pass for K <= 8 * 8 * blockDim.
#include "qqq_common.cuh"

namespace qqq {

constexpr int kQuantThreads = 256;
constexpr int kQuantMaxChunks = 8;  // 16-byte chunks (8 halves) cached per thread

__device__ __forceinline__ uint32_t habs2_u32(uint32_t v) { return v & 0x7FFF7FFFu; }

__device__ __forceinline__ uint32_t quant4(uint32_t lo, uint32_t hi, float s) {
  // lo, hi: two half2 (4 consecutive halves) -> 4 int8 packed little-endian
  const __half2 a = *reinterpret_cast<const __half2*>(&lo);
  const __half2 b = *reinterpret_cast<const __half2*>(&hi);
  float f[4] = {__low2float(a), __high2float(a), __low2float(b), __high2float(b)};
  uint32_t out = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float q = rintf(__fdiv_rn(f[i], s));
    // NaN (all-zero row: 0/0) -> 0: the reference's float->int8 cast of NaN is undefined; CUDA's cvt gives 0
    int qi = (q == q) ? __float2int_rn(fminf(fmaxf(q, -128.f), 127.f)) : 0;
    out |= (uint32_t)(qi & 0xFF) << (8 * i);
  }
  return out;
}

__global__ void __launch_bounds__(kQuantThreads) act_quant_kernel(const uint4* __restrict__ x, uint2* __restrict__ q,
                                                                 float* __restrict__ s1, int K8 /* K/8 */) {
  const int row = blockIdx.x;
  const uint4* xr = x + (size_t)row * K8;
  uint2* qr = q + (size_t)row * K8;
  uint4 cache[kQuantMaxChunks];
  uint32_t m = 0;  // running max of |x| as raw fp16 bits (monotone for non-negative halves), both lanes
  auto upd = [&](const uint4& v) {
    const uint32_t w[4] = {habs2_u32(v.x), habs2_u32(v.y), habs2_u32(v.z), habs2_u32(v.w)};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __half2 a = *reinterpret_cast<const __half2*>(&w[i]);
      __half2 b = *reinterpret_cast<const __half2*>(&m);
      __half2 c = __hmax2_nan(a, b);
      m = *reinterpret_cast<uint32_t*>(&c);
    }
  };
#pragma unroll
  for (int j = 0; j < kQuantMaxChunks; ++j) {
    const int i = threadIdx.x + j * kQuantThreads;
    if (i < K8) {
      cache[j] = __ldg(xr + i);
      upd(cache[j]);
    }
  }
  for (int i = threadIdx.x + kQuantMaxChunks * kQuantThreads; i < K8; i += kQuantThreads) upd(__ldg(xr + i));
  // reduce the two lanes, then warp, then block
  __half2 mh = *reinterpret_cast<__half2*>(&m);
  __half mx = __hmax_nan(__low2half(mh), __high2half(mh));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    __half other = __shfl_xor_sync(0xffffffffu, mx, o);
    mx = __hmax_nan(mx, other);
  }
  __shared__ __half red[kQuantThreads / 32];
  __shared__ float s_sh;
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    __half t = red[0];
#pragma unroll
    for (int i = 1; i < kQuantThreads / 32; ++i) t = __hmax_nan(t, red[i]);
    const float inv127 = (float)(1.0 / 127.0);
    const __half sh = __float2half_rn(__fmul_rn(__half2float(t), inv127));
    const float s = __half2float(sh);
    s_sh = s;
    s1[row] = s;
  }
  __syncthreads();
  const float s = s_sh;
#pragma unroll
  for (int j = 0; j < kQuantMaxChunks; ++j) {
    const int i = threadIdx.x + j * kQuantThreads;
    if (i < K8) qr[i] = make_uint2(quant4(cache[j].x, cache[j].y, s), quant4(cache[j].z, cache[j].w, s));
  }
  for (int i = threadIdx.x + kQuantMaxChunks * kQuantThreads; i < K8; i += kQuantThreads) {
    const uint4 v = __ldg(xr + i);
    qr[i] = make_uint2(quant4(v.x, v.y, s), quant4(v.z, v.w, s));
  }
}

cudaError_t launch_act_quant(const void* x, void* q, void* s1, int M, int K, cudaStream_t stream) {
  act_quant_kernel<<<M, kQuantThreads, 0, stream>>>(reinterpret_cast<const uint4*>(x), reinterpret_cast<uint2*>(q),
                                                    reinterpret_cast<float*>(s1), K / 8);
  return cudaGetLastError();
}

}  // namespace qqq
