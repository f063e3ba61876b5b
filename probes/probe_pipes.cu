// Per-SMSP issue rate of the instructions the per-group unpack is made of (dev tooling).
// Each warp runs 8 independent dependency chains of one op; cycles per warp-instruction are reported for 1 and 4
// warps per SM sub-partition.   nvcc -arch=sm_100a -O3 -o probe_pipes probe_pipes.cu
#include <cuda_fp16.h>
#include <cstdint>
#include <cstdio>

constexpr int kIters = 2048;
constexpr int kChains = 8;

template <int OP>
__device__ __forceinline__ uint32_t op(uint32_t x, uint32_t a, uint32_t b) {
  uint32_t d;
  if (OP == 0) asm volatile("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(a), "r"(b));           // HFMA2 rrr
  if (OP == 1) { const uint32_t c = 0x65006500u; asm volatile("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(a), "r"(c)); }  // HFMA2 rr,imm
  if (OP == 2) { const uint32_t c = 0x64086408u; asm volatile("sub.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(x), "r"(c)); }              // HADD2 imm
  if (OP == 3) asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(a), "r"(b));               // FFMA rrr
  if (OP == 4) asm volatile("fma.rn.f32 %0, %1, %2, 0f4B400000;" : "=r"(d) : "r"(x), "r"(a));             // FFMA rr,imm
  if (OP == 5) asm volatile("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(x), "r"(a), "r"(b));          // LOP3
  if (OP == 6) asm volatile("prmt.b32 %0, %1, %2, 0x6420;" : "=r"(d) : "r"(x), "r"(a));                    // PRMT
  if (OP == 7) asm volatile("mad.lo.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(a), "r"(b));              // IMAD
  if (OP == 8) asm volatile("shr.u32 %0, %1, 8;" : "=r"(d) : "r"(x));                                      // SHF
  if (OP == 9) asm volatile("add.f32 %0, %1, 0fCB000008;" : "=r"(d) : "r"(x));                            // FADD imm
  if (OP == 10) asm volatile("cvt.rn.f32.s32 %0, %1;" : "=r"(d) : "r"(x));                                 // I2FP.F32.S32
  if (OP == 11) { unsigned short h; asm volatile("cvt.rn.f16.f32 %0, %1;" : "=h"(h) : "r"(x)); d = h; }   // F2FP (one value)
  if (OP == 12) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "r"(x), "r"(a));                   // F2FP.PACK_AB
  if (OP == 13) asm volatile("cvt.rni.s32.f32 %0, %1;" : "=r"(d) : "r"(x));                                // F2I
  if (OP == 14) { unsigned short lo16, hi16; asm volatile("mov.b32 {%0,%1}, %2;" : "=h"(lo16), "=h"(hi16) : "r"(x));
                  asm volatile("cvt.f32.f16 %0, %1;" : "=r"(d) : "h"(lo16)); }                              // HADD2.F32 (half -> float)
  return d;
}

template <int OP>
__global__ void k(uint32_t* out, long long* cyc, uint32_t a, uint32_t b) {
  uint32_t x[kChains];
#pragma unroll
  for (int c = 0; c < kChains; ++c) x[c] = threadIdx.x + c;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < kIters; ++i) {
#pragma unroll
    for (int c = 0; c < kChains; ++c) x[c] = op<OP>(x[c], a, b);
  }
  const long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int c = 0; c < kChains; ++c) s ^= x[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if ((threadIdx.x & 31) == 0) cyc[threadIdx.x >> 5] = t1 - t0;
}

template <int OP>
void run(const char* name) {
  uint32_t* out;
  long long* cyc;
  cudaMalloc(&out, 1024 * 4);
  cudaMalloc(&cyc, 32 * 8);
  for (int warps_per_smsp : {1, 2, 4}) {
    const int threads = 128 * warps_per_smsp;
    k<OP><<<1, threads>>>(out, cyc, 0x3C003C00u, 0x00000000u);
    k<OP><<<1, threads>>>(out, cyc, 0x3C003C00u, 0x00000000u);
    cudaDeviceSynchronize();
    long long h[32];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int w = 0; w < threads / 32; ++w) mx = h[w] > mx ? h[w] : mx;
    // instructions issued per SMSP = warps_per_smsp * kIters * kChains
    printf("%-14s warps/SMSP=%d  cycles per warp-instruction per SMSP = %.2f\n", name, warps_per_smsp,
           (double)mx / ((double)warps_per_smsp * kIters * kChains));
  }
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run<0>("HFMA2 rrr");
  run<1>("HFMA2 rr,imm");
  run<2>("HADD2 r,imm");
  run<3>("FFMA rrr");
  run<4>("FFMA rr,imm");
  run<5>("LOP3");
  run<6>("PRMT");
  run<7>("IMAD");
  run<8>("SHF");
  run<9>("FADD r,imm");
  run<10>("I2FP.F32.S32");
  run<11>("F2FP f16<-f32");
  run<12>("F2FP.PACK_AB");
  run<13>("F2I.S32.F32");
  run<14>("cvt f32<-f16");
  return 0;
}
