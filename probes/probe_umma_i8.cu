// Probe: tcgen05.mma kind::i8 operand paths on sm_100a (development tooling, not the product path).
//   D[128,N] (s32, TMEM) = A[128,K] (s8) * B[N,K]^T (s8), K = 128.
// modes:
//   0  SS  A: smem SW128 (thread-written)       B: smem SW128 (thread-written)
//   1  SS  A: smem no-swizzle core-matrix layout B: smem SW128
//   2  TS  A: TMEM via tcgen05.st.32x32b         B: smem SW128
//   3  TS  A: TMEM via tcgen05.st.16x128b.x2     B: smem SW128   (mma-fragment-like store, +16 lane halves)
//   4  SS  A: smem SW128                         B: smem SW128 via TMA (cuTensorMap SWIZZLE_128B)
//   5  SS  A: no-swizzle                         B: no-swizzle
//   10 bench SS   11 bench TS  (M=128,N=256,K=32 MMAs back-to-back on every SM)
// usage: probe_umma_i8 <mode> [N]
#include "umma_common.cuh"
#include <vector>
#include <cstring>

constexpr int KT = 128;  // bytes of K per tile

template <int MODE>
__global__ void __launch_bounds__(128, 1)
probe_kernel(const int8_t* __restrict__ A, const int8_t* __restrict__ B, int32_t* __restrict__ D,
             uint32_t* __restrict__ tmemA_dump, int N, const __grid_constant__ CUtensorMap tmapB) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                 // 128 x 128 B = 16 KB
  uint8_t* sB = smem + 16384;         // up to 256 x 128 B = 32 KB
  __shared__ uint64_t bar_mma, bar_tma;
  __shared__ uint32_t tmem_base_s;

  const int t = threadIdx.x, w = t >> 5, l = t & 31;
  constexpr bool A_TMEM = (MODE == 2 || MODE == 3);
  constexpr bool A_NOSW = (MODE == 1 || MODE == 5);
  constexpr bool B_NOSW = (MODE == 5);
  constexpr bool B_TMA = (MODE == 4);

  if (w == 0) tmem_alloc(smem_u32(&tmem_base_s), 512);
  if (t == 0) {
    mbar_init(smem_u32(&bar_mma), 1);
    mbar_init(smem_u32(&bar_tma), 1);
    mbar_fence_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t tmem_d = tmem;         // columns [0, N)
  const uint32_t tmem_a = tmem + 256;   // columns [256, 256+32)

  // ---- stage B (N rows x 128 B) ----
  if constexpr (B_TMA) {
    if (t == 0) {
      mbar_expect_tx(smem_u32(&bar_tma), N * KT);
      tma_load_2d(smem_u32(sB), &tmapB, smem_u32(&bar_tma), 0, 0);
    }
  } else {
    for (int r = t; r < N; r += 128) {
      for (int c = 0; c < 8; ++c) {
        uint4 v = *reinterpret_cast<const uint4*>(B + (size_t)r * KT + c * 16);
        uint32_t off = B_NOSW ? ((r >> 3) * 1024 + c * 128 + (r & 7) * 16) : (r * 128 + ((c ^ (r & 7)) * 16));
        *reinterpret_cast<uint4*>(sB + off) = v;
      }
    }
  }
  // ---- stage A (128 rows x 128 B) ----
  if constexpr (!A_TMEM) {
    const int r = t;
    for (int c = 0; c < 8; ++c) {
      uint4 v = *reinterpret_cast<const uint4*>(A + (size_t)r * KT + c * 16);
      uint32_t off = A_NOSW ? ((r >> 3) * 1024 + c * 128 + (r & 7) * 16) : (r * 128 + ((c ^ (r & 7)) * 16));
      *reinterpret_cast<uint4*>(sA + off) = v;
    }
  } else if constexpr (MODE == 2) {
    // thread t owns TMEM lane t (= row t): 32 columns = 128 bytes of K
    const uint32_t* arow = reinterpret_cast<const uint32_t*>(A + (size_t)t * KT);
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t r[8];
      for (int i = 0; i < 8; ++i) r[i] = arow[ks * 8 + i];
      tmem_st_32x32b_x8(tmem_a + ((uint32_t)(32 * w) << 16) + ks * 8, r);
    }
    tmem_wait_st();
  } else {
    // MODE 3: 16x128b.x2; lane l: c = l/4 (row in 8), kq = l%4 (which 4-byte k group of a 16-byte slab)
    const int c = l >> 2, kq = l & 3;
    for (int h = 0; h < 2; ++h) {
      const int row0 = 32 * w + 16 * h + c;
      const uint32_t* a0 = reinterpret_cast<const uint32_t*>(A + (size_t)row0 * KT);
      const uint32_t* a1 = reinterpret_cast<const uint32_t*>(A + (size_t)(row0 + 8) * KT);
      for (int ks = 0; ks < 4; ++ks) {
        // k-step ks covers bytes [32ks, 32ks+32): slab0 = words 8ks+0..3, slab1 = words 8ks+4..7
        uint32_t r0 = a0[8 * ks + kq], r1 = a1[8 * ks + kq], r2 = a0[8 * ks + 4 + kq], r3 = a1[8 * ks + 4 + kq];
        tmem_st_16x128b_x2(tmem_a + ((uint32_t)(32 * w + 16 * h) << 16) + ks * 8, r0, r1, r2, r3);
      }
    }
    tmem_wait_st();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if constexpr (A_TMEM) {
    // read back what landed in TMEM (32x32b view) for diagnosis
    for (int cb = 0; cb < 4; ++cb) {
      uint32_t r[8];
      tmem_ld_32x32b_x8(tmem_a + ((uint32_t)(32 * w) << 16) + cb * 8, r);
      tmem_wait_ld();
      for (int i = 0; i < 8; ++i) tmemA_dump[t * 32 + cb * 8 + i] = r[i];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }

  if (t == 0) {
    if constexpr (B_TMA) mbar_wait(smem_u32(&bar_tma), 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_i8(128, N);
    for (int ks = 0; ks < 4; ++ks) {
      uint64_t db = B_NOSW ? make_smem_desc(smem_u32(sB) + ks * 256, 128, 1024, 0)
                           : make_smem_desc(smem_u32(sB) + ks * 32, 16, 1024, 2);
      if constexpr (A_TMEM) {
        umma_i8_ts(tmem_d, tmem_a + ks * 8, db, idesc, ks > 0);
      } else {
        uint64_t da = A_NOSW ? make_smem_desc(smem_u32(sA) + ks * 256, 128, 1024, 0)
                             : make_smem_desc(smem_u32(sA) + ks * 32, 16, 1024, 2);
        umma_i8_ss(tmem_d, da, db, idesc, ks > 0);
      }
    }
    umma_commit(smem_u32(&bar_mma));
  }
  mbar_wait(smem_u32(&bar_mma), 0);
  tc_fence_after();
  for (int cb = 0; cb < N / 8; ++cb) {
    uint32_t r[8];
    tmem_ld_32x32b_x8(tmem_d + ((uint32_t)(32 * w) << 16) + cb * 8, r);
    tmem_wait_ld();
    for (int i = 0; i < 8; ++i) D[(size_t)t * N + cb * 8 + i] = (int32_t)r[i];
  }
  tc_fence_before();
  __syncthreads();
  if (w == 0) tmem_dealloc(tmem, 512);
}

// ---- throughput: back-to-back MMAs (M=128, N=256, K=32) -------------------------------------------
template <bool TS>
__global__ void __launch_bounds__(128, 1) bench_kernel(int iters, int n_mma) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int t = threadIdx.x, w = t >> 5;
  for (int i = t; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u * (i & 3);
  if (w == 0) tmem_alloc(smem_u32(&tmem_base_s), 512);
  if (t == 0) {
    mbar_init(smem_u32(&bar), 1);
    mbar_fence_init();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (t == 0) {
    const uint32_t idesc = make_idesc_i8(128, n_mma);
    uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
      for (int j = 0; j < 64; ++j) {
        const int ks = j & 3;
        uint64_t db = make_smem_desc(smem_u32(smem + 16384) + ks * 32, 16, 1024, 2);
        if (TS) {
          umma_i8_ts(tmem, tmem + 256 + ks * 8, db, idesc, 1);
        } else {
          uint64_t da = make_smem_desc(smem_u32(smem) + ks * 32, 16, 1024, 2);
          umma_i8_ss(tmem, da, db, idesc, 1);
        }
      }
      umma_commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), phase);
      phase ^= 1;
    }
  }
  __syncthreads();
  tc_fence_after();
  if (w == 0) tmem_dealloc(tmem, 512);
}

static uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

int main(int argc, char** argv) {
  int mode = argc > 1 ? atoi(argv[1]) : 0;
  int N = argc > 2 ? atoi(argv[2]) : 64;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s sm_%d%d SMs=%d mode=%d N=%d\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount, mode, N);

  if (mode >= 10) {
    const int smem_bytes = 16384 + 32768 + 1024;
    auto run = [&](auto kern, const char* name) {
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
      const int iters = 512;
      kern<<<prop.multiProcessorCount, 128, smem_bytes>>>(8, N);
      CK(cudaDeviceSynchronize());
      cudaEvent_t e0, e1;
      CK(cudaEventCreate(&e0));
      CK(cudaEventCreate(&e1));
      float best = 1e30f;
      for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0));
        kern<<<prop.multiProcessorCount, 128, smem_bytes>>>(iters, N);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
      }
      double ops = 2.0 * 128 * N * 32 * 64.0 * iters * prop.multiProcessorCount;
      printf("BENCH %s N=%d: %.3f ms  %.1f TOP/s (int8 dense, all SMs, cta_group::1)\n", name, N, best, ops / best * 1e-9);
    };
    if (mode == 10) run(bench_kernel<false>, "SS");
    else run(bench_kernel<true>, "TS");
    return 0;
  }

  std::vector<int8_t> hA(128 * KT), hB((size_t)N * KT);
  uint32_t s = 12345;
  for (auto& v : hA) v = (int8_t)(lcg(s) & 0xFF);
  for (auto& v : hB) v = (int8_t)(lcg(s) & 0xFF);
  std::vector<int32_t> ref((size_t)128 * N), hD((size_t)128 * N);
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      int acc = 0;
      for (int k = 0; k < KT; ++k) acc += (int)hA[m * KT + k] * (int)hB[n * KT + k];
      ref[(size_t)m * N + n] = acc;
    }
  int8_t *dA, *dB;
  int32_t* dD;
  uint32_t* dDump;
  CK(cudaMalloc(&dA, hA.size()));
  CK(cudaMalloc(&dB, hB.size()));
  CK(cudaMalloc(&dD, hD.size() * 4));
  CK(cudaMalloc(&dDump, 128 * 32 * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xFF, hD.size() * 4));
  CK(cudaMemset(dDump, 0, 128 * 32 * 4));

  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  {
    auto enc = get_encode_tiled();
    cuuint64_t gdim[2] = {(cuuint64_t)KT, (cuuint64_t)N};
    cuuint64_t gstr[1] = {(cuuint64_t)KT};
    cuuint32_t box[2] = {(cuuint32_t)KT, (cuuint32_t)N};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, dB, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { fprintf(stderr, "encodeTiled failed %d\n", (int)r); return 2; }
  }
  const int smem_bytes = 16384 + 32768 + 1024;
  auto launch = [&](auto kern) {
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    kern<<<1, 128, smem_bytes>>>(dA, dB, dD, dDump, N, tmap);
  };
  switch (mode) {
    case 0: launch(probe_kernel<0>); break;
    case 1: launch(probe_kernel<1>); break;
    case 2: launch(probe_kernel<2>); break;
    case 3: launch(probe_kernel<3>); break;
    case 4: launch(probe_kernel<4>); break;
    case 5: launch(probe_kernel<5>); break;
    default: fprintf(stderr, "bad mode\n"); return 2;
  }
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
  size_t bad = 0;
  for (size_t i = 0; i < hD.size(); ++i) bad += (hD[i] != ref[i]);
  printf("RESULT mode=%d N=%d mismatches=%zu/%zu  %s\n", mode, N, bad, hD.size(), bad ? "FAIL" : "PASS");
  if (bad) {
    for (int m = 0; m < 4; ++m) {
      printf(" row %d got:", m);
      for (int n = 0; n < 8; ++n) printf(" %d", hD[(size_t)m * N + n]);
      printf("  ref:");
      for (int n = 0; n < 8; ++n) printf(" %d", ref[(size_t)m * N + n]);
      printf("\n");
    }
    // which rows are right?
    int okrows = 0;
    for (int m = 0; m < 128; ++m) {
      bool ok = true;
      for (int n = 0; n < N; ++n) ok &= hD[(size_t)m * N + n] == ref[(size_t)m * N + n];
      okrows += ok;
    }
    printf(" rows fully correct: %d/128\n", okrows);
  }
  if (mode == 2 || mode == 3) {
    std::vector<uint32_t> dump(128 * 32);
    CK(cudaMemcpy(dump.data(), dDump, dump.size() * 4, cudaMemcpyDeviceToHost));
    size_t badA = 0;
    for (int r = 0; r < 128; ++r)
      for (int c = 0; c < 32; ++c) {
        uint32_t want;
        memcpy(&want, &hA[r * KT + c * 4], 4);
        badA += dump[r * 32 + c] != want;
      }
    printf(" TMEM-A readback (32x32b view) mismatches=%zu/4096\n", badA);
  }
  return bad ? 1 : 0;
}
