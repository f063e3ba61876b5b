// W4A8 GEMM for B200 (sm_100a):  D[M,N] fp16 = ((int32)(A8[M,K] x W8[K,N]) * s2[n]) * s1[m]
//
// Replaces the reference's Marlin-derived kernel (csrc/qqq_gemm.cu:240-820) behind the same boundary; it is
// NOT a port.  The reference keeps weights in registers for mma.sync; here the GEMM is "swap-AB" on the
// 5th-gen tensor cores:
//
//   UMMA-M (128 TMEM lanes)  = 128 output channels of the weight tile
//   UMMA-N (8..256 columns)  = the token tile (the whole batch at decode)
//   UMMA-K = 32 int8, 4 per 128-deep k-block
//
//   TMA producer warp : packed int4 tile [8 x 1 KB] straight from the reference layout (B int32 [K/16,2N]),
//                       int8 token tile [n_tok x 128 B] (128B swizzle), group scales row (per-group only)
//   8 unpack warps    : LDS.64 packed words -> int8 (per-channel: 2 logic ops per word; per-group: exact
//                       fp16 FMA rounding of the reference, csrc/qqq_gemm.cu:167-210) -> tcgen05.st.16x128b
//                       into a TMEM ring.  The reference word layout (one word = 4 k x {n, n+8}) IS the
//                       16x128b store fragment, so the weight tile becomes the UMMA A operand in TMEM
//                       without any shuffle, shared-memory round trip or load-time repack.
//   MMA warp (1 thr)  : tcgen05.mma.cta_group::1.kind::i8, A from TMEM, B (tokens) from smem descriptor,
//                       int32 accumulators in TMEM (double-buffered when n_tok <= 128)
//   4 epilogue warps  : tcgen05.ld -> fp32 scale by s2[n] then s1[m] (reference order, :695-700) -> fp16 -> D
//
// Scheduling: persistent CTAs, static stream-K over (tile, k-block) units.  A tile whose k-range is shared by
// several CTAs is reduced through `C`: every contributor stores its int32 partial tile into its own slot (plain
// coalesced stores), the last CTA to arrive (lock counter in `workspace`) sums the slots in a fixed order (exact,
// deterministic), applies the scales, writes D and resets the lock.
#include "qqq_common.cuh"
#include "qqq_gemm_sm100.h"

namespace qqq {

namespace {

constexpr int kThreads = 448;  // 14 warps: 0 TMA, 1 MMA, 2-9 unpack, 10-13 epilogue
constexpr int kUnpackWarp0 = 2;
constexpr int kEpiWarp0 = 10;
constexpr int kTmemColsA0 = 256;  // A ring lives in columns [256, 512)

struct Ring {
  int idx = 0;
  uint32_t phase = 0;
  int n;
  __device__ explicit Ring(int n_) : n(n_) {}
  __device__ __forceinline__ void advance() {
    if (++idx == n) {
      idx = 0;
      phase ^= 1;
    }
  }
};

// position of channel n inside the permuted s_channel vector (reference scale_perm_single,
// QQQ/gptq/qlinear/qlinear_marlin.py:173-175): permuted[32b + 8i + jj] = natural[32b + 2i + {0,1,8,9,16,17,24,25}[jj]]
__device__ __forceinline__ int s2_position(int n) {
  const int v = n & 31;
  return (n & ~31) + 8 * ((v & 7) >> 1) + 2 * (v >> 3) + (v & 1);
}

// ---- nibble -> int8 -------------------------------------------------------------------------------------
// per-channel (reference dequant_per_channel, :146-151 / :540-542): the nibble stays in the high half of the
// byte, W8 = 16*w4.   blk0 (channel c): odd nibbles;  blk1 (channel c+8): even nibbles.
__device__ __forceinline__ void unpack_pc(uint32_t w, uint32_t& b0, uint32_t& b1) {
  b0 = w & 0xF0F0F0F0u;
  b1 = (w << 4) & 0xF0F0F0F0u;
}
// per-group (reference dequant_per_group, :167-210): W8 = RNE((v-8)*s) via ONE fp16 FMA per pair so the
// rounding is identical.  We add 1280 (0x6500) instead of the reference's 1152 (0x6480): 1280 = 1024+256, so
// the low mantissa byte is RNE((v-8)*s) mod 256, i.e. already two's complement — the reference's final
// `^ 0x80808080` disappears.  Identical for every |RNE((v-8)*s)| <= 128, which pack() guarantees (:209-217).
__device__ __forceinline__ uint32_t unpack_pg4(uint32_t w, uint32_t s2h) {
  // w holds the 4 nibbles of ONE channel at bits [0,4) [16,20) (k r0,r1) and [4,8) [20,24) (k r2,r3)
  uint32_t lo = (w & 0x000F000Fu) | 0x64006400u;  // (1024 + v_r0, 1024 + v_r1)
  uint32_t hi = (w & 0x00F000F0u) | 0x64006400u;  // (1024 + 16 v_r2, 1024 + 16 v_r3)
  const uint32_t k1032 = 0x64086408u, k1_16 = 0x2C002C00u, km72 = 0xD480D480u, k1280 = 0x65006500u;
  __half2 x01 = __hsub2(*reinterpret_cast<__half2*>(&lo), *reinterpret_cast<const __half2*>(&k1032));
  __half2 x23 = __hfma2(*reinterpret_cast<__half2*>(&hi), *reinterpret_cast<const __half2*>(&k1_16),
                        *reinterpret_cast<const __half2*>(&km72));
  const __half2 s = *reinterpret_cast<const __half2*>(&s2h);
  __half2 y01 = __hfma2(x01, s, *reinterpret_cast<const __half2*>(&k1280));
  __half2 y23 = __hfma2(x23, s, *reinterpret_cast<const __half2*>(&k1280));
  uint32_t out;
  asm("prmt.b32 %0, %1, %2, 0x6420;" : "=r"(out) : "r"(*reinterpret_cast<uint32_t*>(&y01)), "r"(*reinterpret_cast<uint32_t*>(&y23)));
  return out;
}

template <bool GROUPED>
__global__ void __launch_bounds__(kThreads, 1)
qqq_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled TMA/UMMA tiles need 1024-byte alignment
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int NS = p.num_stages;
  const int KSUB = p.ksub;                 // 128-deep k sub-blocks per pipeline stage ("unit")
  const int tok_bytes = p.n_tok * 128;     // one sub-block of tokens
  const int stage_b = KSUB * kStageB, stage_t = KSUB * tok_bytes, stage_s = KSUB * kStageS;
  uint8_t* sB = smem;
  uint8_t* sT = sB + NS * stage_b;
  uint8_t* sS = sT + NS * stage_t;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sS + NS * stage_s);
  const uint32_t bar_full = smem_u32(bars);
  const uint32_t bar_empty = bar_full + 8 * NS;
  const uint32_t bar_afull = bar_empty + 8 * NS;
  const uint32_t bar_aempty = bar_afull + 8 * kASlots;
  const uint32_t bar_dfull = bar_aempty + 8 * kASlots;
  const uint32_t bar_dempty = bar_dfull + 8 * 2;
  uint32_t* misc = reinterpret_cast<uint32_t*>(bars + 2 * NS + 2 * kASlots + 4);  // [0] tmem base, [1] "last" flag
  float* s1_sm = reinterpret_cast<float*>(misc + 4);                               // [kMaxTok] per-token scales

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KB = p.k_units;  // units (KSUB sub-blocks each) per tile
  const int u_begin = min((long long)blockIdx.x * p.units_per_cta, (long long)p.total_units);
  const int u_end = min((long long)u_begin + p.units_per_cta, (long long)p.total_units);
  const int ndbuf = p.n_tok <= 128 ? 2 : 1;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int i = 0; i < NS; ++i) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, 1);
    }
    for (int i = 0; i < kASlots; ++i) {
      mbar_init(bar_afull + 8 * i, 4);
      mbar_init(bar_aempty + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_dfull + 8 * i, 1);
      mbar_init(bar_dempty + 8 * i, 4);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&misc[0]), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = misc[0];

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    // The whole warp runs the loop (uniform control flow keeps descriptors in uniform registers); one elected
    // lane issues.
    Ring st(NS);
    for (int u = u_begin; u < u_end; ++u) {
      const int tile = u / KB, kb = u - tile * KB;
      const int mt = tile % p.m_tiles, nt = tile / p.m_tiles;
      mbar_wait(bar_empty + 8 * st.idx, st.phase ^ 1);
      if (elect_one()) {
        const uint32_t full = bar_full + 8 * st.idx;
        // k sub-blocks past the end of K are zero-filled by TMA (weights and tokens), so they add nothing
        uint32_t sbytes = 0;
        int nsub_valid = KSUB;
        if (GROUPED) {
          sbytes = (uint32_t)min(kTileN, p.N - nt * kTileN) * 2u;
          nsub_valid = min(KSUB, p.k_blocks - kb * KSUB);
        }
        mbar_expect_tx(full, stage_b + stage_t + (GROUPED ? sbytes * nsub_valid : 0u));
        tma_load_2d(smem_u32(sB + st.idx * stage_b), &tmap_b, full, nt * (2 * kTileN), kb * KSUB * 8, p.hint_b);
        for (int sub = 0; sub < KSUB; ++sub) {
          tma_load_2d(smem_u32(sT + st.idx * stage_t + sub * tok_bytes), &tmap_a, full, (kb * KSUB + sub) * kBlockK,
                      mt * p.n_tok, p.hint_a);
          if (GROUPED && sub < nsub_valid)
            bulk_load_1d(smem_u32(sS + st.idx * stage_s + sub * kStageS),
                         p.s3 + (size_t)(kb * KSUB + sub) * p.N + nt * kTileN, sbytes, full);
        }
      }
      __syncwarp();
      st.advance();
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    Ring st(NS), as(kASlots);
    const uint32_t idesc = make_idesc_i8(kTileN, p.n_tok);
    const uint64_t desc_tok0 = make_smem_desc(smem_u32(sT), 16, 1024, 2);  // + byte offset >> 4 per stage / k-step
    int seg = 0;
    for (int u = u_begin; u < u_end; ++seg) {
      const int tile = u / KB, kb0 = u - tile * KB;
      const int kb1 = min(KB, kb0 + (u_end - u));
      const int dbuf = seg % ndbuf;
      const uint32_t dph = (seg / ndbuf) & 1;
      mbar_wait(bar_dempty + 8 * dbuf, dph ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + dbuf * p.n_tok;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(bar_full + 8 * st.idx, st.phase);
        for (int sub = 0; sub < KSUB; ++sub) {
          mbar_wait(bar_afull + 8 * as.idx, as.phase);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t desc = desc_tok0 + (uint64_t)((st.idx * stage_t + sub * tok_bytes) >> 4);
            const uint32_t tmem_a = tmem_base + kTmemColsA0 + as.idx * 32;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_i8_ts(tmem_d, tmem_a + ks * 8, desc + 2 * ks, idesc, (kb > kb0 || sub > 0 || ks > 0) ? 1u : 0u);
            umma_commit(bar_aempty + 8 * as.idx);
            if (sub == KSUB - 1) {
              umma_commit(bar_empty + 8 * st.idx);
              // same thread as the MMAs above: tcgen05.commit tracks the issuing thread's operations
              if (kb == kb1 - 1) umma_commit(bar_dfull + 8 * dbuf);
            }
          }
          __syncwarp();
          as.advance();
        }
        st.advance();
      }
      u += kb1 - kb0;
    }
  } else if (warp < kEpiWarp0) {
    // ===================================== unpack warps =====================================
    // Two groups of 4 warps alternate k-blocks; inside a group warp <-> TMEM lane quadrant q = warp % 4:
    // channels [32q, 32q+32) of the tile = 64-channel block nb = q/2, 16-wide n-tiles j = 2(q%2) + {0,1}.
    const int grp = (warp - kUnpackWarp0) >> 2;
    const int q = warp & 3, nb = q >> 1, jp = q & 1;
    const int c = lane >> 2;
    Ring st(NS), as(kASlots);
    int it = 0;  // running sub-block index: the two groups take alternate sub-blocks
    for (int u = u_begin; u < u_end; ++u) {
      mbar_wait(bar_full + 8 * st.idx, st.phase);
      for (int sub = 0; sub < KSUB; ++sub, ++it) {
        if ((it & 1) == grp) {
          mbar_wait(bar_aempty + 8 * as.idx, as.phase ^ 1);
          tc_fence_after();
          const uint8_t* src = sB + st.idx * stage_b + sub * kStageB + nb * 512 + lane * 16 + jp * 8;
          uint2 w[8];
#pragma unroll
          for (int kt = 0; kt < 8; ++kt) w[kt] = *reinterpret_cast<const uint2*>(src + kt * 1024);
          uint2 sc = make_uint2(0, 0);
          if (GROUPED)
            sc = *reinterpret_cast<const uint2*>(sS + st.idx * stage_s + sub * kStageS + nb * 128 + c * 16 + jp * 8);
          const uint32_t tmem_a = tmem_base + kTmemColsA0 + as.idx * 32 + ((uint32_t)(32 * q) << 16);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t s_b0 = 0, s_b1 = 0;
            if (GROUPED) {
              const uint32_t sp = h ? sc.y : sc.x;  // half2: (scale of channel c [blk0], scale of channel c+8 [blk1])
              s_b0 = __byte_perm(sp, sp, 0x1010);
              s_b1 = __byte_perm(sp, sp, 0x3232);
            }
#pragma unroll
            for (int part = 0; part < 2; ++part) {
              uint32_t r[8];
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const uint32_t word = h ? w[4 * part + g].y : w[4 * part + g].x;
                if (GROUPED) {
                  r[2 * g] = unpack_pg4(word, s_b0);
                  r[2 * g + 1] = unpack_pg4(word >> 8, s_b1);
                } else {
                  unpack_pc(word, r[2 * g], r[2 * g + 1]);
                }
              }
              tmem_st_16x128b_x4(tmem_a + ((uint32_t)(16 * h) << 16) + part * 16, r);
            }
          }
          tmem_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_afull + 8 * as.idx);
        }
        as.advance();
      }
      st.advance();
    }
  } else {
    // ===================================== epilogue warps ===================================
    const int q = warp & 3;
    const int epi_tid = threadIdx.x - kEpiWarp0 * 32;
    const int m_pad = p.m_tiles * p.n_tok;  // rows of one split-K slot in C
    int seg = 0, staged_mt = -1;
    for (int u = u_begin; u < u_end; ++seg) {
      const int tile = u / KB, kb0 = u - tile * KB;
      const int kb1 = min(KB, kb0 + (u_end - u));
      const int mt = tile % p.m_tiles, nt = tile / p.m_tiles;
      const int dbuf = seg % ndbuf;
      const uint32_t dph = (seg / ndbuf) & 1;
      const int n = nt * kTileN + 32 * q + lane;
      const bool n_ok = n < p.N;
      const int m0 = mt * p.n_tok;
      const int rows = min(p.n_tok, p.M - m0);  // valid token rows of this tile
      const bool whole = (kb0 == 0 && kb1 == KB);
      const int first_cta = (tile * KB) / p.units_per_cta;
      const int parts = (tile * KB + KB - 1) / p.units_per_cta - first_cta + 1;
      const int part = (int)blockIdx.x - first_cta;
      const float s2v = n_ok ? __ldg(p.s2 + s2_position(n)) : 0.f;
      __half* __restrict__ dcol = p.D + n;
      int* __restrict__ ccol = p.C + n;

      // per-token scales of this token tile -> smem (once per tile change), so the store loop has no global loads
      if (mt != staged_mt) {
        named_bar_sync(1, 128);  // previous users of s1_sm are done
        for (int i = epi_tid; i < p.n_tok; i += 128) s1_sm[i] = (m0 + i < p.M) ? __ldg(p.s1 + m0 + i) : 0.f;
        named_bar_sync(1, 128);
        staged_mt = mt;
      }

      mbar_wait(bar_dfull + 8 * dbuf, dph);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + dbuf * p.n_tok + ((uint32_t)(32 * q) << 16);
      for (int c16 = 0; c16 * 16 < rows; ++c16) {
        uint32_t r[16];
        tmem_ld_32x32b_x16(tmem_d + c16 * 16, r);
        tmem_wait_ld();
        if (n_ok) {
          const int mb = c16 * 16;
          if (whole) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              if (mb + i < rows) {
                const float v = (__int2float_rn((int)r[i]) * s2v) * s1_sm[mb + i];
                dcol[(size_t)(m0 + mb + i) * p.N] = __float2half_rn(v);
              }
            }
          } else {
            // split-K: this CTA's partial sums go to its own slot of C (plain coalesced stores, no atomics)
            int* __restrict__ slot = ccol + (size_t)(part * m_pad + m0 + mb) * p.N;
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (mb + i < rows) slot[(size_t)i * p.N] = (int)r[i];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_dempty + 8 * dbuf);  // accumulator buffer may be overwritten

      if (!whole) {
        // The last CTA to arrive on the tile's lock sums the slots in a fixed order (integer: exact), applies the
        // scales, writes D and resets the lock.  C itself needs neither zeroing nor restoring.
        int* lock = p.locks + nt + p.n_tiles * mt;
        __threadfence();
        named_bar_sync(1, 128);
        if (epi_tid == 0) {
          const int old = atomicAdd(lock, 1);
          misc[1] = (old == parts - 1) ? 1u : 0u;
        }
        named_bar_sync(1, 128);
        const bool last = misc[1] != 0;
        if (last) {
          __threadfence();
          if (n_ok) {
            for (int i0 = 0; i0 < rows; i0 += 16) {
              int acc[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) acc[j] = 0;
              for (int pp = 0; pp < parts; ++pp) {
                const int* __restrict__ src = ccol + (size_t)(pp * m_pad + m0 + i0) * p.N;
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  if (i0 + j < rows) acc[j] += __ldcg(src + (size_t)j * p.N);
              }
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                if (i0 + j < rows) {
                  const float v = (__int2float_rn(acc[j]) * s2v) * s1_sm[i0 + j];
                  dcol[(size_t)(m0 + i0 + j) * p.N] = __float2half_rn(v);
                }
              }
            }
          }
          if (epi_tid == 0) *lock = 0;
        }
        named_bar_sync(1, 128);  // misc[1] is reused by the next segment
      }
      u += kb1 - kb0;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace

size_t gemm_smem_bytes(int num_stages, int n_tok, int ksub) {
  return 1024 + (size_t)num_stages * ksub * (kStageB + n_tok * 128 + kStageS) + 8 * (2 * num_stages + 2 * kASlots + 4) +
         16 + 4 * kMaxTok;
}

cudaError_t launch_gemm(const CUtensorMap& tmap_a, const CUtensorMap& tmap_b, const GemmParams& p, bool grouped,
                        int grid, int dev, cudaStream_t stream) {
  static bool attr_set[2][64] = {};  // the opt-in shared-memory attribute is per device
  const size_t smem = gemm_smem_bytes(p.num_stages, p.n_tok, p.ksub);
  auto kern = grouped ? qqq_gemm_kernel<true> : qqq_gemm_kernel<false>;
  if (dev < 0 || dev >= 64 || !attr_set[grouped][dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemBytes);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) attr_set[grouped][dev] = true;
  }
  kern<<<grid, kThreads, smem, stream>>>(tmap_a, tmap_b, p);
  return cudaGetLastError();
}

}  // namespace qqq
