// W4A8 GEMM for B200 (sm_100a):  D[M,N] fp16 = ((int32)(A8[M,K] x W8[K,N]) * s2[n]) * s1[m]
//
// Replaces the reference's Marlin-derived kernel (csrc/qqq_gemm.cu:240-820) behind the same boundary; it is
// NOT a port.  The reference keeps weights in registers for mma.sync; here the GEMM is "swap-AB" on the
// 5th-gen tensor cores:
//
//   UMMA-M (128 TMEM lanes)  = 128 output channels of the weight tile
//   UMMA-N (16..256 columns) = the token tile (the whole batch at decode)
//   UMMA-K = 32 int8, 4 per 128-deep k sub-block
//
//   warp 0  weights producer : TMA of the packed int4 tile [8*KSUB rows x 1 KB] straight from the reference
//                              layout (B int32 [K/16,2N]) + the group-scale rows (per-group only)  -> ring W
//   warp 2  tokens producer  : TMA of the int8 token tile [n_tok x 128 B] (128B swizzle)           -> ring T
//   8-12 unpack warps (2-3 groups of 4, one warp per TMEM lane quadrant): LDS.64 packed words -> int8 in registers
//                              (per-channel: 2 logic ops per word; per-group: the reference's exact fp16-FMA
//                              rounding, csrc/qqq_gemm.cu:167-210) -> tcgen05.st.16x128b into a TMEM ring.
//                              The reference word layout (one word = 4 k x {n, n+8}) IS the 16x128b store
//                              fragment, so the weight tile becomes the UMMA A operand in TMEM without any
//                              shuffle, shared-memory round trip or load-time repack.
//   warp 1  MMA issuer       : tcgen05.mma.cta_group::1.kind::i8, A from TMEM, B (tokens) from a smem descriptor,
//                              int32 accumulators in TMEM (double-buffered when n_tok <= kDbufMaxTok = 208)
//   4-8 epilogue warps       : tcgen05.ld -> fp32 * s2[n] * s1[m] (reference order, :695-700) -> fp16 -> D
//
// The two smem rings are decoupled: weight stages are released by the unpack warps as soon as they are in
// registers, token stages by tcgen05.commit when the MMAs that read them retire; the TMEM ring lets the unpack
// run several k-blocks ahead of the MMA.
//
// Scheduling: persistent CTAs, static stream-K over (tile, k-unit) units.  A tile whose k-range is shared by
// several CTAs is reduced through `C`: every contributor stores its int32 partial tile into its own slot (plain
// coalesced stores), the last CTA to arrive (lock counter in `workspace`) sums the slots in a fixed order (exact,
// deterministic), applies the scales, writes D and resets the lock.
//
// Template variants of the one kernel: kPair (cluster of 2, cta_group::2: each CTA loads half of the token tile, the
// leader issues UMMA M = 256 for both), kReduce (tensor-parallel row shards: the epilogue adds into a multicast D with
// multimem.red instead of storing), kAcc (raw int32 accumulators out, for the bit-exact tensor-parallel mode); kOutScatter
// (row goes to the partial-sum slot of the rank that owns it: the reduce-scatter half of the fused tensor-parallel exchange).
#include "qqq_common.cuh"
#include "qqq_gemm_sm100.h"

namespace qqq {

// Optional per-role timeline for development (probes/trace_timeline.py builds a separate library with -DQQQ_TRACE;
// the product build compiles these hooks away).
#ifdef QQQ_TRACE
__device__ unsigned long long* g_trace = nullptr;  // [20 roles][2048 events] of clock64()
#define QQQ_TR_INIT() unsigned long long* tr_ = (blockIdx.x == QQQ_TRACE_CTA) ? g_trace : nullptr
#define QQQ_TR(role, idx)                                                            \
  do {                                                                               \
    const int i_ = (idx);                                                            \
    if (tr_ != nullptr && i_ >= 0 && i_ < 2048) tr_[(role) * 2048 + i_] = clock64(); \
  } while (0)
#else
#define QQQ_TR_INIT()
#define QQQ_TR(role, idx)
#endif

namespace {

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

// Two-phase static schedule.  Phase A: the `a_tiles` remainder tiles that do not fill a whole wave are cut along K
// into `a_upc`-unit slices, one slice per CTA (stream-K; a slice may straddle a tile boundary) — processed FIRST, so
// their split-K fix-up overlaps the rest of the CTA's work.  Phase B: whole tiles, dealt round-robin: CTA c takes tiles
// a_tiles + c, a_tiles + c + grid, ...  Tile ids run token-tile-fastest (tile = mt + m_tiles * column), so at any moment
// neighbouring CTAs work on the token tiles of the SAME weight column: its packed weights come from DRAM once and from
// L2 for the other m_tiles - 1 readers (a CTA walking the token tiles of its own column one after the other re-read
// them from DRAM: 2.4x the algorithmic traffic at M = 1024, N = 21760).  Every warp role walks the same segment list.
struct Sched {
  int KU, a_begin, a_end, n_a, b_first, b_step, n_b;
  __device__ Sched(const GemmParams& p, int cta) {
    KU = p.k_units;
    a_begin = min(cta * p.a_upc, p.a_units);
    a_end = min(a_begin + p.a_upc, p.a_units);
    n_a = a_end > a_begin ? (a_end - 1) / KU - a_begin / KU + 1 : 0;
    b_first = p.a_tiles + cta;
    b_step = p.b_step;
    const int total = p.a_tiles + p.b_tiles;
    n_b = b_first < total ? (total - b_first + b_step - 1) / b_step : 0;
  }
  __device__ __forceinline__ int num_segments() const { return n_a + n_b; }
  __device__ __forceinline__ int num_units() const { return (a_end - a_begin) + n_b * KU; }
  __device__ __forceinline__ void segment(int i, int& tile, int& kb0, int& kb1) const {
    if (i < n_a) {
      tile = a_begin / KU + i;
      kb0 = i == 0 ? a_begin - tile * KU : 0;
      kb1 = min(KU, a_end - tile * KU);
    } else {
      tile = b_first + (i - n_a) * b_step;
      kb0 = 0;
      kb1 = KU;
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
__device__ __forceinline__ uint32_t and_or(uint32_t a, uint32_t b, uint32_t c) {  // (a & b) | c
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint32_t unpack_pg4(uint32_t w, uint32_t s2h) {
  // w holds the 4 nibbles of ONE channel at bits [0,4) [16,20) (k r0,r1) and [4,8) [20,24) (k r2,r3)
  // (w & mask) | 0x6400 as ONE three-input LOP3 with both constants in registers: with two immediates the
  // compiler emits an AND and an OR, and the ALU pipe (not the fp16 pipe) becomes the bound of the whole kernel
  uint32_t lo = and_or(w, 0x000F000Fu, 0x64006400u);  // (1024 + v_r0, 1024 + v_r1)
  uint32_t hi = and_or(w, 0x00F000F0u, 0x64006400u);  // (1024 + 16 v_r2, 1024 + 16 v_r3)
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

// How a finished 16-byte piece (8 channels of token row m) of the output leaves the SM.
//   kOutStore  : a plain store into D.
//   kOutReduce : (tensor-parallel row shards, qqq_gemm_reduce_sm100a) D is a MULTICAST address bound to the same buffer on
//                every rank; `multimem.red` adds the eight fp16 values into every replica (the NVSwitch fans the reduction
//                out), so once all ranks' kernels have finished every replica holds the all-reduced output — one-shot
//                all-reduce in the epilogue.
//   kOutScatter: (qqq_gemm_scatter_sm100a) the row goes to the rank that OWNS it (owner = m / tp_rows), into the slot of this
//                rank in the owner's partial-sum buffer [world][tp_rows][N] — a peer store over NVLink, issued tile by tile
//                while the GEMM runs: the reduce-scatter half of the exchange; tp_reduce_quant.cu is the other half.
//   kOutAcc    : raw int32 accumulators (handled in the epilogue loop, not here).
enum : int { kOutStore = 0, kOutReduce = 1, kOutAcc = 2, kOutScatter = 3 };

template <int kOut>
__device__ __forceinline__ __half* out_row(const GemmParams& p, int m) {
  if constexpr (kOut == kOutScatter) {
    const int owner = m / p.tp_rows;
    return p.tp_part[owner] + (size_t)(p.tp_rank * p.tp_rows + (m - owner * p.tp_rows)) * p.N;
  } else {
    return p.D + (size_t)m * p.N;
  }
}
__device__ __forceinline__ uint4 hadd2_x4(const uint4& a, const uint4& b) {
  uint4 o;
  const uint32_t* pa = &a.x;
  const uint32_t* pb = &b.x;
  uint32_t* po = &o.x;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __half2 h = __hadd2(*reinterpret_cast<const __half2*>(pa + i), *reinterpret_cast<const __half2*>(pb + i));
    po[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  return o;
}
template <int kOut>
__device__ __forceinline__ void emit_d(__half* dp, const uint4& v) {
  if constexpr (kOut == kOutReduce) {
    asm volatile("multimem.red.relaxed.sys.global.add.v4.f16x2 [%0], {%1, %2, %3, %4};" ::"l"(dp), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w)
                 : "memory");
  } else {
    *reinterpret_cast<uint4*>(dp) = v;
  }
}


// kAcc (qqq_gemm_acc_sm100a): the finished tile leaves as the raw int32 accumulators, row-major [M, N] int32, no scales —
// for the bit-exact tensor-parallel mode (int32 partial sums are all-reduced, the scales applied once afterwards).
template <bool GROUPED, bool kPair, int kOut = kOutStore>
__global__ void __launch_bounds__(kThreads, 1)
qqq_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                const GemmParams p) {
  constexpr bool kAcc = kOut == kOutAcc;
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled TMA/UMMA tiles need 1024-byte alignment
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int NSW = p.stages_w, NST = p.stages_t;
  const int KSUB = p.ksub;                  // 128-deep k sub-blocks per pipeline stage ("unit")
  // TMEM columns: accumulator buffers first (two when they leave >= 128 columns for the weight ring, so the drain of
  // one tile overlaps the MMAs of the next), then the ring of unpacked weight tiles (one slot = one unit = 32*KSUB columns)
  const int ndbuf = p.n_tok <= kDbufMaxTok ? 2 : 1;
  const int tmem_a0 = ndbuf * p.n_tok;
  constexpr int tmem_cols = 512;
  const int NA = min(kMaxASlots, (tmem_cols - tmem_a0) / (32 * KSUB));
  // CTA pair (cluster of 2, tcgen05 cta_group::2): the two CTAs take adjacent 128-channel tiles of the same token tile
  // and the same k-range; one MMA instruction of the leader (rank 0) drives both tensor cores (UMMA M = 256), every
  // CTA loads only its half of the token tile (rows [rank*n_tok/2, ...)), halving the L2->SM token traffic per SM.
  constexpr int PAIR = kPair ? 1 : 0;  // a template parameter: kernels with cta_group::2 instructions need a cluster launch
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const int tok_rows = PAIR ? p.n_tok / 2 : p.n_tok;  // token rows this CTA holds per sub-block
  const int tok_bytes = tok_rows * 128;               // one sub-block of tokens in this CTA's shared memory
  const int stage_w = KSUB * kStageB, stage_t = KSUB * tok_bytes, stage_s = KSUB * kStageS;
  uint8_t* sStage = smem;                   // epilogue staging: one [16][32] fp16 tile per epilogue warp
  uint8_t* sT = sStage + kEpiStageBytes;
  uint8_t* sW = sT + NST * stage_t;
  uint8_t* sS = sW + NSW * stage_w;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sS + NSW * stage_s);
  const uint32_t bar_fullw = smem_u32(bars);
  const uint32_t bar_emptyw = bar_fullw + 8 * NSW;
  const uint32_t bar_fullt = bar_emptyw + 8 * NSW;
  const uint32_t bar_emptyt = bar_fullt + 8 * NST;
  const uint32_t bar_afull = bar_emptyt + 8 * NST;
  const uint32_t bar_aempty = bar_afull + 8 * kMaxASlots;
  const uint32_t bar_dfull = bar_aempty + 8 * kMaxASlots;
  const uint32_t bar_dempty = bar_dfull + 8 * 2;
  uint32_t* misc = reinterpret_cast<uint32_t*>(bars + 2 * NSW + 2 * NST + 2 * kMaxASlots + 4);  // [0] tmem base, [1] flag
  float* s1_sm = reinterpret_cast<float*>(misc + 4);                                            // [kMaxTok]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = p.unpack_groups;                 // 2 or 3 groups of unpack warps (host policy in qqq_c_api.cu)
  const int epi_warp0 = kUnpackWarp0 + 4 * G;    // warps [epi_warp0, kWarps) drain the accumulators
  const int n_epi = kWarps - epi_warp0;          // 8 or 4 epilogue warps
  const int n_epi_thr = 32 * n_epi;
  const int KU = p.k_units;  // units per tile
  const Sched sched(p, (int)blockIdx.x >> PAIR);  // the schedule is over (super-)tiles: both CTAs of a pair walk it
  const int n_seg = sched.num_segments();
  QQQ_TR_INIT();

  grid_launch_dependents();
  if (warp == 0) {
    // barrier init spread over the lanes of warp 0
    if (lane == 0) {
      tma_prefetch_desc(&tmap_a);
      tma_prefetch_desc(&tmap_b);
    }
    for (int i = lane; i < NSW; i += 32) {
      mbar_init(bar_fullw + 8 * i, 1);
      mbar_init(bar_emptyw + 8 * i, 4 * KSUB);
    }
    for (int i = lane; i < NST; i += 32) {
      mbar_init(bar_fullt + 8 * i, 1);
      mbar_init(bar_emptyt + 8 * i, 1);
    }
    // pair: the leader's weight-full and accumulator-empty barriers collect the arrivals of both CTAs
    for (int i = lane; i < kMaxASlots; i += 32) {
      mbar_init(bar_afull + 8 * i, (4 * KSUB) << PAIR);
      mbar_init(bar_aempty + 8 * i, 1);
    }
    if (lane < 2) {
      mbar_init(bar_dfull + 8 * lane, 1);
      mbar_init(bar_dempty + 8 * lane, n_epi << PAIR);
    }
    mbar_fence_init();
    __syncwarp();
  }
  // Weight stages that fit the (still empty) ring are requested right away, before TMEM allocation and the
  // CTA-wide barrier: the first DRAM round trip overlaps the rest of the prologue.
  auto issue_weights = [&](int stage, int nt, int kb) {
    const uint32_t full = bar_fullw + 8 * stage;
    uint32_t sbytes = 0;
    int nsub_valid = 0;
    if (GROUPED) {
      sbytes = (uint32_t)min(kTileN, p.N - nt * kTileN) * 2u;
      nsub_valid = min(KSUB, p.k_blocks - kb * KSUB);
    }
    mbar_expect_tx(full, stage_w + sbytes * nsub_valid);
    tma_load_2d(smem_u32(sW + stage * stage_w), &tmap_b, full, nt * (2 * kTileN), kb * KSUB * 8, p.hint_b);
    if (GROUPED) {
      for (int sub = 0; sub < nsub_valid; ++sub)
        bulk_load_1d(smem_u32(sS + stage * stage_s + sub * kStageS),
                     p.s3 + (size_t)(kb * KSUB + sub) * p.N + nt * kTileN, sbytes, full);
    }
  };
  // Weight producer state (warp 0): the first ring of stages is requested before TMEM allocation and the CTA-wide
  // barrier, the rest in the role loop below.
  // a scheduled tile is (n super-tile, token tile): this CTA's 128-channel tile is n-tile 2*snt + rank of a pair
  auto nt_of = [&](int tile) { return ((tile / p.m_tiles) << PAIR) + (int)rank; };
  int w_seg = 0, w_tile = 0, w_kb = 0, w_kb1 = 0, w_count = 0;
  Ring w_st(NSW);
  auto w_next = [&]() {  // advance to the next unit; false when the CTA's work is exhausted
    while (w_kb >= w_kb1) {
      if (w_seg >= n_seg) return false;
      sched.segment(w_seg++, w_tile, w_kb, w_kb1);
    }
    return true;
  };
  if (warp == 0) {
    while (w_count < NSW && w_next()) {
      if (elect_one()) issue_weights(w_st.idx, nt_of(w_tile), w_kb);
      __syncwarp();
      w_st.advance();
      ++w_kb;
      ++w_count;
    }
  }
  if (warp == 1) {
    if (PAIR)
      tmem_alloc_pair(smem_u32(&misc[0]), tmem_cols);
    else
      tmem_alloc(smem_u32(&misc[0]), tmem_cols);
  }
  tc_fence_before();
  if (PAIR)
    cluster_sync_all();  // the peer's barriers are initialised and its TMEM is allocated before anything remote
  else
    __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = misc[0];
  // pair: barriers of the leader CTA that this CTA signals
  const uint32_t lead_fullt = PAIR ? mapa_shared(bar_fullt, 0) : bar_fullt;
  const uint32_t lead_afull = PAIR ? mapa_shared(bar_afull, 0) : bar_afull;
  const uint32_t lead_dempty = PAIR ? mapa_shared(bar_dempty, 0) : bar_dempty;
  if (threadIdx.x == 0) QQQ_TR(12, 0);

  if (warp == 0) {
    // ===================================== weights producer =================================
    // The whole warp runs the loop (uniform control flow keeps descriptors in uniform registers); one elected
    // lane issues.  k sub-blocks past the end of K are zero-filled by TMA, so they add nothing.
    while (w_next()) {
      mbar_wait(bar_emptyw + 8 * w_st.idx, w_st.phase ^ 1);
      if (elect_one()) {
        QQQ_TR(0, w_count);
        issue_weights(w_st.idx, nt_of(w_tile), w_kb);
        QQQ_TR(1, w_count);
      }
      __syncwarp();
      w_st.advance();
      ++w_kb;
      ++w_count;
    }
  } else if (warp == 2) {
    // ===================================== tokens producer ==================================
    grid_dependency_wait();  // A8 is produced by the preceding kernel (activation quant); weights are not
    Ring st(NST);
#ifdef QQQ_TRACE
    int t_count = 0;
#endif
    for (int sg = 0; sg < n_seg; ++sg) {
      int tile, kb0, kb1;
      sched.segment(sg, tile, kb0, kb1);
      const int mt = tile % p.m_tiles;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(bar_emptyt + 8 * st.idx, st.phase ^ 1);
        if (elect_one()) {
          QQQ_TR(18, t_count++);
          const uint32_t full = bar_fullt + 8 * st.idx;
          if (!PAIR) {
            mbar_expect_tx(full, stage_t);
            for (int sub = 0; sub < KSUB; ++sub)
              tma_load_2d(smem_u32(sT + st.idx * stage_t + sub * tok_bytes), &tmap_a, full, (kb * KSUB + sub) * kBlockK,
                          mt * p.n_tok, p.hint_a);
          } else {
            // both halves of the token tile are signalled on the leader's barrier (the MMA issuer waits there)
            if (rank == 0) mbar_expect_tx(full, 2 * stage_t);
            for (int sub = 0; sub < KSUB; ++sub)
              tma_load_2d_pair(smem_u32(sT + st.idx * stage_t + sub * tok_bytes), &tmap_a, lead_fullt + 8 * st.idx,
                               (kb * KSUB + sub) * kBlockK, mt * p.n_tok + (int)rank * tok_rows, p.hint_a);
          }
        }
        __syncwarp();
        st.advance();
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ===================================== MMA issuer (pair: the leader CTA only) ===========
    Ring st(NST), as(NA);
    const uint32_t idesc = make_idesc_i8(kTileN << PAIR, p.n_tok);
    const uint64_t desc_tok0 = make_smem_desc(smem_u32(sT), 16, 1024, 2);  // + (byte offset >> 4) per stage / k-step
    int ucount = 0;
    for (int seg = 0; seg < n_seg; ++seg) {
      int tile, kb0, kb1;
      sched.segment(seg, tile, kb0, kb1);
      const int dbuf = seg % ndbuf;
      const uint32_t dph = (seg / ndbuf) & 1;
      mbar_wait(bar_dempty + 8 * dbuf, dph ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + dbuf * p.n_tok;
      for (int kb = kb0; kb < kb1; ++kb, ++ucount) {
#ifdef QQQ_TRACE  // which operand the issuer waits for: tokens landed (16), then weights unpacked (17)
        mbar_wait(bar_fullt + 8 * st.idx, st.phase);
        if (lane == 0) QQQ_TR(16, ucount);
        mbar_wait(bar_afull + 8 * as.idx, as.phase);
        if (lane == 0) QQQ_TR(17, ucount);
#else
        mbar_wait2(bar_fullt + 8 * st.idx, st.phase, bar_afull + 8 * as.idx, as.phase);
#endif
        tc_fence_after();
        if (elect_one()) {
          QQQ_TR(5, ucount);
          uint64_t desc = desc_tok0 + (uint64_t)((st.idx * stage_t) >> 4);
          uint32_t tmem_a = tmem_base + tmem_a0 + as.idx * 32 * KSUB;
          for (int sub = 0; sub < KSUB; ++sub) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint32_t acc = (kb > kb0 || sub > 0 || ks > 0) ? 1u : 0u;
              if (PAIR)
                umma_i8_ts_pair(tmem_d, tmem_a + ks * 8, desc + 2 * ks, idesc, acc);
              else
                umma_i8_ts(tmem_d, tmem_a + ks * 8, desc + 2 * ks, idesc, acc);
            }
            desc += (uint64_t)(tok_bytes >> 4);
            tmem_a += 32;
          }
          // same thread as the MMAs above: tcgen05.commit tracks the issuing thread's operations
          if (PAIR) {  // the barriers at these offsets in BOTH CTAs of the pair
            umma_commit_pair(bar_aempty + 8 * as.idx, 3);
            umma_commit_pair(bar_emptyt + 8 * st.idx, 3);
            if (kb == kb1 - 1) umma_commit_pair(bar_dfull + 8 * dbuf, 3);
          } else {
            umma_commit(bar_aempty + 8 * as.idx);
            umma_commit(bar_emptyt + 8 * st.idx);
            if (kb == kb1 - 1) umma_commit(bar_dfull + 8 * dbuf);
          }
          QQQ_TR(6, ucount);
        }
        st.advance();
        as.advance();
      }
    }
  } else if (warp >= kUnpackWarp0 && warp < epi_warp0) {
    // ===================================== unpack warps =====================================
    // G groups of 4 warps take k sub-blocks round-robin; inside a group warp <-> TMEM lane quadrant
    // q = warp % 4: channels [32q, 32q+32) of the tile = 64-channel block nb = q/2, 16-wide n-tiles j = 2(q%2)+{0,1}.
    const int grp = (warp - kUnpackWarp0) >> 2;
    const int q = warp & 3, nb = q >> 1, jp = q & 1;
    const int c = lane >> 2;
    Ring st(NSW), as(NA);
    int turn = 0;  // which group owns the next sub-block
    int itn = 0;
    const int n_units = sched.num_units();
    for (int u = 0; u < n_units; ++u) {
      bool stage_ready = false, slot_ready = false;
      for (int sub = 0; sub < KSUB; ++sub, ++itn) {
        if (turn == grp) {
          if (!stage_ready) {
            mbar_wait(bar_fullw + 8 * st.idx, st.phase);
            stage_ready = true;
          }
          if (q == 0 && lane == 0) QQQ_TR(2, itn);
          const uint8_t* src = sW + st.idx * stage_w + sub * kStageB + nb * 512 + lane * 16 + jp * 8;
          uint2 w[8];
#pragma unroll
          for (int kt = 0; kt < 8; ++kt) w[kt] = *reinterpret_cast<const uint2*>(src + kt * 1024);
          uint2 sc = make_uint2(0, 0);
          if (GROUPED)
            sc = *reinterpret_cast<const uint2*>(sS + st.idx * stage_s + sub * kStageS + nb * 128 + c * 16 + jp * 8);
          if (!slot_ready) {  // the loads above are in flight while we wait for the TMEM slot
            mbar_wait(bar_aempty + 8 * as.idx, as.phase ^ 1);
            tc_fence_after();
            slot_ready = true;
          }
          if (q == 0 && lane == 0) QQQ_TR(3, itn);
          const uint32_t tmem_a =
              tmem_base + tmem_a0 + as.idx * 32 * KSUB + sub * 32 + ((uint32_t)(32 * q) << 16);
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
          if (q == 0 && lane == 0) QQQ_TR(10, itn);
          tmem_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (PAIR)
              mbar_arrive_cluster(lead_afull + 8 * as.idx);
            else
              mbar_arrive(bar_afull + 8 * as.idx);
            mbar_arrive(bar_emptyw + 8 * st.idx);  // this warp's reads of the weight stage are done
          }
          if (q == 0 && lane == 0) QQQ_TR(4, itn);
        }
        turn = (turn == G - 1) ? 0 : turn + 1;
      }
      st.advance();
      as.advance();
    }
  } else if (warp >= epi_warp0) {
    // ===================================== epilogue warps ===================================
    const int q = warp & 3;
    const int epi_tid = threadIdx.x - epi_warp0 * 32;
    const int eh = (warp - epi_warp0) >> 2;  // which of the n_epi/4 warps of this quadrant: takes every (n_epi/4)-th chunk
    const int mstep = 16 * (n_epi >> 2);
    // D leaves through a warp-private shared-memory tile: the warp holds a chunk as [channel = lane][16 tokens]; it
    // writes it as [16 tokens][32 channels] fp16 (row = 64 B, conflict-free), then every lane re-reads 16 B = 8
    // channels of one token and stores them: 2 vector stores per lane and chunk (each instruction covers 8 token
    // rows x 64 B) instead of 16 two-byte stores on a serial address chain.  (Measured alternatives, profiles/r01:
    // a TMA store per 4-warp group or per warp, and half2 token-pair staging, were all slower: the drain is bound by
    // the issue rate of the two epilogue warps of a sub-partition, ~730 cycles per 16-token chunk.)
    unsigned short* stg = reinterpret_cast<unsigned short*>(sStage + (warp - epi_warp0) * kStageD);
    const int st_tok = lane >> 2, st_part = lane & 3;  // this lane's token row (and +8) and 8-channel part of the tile
    const size_t tile_ints = (size_t)p.n_tok * kTileN;  // one partial tile in C
    grid_dependency_wait();  // s1 comes from the preceding kernel; D / C / lock words may still be in use by it
    int staged_mt = -1;
#ifdef QQQ_TRACE
    int echunk = 0;
#endif
    for (int seg = 0; seg < n_seg; ++seg) {
      int tile, kb0, kb1;
      sched.segment(seg, tile, kb0, kb1);
      const int nt = nt_of(tile), mt = tile % p.m_tiles;
      const int rtile = (tile << PAIR) + (int)rank;  // unique id of this CTA's 128-channel tile (scratch blocks)
      const int dbuf = seg % ndbuf;
      const uint32_t dph = (seg / ndbuf) & 1;
      const int n = nt * kTileN + 32 * q + lane;
      const bool n_ok = n < p.N;
      const bool q_ok = nt * kTileN + 32 * q < p.N;  // warp-uniform
      const int m0 = mt * p.n_tok;
      const int rows = min(p.n_tok, p.M - m0);  // valid token rows of this tile
      const bool whole = (kb0 == 0 && kb1 == KU);
      // contributors of a phase-A tile: the CTAs whose slice [b*a_upc, (b+1)*a_upc) meets the tile's units
      const int parts = whole ? 1 : (tile * KU + KU - 1) / p.a_upc - (tile * KU) / p.a_upc + 1;
      float s2v = 0.f;
      if constexpr (!kAcc) s2v = n_ok ? __ldg(p.s2 + s2_position(n)) : 0.f;  // kAcc: no scales (s1 / s2 are null)
      // optional bias, added in fp16 after the fp16 rounding of D like the reference's eager `D + bias`
      // (qlinear_marlin.py:287) — one rounding more, same bits.  Applied on the store side, where a lane holds 8 channels
      // of a token row: 8 half2 adds per chunk instead of 16 scalar ones inside the conversion loop.
      const bool has_bias = p.bias != nullptr;
      uint4 bias_v = make_uint4(0, 0, 0, 0);
      if (has_bias && q_ok) bias_v = __ldg(reinterpret_cast<const uint4*>(p.bias + nt * kTileN + 32 * q + 8 * st_part));
      // Partial tiles of split-K live in C as compact [n_tok][128] int32 blocks, block index =
      // (ticket * a_tiles + tile): a chunk's 16 rows are 512 B apart, so the 16 stores / loads of a lane use one base
      // register and immediate offsets (no serial address chain), and each of them is one full 128-byte line per warp.
      int* __restrict__ cbase = p.C + (size_t)rtile * tile_ints + 32 * q + lane;
      const size_t ticket_stride = (size_t)(p.a_tiles << PAIR) * tile_ints;  // split tiles are tiles [0, a_tiles)

      // per-token scales of this token tile -> smem (once per tile change), so the store loop has no global loads
      if (mt != staged_mt) {
        named_bar_sync(1, n_epi_thr);  // previous users of s1_sm are done
        if constexpr (!kAcc)
          for (int i = epi_tid; i < p.n_tok; i += n_epi_thr) s1_sm[i] = (m0 + i < p.M) ? __ldg(p.s1 + m0 + i) : 0.f;
        named_bar_sync(1, n_epi_thr);
        staged_mt = mt;
      }

      mbar_wait(bar_dfull + 8 * dbuf, dph);
      tc_fence_after();
      if (epi_tid == 0) QQQ_TR(7, seg);
      const uint32_t tmem_d = tmem_base + dbuf * p.n_tok + ((uint32_t)(32 * q) << 16);

      // Split-K tiles: contributors take a ticket on the tile's lock word (low half = tickets, high half = partials
      // published).  Tickets 0..parts-2 publish their int32 partial tile to block[ticket] of C (plain coalesced
      // stores); the last ticket keeps its partial in TMEM, waits until the others are published (they arrived
      // earlier, so this is normally immediate), adds them and finishes the tile.  Integer sums: exact in any order.
      int ticket = 0;
      int* lock = p.locks + nt + p.n_tiles * mt;
      if (!whole) {
        named_bar_sync(1, n_epi_thr);
        if (epi_tid == 0) misc[1] = (uint32_t)atomicAdd(lock, 1) & 0xFFFFu;
        named_bar_sync(1, n_epi_thr);
        ticket = (int)misc[1];
        if (ticket == parts - 1) {
          if (epi_tid == 0) {
            while ((ld_acquire_gpu(lock) >> 16) != parts - 1) {
            }
          }
          named_bar_sync(1, n_epi_thr);
          __threadfence();  // order this thread's reads of the published partials after the acquire above
        }
      }
      const bool finish = whole || ticket == parts - 1;  // this CTA writes D for the tile
      const int others = (!whole && finish) ? parts - 1 : 0;  // published partial tiles the finisher adds

      // partial sums published by the other contributors, prefetched one 16-row chunk ahead (rows past `rows`
      // of a published block hold sums of zero-filled tokens, i.e. zeros: no predication needed)
      int pre[16];
      auto fetch_partials = [&](int mb, int* acc) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = 0;
        for (int pp = 0; pp < others; ++pp) {
          const int* __restrict__ src = cbase + (size_t)pp * ticket_stride + (size_t)mb * kTileN;
#pragma unroll
          for (int i = 0; i < 16; ++i) acc[i] += __ldcg(src + i * kTileN);
        }
      };
      if (others > 0 && 16 * eh < rows) fetch_partials(16 * eh, pre);

      // One 16-token chunk of this lane's channel: add the published partials (finisher), then either scale,
      // transpose through the warp's tile and store fp16 rows of D, or publish the int32 partial.
      auto process = [&](uint32_t(&r)[16], int mb) {
        if (others > 0) {
#pragma unroll
          for (int i = 0; i < 16; ++i) r[i] += (uint32_t)pre[i];
          if (mb + mstep < rows) fetch_partials(mb + mstep, pre);
        }
        if (finish && kAcc) {
          // lane = channel, i = token: every store instruction of the warp writes 32 consecutive int32 of one row
          if (n_ok) {
            int* __restrict__ d32 = reinterpret_cast<int*>(p.D) + (size_t)(m0 + mb) * p.N + n;
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (mb + i < rows) d32[(size_t)i * p.N] = (int)r[i];
          }
        } else if (finish) {
          unsigned short* sp = stg + lane;
          const float4* s4 = reinterpret_cast<const float4*>(s1_sm + mb);
#pragma unroll
          for (int g4 = 0; g4 < 4; ++g4) {
            const float4 sv = s4[g4];
            const float sa[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int i = 4 * g4 + j;
              sp[i * 32] = __half_as_ushort(__float2half_rn((__int2float_rn((int)r[i]) * s2v) * sa[j]));
            }
          }
          if (epi_tid == 0) QQQ_TR(9, echunk);
          __syncwarp();
          if (epi_tid == 0) QQQ_TR(15, echunk);
          if (q_ok) {  // the whole 32-channel quadrant is inside N (N % 64 == 0) or outside
            const uint4* rp = reinterpret_cast<const uint4*>(stg) + lane;  // row st_tok, part st_part
            uint4 v0 = rp[0], v1 = rp[32];                                  // rows st_tok and st_tok + 8
            if (has_bias) {
              v0 = hadd2_x4(v0, bias_v);
              v1 = hadd2_x4(v1, bias_v);
            }
            const int col = nt * kTileN + 32 * q + 8 * st_part;
            if constexpr (kOut == kOutScatter) {
              if (mb + st_tok < rows) emit_d<kOut>(out_row<kOut>(p, m0 + mb + st_tok) + col, v0);
              if (mb + st_tok + 8 < rows) emit_d<kOut>(out_row<kOut>(p, m0 + mb + st_tok + 8) + col, v1);
            } else {
              __half* dp = p.D + (size_t)(m0 + mb + st_tok) * p.N + col;
              if (mb + st_tok < rows) emit_d<kOut>(dp, v0);
              if (mb + st_tok + 8 < rows) emit_d<kOut>(dp + (size_t)8 * p.N, v1);
            }
          }
          __syncwarp();  // the tile is rewritten by the next chunk
        } else {
          int* __restrict__ dst = cbase + (size_t)ticket * ticket_stride + (size_t)mb * kTileN;
#pragma unroll
          for (int i = 0; i < 16; ++i) dst[i * kTileN] = (int)r[i];
        }
      };
      // Software-pipelined drain: the TMEM load of the next chunk is in flight while this one is converted and
      // stored (TMEM reads run at 64 B/clk per SM: 2048 cycles for a 128 x 256 accumulator).
      {
        uint32_t ra[16], rb[16];
        int mb = 16 * eh;
        if (mb < rows) tmem_ld_32x32b_x16(tmem_d + mb, ra);
        while (mb < rows) {
          tmem_wait_ld();
          if (epi_tid == 0) QQQ_TR(13, echunk);
          const int mb2 = mb + mstep;
          if (mb2 < rows) tmem_ld_32x32b_x16(tmem_d + mb2, rb);
          process(ra, mb);
          if (epi_tid == 0) QQQ_TR(14, echunk++);
          if (mb2 >= rows) break;
          tmem_wait_ld();
          if (epi_tid == 0) QQQ_TR(13, echunk);
          mb = mb2 + mstep;
          if (mb < rows) tmem_ld_32x32b_x16(tmem_d + mb, ra);
          process(rb, mb2);
          if (epi_tid == 0) QQQ_TR(14, echunk++);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {  // accumulator buffer may be overwritten (pair: the leader's MMA issuer waits for both CTAs)
        if (PAIR)
          mbar_arrive_cluster(lead_dempty + 8 * dbuf);
        else
          mbar_arrive(bar_dempty + 8 * dbuf);
      }
      if (epi_tid == 0) QQQ_TR(8, seg);

      if (!whole) {
        if (finish) {
          named_bar_sync(1, n_epi_thr);  // every warp has read the published partials
          if (epi_tid == 0) *lock = 0;
        } else {
          __threadfence();  // partial tile visible device-wide before it is announced
          named_bar_sync(1, n_epi_thr);
          if (epi_tid == 0) atomicAdd(lock, 0x10000);
        }
      }
      if (epi_tid == 0) QQQ_TR(11, seg);
    }
  }

  tc_fence_before();
  if (PAIR)
    cluster_sync_all();  // neither CTA may exit (or free TMEM) while the other can still signal it or use its operands
  else
    __syncthreads();
  if (threadIdx.x == 0) QQQ_TR(12, 1);
  if (warp == 1) {
    if (PAIR)
      tmem_dealloc_pair(tmem_base, tmem_cols);
    else
      tmem_dealloc(tmem_base, tmem_cols);
  }
}

}  // namespace

size_t gemm_smem_bytes(const GemmParams& p) {
  return 1024 + kEpiStageBytes + (size_t)p.stages_t * p.ksub * (p.n_tok >> p.pair) * 128 +
         (size_t)p.stages_w * p.ksub * (kStageB + kStageS) +
         8 * (2 * p.stages_w + 2 * p.stages_t + 2 * kMaxASlots + 4) + 16 + 4 * kMaxTok;
}

cudaError_t launch_gemm(const CUtensorMap& tmap_a, const CUtensorMap& tmap_b, const GemmParams& p, bool grouped,
                        int grid, int dev, cudaStream_t stream, bool pdl) {
  static bool attr_set[10][64] = {};  // the opt-in shared-memory attribute is per device
  const size_t smem = gemm_smem_bytes(p);
  if (p.out_mode != kOutStore && p.pair) return cudaErrorInvalidValue;  // the planner never pairs these launches
  if (p.out_mode < 0 || p.out_mode > kOutScatter) return cudaErrorInvalidValue;
  using Kern = void (*)(const CUtensorMap, const CUtensorMap, const GemmParams);
  static const Kern table[10] = {
      qqq_gemm_kernel<false, false>, qqq_gemm_kernel<true, false>, qqq_gemm_kernel<false, true>, qqq_gemm_kernel<true, true>,
      qqq_gemm_kernel<false, false, kOutReduce>, qqq_gemm_kernel<true, false, kOutReduce>,
      qqq_gemm_kernel<false, false, kOutAcc>, qqq_gemm_kernel<true, false, kOutAcc>,
      qqq_gemm_kernel<false, false, kOutScatter>, qqq_gemm_kernel<true, false, kOutScatter>};
  const int g1 = grouped ? 1 : 0;
  const int variant = p.out_mode != kOutStore ? 2 + 2 * p.out_mode + g1 : g1 + (p.pair ? 2 : 0);
  const Kern kern = table[variant];
  if (dev < 0 || dev >= 64 || !attr_set[variant][dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemBytes);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) attr_set[variant][dev] = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // overlap launch + prologue + weight prefetch
    attr[na].val.programmaticStreamSerializationAllowed = 1;           // with the tail of the preceding kernel
    ++na;
  }
  if (p.pair) {
    attr[na].id = cudaLaunchAttributeClusterDimension;  // CTA pairs: two SMs of one TPC
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kern, tmap_a, tmap_b, p);
}

#ifdef QQQ_TRACE
extern "C" int qqq_trace_set(void* buf) {
  return (int)cudaMemcpyToSymbol(g_trace, &buf, sizeof(buf));
}
#endif

}  // namespace qqq
