// Host/device shared declarations of the sm_100a W4A8 GEMM (see qqq_gemm_sm100.cu).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>

namespace qqq {

constexpr int kTileN = 128;      // output channels per CTA tile  (UMMA M)
constexpr int kBlockK = 128;     // reduction depth of one k sub-block (= the per-group quantisation group)
constexpr int kStageB = 8192;    // packed int4 bytes per sub-block: 8 rows of B x 256 words
constexpr int kStageS = 256;     // group-scale bytes per sub-block: 128 channels x fp16
constexpr int kMaxASlots = 8;    // ring slots of 32*ksub columns each
constexpr int kMaxStages = 16;
constexpr int kMaxTok = 256;     // token tile (UMMA N) upper bound
// Token tiles up to this size get two accumulator buffers in TMEM (the drain of one tile overlaps the MMAs of the
// next); the rest of the 512 columns is the unpacked-weight ring: 208 leaves 3 slots of 32 columns (5 x 208 tokens cover
// M = 1024), 192 leaves 4.
#ifndef QQQ_DBUF_MAX_TOK
#define QQQ_DBUF_MAX_TOK 208
#endif
constexpr int kDbufMaxTok = QQQ_DBUF_MAX_TOK;
constexpr int kMaxSmemBytes = 232448;  // 227 KB opt-in limit per CTA on sm_100
constexpr int kStageD = 1024;          // one epilogue staging tile: 16 tokens x 32 channels fp16 (one warp's chunk)
constexpr int kEpiStageBytes = 8 * kStageD;  // up to 8 epilogue warps
// warp roles: 0 weights TMA, 1 MMA (+TMEM alloc), 2 tokens TMA, 3 idle, then 4*G unpack warps (G groups x 4 TMEM
// lane quadrants) and the remaining 16-4G warps as epilogue (G = 2: 8 epilogue warps, G = 3: 4).  24 warps
// (G up to 4) were measured and did not help: all warps of a TMEM quadrant share one SM sub-partition, so extra
// groups add latency hiding but no ALU throughput (profiles/r01/sweep_groups_24warps.log).
constexpr int kUnpackWarp0 = 4;
constexpr int kWarps = 20;
constexpr int kThreads = 32 * kWarps;

struct GemmParams {
  int32_t* C;        // split-K scratch: compact [n_tok][128] int32 partial tiles, block = ticket * tiles + tile
  __half* D;         // output [M, N]
  const float* s1;   // [M]
  const float* s2;   // [N] permuted (reference scale_perm_single)
  const __half* s3;  // [K/128, N] permuted (reference scale_perm) or nullptr
  const __half* bias;  // [N] natural channel order, added to D in fp16 (QuantLinear.forward's `D + bias`), or nullptr
  int* locks;        // [n_tiles * m_tiles] zero in / zero out
  int M, N, K;
  int n_tok;         // token tile, multiple of 16, <= 256
  int m_tiles, n_tiles, k_blocks;
  int ksub;          // 128-deep k sub-blocks per pipeline stage (1, 2 or 4): amortises barrier traffic at small n_tok
  int k_units;       // ceil(k_blocks / ksub): pipeline stages ("units") per tile
  int stages_w, stages_t;  // depth of the weight / token smem rings
  int unpack_groups;       // 2 or 3 groups of 4 unpack warps; the other 16-4G non-control warps are epilogue warps
  int pair;                // 1: CTA pairs (cluster of 2, cta_group::2): scheduled tiles are 256 channels wide, each CTA of
                           // a pair owns 128 of them and loads half of the token tile; n_tiles stays in 128-channel tiles
  // two-phase schedule (see Sched in qqq_gemm_sm100.cu); a unit = (tile, k-unit), tile = mt + m_tiles*nt
  int a_tiles;  // tiles [0, a_tiles) are cut along K: CTA b owns phase-A units [b*a_upc, (b+1)*a_upc) of a_units
  int a_units;  // a_tiles * k_units
  int a_upc;    // >= 1 (1 when a_units == 0)
  int b_tiles;  // tiles [a_tiles, a_tiles + b_tiles) are processed whole, dealt round-robin: CTA b owns a_tiles + b + i*b_step
  int b_tpc;    // ceil(b_tiles / b_step): whole tiles of the busiest CTA
  int b_step;   // number of schedule indices (CTAs, or CTA pairs) the whole tiles are dealt to
  uint64_t hint_a, hint_b;  // L2 eviction policies for the token / weight streams
  int out_mode;             // 0: store D;  1: D is a multicast address and the epilogue adds into it (multimem.red);
                            // 2: D is int32 [M, N] and receives the raw accumulators (no scales);  3: scatter (below)
  // out_mode 3 (tensor-parallel row shards): token row m goes to rank m / tp_rows, into slot tp_rank of that rank's
  // partial-sum buffer tp_part[owner] = fp16 [world][tp_rows][N] (peer-mapped pointers)
  int tp_rank, tp_rows;
  __half* tp_part[8];
};

size_t gemm_smem_bytes(const GemmParams& p);
cudaError_t launch_gemm(const CUtensorMap& tmap_a, const CUtensorMap& tmap_b, const GemmParams& p, bool grouped,
                        int grid, int dev, cudaStream_t stream, bool pdl);

}  // namespace qqq
