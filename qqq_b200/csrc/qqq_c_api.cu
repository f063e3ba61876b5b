// C ABI of libqqq_b200.so (declared in include/qqq_b200.h): raw pointers + ints in, int rc out; no allocation,
// no synchronisation, no torch.  Mirrors the shape of the reference's inner `qqq_cuda()` (csrc/qqq_gemm.cu:950-969).
#include "../../include/qqq_b200.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>

#include "qqq_common.cuh"
#include "qqq_gemm_sm100.h"

namespace {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

void set_err(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) {
      ok = false;
      return;
    }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() {
    int cur = -1;
    if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
  }
};

struct DeviceInfo {
  bool valid = false;
  int cc_major = 0;
  int sms = 0;
};
DeviceInfo g_dev[64];
std::mutex g_mu;

const DeviceInfo* device_info(int dev) {
  if (dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_dev[dev].valid) {
    int maj = 0, sms = 0;
    if (cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return nullptr;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return nullptr;
    g_dev[dev].cc_major = maj;
    g_dev[dev].sms = sms;
    g_dev[dev].valid = true;
  }
  return &g_dev[dev];
}

// cuTensorMapEncodeTiled through the runtime's driver entry point: no link-time libcuda dependency, so the
// library still loads (and its symbols can be checked) on a machine without a GPU driver.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// Tensor maps depend only on (base pointer, dims, box): weights are long-lived module buffers and activation
// buffers come back from the caching allocator, so a small cache removes the driver call from the hot path.
struct MapKey {
  const void* base;
  uint64_t d0, d1;
  uint32_t b0, b1, kind;
  bool operator==(const MapKey& o) const {
    return base == o.base && d0 == o.d0 && d1 == o.d1 && b0 == o.b0 && b1 == o.b1 && kind == o.kind;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    uint64_t h = reinterpret_cast<uint64_t>(k.base) * 0x9E3779B97F4A7C15ull;
    h ^= (k.d0 * 0xC2B2AE3D27D4EB4Full) ^ (k.d1 << 17) ^ ((uint64_t)k.b0 << 40) ^ ((uint64_t)k.b1 << 48) ^ k.kind;
    return (size_t)(h ^ (h >> 29));
  }
};
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;
std::mutex g_maps_mu;

bool encode_2d_uncached(CUtensorMap* m, CUtensorMapDataType dt, const void* base, uint64_t d0, uint64_t d1,
                        uint64_t row_bytes, uint32_t b0, uint32_t b1, CUtensorMapSwizzle sw);

bool encode_2d(CUtensorMap* m, CUtensorMapDataType dt, const void* base, uint64_t d0, uint64_t d1, uint64_t row_bytes,
               uint32_t b0, uint32_t b1, CUtensorMapSwizzle sw) {
  const MapKey key{base, d0, d1, b0, b1, (uint32_t)dt * 16u + (uint32_t)sw};
  {
    std::lock_guard<std::mutex> lk(g_maps_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) {
      *m = it->second;
      return true;
    }
  }
  if (!encode_2d_uncached(m, dt, base, d0, d1, row_bytes, b0, b1, sw)) return false;
  std::lock_guard<std::mutex> lk(g_maps_mu);
  if (g_maps.size() > 8192) g_maps.clear();
  g_maps.emplace(key, *m);
  return true;
}

bool encode_2d_uncached(CUtensorMap* m, CUtensorMapDataType dt, const void* base, uint64_t d0, uint64_t d1,
                        uint64_t row_bytes, uint32_t b0, uint32_t b1, CUtensorMapSwizzle sw) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) {
    set_err("cuTensorMapEncodeTiled entry point unavailable");
    return false;
  }
  cuuint64_t gdim[2] = {d0, d1};
  cuuint64_t gstr[1] = {row_bytes};
  cuuint32_t box[2] = {b0, b1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, dt, 2, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_err("cuTensorMapEncodeTiled failed (CUresult %d) dims=(%llu,%llu) box=(%u,%u)", (int)r,
            (unsigned long long)d0, (unsigned long long)d1, b0, b1);
    return false;
  }
  return true;
}

// Programmatic dependent launch (on by default; QQQ_B200_PDL=0 disables it): the kernels call griddepcontrol.wait
// before touching anything the preceding kernel in the stream may produce or still use.
bool use_pdl() {
  static const bool on = !(getenv("QQQ_B200_PDL") && atoi(getenv("QQQ_B200_PDL")) == 0);
  return on;
}

// The reference's tile-shape validity rule (csrc/qqq_gemm.cu:867-916), applied for error parity only.
bool reference_shape_ok(int n, int k, int thread_k, int thread_n) {
  static const int cfg[4][2] = {{128, 128}, {128, 64}, {64, 256}, {64, 128}};
  if (thread_k != -1 && thread_n != -1) {
    if (thread_k != 128 && thread_k != 64) return false;
    if (thread_n < 64) return false;
    return k % thread_k == 0 && n % thread_n == 0;
  }
  for (auto& c : cfg)
    if (k % c[0] == 0 && n % c[1] == 0) return true;
  return false;
}

// Pure host-side planning (no CUDA calls): tiling, pipeline depths and the two-phase schedule for a problem on
// `sm_count` SMs.  Exported as qqq_b200_plan() so the schedule can be checked exhaustively on a CPU-only machine.
int plan_gemm(int M, int N, int K, bool grouped, int sm_count, int max_par, bool has_scratch, qqq::GemmParams& p,
              int* grid_out, bool allow_pair = true) {
  using namespace qqq;
  p.M = M;
  p.N = N;
  p.K = K;
  // token tiling: the whole batch in one UMMA-N tile up to 256 tokens, otherwise equal tiles of <= 256 ...
  p.n_tiles = (N + kTileN - 1) / kTileN;
  p.k_blocks = (K + kBlockK - 1) / kBlockK;
  auto tile_tokens = [&](int cap, int* m_tiles) {
    *m_tiles = (M + cap - 1) / cap;
    const int per_tile = (M + *m_tiles - 1) / *m_tiles;
    return (per_tile + 15) / 16 * 16;
  };
  auto kb_cycles = [](int n) {  // measured cycles per 128-deep k-block of a tile of n tokens (see below)
    return n <= 128 ? 350.0 + 0.94 * n : (n <= 208 ? 470.0 + 1.625 * (n - 128) : 590.0);
  };
  p.n_tok = tile_tokens(kMaxTok, &p.m_tiles);
  if (M > 64) {
    // ... unless smaller tiles fill the SMs so much better that they win despite their higher cost per token.
    // Cost model from the round-2 measurements (profiles/r02/call_b, call_f): a 128-deep k-block of a tile costs
    // ~350 + 0.94 * n_tok cycles up to 128 tokens, ~600 at 208 and ~590 at 256 (operand traffic L2 -> SM and hand-offs, not
    // the 2 * n_tok cycles of the tensor pipe, bound the main loop); the accumulator drain costs ~29 cycles per token and
    // is exposed once per tile when the accumulator is single-buffered (n_tok > kDbufMaxTok), once per CTA otherwise.
    // Whole-tile waves are compared; stream-K (below) only smooths the remainder.  208 tokens = the largest tile that
    // still leaves room for two accumulators and a 3-slot weight ring in TMEM (5 x 208 covers M = 1024).
    double best = 0.0;
    const int caps[4] = {kMaxTok, kDbufMaxTok, 128, 64};
    for (int ci = 0; ci < 4; ++ci) {
      int mt = 0;
      const int nt = tile_tokens(caps[ci], &mt);
      // 256-token tiles run as CTA pairs where the policy below turns pairs on (same rule): cheaper k-blocks, waves of pairs
      const bool pr = allow_pair && nt == kMaxTok && p.n_tiles % 2 == 0 && sm_count >= 2 && mt >= 2 &&
                      2ll * mt * p.n_tiles >= 5ll * sm_count;
      const long long units_w = pr ? (long long)mt * (p.n_tiles / 2) : (long long)mt * p.n_tiles;
      const int lanes = pr ? sm_count / 2 : sm_count;
      const double waves = (double)((units_w + lanes - 1) / lanes);
      const double drain = 29.0 * nt + 100.0;
      const bool dbuf = nt <= kDbufMaxTok;
      double cost = waves * p.k_blocks * (pr ? 552.0 : kb_cycles(nt)) + (dbuf ? 1.0 : waves) * drain;
      cost *= 1.0 + 0.03 * ci;  // near ties go to the larger tile (fewer re-reads of the weights)
      if (ci == 0 || cost < best) {
        best = cost;
        p.n_tok = nt;
        p.m_tiles = mt;
      }
    }
  }
  // Tuning overrides for experiments (not part of the ABI): QQQ_B200_NTOK sets the token-tile cap, QQQ_B200_KSUB /
  // QQQ_B200_NST force the stage depth in k and the token ring depth.
  static const int env_ntok = getenv("QQQ_B200_NTOK") ? atoi(getenv("QQQ_B200_NTOK")) : 0;
  static const int env_ksub = getenv("QQQ_B200_KSUB") ? atoi(getenv("QQQ_B200_KSUB")) : 0;
  static const int env_nst = getenv("QQQ_B200_NST") ? atoi(getenv("QQQ_B200_NST")) : 0;
  if (env_ntok >= 16 && env_ntok <= kMaxTok && env_ntok % 16 == 0) p.n_tok = tile_tokens(env_ntok, &p.m_tiles);
  // several k-blocks per pipeline stage so that barrier round trips and the single-thread MMA issue loop are
  // amortised over >= 512 tensor-pipe cycles (an MMA of N tokens takes ~N/2 cycles, 4 per k-block)
  p.ksub = p.n_tok <= 64 ? 4 : (p.n_tok <= 128 ? 2 : 1);
  if (env_ksub == 1 || env_ksub == 2 || env_ksub == 4) p.ksub = env_ksub;
  while (p.ksub > 1 && p.k_blocks < p.ksub) p.ksub >>= 1;
  p.k_units = (p.k_blocks + p.ksub - 1) / p.ksub;
  // CTA pairs (cluster of 2, cta_group::2, see qqq_gemm_sm100.cu): two CTAs share a token tile and each loads half of
  // it.  Measured (profiles/r01/pair_mode.log): +7-9 % when every SM walks several 256-token tiles (M >= 1024 at
  // N = 21760), neutral at ~2 tiles per SM, slower when there is at most one tile per SM or the tiles are small.
  // QQQ_B200_PAIR=0/1 overrides the policy (1: wherever the shape allows it).
  static const int env_pair = getenv("QQQ_B200_PAIR") ? atoi(getenv("QQQ_B200_PAIR")) : -1;
  const bool pair_ok = allow_pair && p.n_tiles % 2 == 0 && p.n_tok % 32 == 0 && sm_count >= 2;
  const bool pair_auto = p.n_tok == kMaxTok && p.m_tiles >= 2 && 2ll * p.m_tiles * p.n_tiles >= 5ll * sm_count;
  p.pair = (pair_ok && (env_pair == 1 || (env_pair != 0 && pair_auto))) ? 1 : 0;
  const int sched_cols = p.n_tiles >> p.pair;   // scheduled (super-)tiles per token tile
  const int sched_ctas = sm_count >> p.pair;    // CTAs (pairs) the schedule is distributed over
  const long long tiles = (long long)p.m_tiles * sched_cols;
  const long long units = tiles * p.k_units;
  if (units >= (1ll << 31)) {
    set_err("problem too large");
    return QQQ_ERR_PROB_SHAPE;
  }
  // Unpack groups.  Decode-size tiles: 3 groups + 4 epilogue warps; from 128 tokens up the accumulator drain is the
  // exposed part: 2 groups + 8 epilogue warps (measured for both modes: the per-group rescale is bound by the issue
  // rate of the sub-partition, which a third warp on the same sub-partition cannot widen).
  static const int env_grp = getenv("QQQ_B200_GROUPS") ? atoi(getenv("QQQ_B200_GROUPS")) : 0;
  const int g_auto = p.n_tok <= 64 ? 3 : 2;
  p.unpack_groups = (env_grp >= 2 && env_grp <= 3) ? env_grp : g_auto;
  // Weight-ring depth must be a multiple of `period`.  Sub-block i = ksub*unit + sub is unpacked by group i % G,
  // and a group only waits on the full-barrier of the stages it unpacks from.  mbarrier waits are by phase PARITY:
  // a group that waits for unit w on a stage whose previous occupant (unit w - depth) it never waited on can find
  // that barrier still one phase behind (previous data not landed yet) and would then pass on the stale parity.
  // With depth % period == 0 every stage is always consumed by the same groups, in order, so each wait follows the
  // wait on the previous occupant.  (ksub >= G: every group touches every unit, period 1.)
  const int G = p.unpack_groups;
  auto gcd = [](int a, int b) { while (b) { const int t = a % b; a = b; b = t; } return a; };
  const int period = p.ksub >= G ? 1 : G / gcd(p.ksub, G);
  // smem rings.  Depth = latency x consumption rate: when one token tile covers M the weights stream from DRAM
  // (long latency, ~160 KB in flight); with several token tiles they mostly hit L2 (~64 KB).  The token ring
  // (L2-resident data) takes the rest, at least 3 and at most 6 stages.
  const int stage_t = p.ksub * (p.n_tok >> p.pair) * 128, stage_w = p.ksub * (kStageB + kStageS);
  const int budget =
      kMaxSmemBytes - 1024 - kEpiStageBytes - 8 * (4 * kMaxStages + 2 * kMaxASlots + 4) - 16 - 4 * kMaxTok;
  const int inflight = 163840;  // bytes of weights (or tokens) in flight per CTA
  const int max_w = kMaxStages / period * period;
  auto round_w = [&](int n) { n = n > max_w ? max_w : n; return n / period * period; };
  int nsw = 0, nst = 0;
  if (p.m_tiles > 1) {  // tensor-bound regime: the token ring comes first (~160 KB, (L2 latency + 512) / stages <= 512)
    nst = (inflight + stage_t - 1) / stage_t;
    nst = nst < 3 ? 3 : (nst > 6 ? 6 : nst);
    if (env_nst >= 2 && env_nst <= kMaxStages) nst = env_nst;
    for (; nst >= 2; --nst) {
      nsw = round_w((budget - nst * stage_t) / stage_w);
      if (nsw >= 4 || (nst <= 3 && nsw >= 2)) break;
    }
  } else {
    nsw = round_w((inflight + stage_w - 1) / stage_w + period - 1);
    if (nsw < period) nsw = period;
    for (; nsw >= period && nsw >= 2; nsw -= period) {
      nst = (budget - nsw * stage_w) / stage_t;
      if (nst >= 3) break;
    }
    if (nst > 6) nst = 6;
    if (env_nst >= 2 && env_nst <= kMaxStages && nsw * stage_w + env_nst * stage_t <= budget) nst = env_nst;
  }
  if (nst < 2 || nsw < 2 || nsw % period != 0 || nsw * stage_w + nst * stage_t > budget) {
    set_err("internal: no room for the smem rings (n_tok=%d ksub=%d)", p.n_tok, p.ksub);
    return QQQ_ERR_KERN_SHAPE;
  }
  p.stages_t = nst;
  p.stages_w = nsw;

  int grid = sched_ctas;
  if ((long long)grid > units) grid = (int)units;
  // Schedule.  Whole tiles go round the CTAs in waves; the remainder tiles that would leave SMs idle in a last
  // partial wave are instead cut along K over all CTAs (stream-K) when the caller's scratch allows it: C (64*max_par
  // rows of N int32) holds compact [n_tok][128] partial tiles, block index = ticket*a_tiles + tile, i.e. one set of
  // a_tiles blocks per contributor but the last; plus one lock word per tile.
  const long long whole_per_cta = tiles / grid;
  const long long rem = tiles - whole_per_cta * grid;
  long long a_tiles = 0, a_upc = 1;
  static const int env_split = getenv("QQQ_B200_SPLIT") ? atoi(getenv("QQQ_B200_SPLIT")) : -1;  // experiments
  if (rem > 0 && has_scratch && (tiles << p.pair) <= (long long)(N / 128) * max_par && env_split != 0) {
    const long long tile_ints = ((long long)p.n_tok * kTileN) << p.pair;  // partial tile(s) of one scheduled tile
    const long long c_ints = 64ll * max_par * N;              // capacity of C
    auto parts_max = [&](long long upc) { return (upc % p.k_units == 0) ? 1ll : (p.k_units - 1) / upc + 2; };
    // (1) cut only the remainder tiles, over all CTAs, ahead of the whole tiles (fix-up hidden behind the rest)
    const long long upc_rem = (rem * p.k_units + grid - 1) / grid;
    // (2) fallback when C is too small for that many contributors per tile: one contiguous stream-K range per CTA
    //     over ALL tiles (at most 2-3 contributors per tile, but the fix-up of a CTA's last tile is exposed)
    const long long upc_all = (units + grid - 1) / grid;
    // Cycle model from the in-kernel timelines (profiles/): a k-block of a single token tile costs 4 MMAs of
    // max(48 issue, n_tok/2 pipe) cycles + ~48 of hand-off and not less than ~420 when the weights stream from DRAM; with
    // several token tiles the measured kb_cycles(); a drain costs ~730 cycles per 16-token chunk and epilogue warp + ~1400.
    // One token tile (decode, M <= 256): the weights stream from DRAM, 8 KB per k-block and CTA against 3334 B/clk of HBM
    // for the whole chip (x1.15 measured) — 419 cycles with 148 CTAs streaming, but only the unpack floor (~300 per-group,
    // ~200 per-channel) when a few CTAs have the memory system to themselves: (32, 4096, 1024) runs 32 k-blocks per CTA on
    // 8 CTAs in 9.0 us, faster than 64 CTAs with 4 k-blocks each and a fix-up (16.2 us).
    auto kb_single = [&](long long active) {
      const double base = 4.0 * (p.n_tok / 2 > 48 ? p.n_tok / 2 : 48) + 48.0;
      const double unpack = grouped ? 300.0 : 200.0, hbm = 2.83 * (double)active;
      return base > unpack ? (base > hbm ? base : hbm) : (unpack > hbm ? unpack : hbm);
    };
    const long long active_whole = tiles < grid ? tiles : grid;
    const double c_kb_multi = p.pair ? 552.0 : kb_cycles(p.n_tok);
    const double t_u = p.ksub * (p.m_tiles > 1 ? c_kb_multi : kb_single(grid));              // all CTAs busy (split)
    const double t_u_whole = p.ksub * (p.m_tiles > 1 ? c_kb_multi : kb_single(active_whole));  // one CTA per whole tile
    const int n_epi = kWarps - kUnpackWarp0 - 4 * p.unpack_groups;
    const double t_d = 730.0 * ((p.n_tok / 16 + n_epi / 4 - 1) / (n_epi / 4)) + 1400.0;
    // A CTA that ends on a cut tile publishes a partial tile or finishes one: int32 partials (2x the bytes of the fp16
    // output) go through L2 both ways, the publisher's fence + announcement costs ~5 k cycles after its drain and the
    // finisher can only start when the slowest contributor has announced.  Measured end to end (profiles/r02/call_m,
    // call_n): ~10 k cycles + 40 per token on top of the plain drain — (128, 4096, 4096) took 23.0 us as stream-K over 148
    // CTAs against 10.7 us as 32 whole tiles, (32, 4096, 1024) 16.2 against 9.0 us, while (128, 8192, 21760) wins 29.0 vs
    // 33.3 us and (32, 14336, 4096) 18.2 vs 31.8 us because they save dozens of k-blocks per CTA.
    const double t_fix = 10000.0 + 40.0 * p.n_tok;
    // (1) multiplies the contributors per tile (and with them the partial-tile traffic of the fix-up), so it is
    // used for decode-size tiles only; measured: wins at n_tok <= 32, loses from 64 tokens up.
    if ((p.n_tok <= 32 || env_split == 2) && (parts_max(upc_rem) - 1) * rem * tile_ints <= c_ints) {
      // with whole tiles behind the slices the fix-up is hidden; with fewer tiles than CTAs it is the CTA's tail
      const double cost_whole = (double)p.k_units * t_u_whole + t_d;
      const double cost_split = (double)upc_rem * t_u + t_d + t_fix;
      if (whole_per_cta > 0 || env_split == 1 || env_split == 2 || cost_split < cost_whole) {
        a_tiles = rem;
        a_upc = upc_rem;
      }
    } else if ((parts_max(upc_all) - 1) * tiles * tile_ints <= c_ints) {
      // Stream-K over all tiles balances the SMs but costs every CTA one more accumulator drain (its partial of a
      // straddled tile), which is exposed when the accumulator is single-buffered.
      const bool dbuf = p.n_tok <= kDbufMaxTok;  // double-buffered accumulators: only the last drain of a CTA is exposed
      const long long waves = (tiles + grid - 1) / grid;
      const double cost_whole = (double)waves * p.k_units * t_u_whole + (dbuf ? 1.0 : (double)waves) * t_d;
      const long long segs = (upc_all + p.k_units - 1) / p.k_units + 1;
      const double cost_split = (double)upc_all * t_u + (dbuf ? 1.0 : (double)segs) * t_d + t_fix;
      if (env_split == 1 || cost_split < cost_whole) {
        a_tiles = tiles;
        a_upc = upc_all;
      }
    }
  }
  p.a_tiles = (int)a_tiles;
  p.a_units = (int)(a_tiles * p.k_units);
  p.a_upc = (int)a_upc;
  p.b_tiles = (int)(tiles - a_tiles);
  {
    // whole tiles are dealt round-robin to min(grid, b_tiles) schedule indices; phase A uses ceil(a_units / a_upc)
    const int used_b = p.b_tiles < grid ? p.b_tiles : grid;
    const int used_a = p.a_units > 0 ? (int)((p.a_units + a_upc - 1) / a_upc) : 0;
    grid = used_b > used_a ? used_b : used_a;
    p.b_step = grid;
    p.b_tpc = p.b_tiles > 0 ? (p.b_tiles + grid - 1) / grid : 0;
  }
  *grid_out = grid << p.pair;
  return QQQ_OK;
}

}  // namespace

extern "C" {

int qqq_b200_version(void) { return 101; }
const char* qqq_b200_last_error(void) { return g_err; }
long long qqq_b200_launch_count(void) { return g_launches.load(); }

int qqq_b200_plan(int prob_m, int prob_n, int prob_k, int groupsize, int sm_count, int max_par, int* out /* [20] */) {
  qqq::GemmParams p;
  memset(&p, 0, sizeof(p));
  int grid = 0;
  if (prob_m <= 0 || prob_n <= 0 || prob_k <= 0 || sm_count <= 0 || out == nullptr) return QQQ_ERR_PROB_SHAPE;
  const int rc = plan_gemm(prob_m, prob_n, prob_k, groupsize == 128, sm_count, max_par, true, p, &grid);
  if (rc != QQQ_OK) return rc;
  const int v[20] = {grid,      p.n_tok,   p.m_tiles, p.n_tiles,  p.k_blocks, p.ksub,    p.k_units,      p.a_tiles,
                     p.a_units, p.a_upc,   p.b_tiles, p.b_tpc,    p.stages_w, p.stages_t, p.unpack_groups,
                     (int)qqq::gemm_smem_bytes(p), p.pair, p.b_step, 0, 0};
  for (int i = 0; i < 20; ++i) out[i] = v[i];
  return QQQ_OK;
}

struct TpScatter {
  int rank, world, rows;
  void* const* part;
};

static int gemm_impl(const void* A, const void* B, void* C, void* D, const void* s1, const void* s2, const void* s3,
                     int prob_m, int prob_n, int prob_k, void* workspace, int groupsize, int dev, void* stream_,
                     int thread_k, int thread_n, int sms, int max_par, int out_mode, const TpScatter* tp = nullptr,
                     const void* bias = nullptr) {
  using namespace qqq;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int M = prob_m, N = prob_n, K = prob_k;
  if (M < 0 || N < 0 || K < 0) {
    set_err("negative problem size");
    return QQQ_ERR_PROB_SHAPE;
  }
  // reference: shape errors are reported before the empty-problem early-out (csrc/qqq_gemm.cu:996-1003)
  if (!reference_shape_ok(N, K, thread_k, thread_n) || (groupsize != -1 && (groupsize <= 0 || K % groupsize != 0))) {
    set_err("problem (m=%d, n=%d, k=%d) not compatible with thread_k=%d, thread_n=%d, groupsize=%d", M, N, K,
            thread_k, thread_n, groupsize);
    return QQQ_ERR_PROB_SHAPE;
  }
  if (groupsize != -1 && groupsize != 128) {
    // the reference instantiates group_blocks in {-1, 8} only (csrc/qqq_gemm.cu:935-945); QuantLinear passes a
    // single group spanning K as the per-channel format (empty s3 -> groupsize -1)
    set_err("groupsize %d not supported (only -1 and 128)", groupsize);
    return QQQ_ERR_KERN_SHAPE;
  }
  if (M == 0 || N == 0 || K == 0) return QQQ_OK;
  const bool grouped = groupsize == 128;
  if (grouped && s3 == nullptr) {
    set_err("s3 is null with groupsize=128");
    return QQQ_ERR_PROB_SHAPE;
  }
  if (grouped && K % 128 != 0) {
    set_err("per-group needs K %% 128 == 0");
    return QQQ_ERR_PROB_SHAPE;
  }
  if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(D) |
       reinterpret_cast<uintptr_t>(s3)) & 15) {
    set_err("A, B, D and s3 must be 16-byte aligned");
    return QQQ_ERR_PROB_SHAPE;
  }
  const DeviceInfo* di = device_info(dev);
  if (!di) {
    set_err("cannot query device %d: %s", dev, cudaGetErrorString(cudaGetLastError()));
    return QQQ_ERR_CUDA;
  }
  if (di->cc_major != 10) {
    set_err("device %d has compute capability %d.x; this library is sm_100a only", dev, di->cc_major);
    return QQQ_ERR_DEVICE;
  }
  DeviceGuard guard(dev);
  if (!guard.ok) {
    set_err("cudaSetDevice(%d) failed", dev);
    return QQQ_ERR_CUDA;
  }

  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.C = reinterpret_cast<int32_t*>(C);
  p.D = reinterpret_cast<__half*>(D);
  p.s1 = reinterpret_cast<const float*>(s1);
  p.s2 = reinterpret_cast<const float*>(s2);
  p.s3 = grouped ? reinterpret_cast<const __half*>(s3) : nullptr;
  p.bias = reinterpret_cast<const __half*>(bias);
  p.locks = reinterpret_cast<int*>(workspace);
  p.M = M;
  p.N = N;
  p.K = K;
  const int sm_count = (sms > 0 && sms < di->sms) ? sms : di->sms;
  int grid = 0;
  {
    const int rc = plan_gemm(M, N, K, grouped, sm_count, max_par, C != nullptr && workspace != nullptr, p, &grid,
                             /*allow_pair=*/out_mode == 0);
    if (rc != QQQ_OK) return rc;
  }
  p.out_mode = out_mode;
  if (tp != nullptr) {
    p.tp_rank = tp->rank;
    p.tp_rows = tp->rows;
    for (int r = 0; r < 8; ++r) p.tp_part[r] = r < tp->world ? reinterpret_cast<__half*>(tp->part[r]) : nullptr;
  }
  // weights are streamed once when a single token tile covers M; tokens are re-read by every CTA
  p.hint_b = p.m_tiles == 1 ? kEvictFirst : kEvictNormal;
  p.hint_a = kEvictLast;
  static const int env_hints = getenv("QQQ_B200_HINTS") ? atoi(getenv("QQQ_B200_HINTS")) : 1;
  if (env_hints == 0) p.hint_a = p.hint_b = kEvictNormal;  // experiments

  CUtensorMap tmap_a, tmap_b;
  if (!encode_2d(&tmap_a, CU_TENSOR_MAP_DATA_TYPE_UINT8, A, (uint64_t)K, (uint64_t)M, (uint64_t)K, kBlockK,
                 (uint32_t)(p.n_tok >> p.pair), CU_TENSOR_MAP_SWIZZLE_128B))
    return QQQ_ERR_CUDA;
  if (!encode_2d(&tmap_b, CU_TENSOR_MAP_DATA_TYPE_INT32, B, (uint64_t)2 * N, (uint64_t)(K / 16), (uint64_t)N * 8,
                 2 * kTileN, 8 * p.ksub, CU_TENSOR_MAP_SWIZZLE_NONE))
    return QQQ_ERR_CUDA;

  cudaError_t e = launch_gemm(tmap_a, tmap_b, p, grouped, grid, dev, stream, use_pdl());
  if (e != cudaSuccess) {
    set_err("kernel launch failed: %s", cudaGetErrorString(e));
    return QQQ_ERR_CUDA;
  }
  g_launches.fetch_add(1);
  return QQQ_OK;
}

int qqq_gemm_sm100a(const void* A, const void* B, void* C, void* D, const void* s1, const void* s2, const void* s3,
                    int prob_m, int prob_n, int prob_k, void* workspace, int groupsize, int dev, void* stream_,
                    int thread_k, int thread_n, int sms, int max_par) {
  return gemm_impl(A, B, C, D, s1, s2, s3, prob_m, prob_n, prob_k, workspace, groupsize, dev, stream_, thread_k, thread_n,
                   sms, max_par, 0);
}

int qqq_gemm_bias_sm100a(const void* A, const void* B, void* C, void* D, const void* s1, const void* s2, const void* s3,
                         const void* bias, int prob_m, int prob_n, int prob_k, void* workspace, int groupsize, int dev,
                         void* stream_, int sms, int max_par) {
  return gemm_impl(A, B, C, D, s1, s2, s3, prob_m, prob_n, prob_k, workspace, groupsize, dev, stream_, -1, -1, sms, max_par, 0,
                   nullptr, bias);
}

int qqq_gemm_reduce_sm100a(const void* A, const void* B, void* C, void* D_multicast, const void* s1, const void* s2,
                           const void* s3, int prob_m, int prob_n, int prob_k, void* workspace, int groupsize, int dev,
                           void* stream_, int thread_k, int thread_n, int sms, int max_par) {
  return gemm_impl(A, B, C, D_multicast, s1, s2, s3, prob_m, prob_n, prob_k, workspace, groupsize, dev, stream_, thread_k,
                   thread_n, sms, max_par, 1);
}

int qqq_gemm_acc_sm100a(const void* A, const void* B, void* C, void* D_int32, const void* s3, int prob_m, int prob_n,
                        int prob_k, void* workspace, int groupsize, int dev, void* stream_, int sms, int max_par) {
  return gemm_impl(A, B, C, D_int32, nullptr, nullptr, s3, prob_m, prob_n, prob_k, workspace, groupsize, dev, stream_, -1, -1,
                   sms, max_par, 2);
}

int qqq_gemm_scatter_sm100a(const void* A, const void* B, void* C, void* const* peer_partials, const void* s1, const void* s2,
                            const void* s3, int prob_m, int prob_n, int prob_k, void* workspace, int groupsize, int dev,
                            void* stream_, int sms, int max_par, int tp_rank, int tp_world, int tp_rows) {
  if (peer_partials == nullptr || tp_world < 1 || tp_world > 8 || tp_rank < 0 || tp_rank >= tp_world || tp_rows < 1 ||
      (long long)tp_rows * tp_world < prob_m) {
    set_err("gemm_scatter: need 1 <= world <= 8, 0 <= rank < world, rows * world >= m (got rank=%d world=%d rows=%d m=%d)",
            tp_rank, tp_world, tp_rows, prob_m);
    return QQQ_ERR_PROB_SHAPE;
  }
  for (int r = 0; r < tp_world; ++r)
    if (peer_partials[r] == nullptr || (reinterpret_cast<uintptr_t>(peer_partials[r]) & 15)) {
      set_err("gemm_scatter: partial-sum buffer of rank %d is null or not 16-byte aligned", r);
      return QQQ_ERR_PROB_SHAPE;
    }
  const TpScatter tp{tp_rank, tp_world, tp_rows, peer_partials};
  // D is unused; the alignment check of gemm_impl still wants a 16-byte aligned pointer
  return gemm_impl(A, B, C, peer_partials[tp_rank], s1, s2, s3, prob_m, prob_n, prob_k, workspace, groupsize, dev, stream_, -1,
                   -1, sms, max_par, 3, &tp);
}

int qqq_tp_reduce_quant_sm100a(const void* partials, void* const* a8_dst, void* a8_multicast, void* const* s1_dst,
                               void* s1_multicast, void* h_out, const void* bias, void* flags, void* const* peer_flags,
                               int tp_rank, int tp_world, int tp_rows, int prob_m, int prob_n, int dev, void* stream_) {
  if (tp_world < 1 || tp_world > 8 || tp_rank < 0 || tp_rank >= tp_world || tp_rows < 1 ||
      (long long)tp_rows * tp_world < prob_m || prob_m < 0 || prob_n <= 0 || prob_n % 16 != 0) {
    set_err("tp_reduce_quant: bad geometry (rank=%d world=%d rows=%d m=%d n=%d)", tp_rank, tp_world, tp_rows, prob_m, prob_n);
    return QQQ_ERR_PROB_SHAPE;
  }
  if (partials == nullptr || flags == nullptr || peer_flags == nullptr || a8_dst == nullptr || s1_dst == nullptr ||
      ((a8_multicast == nullptr) != (s1_multicast == nullptr))) {
    set_err("tp_reduce_quant: null buffer (the two multicast addresses come together or not at all)");
    return QQQ_ERR_PROB_SHAPE;
  }
  for (int r = 0; r < tp_world; ++r)
    if (a8_dst[r] == nullptr || s1_dst[r] == nullptr || peer_flags[r] == nullptr ||
        (reinterpret_cast<uintptr_t>(a8_dst[r]) & 15)) {
      set_err("tp_reduce_quant: buffers of rank %d are null or misaligned", r);
      return QQQ_ERR_PROB_SHAPE;
    }
  if ((reinterpret_cast<uintptr_t>(partials) | reinterpret_cast<uintptr_t>(h_out) | reinterpret_cast<uintptr_t>(bias) |
       reinterpret_cast<uintptr_t>(a8_multicast)) & 15) {
    set_err("tp_reduce_quant: partials, h_out, bias and the multicast address must be 16-byte aligned");
    return QQQ_ERR_PROB_SHAPE;
  }
  const DeviceInfo* di = device_info(dev);
  if (!di) {
    set_err("cannot query device %d", dev);
    return QQQ_ERR_CUDA;
  }
  if (di->cc_major != 10) {
    set_err("device %d has compute capability %d.x; this library is sm_100a only", dev, di->cc_major);
    return QQQ_ERR_DEVICE;
  }
  DeviceGuard guard(dev);
  if (!guard.ok) return QQQ_ERR_CUDA;
  cudaError_t e = qqq::launch_tp_reduce_quant(partials, a8_dst, a8_multicast, s1_dst, s1_multicast, h_out, bias, flags,
                                              peer_flags, tp_rank, tp_world, tp_rows, prob_m, prob_n,
                                              reinterpret_cast<cudaStream_t>(stream_), use_pdl());
  if (e != cudaSuccess) {
    set_err("tp_reduce_quant launch failed: %s", cudaGetErrorString(e));
    return QQQ_ERR_CUDA;
  }
  g_launches.fetch_add(1);
  return QQQ_OK;
}

int qqq_act_quant_strided_sm100a(const void* x, long long ldx, void* q, void* s1, int prob_m, int prob_k, int dev,
                                 void* stream_) {
  if (prob_m < 0 || prob_k <= 0 || prob_k % 8 != 0 || ldx < prob_k || ldx % 8 != 0) {
    set_err("act_quant: K and the row stride must be positive multiples of 8, stride >= K (got m=%d k=%d ld=%lld)", prob_m,
            prob_k, ldx);
    return QQQ_ERR_PROB_SHAPE;
  }
  if (prob_m == 0) return QQQ_OK;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(q) & 7)) {
    set_err("act_quant: x must be 16-byte and q 8-byte aligned");
    return QQQ_ERR_PROB_SHAPE;
  }
  const DeviceInfo* di = device_info(dev);
  if (!di) {
    set_err("cannot query device %d", dev);
    return QQQ_ERR_CUDA;
  }
  if (di->cc_major != 10) {
    set_err("device %d has compute capability %d.x; this library is sm_100a only", dev, di->cc_major);
    return QQQ_ERR_DEVICE;
  }
  DeviceGuard guard(dev);
  if (!guard.ok) return QQQ_ERR_CUDA;
  cudaError_t e =
      qqq::launch_act_quant(x, ldx, q, s1, prob_m, prob_k, reinterpret_cast<cudaStream_t>(stream_), use_pdl());
  if (e != cudaSuccess) {
    set_err("act_quant launch failed: %s", cudaGetErrorString(e));
    return QQQ_ERR_CUDA;
  }
  g_launches.fetch_add(1);
  return QQQ_OK;
}

int qqq_act_quant_sm100a(const void* x, void* q, void* s1, int prob_m, int prob_k, int dev, void* stream_) {
  return qqq_act_quant_strided_sm100a(x, prob_k, q, s1, prob_m, prob_k, dev, stream_);
}

}  // extern "C"
