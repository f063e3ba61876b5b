// Tensor-parallel row shards, second half of the fused exchange (new work: the reference has no tensor parallelism,
// SURVEY.md §8e / row N4).  The first half is the epilogue of the row-shard GEMM (qqq_gemm_kernel<.., kOutScatter>):
// every rank stores the fp16 rows of its partial output straight into the partial-sum slot of the rank that OWNS those
// token rows (peer stores over NVLink, tile by tile, while the GEMM is still running).  This kernel then does, on the
// owner, for its rows_cap = ceil(M / world) rows:
//
//   wait until every rank's GEMM has delivered (arrive-1 flags)                     cross-rank barrier, in-kernel
//   h[m, :]   = fp16( fp32 sum over source ranks 0..world-1, in that order )        reduce-scatter, deterministic
//   (+ bias)                                                                        QuantLinear.forward's `D + bias`
//   s1[m]     = fp32(fp16(max|h[m,:]| / 127));  a8[m,:] = int8(rint(h / s1))        the reference's dynamic_quant
//                                                                                   (qlinear_marlin.py:265-268)
//   a8[m, :], s1[m] -> EVERY rank's gathered buffers (multimem.st, or one store per peer)   all-gather, 1 byte/elt
//   signal "my rows are delivered" (arrive-2) and wait for everybody else's
//
// so that when the kernel completes, every rank holds the int8 activations + per-token scales of the all-reduced
// output for all M tokens — exactly what the column-parallel linears of the next block consume — and the fp16 hidden
// state stays sequence-sharded (h_out, optional).  Compared with "GEMM, NCCL all-reduce of fp16 [M, N], activation quant
// replicated on every rank": 1.5 instead of 4 bytes per element cross NVLink, the quant runs on M / world rows, and no
// collective library call sits between the GEMMs (NCCL's floor on this box: 14 us per call, 48 us at 8 MB / 2 ranks).
//
// The sum is taken in fp32 over the fp16 partials in source-rank order and rounded once: bit-reproducible (the tests
// compare with a torch restatement bit for bit) and at least as accurate as an fp16 ring all-reduce.
//
// Flags: one block of 32 words per rank in symmetric memory, written by peers with st.release.sys and polled locally
// with ld.acquire.sys.  [0..7] arrive-1 per source rank, [8..15] arrive-2 per source rank, [16] epoch (number of
// completed calls), [17] finished-CTA counter, [18] time-outs seen.  Values are epochs (monotonic), so nothing is reset
// between calls and CUDA-graph replays need no host state.  Every wait is bounded (kSpinTimeoutNs): a missing peer shows
// up as flags[18] != 0 and garbage output, never as a hung GPU.
#include "../../include/qqq_b200.h"
#include <cstdlib>

#include "quant_common.cuh"

namespace qqq {

constexpr int kTpThreads = 256;     // threads per token row
constexpr int kTpRowsPerCta = 2;    // rows a CTA works on at a time
constexpr int kTpMaxWorld = 8;
constexpr unsigned long long kSpinTimeoutNs = 2000000000ull;  // 2 s

struct TpReduceParams {
  const __half* part;           // local [world][rows_cap][N]: slot s = rank s's partial output for MY rows
  int8_t* a8_dst[kTpMaxWorld];  // gathered int8 [world * rows_cap][N] on every rank (peer pointers) ...
  float* s1_dst[kTpMaxWorld];   // ... and the gathered per-token scales [world * rows_cap]
  int8_t* a8_mc;                // multicast addresses of the same two buffers, or null (then: one store per rank)
  float* s1_mc;
  __half* h_out;                // optional local fp16 [rows_cap][N]: my rows of the reduced output (sequence-sharded)
  const __half* bias;           // optional [N]
  uint32_t* flags;              // my flag block
  uint32_t* peer_flags[kTpMaxWorld];
  int rank, world, rows_cap, M, N;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// wait until *flag has reached `epoch`; false on time-out
__device__ __forceinline__ bool wait_epoch(const uint32_t* flag, uint32_t epoch) {
  if ((int)(ld_acquire_sys(flag) - epoch) >= 0) return true;
  const unsigned long long t0 = global_timer_ns();
  for (;;) {
#pragma unroll 1
    for (int i = 0; i < 64; ++i)
      if ((int)(ld_acquire_sys(flag) - epoch) >= 0) return true;
    if (global_timer_ns() - t0 > kSpinTimeoutNs) return false;
  }
}
__device__ __forceinline__ void multimem_st_16(void* mc, const uint4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void multimem_st_4(void* mc, uint32_t v) {
  asm volatile("multimem.st.relaxed.sys.global.b32 [%0], %1;" ::"l"(mc), "r"(v) : "memory");
}

// fp32 sum of one 16-byte chunk (8 halves) over the source ranks, in rank order, rounded to fp16 once (+ bias, fp16 add)
__device__ __forceinline__ uint4 sum_chunk(const uint4* __restrict__ part, size_t slot_stride16, int world, size_t idx,
                                           const uint4* __restrict__ bias16, int col16) {
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int s = 0; s < world; ++s) {
    const uint4 v = __ldcg(part + (size_t)s * slot_stride16 + idx);  // written by peers: never through the non-coherent path
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __half2 h = *reinterpret_cast<const __half2*>(&w[i]);
      acc[2 * i] = __fadd_rn(acc[2 * i], __low2float(h));
      acc[2 * i + 1] = __fadd_rn(acc[2 * i + 1], __high2float(h));
    }
  }
  uint32_t o[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __half2 h = __floats2half2_rn(acc[2 * i], acc[2 * i + 1]);
    o[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  if (bias16 != nullptr) {
    const uint4 b = __ldg(bias16 + col16);
    const uint32_t bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __half2 h = __hadd2_rn(*reinterpret_cast<__half2*>(&o[i]), *reinterpret_cast<const __half2*>(&bw[i]));
      o[i] = *reinterpret_cast<uint32_t*>(&h);
    }
  }
  return make_uint4(o[0], o[1], o[2], o[3]);
}

// kTpRowsPerCta token rows at a time per CTA (one group of kTpThreads threads per row, rows strided over the groups of the
// grid): system-scope fences are the expensive part of this kernel and there is one per CTA, so few, fat CTAs.  A thread owns NCH units of 16 consecutive channels (two 16-byte chunks of
// fp16 in, ONE 16-byte piece of int8 out: NVLink moves 16-byte multicast stores at several times the rate of 8-byte
// ones); the units stay in registers between the reduction / max pass and the quantise pass (NCH == 0: any N, the second
// pass recomputes the sum).
template <int NCH>
__global__ void __launch_bounds__(kTpThreads * kTpRowsPerCta) tp_reduce_quant_kernel(const TpReduceParams p) {
  grid_launch_dependents();  // the next GEMM may run its prologue and prefetch its weights under this kernel
  grid_dependency_wait();    // the row-shard GEMM of this rank has completed: its peer stores are performed
  __shared__ uint32_t sh_epoch;
  __shared__ uint32_t sh_last;
  __shared__ __half red_all[kTpRowsPerCta][kTpThreads / 32];
  const int group = threadIdx.x / kTpThreads;  // which of the CTA's concurrent rows this thread works on
  __half* red = red_all[group];
  const int tid = threadIdx.x % kTpThreads;    // position inside the row group
  if (threadIdx.x == 0) sh_epoch = *reinterpret_cast<volatile uint32_t*>(p.flags + 16) + 1u;
  __syncthreads();
  const uint32_t epoch = sh_epoch;
  // arrive-1: "my GEMM's rows are in your slots" to every rank (one CTA sends), then every CTA waits for all senders
  if (blockIdx.x == 0 && threadIdx.x < p.world) st_release_sys(p.peer_flags[threadIdx.x] + p.rank, epoch);
  if (threadIdx.x < p.world) {
    if (!wait_epoch(p.flags + threadIdx.x, epoch)) atomicAdd(p.flags + 18, 1u);
  }
  __syncthreads();

  const int N8 = p.N >> 3;    // 16-byte chunks of fp16 per row
  const int N16 = p.N >> 4;   // units of 16 channels per row (N % 16 == 0)
  const int my_rows = max(0, min(p.rows_cap, p.M - p.rank * p.rows_cap));
  const size_t slot_stride16 = (size_t)p.rows_cap * N8;
  const uint4* part16 = reinterpret_cast<const uint4*>(p.part);
  const uint4* bias16 = reinterpret_cast<const uint4*>(p.bias);
  constexpr int NC = NCH > 0 ? NCH : 1;
  for (int row = blockIdx.x * kTpRowsPerCta + group; row < my_rows; row += gridDim.x * kTpRowsPerCta) {
    const size_t row16 = (size_t)row * N8;
    const size_t grow = (size_t)p.rank * p.rows_cap + row;  // row of the gathered buffers
    uint4 cache[2 * NC];
    uint32_t m = 0;
    auto load_unit = [&](int u, uint4& a, uint4& b) {
      a = sum_chunk(part16, slot_stride16, p.world, row16 + 2 * u, bias16, 2 * u);
      b = sum_chunk(part16, slot_stride16, p.world, row16 + 2 * u + 1, bias16, 2 * u + 1);
      m = hmax2_u32(hmax2_u32(m, habs2_u32(a.x)), habs2_u32(a.y));
      m = hmax2_u32(hmax2_u32(m, habs2_u32(a.z)), habs2_u32(a.w));
      m = hmax2_u32(hmax2_u32(m, habs2_u32(b.x)), habs2_u32(b.y));
      m = hmax2_u32(hmax2_u32(m, habs2_u32(b.z)), habs2_u32(b.w));
    };
    if (NCH > 0) {
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        const int u = tid + j * kTpThreads;
        if (u < N16) {
          load_unit(u, cache[2 * j], cache[2 * j + 1]);
        } else {
          cache[2 * j] = cache[2 * j + 1] = make_uint4(0, 0, 0, 0);
        }
      }
    } else {
      for (int u = tid; u < N16; u += kTpThreads) {
        uint4 a, b;
        load_unit(u, a, b);
      }
    }
    __half2 mh = *reinterpret_cast<__half2*>(&m);
    __half mx = __hmax_nan(__low2half(mh), __high2half(mh));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = __hmax_nan(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    named_bar_sync(1 + group, kTpThreads);  // `red` of the previous row has been read by everybody in the row group
    if ((tid & 31) == 0) red[tid >> 5] = mx;
    named_bar_sync(1 + group, kTpThreads);
    __half t = red[0];
#pragma unroll
    for (int i = 1; i < kTpThreads / 32; ++i) t = __hmax_nan(t, red[i]);
    const float s = token_scale(t);
    if (tid == 0) {
      if (p.s1_mc != nullptr) {
        multimem_st_4(p.s1_mc + grow, __float_as_uint(s));
      } else {
        for (int r = 0; r < p.world; ++r) p.s1_dst[r][grow] = s;
      }
    }
    const bool fast = s > 0.f && s < __int_as_float(0x7F800000);
    const float rcp = __frcp_rn(s);
    auto emit = [&](int u, const uint4& a, const uint4& b) {
      const uint4 q = fast ? make_uint4(quant4<true>(a.x, a.y, s, rcp), quant4<true>(a.z, a.w, s, rcp),
                                        quant4<true>(b.x, b.y, s, rcp), quant4<true>(b.z, b.w, s, rcp))
                           : make_uint4(quant4<false>(a.x, a.y, s, rcp), quant4<false>(a.z, a.w, s, rcp),
                                        quant4<false>(b.x, b.y, s, rcp), quant4<false>(b.z, b.w, s, rcp));
      const size_t off = grow * (size_t)p.N + (size_t)u * 16;
      if (p.a8_mc != nullptr) {
        multimem_st_16(p.a8_mc + off, q);
      } else {
        for (int r = 0; r < p.world; ++r) *reinterpret_cast<uint4*>(p.a8_dst[r] + off) = q;
      }
      if (p.h_out != nullptr) {
        reinterpret_cast<uint4*>(p.h_out)[row16 + 2 * u] = a;
        reinterpret_cast<uint4*>(p.h_out)[row16 + 2 * u + 1] = b;
      }
    };
    if (NCH > 0) {
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        const int u = tid + j * kTpThreads;
        if (u < N16) emit(u, cache[2 * j], cache[2 * j + 1]);
      }
    } else {
      for (int u = tid; u < N16; u += kTpThreads) {
        uint4 a, b;
        uint32_t keep = m;
        load_unit(u, a, b);
        m = keep;
        emit(u, a, b);
      }
    }
  }

  // arrive-2: the last CTA of this rank to finish tells every rank "my rows are delivered", then waits for the others,
  // so that the kernel's completion means: all M rows of a8 / s1 are in place on this rank.  A system-scope fence costs
  // 2-3 us on this fabric (profiles/r02/call_g), so there is exactly ONE on the critical path: every CTA fences its own
  // stores (one thread, after the CTA-wide barrier: cumulativity covers the other threads' stores) BEFORE it bumps the
  // finished-CTA counter; whoever sees the counter complete therefore knows that every CTA's rows have been performed
  // at system scope and can raise the flags with plain system-scope stores.
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    const uint32_t done = atomicAdd(p.flags + 17, 1u);
    sh_last = (done == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (sh_last) {
    if (threadIdx.x == 0) {
      p.flags[17] = 0;
      p.flags[16] = epoch;
    }
    if (threadIdx.x < p.world) {
      st_relaxed_sys(p.peer_flags[threadIdx.x] + 8 + p.rank, epoch);
      if (!wait_epoch(p.flags + 8 + threadIdx.x, epoch)) atomicAdd(p.flags + 18, 1u);
    }
  }
}

template <int NCH>
static cudaError_t launch_tp(const TpReduceParams& p, int grid, cudaStream_t stream, bool pdl) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kTpThreads * kTpRowsPerCta);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, tp_reduce_quant_kernel<NCH>, p);
}

cudaError_t launch_tp_reduce_quant(const void* part, void* const* a8_dst, void* a8_mc, void* const* s1_dst, void* s1_mc,
                                   void* h_out, const void* bias, void* flags, void* const* peer_flags, int rank, int world,
                                   int rows_cap, int M, int N, cudaStream_t stream, bool pdl) {
  TpReduceParams p;
  p.part = reinterpret_cast<const __half*>(part);
  for (int r = 0; r < kTpMaxWorld; ++r) {
    p.a8_dst[r] = r < world ? reinterpret_cast<int8_t*>(a8_dst[r]) : nullptr;
    p.s1_dst[r] = r < world ? reinterpret_cast<float*>(s1_dst[r]) : nullptr;
    p.peer_flags[r] = r < world ? reinterpret_cast<uint32_t*>(peer_flags[r]) : nullptr;
  }
  p.a8_mc = reinterpret_cast<int8_t*>(a8_mc);
  p.s1_mc = reinterpret_cast<float*>(s1_mc);
  p.h_out = reinterpret_cast<__half*>(h_out);
  p.bias = reinterpret_cast<const __half*>(bias);
  p.flags = reinterpret_cast<uint32_t*>(flags);
  p.rank = rank;
  p.world = world;
  p.rows_cap = rows_cap;
  p.M = M;
  p.N = N;
  // one CTA per SM measured best (fewer fences and counter bumps; rows are grid-strided); QQQ_B200_TPGRID overrides
  static const int env_grid = getenv("QQQ_B200_TPGRID") ? atoi(getenv("QQQ_B200_TPGRID")) : 148;
  const int my_rows = M - rank * rows_cap < 0 ? 0 : (M - rank * rows_cap < rows_cap ? M - rank * rows_cap : rows_cap);
  // every rank launches at least one CTA: it still takes part in both flag exchanges
  const int want = (my_rows + kTpRowsPerCta - 1) / kTpRowsPerCta;
  const int grid = want < 1 ? 1 : (want > env_grid ? env_grid : want);
  const int nch = (N / 16 + kTpThreads - 1) / kTpThreads;  // 16-channel units per thread
  if (nch <= 1) return launch_tp<1>(p, grid, stream, pdl);
  if (nch <= 2) return launch_tp<2>(p, grid, stream, pdl);
  return launch_tp<0>(p, grid, stream, pdl);
}

}  // namespace qqq
