/*
 * qqq_b200.h — C ABI of libqqq_b200.so: the B200 (sm_100a) W4A8 GEMM behind QQQ's `qqq_gemm()` boundary.
 *
 * Every entry point takes raw device pointers, plain ints and a CUDA stream; none allocates, none
 * synchronises, none touches torch.  Citations are into the reference tree (HandH1998/QQQ).
 *
 * Replaces:
 *   csrc/qqq_gemm.cu:950-1046   int qqq_cuda(...)            -> qqq_gemm_sm100a()          (same argument list)
 *   csrc/qqq_gemm.cu:1048-1106  void qqq_gemm(torch::Tensor…) -> python wrapper qqq_b200.ops.qqq_gemm over this ABI
 *   csrc/pybind.cpp:3-5         QQQ._CUDA.qqq_gemm            -> see INTEGRATION.md for the two-line binding
 *   QQQ/gptq/qlinear/qlinear_marlin.py:265-268 dynamic_quant  -> qqq_act_quant_sm100a()     (row "N2" of SURVEY §8f)
 */
#ifndef QQQ_B200_H_
#define QQQ_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

/* Return codes: 0 and 1/2 mirror ERR_PROB_SHAPE / ERR_KERN_SHAPE of csrc/qqq_gemm.cu:947-948. */
#define QQQ_OK 0
#define QQQ_ERR_PROB_SHAPE 1 /* (M,N,K,groupsize,thread_k,thread_n) not supported                       */
#define QQQ_ERR_KERN_SHAPE 2 /* no kernel instantiation for the requested configuration                */
#define QQQ_ERR_WORKSPACE 3  /* max_par too small for the scratch contract (see below)                 */
#define QQQ_ERR_CUDA 4       /* a CUDA runtime/driver call failed; see qqq_b200_last_error()           */
#define QQQ_ERR_DEVICE 5     /* device is not compute capability 10.x                                  */

/*
 * D[M,N] (fp16) = ((int32)(A[M,K] (int8) x W8[K,N]) * s2[n]) * s1[m]          csrc/qqq_gemm.cu:695-700
 *   per-channel (groupsize == -1): W8 = nibble << 4 (= 16*w4), s2 = s_w/16     csrc/qqq_gemm.cu:146-151
 *   per-group   (groupsize == 128): W8 = RNE((nibble-8) * s3[k/128][n])        csrc/qqq_gemm.cu:167-210
 *
 *   A   int8  [M,K] row-major                      B   int32 [K/16, 2N] reference ("Marlin") packing, untouched
 *   C   int32 [>= 64*max_par, N] scratch           D   fp16  [M,N] row-major (written)
 *   s1  fp32  [M]   per-token scales               s2  fp32  [N] per-channel scales, reference permutation
 *   s3  fp16  [K/groupsize, N] reference permutation, or NULL/ignored when groupsize == -1
 *   workspace int32 [>= N/128*max_par]
 *
 * Scratch contract (identical to the reference's): `workspace` must be all-zero on entry and is returned all-zero
 * (csrc/qqq_gemm.cu:213-237); `C` is pure scratch (need not be zero, is not restored).  C holds split-K partial
 * tiles: each contributing CTA stores its int32 partial into its own slot, the last arriver (lock word in
 * `workspace`) sums the slots in a fixed order.  Results never depend on the partition or on arrival order.
 *
 * thread_k, thread_n, sms: -1 = auto.  thread_k/thread_n are validated like the reference
 * (csrc/qqq_gemm.cu:867-916) and otherwise ignored: this kernel has its own tiling.  sms caps the grid.
 * Alignment: A, B, D, s3 16-byte aligned.  Shapes: the reference's rule, (K % 128 == 0 and N % 64 == 0) or (K % 64 == 0 and
 * N % 128 == 0) (csrc/qqq_gemm.cu:847-865,899-916); per-group additionally K % 128 == 0.  Any M >= 0.
 * The launch goes to `stream` on device `dev` (a device guard is applied; the reference has none).
 */
int qqq_gemm_sm100a(const void* A, const void* B, void* C, void* D, const void* s1, const void* s2,
                    const void* s3, int prob_m, int prob_n, int prob_k, void* workspace, int groupsize,
                    int dev, void* stream /* cudaStream_t */, int thread_k, int thread_n, int sms, int max_par);

/*
 * qqq_gemm_sm100a with the bias add of QuantLinear.forward (QQQ/gptq/qlinear/qlinear_marlin.py:286-288, an eager `D + bias`
 * there) folded into the epilogue: D = fp16(fp16((acc * s2) * s1) + bias[n]) — the add is an fp16 add on the rounded
 * output, so the bits equal the reference's two-step result.  bias fp16 [N] in natural channel order, or NULL.
 */
int qqq_gemm_bias_sm100a(const void* A, const void* B, void* C, void* D, const void* s1, const void* s2, const void* s3,
                         const void* bias, int prob_m, int prob_n, int prob_k, void* workspace, int groupsize, int dev,
                         void* stream /* cudaStream_t */, int sms, int max_par);

/*
 * Tensor-parallel row shards (new work: the reference has no tensor parallelism): the same GEMM on this
 * rank's K-shard, but every 16-byte piece of the output is ADDED (fp16, `multimem.red`) into `D_multicast` instead of
 * stored.  `D_multicast` is the multicast address (cuMulticast* / NVLS, e.g. torch symmetric memory's `multicast_ptr`) of
 * an fp16 [M,N] buffer replicated on every rank of the tensor-parallel group: the NVSwitch applies each add to every
 * replica, so when all ranks' launches have completed, each replica holds the all-reduced output — the all-reduce
 * that follows o_proj / down_proj rides in the GEMM epilogue.  The caller zeroes its replica beforehand and separates
 * producers from consumers with a cross-rank barrier (qqq_b200/tp.py: FusedRowParallelQuantLinear).  All other
 * arguments, the scratch contract and the return codes are those of qqq_gemm_sm100a; s1 are this rank's per-shard
 * token scales.
 */
int qqq_gemm_reduce_sm100a(const void* A, const void* B, void* C, void* D_multicast, const void* s1, const void* s2,
                           const void* s3, int prob_m, int prob_n, int prob_k, void* workspace, int groupsize,
                           int dev, void* stream /* cudaStream_t */, int thread_k, int thread_n, int sms, int max_par);

/*
 * Tensor-parallel row shards as a two-kernel fused exchange — reduce-scatter in the GEMM epilogue, all-gather in the
 * quantisation of the next block's input (new work; replaces "GEMM, NCCL all-reduce, replicated
 * activation quant").  Token rows are owned in blocks of tp_rows = ceil(M / world) rows: rank r owns [r*tp_rows, ...).
 *
 * qqq_gemm_scatter_sm100a: the GEMM of qqq_gemm_sm100a on this rank's K-shard; the fp16 output row m is stored into
 *   peer_partials[m / tp_rows] + ((tp_rank * tp_rows + m % tp_rows) * N): slot `tp_rank` of the owner's partial-sum buffer
 *   (fp16 [world][tp_rows][N], peer-mapped device pointers, e.g. torch symmetric memory's buffer_ptrs; a HOST array of
 *   tp_world pointers, read during the call).  Scratch contract and return codes as qqq_gemm_sm100a.
 * qqq_tp_reduce_quant_sm100a (run by every rank after its scatter GEMM, same stream): waits in-kernel until all ranks'
 *   GEMMs have delivered, sums the world slots of its own rows in fp32 in rank order, rounds once to fp16 (+ bias), applies
 *   the reference's per-token quantisation (qlinear_marlin.py:265-268) and writes int8 rows + fp32 scales into the gathered
 *   buffers a8 [world*tp_rows][N] / s1 [world*tp_rows] of EVERY rank (multicast address when given, else one store per
 *   entry of a8_dst / s1_dst), optionally its fp16 rows into h_out [tp_rows][N]; returns (stream-ordered) when all ranks'
 *   rows have arrived here.  `flags`: 32 zero-initialised uint32 of this rank in peer-mapped memory, peer_flags[r] the same
 *   block on rank r; flags[18] counts waits that timed out (2 s) — never a hang.  All ranks must issue the same sequence of
 *   calls.
 */
int qqq_gemm_scatter_sm100a(const void* A, const void* B, void* C, void* const* peer_partials, const void* s1,
                            const void* s2, const void* s3, int prob_m, int prob_n, int prob_k, void* workspace,
                            int groupsize, int dev, void* stream /* cudaStream_t */, int sms, int max_par, int tp_rank,
                            int tp_world, int tp_rows);
int qqq_tp_reduce_quant_sm100a(const void* partials, void* const* a8_dst, void* a8_multicast, void* const* s1_dst,
                               void* s1_multicast, void* h_out, const void* bias, void* flags, void* const* peer_flags,
                               int tp_rank, int tp_world, int tp_rows, int prob_m, int prob_n, int dev,
                               void* stream /* cudaStream_t */);

/*
 * The same GEMM without its epilogue scales: D_int32[M,N] (int32, row-major) = A[M,K] (int8) x W8[K,N], the exact integer
 * accumulators.  For the bit-exact tensor-parallel mode of row shards: every rank quantises its K-shard
 * of the activations with the SHARED per-token scale (all-reduce-max of the row maxima), the int32 partial sums are
 * all-reduced (exact), and f16((f32(acc) * s2[n]) * s1[m]) is applied once — the result equals the 1-GPU output bit for
 * bit (qqq_b200/tp.py: ExactRowParallelQuantLinear).  s1 / s2 are not arguments; scratch contract as qqq_gemm_sm100a.
 */
int qqq_gemm_acc_sm100a(const void* A, const void* B, void* C, void* D_int32, const void* s3, int prob_m, int prob_n,
                        int prob_k, void* workspace, int groupsize, int dev, void* stream /* cudaStream_t */, int sms,
                        int max_par);

/*
 * Per-token dynamic int8 quantisation of activations, bit-identical to
 * QQQ/gptq/qlinear/qlinear_marlin.py:265-268:
 *   s1[m] = fp32( fp16( max_k |x[m,k]| / 127 ) );   q[m,k] = int8( clamp( rint( x[m,k] / s1[m] ), -128, 127 ) )
 *   x fp16 [M,K] row-major (K % 8 == 0, 16-byte aligned)  ->  q int8 [M,K], s1 fp32 [M]
 */
int qqq_act_quant_sm100a(const void* x, void* q, void* s1, int prob_m, int prob_k, int dev, void* stream);
/* Same, for x given as a column slice of a wider row-major matrix (row stride ldx halves, ldx % 8 == 0): lets the
 * output slices of a merged QKV / gate-up GEMM be quantised in place, without a gather copy. */
int qqq_act_quant_strided_sm100a(const void* x, long long ldx, void* q, void* s1, int prob_m, int prob_k, int dev,
                                 void* stream);

/* Library/ABI version (major*100 + minor) and the last error string of the calling thread. */
int qqq_b200_version(void);
const char* qqq_b200_last_error(void);

/* Introspection used by bench.py / tests: number of kernel launches issued by this library so far. */
long long qqq_b200_launch_count(void);

/* The tiling / schedule the library would use for a problem on `sm_count` SMs (pure host computation, no GPU):
 * out[20] = {grid, n_tok, m_tiles, n_tiles, k_blocks, ksub, k_units, a_tiles, a_units, a_upc, b_tiles, b_tpc,
 *            stages_w, stages_t, unpack_groups, smem_bytes, pair, b_step, 0, 0}.  Tiles are 128 channels x n_tok tokens; a
 * SCHEDULED tile is one tile, or with pair = 1 two adjacent 128-channel tiles handled by a CTA pair (cluster of 2,
 * cta_group::2; grid is then even and CTAs 2c, 2c+1 walk the schedule of index c).  Scheduled tile id = mt +
 * m_tiles * column; a unit is `ksub` 128-deep k-blocks of one scheduled tile.  Scheduled tiles [0,a_tiles) are cut
 * along K: schedule index b owns their units [b*a_upc, (b+1)*a_upc) (processed first); tiles [a_tiles,
 * a_tiles+b_tiles) are whole and dealt round-robin: index b owns a_tiles + b + i*b_step (at most b_tpc of them), so that
 * concurrently running CTAs share weight columns.  Used by the CPU tests to check that every unit
 * is covered exactly once. */
int qqq_b200_plan(int prob_m, int prob_n, int prob_k, int groupsize, int sm_count, int max_par, int* out);

#ifdef __cplusplus
}
#endif
#endif /* QQQ_B200_H_ */
