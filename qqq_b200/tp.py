"""Tensor-parallel sharding of QuantLinear (new work — the reference has no TP).

Packed tensors shard by pure slicing, no repack:
  column-parallel (split N; q,k,v,gate,up):  B[:, 2*n0:2*n1], s_channel[:, n0:n1], s_group[:, n0:n1]   (n0,n1 % 64 == 0)
      -> each rank's columns are bit-identical to the 1-GPU columns; no collective if the consumer is sharded alike.
  row-parallel (split K; o_proj, down_proj):  B[k0/16:k1/16], s_group[k0/128:k1/128], s_channel replicated
      -> one all-reduce of the fp16 [M, N] partial outputs per layer (NCCL over NVLink).  Each rank quantises its own
         K-shard of the activations (per-shard s1), so the sum is NOT bit-equal to the 1-GPU result (it is usually
         closer to the fp result); tolerance parity is tested against the full-K oracle.
"""
from __future__ import annotations

import torch
import torch.distributed as dist
import torch.nn as nn

from . import ops
from .qlinear import QuantizedActivation, QuantLinear


def split_sizes(total: int, world: int, unit: int):
    """Split `total` (a multiple of `unit`) into `world` contiguous shards, each a multiple of `unit`,
    sizes differing by at most one unit (larger shards first)."""
    assert total % unit == 0
    blocks = total // unit
    base, rem = divmod(blocks, world)
    sizes = [(base + (1 if r < rem else 0)) * unit for r in range(world)]
    offs = [0]
    for s in sizes:
        offs.append(offs[-1] + s)
    return sizes, offs


def shard_quant_linear(ql: QuantLinear, rank: int, world: int, mode: str) -> QuantLinear:
    """Return a new QuantLinear holding rank's shard of `ql` ('column' = split outfeatures, 'row' = split infeatures)."""
    K, N = ql.infeatures, ql.outfeatures
    per_group = ql.per_group
    gs = ql.group_size if per_group else -1
    if mode == "column":
        _, offs = split_sizes(N, world, 64)
        n0, n1 = offs[rank], offs[rank + 1]
        out = QuantLinear(ql.bits, gs, K, n1 - n0, bias=ql.bias is not None)
        out.B = ql.B[:, 2 * n0:2 * n1].contiguous()
        out.s_channel = ql.s_channel[:, n0:n1].contiguous()
        if per_group:
            out.s_group = ql.s_group[:, n0:n1].contiguous()
        else:
            out.s_group = ql.s_group
        if ql.bias is not None:
            out.bias = ql.bias[n0:n1].contiguous()
    elif mode == "row":
        unit = 128 if per_group else 64
        _, offs = split_sizes(K, world, unit)
        k0, k1 = offs[rank], offs[rank + 1]
        out = QuantLinear(ql.bits, gs, k1 - k0, N, bias=ql.bias is not None and rank == 0)
        out.B = ql.B[k0 // 16:k1 // 16].contiguous()
        out.s_channel = ql.s_channel
        if per_group:
            out.s_group = ql.s_group[k0 // 128:k1 // 128].contiguous()
        else:
            out.s_group = ql.s_group
        if ql.bias is not None and rank == 0:
            out.bias = ql.bias
    else:
        raise ValueError(mode)
    dev = ql.B.device
    out.workspace = torch.zeros(max(out.outfeatures // 128 * 16, 16), dtype=torch.int32, device=dev)
    out.reduce_buffer = torch.zeros((out.max_par * 64, out.outfeatures), dtype=torch.int32, device=dev)
    return out


class ColumnParallelQuantLinear(nn.Module):
    """y_local = x @ W[:, shard]; optionally all-gathers the shards along the feature dim."""

    def __init__(self, shard: QuantLinear, gather_output: bool = False, group=None):
        super().__init__()
        self.shard = shard
        self.gather_output = gather_output
        self.group = group

    def forward(self, x):
        """x: the replicated fp16 activation, or the `QuantizedActivation` a ScatterRowParallelQuantLinear delivered."""
        y = self.shard(x)
        if not self.gather_output:
            return y
        world = dist.get_world_size(self.group)
        rank = dist.get_rank(self.group)
        # shards may be uneven (split_sizes: N/64 blocks do not always divide by world): every rank's width is known
        # from the split rule, so uneven gathers need no size exchange
        widths = torch.tensor([y.shape[-1]], device=y.device)
        all_w = [torch.empty_like(widths) for _ in range(world)]
        dist.all_gather(all_w, widths, group=self.group)
        all_w = [int(w.item()) for w in all_w]
        wmax = max(all_w)
        y2 = y.reshape(-1, y.shape[-1])
        if all_w[rank] < wmax:
            y2 = torch.nn.functional.pad(y2, (0, wmax - all_w[rank]))
        parts = [torch.empty_like(y2) for _ in range(world)]
        dist.all_gather(parts, y2.contiguous(), group=self.group)
        out = torch.cat([t[:, :w] for t, w in zip(parts, all_w)], dim=-1)
        return out.reshape(y.shape[:-1] + (out.shape[-1],))


class RowParallelQuantLinear(nn.Module):
    """y = all_reduce_sum(x_local @ W[shard, :])  (fp16 partial sums, one collective per layer)."""

    def __init__(self, shard: QuantLinear, group=None):
        super().__init__()
        self.shard = shard
        self.group = group

    def forward(self, x_local):
        y = self.shard(x_local)
        dist.all_reduce(y, op=dist.ReduceOp.SUM, group=self.group)
        return y


def _replicated_bias(shard: QuantLinear, group=None):
    """`shard_quant_linear(..., "row")` keeps the bias on rank 0 only (it is added once, before the all-reduce).  Modules
    that add it AFTER their reduction need it on every rank: broadcast it from rank 0 of the group (collective call)."""
    dev = shard.B.device
    flag = torch.tensor([1 if shard.bias is not None else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
    if int(flag.item()) == 0:
        return None
    b = shard.bias.clone() if shard.bias is not None else torch.empty(shard.outfeatures, dtype=torch.half, device=dev)
    src = dist.get_global_rank(group, 0) if group is not None else 0
    dist.broadcast(b, src=src, group=group)
    return b


class ExactRowParallelQuantLinear(nn.Module):
    """Row-parallel linear whose output equals the 1-GPU QuantLinear's BIT FOR BIT ("exact" mode):

      1. the per-token scale is shared: all-reduce-MAX of the ranks' row maxima of |x_local| (a [M] vector), then the
         reference's own scale expression on it — the same s1 the 1-GPU module computes from the full row;
      2. every rank quantises its K-shard with that s1 (the reference's eager expression, qlinear_marlin.py:267) and runs
         the GEMM without epilogue scales (qqq_gemm_acc_sm100a): exact int32 partial sums;
      3. all-reduce-SUM of the int32 [M, N] partials (exact in any order; 2x the bytes of the fp16 mode);
      4. f16((f32(acc) * s2[n]) * s1[m]) once, in the reference's order (csrc/qqq_gemm.cu:695-700).

    Costs two collectives and twice the bytes; RowParallelQuantLinear / FusedRowParallelQuantLinear are the fast modes."""

    def __init__(self, shard: QuantLinear, group=None):
        super().__init__()
        self.shard = shard
        self.group = group
        from .qlinear import _scale_perm_positions

        _, sp32 = _scale_perm_positions(shard.s_channel.device)
        inv = torch.empty_like(sp32)
        inv[sp32] = torch.arange(32, device=sp32.device)
        self.register_buffer("_unperm32", inv, persistent=False)  # natural channel n sits at permuted position inv[n % 32]
        self.bias = _replicated_bias(shard, group)

    def forward(self, x_local):
        ql = self.shard
        out_shape = x_local.shape[:-1] + (ql.outfeatures,)
        A = x_local.reshape(-1, x_local.shape[-1]).half()
        amax = A.abs().max(dim=-1, keepdim=True)[0]
        dist.all_reduce(amax, op=dist.ReduceOp.MAX, group=self.group)
        s1 = amax.div(127.0).to(torch.float32)  # the reference's expression on the full-row maximum
        A8 = (A / s1).round().clamp(-128, 127).to(torch.int8)
        acc = torch.empty(A.shape[0], ql.outfeatures, dtype=torch.int32, device=A.device)
        ops.qqq_gemm_acc(A8, ql.B, ql.reduce_buffer, acc, ql.s_group, ql.workspace, ql.max_par)
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=self.group)
        N = ql.outfeatures
        s2 = ql.s_channel.reshape(N // 32, 32)[:, self._unperm32].reshape(1, N)  # natural channel order
        out = ((acc.to(torch.float32) * s2) * s1).to(torch.float16).reshape(out_shape)
        return out + self.bias if self.bias is not None else out


# ---------------------------------------------------------------------------------------------------------------
# GEMM + all-reduce in one kernel
# ---------------------------------------------------------------------------------------------------------------
class _SymmMemBackend:
    """torch symmetric memory (CUDA VMM + NVLS multicast) as plumbing: one fp16 buffer replicated on every rank of
    the group, its multicast address, and a device-side cross-rank barrier on the current stream."""

    def __init__(self, group):
        import torch.distributed._symmetric_memory as symm

        self.symm = symm
        self.group = group if group is not None else dist.group.WORLD
        try:  # older builds need the group enabled explicitly; newer ones deprecate the call
            symm.enable_symm_mem_for_group(self.group.group_name)
        except Exception:
            pass

    def alloc(self, numel: int, device):
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        t = self.symm.empty(numel, dtype=torch.float16, device=device)
        hdl = self.symm.rendezvous(t, self.group)
        mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
        if mc == 0:
            raise RuntimeError("symmetric memory has no multicast address on this system (NVSwitch/NVLS required)")
        return t, mc, (lambda: hdl.barrier(channel=0))


class AllReduceWorkspace:
    """Replicated fp16 output buffers shared by every fused row-parallel layer of a model.

    Protocol of a fused call on every rank (all ranks issue the same calls in the same order):
        zero my replica of the region  ->  barrier  ->  GEMM whose epilogue adds into the multicast address  ->  barrier
    * first barrier: no rank's adds can reach a replica before its owner has zeroed it;
    * second barrier: every rank's kernel has completed, so every replica holds the sum over ranks.
    The protocol carries no state from one call to the next (the region is zeroed when its size is known, by the call
    that uses it), so it is safe under CUDA-graph capture and replay whatever the number of calls per graph and
    whatever the sequence of shapes.  Calls alternate between two buffers only so that an output stays valid while the
    next fused call runs (o_proj's output feeds the residual add while down_proj is already reducing); it is
    overwritten by the call after that.  Consumers that need it longer clone it.
    """

    NBUF = 2

    def __init__(self, max_tokens: int, max_features: int, group=None, device=None, backend=None):
        self.capacity = max_tokens * max_features
        self.backend = backend if backend is not None else _SymmMemBackend(group)
        self.bufs, self.mc, self.barriers = [], [], []
        for _ in range(self.NBUF):
            t, mc, bar = self.backend.alloc(self.capacity, device)
            self.bufs.append(t)
            self.mc.append(mc)
            self.barriers.append(bar)
        self.turn = 0

    def begin(self, M: int, N: int):
        """Zero this rank's replica of an [M, N] region and wait until every rank has done so.
        -> (multicast address to reduce into, this rank's [M, N] view of the result)"""
        n = M * N
        if n > self.capacity:
            raise RuntimeError(f"AllReduceWorkspace: {M} x {N} exceeds the capacity of {self.capacity} elements")
        cur = self.turn
        self.bufs[cur][:n].zero_()
        self.barriers[cur]()
        return self.mc[cur], self.bufs[cur][:n].view(M, N)

    def end(self):
        """Wait until every rank's reducing GEMM has completed."""
        self.barriers[self.turn]()
        self.turn = (self.turn + 1) % self.NBUF


class FusedRowParallelQuantLinear(nn.Module):
    """y = sum over ranks of x_local @ W[shard, :], with the sum performed by the GEMM epilogue itself: each rank's kernel
    adds its fp16 output tiles into a multicast buffer (`multimem.red`, NVSwitch), so no NCCL all-reduce follows the GEMM.
    One-shot: every replica receives every rank's tile, i.e. (world-1) x the output bytes arrive per GPU — the right
    trade at decode and at world = 2; for large prefill batches at world >= 4 the NCCL path moves fewer bytes.
    Numerics: per-shard activation scales like RowParallelQuantLinear, fp16 adds in arrival order (tolerance parity)."""

    def __init__(self, shard: QuantLinear, workspace: AllReduceWorkspace):
        super().__init__()
        self.shard = shard
        object.__setattr__(self, "ws", workspace)
        self.bias = _replicated_bias(shard, getattr(workspace.backend, "group", None))

    def forward(self, x_local):
        ql = self.shard
        out_shape = x_local.shape[:-1] + (ql.outfeatures,)
        A = x_local.reshape(-1, x_local.shape[-1]).half()
        mc, out = self.ws.begin(A.shape[0], ql.outfeatures)  # the barrier's wait overlaps nothing useful later: issue it first
        q, s1 = ql.dynamic_quant(A)
        ops.qqq_gemm_reduce(q, ql.B, ql.reduce_buffer, mc, s1, ql.s_channel, ql.s_group, ql.workspace, ql.outfeatures,
                            ql.max_par)
        self.ws.end()
        out = out.reshape(out_shape)
        return out + self.bias if self.bias is not None else out


# ---------------------------------------------------------------------------------------------------------------
# Row-parallel exchange fused into the kernels on both sides of it: reduce-scatter in the GEMM epilogue (peer stores),
# all-gather in the activation quantisation of the next block (multicast stores).  No collective-library call.
# ---------------------------------------------------------------------------------------------------------------
class _SymmBytesBackend:
    """torch symmetric memory as plumbing: ONE byte buffer per rank, peer-mapped on every rank (+ its NVLS multicast
    address when the fabric has one)."""

    def __init__(self, group=None):
        import torch.distributed._symmetric_memory as symm

        self.symm = symm
        self.group = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        self.world = dist.get_world_size(self.group)
        self._hdls = []  # keep every mapping alive as long as the backend lives

    def alloc(self, nbytes: int, device):
        """-> (local uint8 tensor (zeroed, all ranks synchronised), [address on rank r for r in ranks], multicast address | 0)"""
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        t = self.symm.empty(nbytes, dtype=torch.uint8, device=device)
        hdl = self.symm.rendezvous(t, self.group)
        t.zero_()
        torch.cuda.synchronize(device)
        hdl.barrier(channel=0)
        torch.cuda.synchronize(device)
        self._hdls.append(hdl)
        mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
        return t, [int(a) for a in hdl.buffer_ptrs], mc


class ScatterWorkspace:
    """Symmetric buffers shared by every ScatterRowParallelQuantLinear of a model (they run one after the other):

        partial  fp16 [world][rows][N]   slot s: rank s's partial output for the rows THIS rank owns (written by peers)
        a8       int8 [world*rows][N]    the gathered quantised activations, identical on every rank after a call
        s1       fp32 [world*rows]       their per-token scales
        flags    uint32 [32]             epoch flags of the in-kernel cross-rank synchronisation (see tp_reduce_quant.cu)

    with rows = ceil(M / world) chosen per call.  One buffer set is enough: a rank can only overwrite `partial` / `a8` of a
    peer after that peer has finished reading them (arrive-1 / arrive-2 flags), see the kernel's header."""

    def __init__(self, max_tokens: int, max_features: int, group=None, device=None, backend=None, use_multicast=True):
        self.backend = backend if backend is not None else _SymmBytesBackend(group)
        self.rank, self.world = self.backend.rank, self.backend.world
        self.max_tokens, self.max_features = max_tokens, max_features
        rows = -(-max_tokens // self.world)
        mpad = rows * self.world
        al = lambda n: (n + 255) // 256 * 256  # noqa: E731
        self.off_part = 0
        self.off_a8 = self.off_part + al(2 * mpad * max_features)
        self.off_s1 = self.off_a8 + al(mpad * max_features)
        self.off_flags = self.off_s1 + al(4 * mpad)
        nbytes = self.off_flags + 256
        self.buf, self.ptrs, mc = self.backend.alloc(nbytes, device)
        self.mc = mc if use_multicast else 0
        self.calls = 0

    def geometry(self, M: int, N: int):
        if M > self.max_tokens or N > self.max_features:
            raise RuntimeError(f"ScatterWorkspace: {M} x {N} exceeds the capacity {self.max_tokens} x {self.max_features}")
        return -(-M // self.world)  # rows owned per rank

    def views(self, M: int, N: int):
        """This rank's (a8 [M, N] int8, s1 [M, 1] fp32) views of the gathered buffers."""
        a8 = self.buf[self.off_a8:self.off_a8 + M * N].view(torch.int8).view(M, N)
        s1 = self.buf[self.off_s1:self.off_s1 + 4 * M].view(torch.float32).view(M, 1)
        return a8, s1

    def timeouts(self) -> int:
        """Number of in-kernel waits that gave up (2 s) since the workspace was created: 0 unless a rank went missing."""
        return int(self.buf[self.off_flags:self.off_flags + 128].view(torch.int32)[18].item())


class ScatterRowParallelQuantLinear(nn.Module):
    """Row-parallel linear (o_proj / down_proj; split K) whose exchange is fused into the kernels around it:

        x_local --act-quant--> GEMM on this rank's K-shard; epilogue stores each output row into the partial-sum slot of
                               the rank that owns the row (peer stores over NVLink, qqq_gemm_scatter_sm100a)
                --------------> owner: fp32 sum of the `world` fp16 partials in rank order -> fp16 (+ bias) -> per-token
                               int8 quantisation -> int8 rows + scales multicast to every rank (qqq_tp_reduce_quant_sm100a)

    Returns the `QuantizedActivation` of the all-reduced output for ALL tokens (what the next block's column-parallel
    linears consume: pass it to ColumnParallelQuantLinear / QuantLinear.forward); with keep_hidden=True `self.hidden`
    holds this rank's rows of the fp16 sum (sequence-sharded residual stream).  Numerics: per-shard activation scales
    like RowParallelQuantLinear; the cross-rank sum is fp32 in rank order, rounded once — deterministic, reproduced bit
    for bit by `reference_reduce_quant` below."""

    def __init__(self, shard: QuantLinear, workspace: ScatterWorkspace, keep_hidden: bool = False):
        super().__init__()
        self.shard = shard
        object.__setattr__(self, "ws", workspace)
        self.keep_hidden = keep_hidden
        self.hidden = None
        self.bias = _replicated_bias(shard, getattr(workspace.backend, "group", None))

    def forward(self, x_local):
        ql, ws = self.shard, self.ws
        if isinstance(x_local, QuantizedActivation):
            q, s1, lead = x_local.q, x_local.s1, x_local.lead
        else:
            lead = tuple(x_local.shape[:-1])
            q, s1 = ql.dynamic_quant(x_local.reshape(-1, x_local.shape[-1]).half())
        M, N = q.shape[0], ql.outfeatures
        rows = ws.geometry(M, N)
        base = ws.ptrs
        ops.qqq_gemm_scatter(q, ql.B, ql.reduce_buffer, [b + ws.off_part for b in base], s1, ql.s_channel, ql.s_group,
                             ql.workspace, N, ws.rank, ws.world, rows, ql.max_par)
        h = None
        if self.keep_hidden:
            my_rows = max(0, min(rows, M - ws.rank * rows))
            h = torch.empty(max(my_rows, 1), N, dtype=torch.float16, device=q.device)
        dev = q.get_device() if q.is_cuda else -1
        ops.tp_reduce_quant(base[ws.rank] + ws.off_part, [b + ws.off_a8 for b in base], ws.mc + ws.off_a8 if ws.mc else 0,
                            [b + ws.off_s1 for b in base], ws.mc + ws.off_s1 if ws.mc else 0, h, self.bias,
                            base[ws.rank] + ws.off_flags, [b + ws.off_flags for b in base], ws.rank, ws.world, rows, M, N, dev)
        ws.calls += 1
        if self.keep_hidden:
            self.hidden = h[:max(0, min(rows, M - ws.rank * rows))]
        a8, s1g = ws.views(M, N)
        return QuantizedActivation(a8, s1g, lead)


def reference_reduce_quant(partials, bias=None):
    """Torch restatement of qqq_tp_reduce_quant_sm100a for tests / the bench's parity leg: `partials` = the fp16 [M, N]
    partial outputs of ranks 0..world-1 -> (h fp16 [M, N], a8 int8 [M, N], s1 fp32 [M, 1]).  Same operation order: fp32
    adds in rank order starting from 0, one rounding to fp16, fp16 bias add, then the reference's quantisation expression
    (qlinear_marlin.py:265-268) as it evaluates on a CUDA device."""
    acc = torch.zeros_like(partials[0], dtype=torch.float32)
    for p in partials:
        acc = acc + p.float()
    h = acc.half()
    if bias is not None:
        h = h + bias
    s1 = h.abs().max(dim=-1, keepdim=True)[0].div(127.0).to(torch.float32)
    a8 = (h / s1).round().clamp(-128, 127).to(torch.int8)
    return h, a8, s1
