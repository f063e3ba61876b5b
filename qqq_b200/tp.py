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

from .qlinear import QuantLinear


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
        y = self.shard(x)
        if not self.gather_output:
            return y
        world = dist.get_world_size(self.group)
        parts = [torch.empty_like(y) for _ in range(world)]  # equal shards only
        dist.all_gather(parts, y.contiguous(), group=self.group)
        return torch.cat(parts, dim=-1)


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
