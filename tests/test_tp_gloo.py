"""CPU, world_size=2, gloo: the tensor-parallel host logic (shard slicing of the packed tensors + the collectives).
The per-rank GEMM is evaluated with the CPU oracle on the SHARD's buffers — exactly what the CUDA kernel is
parity-tested against — so this checks that slicing B / s_channel / s_group needs no repack."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _full_module(p, K, N, gs):
    import qqq_b200

    ql = qqq_b200.QuantLinear(4, gs, K, N, bias=False)
    ql.B.copy_(torch.from_numpy(p["B"]))
    ql.s_channel.copy_(torch.from_numpy(p["s2"]))
    if gs != -1:
        ql.s_group.copy_(torch.from_numpy(p["s3"]))
    return ql


def _oracle_forward(ql, x_np):
    from oracle import qqq_oracle as O

    A8, s1 = O.dynamic_quant(x_np, cuda_semantics=True)
    s3 = ql.s_group.numpy() if ql.s_group.numel() else None
    return O.qqq_gemm_oracle(A8, ql.B.numpy(), s1, ql.s_channel.numpy(), s3)


def _worker(rank, world, port, gs, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import qqq_oracle as O
    from qqq_b200 import tp

    M, K, N = 6, 512, 256
    p = O.make_problem(M, K, N, gs, seed=11)
    full = _full_module(p, K, N, gs)
    D_full = _oracle_forward(full, p["x"])

    # column parallel: shards are bit-identical column blocks of the 1-GPU result
    col = tp.shard_quant_linear(full, rank, world, "column")
    y_loc = torch.from_numpy(_oracle_forward(col, p["x"]))
    parts = [torch.empty_like(y_loc) for _ in range(world)]
    dist.all_gather(parts, y_loc)
    D_col = torch.cat(parts, dim=1).numpy()
    ok_col = np.array_equal(D_col.view(np.uint16), D_full.view(np.uint16))

    # row parallel: per-shard activation scales + fp16 all-reduce => tolerance parity vs the full-K result
    row = tp.shard_quant_linear(full, rank, world, "row")
    _, offs = tp.split_sizes(K, world, 128 if gs != -1 else 64)
    x_loc = p["x"][:, offs[rank]:offs[rank + 1]]
    y_row = torch.from_numpy(_oracle_forward(row, x_loc)).float()
    dist.all_reduce(y_row)
    err_row = float(np.abs(y_row.numpy() - D_full.astype(np.float32)).max())
    scale = float(np.abs(D_full.astype(np.float32)).max())

    # exact mode: shared s1 (all-reduce-max of the row absmax) + int32 partial sums => bit-equal
    amax = torch.from_numpy(np.abs(x_loc.astype(np.float32)).max(axis=1))
    dist.all_reduce(amax, op=dist.ReduceOp.MAX)
    A8_full, s1_full = O.dynamic_quant(p["x"], cuda_semantics=True)
    s1_shared = (amax.numpy().astype(np.float16).astype(np.float32) * np.float32(1 / 127)).astype(np.float16).astype(np.float32)
    ok_s1 = np.array_equal(s1_shared.reshape(-1, 1), s1_full)
    A8_loc = A8_full[:, offs[rank]:offs[rank + 1]]
    W8 = O.weights_int8(row.B.numpy(), row.s_group.numpy() if row.s_group.numel() else None)
    acc = torch.from_numpy((A8_loc.astype(np.int64) @ W8.astype(np.int64)))
    dist.all_reduce(acc)
    W8f = O.weights_int8(full.B.numpy(), full.s_group.numpy() if full.s_group.numel() else None)
    ok_acc = np.array_equal(acc.numpy(), A8_full.astype(np.int64) @ W8f.astype(np.int64))
    if rank == 0:
        out.put((ok_col, err_row, scale, ok_s1, ok_acc))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("gs", [-1, 128])
def test_tp2_sharding_gloo(gs):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, gs, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = q.get(timeout=240)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    ok_col, err_row, scale, ok_s1, ok_acc = res
    assert ok_col, "column-parallel shards must reproduce the 1-GPU columns bit-for-bit"
    assert err_row <= 3e-2 * max(scale, 1.0), f"row-parallel error {err_row} vs scale {scale}"
    assert ok_s1 and ok_acc, "exact row-parallel mode (shared s1 + int32 partials) must be bit-equal"


def test_split_sizes():
    from qqq_b200 import tp

    sizes, offs = tp.split_sizes(11008, 8, 64)
    assert sum(sizes) == 11008 and all(s % 64 == 0 for s in sizes) and max(sizes) - min(sizes) <= 64
    assert offs[0] == 0 and offs[-1] == 11008
    sizes, _ = tp.split_sizes(8192, 8, 128)
    assert sizes == [1024] * 8


# ---------------------------------------------------------------------------------------------------------------
# GEMM + all-reduce fusion: the host protocol (double-buffered replicated output, zero -> reduce -> barrier)
# ---------------------------------------------------------------------------------------------------------------
class _EmulatedMulticast:
    """Stands in for symmetric memory on CPU: `alloc` hands out a plain local tensor and a fake 16-byte aligned
    'multicast address'; the patched qqq_gemm_reduce (below) performs what the NVSwitch would: every rank's partial
    output is added into every rank's replica (here: gloo all-reduce of the partials, then a local add)."""

    def __init__(self):
        self.by_addr = {}
        self.barrier_calls = 0

    def alloc(self, numel, device):
        t = torch.full((numel,), 777.0, dtype=torch.float16)  # dirty on purpose: the workspace must zero it
        addr = 4096 * (len(self.by_addr) + 1)
        self.by_addr[addr] = t

        def bar():
            self.barrier_calls += 1
            dist.barrier()

        return t, addr, bar


def _fused_worker(rank, world, port, gs, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import qqq_oracle as O
    from qqq_b200 import ops, tp

    K, N = 512, 256
    backend = _EmulatedMulticast()

    def dq(x):
        q, s = O.dynamic_quant(x.numpy(), cuda_semantics=True)
        return torch.from_numpy(q), torch.from_numpy(s)

    def gemm_reduce(A, B, C, mc, s1, s2, s3, workspace, prob_n, max_par=16, sms=-1):
        part = torch.from_numpy(O.qqq_gemm_oracle(A.numpy(), B.numpy(), s1.numpy(), s2.numpy(),
                                                  s3.numpy() if s3.numel() else None)).float()
        dist.all_reduce(part)  # "every replica receives every rank's tile"
        buf = backend.by_addr[mc]
        view = buf[: part.numel()].view(part.shape)
        view += part.half()

    ops.dynamic_quant = dq
    ops.qqq_gemm_reduce = gemm_reduce
    ws = tp.AllReduceWorkspace(max_tokens=16, max_features=N, backend=backend)
    ok, worst = True, 0.0
    _, offs = tp.split_sizes(K, world, 128 if gs != -1 else 64)
    mods = []
    for seed in (21, 22):  # two different row-parallel layers sharing the workspace, like o_proj / down_proj
        p = O.make_problem(16, K, N, gs, seed=seed)
        full = _full_module(p, K, N, gs)
        mods.append((p, full, tp.FusedRowParallelQuantLinear(tp.shard_quant_linear(full, rank, world, "row"), ws)))
    prev = None
    for it, M in enumerate((16, 5, 9, 16, 1)):  # shrinking and growing batches reuse the two buffers
        p, full, fused = mods[it % 2]
        x = p["x"][:M]
        ref = _oracle_forward(full, x).astype(np.float32)
        y = fused(torch.from_numpy(x[:, offs[rank]:offs[rank + 1]].copy()))
        if prev is not None:  # the previous output must still be intact after the next fused call has run
            ok = ok and torch.equal(prev[0], prev[1])
        prev = (y, y.clone())
        err = float(np.abs(y.float().numpy() - ref).max())
        worst = max(worst, err / max(float(np.abs(ref).max()), 1.0))
        ok = ok and y.shape == (M, N)
    # two barriers per fused call (replicas zeroed / all adds landed), no state carried between calls
    ok = ok and backend.barrier_calls == 2 * 5
    if rank == 0:
        out.put((ok, worst))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("gs", [-1, 128])
def test_fused_row_parallel_protocol_gloo(gs):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_fused_worker, args=(r, 2, port, gs, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    ok, worst = q.get(timeout=240)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert ok, "workspace protocol: shapes, buffer lifetime or barrier count wrong"
    # per-shard activation scales: tolerance parity against the full-K result (relative to the batch's own max, M down to 1)
    assert worst <= 6e-2, f"fused row-parallel relative error {worst}"


# ---------------------------------------------------------------------------------------------------------------
# exact mode: shared s1 + int32 partial sums -> bit-equal to the 1-GPU module
# ---------------------------------------------------------------------------------------------------------------
def _exact_worker(rank, world, port, gs, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import qqq_oracle as O
    from qqq_b200 import ops, tp

    def gemm_acc(A, B, C, D32, s3, workspace, max_par=16, sms=-1):  # the C-ABI call on the oracle's integer model
        W8 = O.weights_int8(B.numpy(), s3.numpy() if s3.numel() else None)
        D32.copy_(torch.from_numpy((A.numpy().astype(np.int64) @ W8.astype(np.int64)).astype(np.int32)))

    ops.qqq_gemm_acc = gemm_acc
    M, K, N = 9, 512, 256
    p = O.make_problem(M, K, N, gs, seed=41)
    full = _full_module(p, K, N, gs)
    full.bias = torch.linspace(-2, 2, N).half()
    # what the 1-GPU module computes (torch CPU evaluates .div(127.0) as a true division: cuda_semantics=False)
    A8, s1 = O.dynamic_quant(p["x"], cuda_semantics=False)
    want = torch.from_numpy(O.qqq_gemm_oracle(A8, p["B"], s1, p["s2"], p["s3"])) + full.bias
    shard = tp.shard_quant_linear(full, rank, world, "row")
    assert (shard.bias is not None) == (rank == 0)
    mod = tp.ExactRowParallelQuantLinear(shard)
    _, offs = tp.split_sizes(K, world, 128 if gs != -1 else 64)
    y = mod(torch.from_numpy(p["x"][:, offs[rank]:offs[rank + 1]].copy()).reshape(3, 3, -1))
    ok = y.shape == (3, 3, N) and torch.equal(y.reshape(M, N).view(torch.int16), want.view(torch.int16))
    flags = torch.tensor([1 if ok else 0])
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put(bool(flags.item()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("gs", [-1, 128])
def test_exact_row_parallel_is_bit_equal_to_one_gpu_gloo(gs):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_exact_worker, args=(r, 2, port, gs, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    ok = q.get(timeout=240)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert ok, "exact row-parallel mode must reproduce the 1-GPU output bit for bit on every rank (bias included)"


# ---------------------------------------------------------------------------------------------------------------
# fused exchange (scatter GEMM + reduce/quant/gather): host protocol on CPU — buffer geometry, ownership of token rows,
# ragged M, bias, hand-over of the QuantizedActivation to column-parallel consumers
# ---------------------------------------------------------------------------------------------------------------
class _EmulatedSymm:
    """Stands in for symmetric memory on CPU: a local byte buffer, fake per-rank 'addresses' and no multicast."""

    def __init__(self):
        self.rank, self.world, self.group = dist.get_rank(), dist.get_world_size(), None

    def alloc(self, nbytes, device):
        self.buf = torch.full((nbytes,), 0x5A, dtype=torch.uint8)  # dirty: nothing may depend on stale contents ...
        self.buf[-256:] = 0                                        # ... except the flag block, which starts at zero
        return self.buf, [(r + 1) << 32 for r in range(self.world)], 0


def _scatter_worker(rank, world, port, gs, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import qqq_oracle as O
    from qqq_b200 import QuantizedActivation, ops, tp

    K, N = 512, 256
    backend = _EmulatedSymm()
    ws = tp.ScatterWorkspace(max_tokens=16, max_features=N, backend=backend)

    def dq(x):
        q, s = O.dynamic_quant(x.numpy(), cuda_semantics=False)
        return torch.from_numpy(q), torch.from_numpy(s)

    def gemm_scatter(A, B, C, peer_partials, s1, s2, s3, workspace, prob_n, tp_rank, tp_world, tp_rows, max_par=16, sms=-1):
        assert peer_partials == [((r + 1) << 32) + ws.off_part for r in range(world)] and (tp_rank, tp_world) == (rank, world)
        part = torch.from_numpy(O.qqq_gemm_oracle(A.numpy(), B.numpy(), s1.numpy(), s2.numpy(),
                                                  s3.numpy() if s3.numel() else None))
        M = part.shape[0]
        padded = torch.zeros(tp_rows * world, prob_n, dtype=torch.float16)
        padded[:M] = part
        allp = [torch.empty_like(padded) for _ in range(world)]
        dist.all_gather(allp, padded)  # "every rank stores rows into their owner's slot"
        slots = backend.buf[ws.off_part:ws.off_part + 2 * world * tp_rows * prob_n].view(torch.float16).view(world, tp_rows, prob_n)
        for s in range(world):
            slots[s] = allp[s][rank * tp_rows:(rank + 1) * tp_rows]

    def reduce_quant(partials_ptr, a8_dst, a8_mc, s1_dst, s1_mc, h_out, bias, flags_ptr, peer_flags, tp_rank, tp_world,
                     tp_rows, prob_m, prob_n, dev):
        assert partials_ptr == ((rank + 1) << 32) + ws.off_part and a8_mc == 0 and s1_mc == 0
        slots = backend.buf[ws.off_part:ws.off_part + 2 * world * tp_rows * prob_n].view(torch.float16).view(world, tp_rows, prob_n)
        acc = torch.zeros(tp_rows, prob_n)
        for s in range(world):
            acc = acc + slots[s].float()
        h = acc.half()
        if bias is not None:
            h = h + bias
        my_rows = max(0, min(tp_rows, prob_m - rank * tp_rows))
        if h_out is not None:
            h_out[:my_rows] = h[:my_rows]
        q, s1 = dq(h)
        allq = [torch.empty_like(q) for _ in range(world)]
        alls = [torch.empty_like(s1) for _ in range(world)]
        dist.all_gather(allq, q)
        dist.all_gather(alls, s1)
        mpad = tp_rows * world
        backend.buf[ws.off_a8:ws.off_a8 + mpad * prob_n].view(torch.int8).view(mpad, prob_n).copy_(torch.cat(allq))
        backend.buf[ws.off_s1:ws.off_s1 + 4 * mpad].view(torch.float32).view(mpad, 1).copy_(torch.cat(alls))

    ops.dynamic_quant = dq
    ops.qqq_gemm_scatter = gemm_scatter
    ops.tp_reduce_quant = reduce_quant
    _, offs = tp.split_sizes(K, world, 128 if gs != -1 else 64)
    ok = True
    for it, M in enumerate((16, 5, 1, 9)):
        p = O.make_problem(M, K, N, gs, seed=70 + it)
        full = _full_module(p, K, N, gs)
        if it % 2:
            full.bias = torch.linspace(-1, 1, N).half()
        shard = tp.shard_quant_linear(full, rank, world, "row")
        mod = tp.ScatterRowParallelQuantLinear(shard, ws, keep_hidden=True)
        x_loc = torch.from_numpy(p["x"][:, offs[rank]:offs[rank + 1]].copy())
        # what it must equal: restatement on the per-rank partial outputs (bias-free), gathered
        b = shard.bias
        shard.bias = None
        part = torch.from_numpy(_oracle_forward_cpu(shard, x_loc.numpy()))
        shard.bias = b
        parts = [torch.empty_like(part) for _ in range(world)]
        dist.all_gather(parts, part)
        h_ref, a8_ref, s1_ref = tp.reference_reduce_quant(parts, mod.bias)
        qa = mod(x_loc.reshape(1, M, -1))
        rows = -(-M // world)
        mine = slice(rank * rows, min(M, (rank + 1) * rows))
        ok = ok and isinstance(qa, QuantizedActivation) and qa.lead == (1, M)
        ok = ok and torch.equal(qa.q, a8_ref) and torch.equal(qa.s1, s1_ref) and torch.equal(mod.hidden, h_ref[mine])
        # the consumer side: a column-parallel linear takes the QuantizedActivation as is
        col = tp.ColumnParallelQuantLinear(tp.shard_quant_linear(_full_module(O.make_problem(M, N, 128, -1, seed=5), N, 128, -1),
                                                                 rank, world, "column"))
        seen = {}
        ops.qqq_gemm = lambda A, B, C, D, s1, s2, s3, w, *a: seen.update(A=A, s1=s1, D=tuple(D.shape))
        y = col(qa)
        ok = ok and seen["A"] is qa.q and seen["s1"] is qa.s1 and tuple(y.shape) == (1, M, 64)
    flags = torch.tensor([1 if ok else 0])
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put(bool(flags.item()))
    dist.barrier()
    dist.destroy_process_group()


def _oracle_forward_cpu(ql, x_np):
    from oracle import qqq_oracle as O

    A8, s1 = O.dynamic_quant(x_np, cuda_semantics=False)
    s3 = ql.s_group.numpy() if ql.s_group.numel() else None
    return O.qqq_gemm_oracle(A8, ql.B.numpy(), s1, ql.s_channel.numpy(), s3)


@pytest.mark.parametrize("gs", [-1, 128])
def test_scatter_row_parallel_protocol_gloo(gs):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_scatter_worker, args=(r, 2, port, gs, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    ok = q.get(timeout=240)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert ok, "fused exchange: geometry, ownership, bias or hand-over to the consumers wrong"


def test_column_parallel_gather_with_uneven_shards_gloo():
    """gather_output with N/64 not divisible by world (e.g. 11008 over 8 ranks) must not assume equal shards."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_uneven_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    ok = q.get(timeout=240)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert ok


def _uneven_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import qqq_oracle as O
    from qqq_b200 import tp

    M, K, N = 5, 256, 192  # 3 blocks of 64 over 2 ranks: 128 + 64
    p = O.make_problem(M, K, N, -1, seed=3)
    full = _full_module(p, K, N, -1)
    shard = tp.shard_quant_linear(full, rank, world, "column")
    assert shard.outfeatures == (128 if rank == 0 else 64)
    mod = tp.ColumnParallelQuantLinear(shard, gather_output=True)
    mod.shard.forward = lambda x: torch.from_numpy(_oracle_forward_cpu(shard, x.numpy()))
    y = mod(torch.from_numpy(p["x"]))
    want = torch.from_numpy(_oracle_forward_cpu(full, p["x"]))
    ok = torch.equal(y, want)
    flags = torch.tensor([1 if ok else 0])
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put(bool(flags.item()))
    dist.barrier()
    dist.destroy_process_group()
