"""GPU x2: the fused row-parallel exchange (tp.ScatterRowParallelQuantLinear = qqq_gemm_scatter_sm100a +
qqq_tp_reduce_quant_sm100a: reduce-scatter in the GEMM epilogue over peer stores, all-gather in the activation quant over
multicast stores) against a torch restatement fed with the per-rank partial outputs of the plain kernel — BIT-exact:
int8 rows, per-token scales and the fp16 hidden rows.  Needs two B200s of one NVLink domain; skipped on a 1-GPU box
(bench.py's `tp_parity` leg repeats the check inside every multi-GPU bench run)."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, gs, multicast, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dev = torch.device(f"cuda:{rank}")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from oracle import qqq_oracle as O
    from qqq_b200 import tp
    from test_tp_gloo import _full_module

    K, N = 1024, 512
    ws = tp.ScatterWorkspace(max_tokens=300, max_features=N, device=dev, use_multicast=multicast)
    res = []
    mods = {}
    for it, M in enumerate((300, 7, 64, 1, 300, 129)):
        p = O.make_problem(M, K, N, gs, seed=60 + it)
        full = _full_module(p, K, N, gs)
        if it % 2 == 1:
            full.bias = torch.linspace(-1, 1, N).half()
        shard = tp.shard_quant_linear(full, rank, world, "row").to(dev)
        _, offs = tp.split_sizes(K, world, 128 if gs != -1 else 64)
        x_loc = torch.from_numpy(p["x"][:, offs[rank]:offs[rank + 1]].copy()).to(dev)
        bias = shard.bias
        shard.bias = None  # the plain partial output, without the bias rank 0 carries
        part = shard(x_loc)
        shard.bias = bias
        parts = [torch.empty_like(part) for _ in range(world)]
        dist.all_gather(parts, part)
        mod = tp.ScatterRowParallelQuantLinear(shard, ws, keep_hidden=True)
        h_ref, a8_ref, s1_ref = tp.reference_reduce_quant(parts, mod.bias)
        qa = mod(x_loc)
        torch.cuda.synchronize()
        rows = -(-M // world)
        mine = slice(rank * rows, min(M, (rank + 1) * rows))
        ok = (torch.equal(qa.q, a8_ref) and torch.equal(qa.s1.view(torch.int32), s1_ref.view(torch.int32))
              and torch.equal(mod.hidden.view(torch.int16), h_ref[mine].view(torch.int16))
              and tuple(qa.q.shape) == (M, N) and int(shard.workspace.abs().sum()) == 0 and ws.timeouts() == 0)
        # tolerance parity against the 1-GPU module on the full K (per-shard activation scales differ by design)
        full = full.to(dev)
        if full.bias is not None:
            full.bias = full.bias.to(dev)  # assigned after construction (bias=False): a plain attribute, not a buffer
        y_one = full(torch.from_numpy(p["x"]).to(dev)).float()
        rel = float((h_ref.float() - y_one).abs().max() / y_one.abs().max().clamp_min(1e-6))
        res.append((M, bool(ok), rel))
        mods[M] = (mod, x_loc, a8_ref, s1_ref)
    # CUDA-graph replay: epochs live on the device, so a captured sequence of fused calls can be replayed
    mod, x_loc, a8_ref, s1_ref = mods[129]
    mod2, x2, a8_2, s1_2 = mods[64]
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        mod(x_loc)
        mod2(x2)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    dist.barrier()
    with torch.cuda.graph(g):
        qa = mod(x_loc)
        keep = (qa.q.clone(), qa.s1.clone())
        qb = mod2(x2)
    ok_graph = True
    for _ in range(3):
        g.replay()
        torch.cuda.synchronize()
        ok_graph = ok_graph and torch.equal(keep[0], a8_ref) and torch.equal(keep[1], s1_ref) and torch.equal(qb.q, a8_2)
    ok_graph = ok_graph and ws.timeouts() == 0
    flag = torch.tensor([1 if (ok_graph and all(r[1] for r in res)) else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put((res, bool(ok_graph), int(flag.item())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("multicast", [True, False])
@pytest.mark.parametrize("gs", [-1, 128])
def test_scatter_row_parallel_bit_exact_vs_restatement(gs, multicast):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, gs, multicast, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    try:
        res, ok_graph, all_ranks = q.get(timeout=120)
        for pr in procs:
            pr.join(timeout=60)
            assert pr.exitcode == 0
    finally:
        for pr in procs:
            if pr.is_alive():
                pr.kill()
    for (M, ok, rel) in res:
        assert ok, f"fused exchange differs from its restatement at M={M}"
        assert rel <= 6e-2, (M, rel)
    assert ok_graph, "CUDA-graph replay of fused calls"
    assert all_ranks == 1, "some rank disagreed"
