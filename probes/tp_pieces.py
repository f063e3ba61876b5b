"""Dev probe (torchrun, N GPUs): where the time of a fused row-parallel layer goes.  torchrun --nproc-per-node N probes/tp_pieces.py"""
import os, sys
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
import qqq_b200
from qqq_b200 import ops, tp


def graph_us(fn, n=20, reps=5):
    fn(); torch.cuda.synchronize()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize(); dist.barrier()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay(); torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) * 1e3 / (reps * n)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


M = 1024
gen = torch.Generator(device=dev).manual_seed(5 + rank)
for (K0, N) in ((4096, 4096), (11008, 4096)):
    K = tp.split_sizes(K0, world, 64)[0][rank]
    ql = qqq_b200.QuantLinear(4, -1, K, N, bias=False).to(dev)
    ql.B = torch.randint(-2**31, 2**31 - 1, (K // 16, 2 * N), dtype=torch.int32, device=dev, generator=gen)
    ql.s_channel = torch.full((1, N), 1e-3, device=dev)
    x = torch.randn(M, K, device=dev, generator=gen).half()
    q, s1 = ops.dynamic_quant(x)
    qa = qqq_b200.QuantizedActivation(q, s1)
    res = {}
    res["act_quant_local"] = graph_us(lambda: ops.dynamic_quant(x))
    res["gemm_plain"] = graph_us(lambda: ql(qa))
    for mc in (True, False):
        ws = tp.ScatterWorkspace(M, N, device=dev, use_multicast=mc)
        mod = tp.ScatterRowParallelQuantLinear(ql, ws)
        rows = ws.geometry(M, N)
        base = ws.ptrs
        def gemm_scatter():
            ops.qqq_gemm_scatter(q, ql.B, ql.reduce_buffer, [b + ws.off_part for b in base], s1, ql.s_channel, ql.s_group,
                                 ql.workspace, N, ws.rank, ws.world, rows, ql.max_par)
        def k1():
            ops.tp_reduce_quant(base[ws.rank] + ws.off_part, [b + ws.off_a8 for b in base], ws.mc + ws.off_a8 if ws.mc else 0,
                                [b + ws.off_s1 for b in base], ws.mc + ws.off_s1 if ws.mc else 0, None, None,
                                base[ws.rank] + ws.off_flags, [b + ws.off_flags for b in base], ws.rank, ws.world, rows, M, N, lr)
        tag = "mc" if mc else "uc"
        if mc:
            res["gemm_scatter"] = graph_us(gemm_scatter)
        res[f"reduce_quant_{tag}"] = graph_us(k1)
        res[f"fused_pair_{tag}"] = graph_us(lambda: mod(qa))
        res[f"quant+fused_pair_{tag}"] = graph_us(lambda: mod(x))
        res[f"timeouts_{tag}"] = ws.timeouts()
    y = ql(qa)
    res["gemm+nccl_allreduce"] = graph_us(lambda: dist.all_reduce(ql(qa)))
    if rank == 0:
        print(f"world={world} M={M} K={K0}/{world} N={N}: " + "  ".join(f"{k}={v:.2f}" if isinstance(v, float) else f"{k}={v}" for k, v in res.items()), flush=True)
torch.cuda.synchronize(); dist.barrier(); sys.stdout.flush(); os._exit(0)
