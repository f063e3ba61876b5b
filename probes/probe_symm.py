"""Dev probe (torchrun, N GPUs): what the box offers for in-kernel collectives, and what NCCL costs at the sizes of the
row-parallel layers.   torchrun --nproc-per-node N probes/probe_symm.py"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)


def p0(*a):
    if rank == 0:
        print(*a, flush=True)


def graph_us(fn, n, reps=5):
    fn()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(2):
            fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) * 1e3 / (reps * n)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


p0("torch", torch.__version__, "nccl", torch.cuda.nccl.version(), "world", world)
for nbytes in (8 << 20, 4 << 20, 1 << 20, 256 << 10, 64 << 10, 8 << 10):
    x = torch.ones(nbytes // 2, dtype=torch.float16, device=dev)
    try:
        t = graph_us(lambda: dist.all_reduce(x), 20)
        p0(f"nccl all_reduce fp16 {nbytes >> 10:6d} KB: {t:8.2f} us  busbw {2 * (world - 1) / world * nbytes / t * 1e-3:7.1f} GB/s")
    except Exception as e:
        p0("nccl all_reduce graph failed", repr(e)[:200])
    y = torch.empty(nbytes // 2 // world, dtype=torch.float16, device=dev)
    try:
        t = graph_us(lambda: dist.reduce_scatter_tensor(y, x), 20)
        p0(f"nccl reduce_scatter   {nbytes >> 10:6d} KB: {t:8.2f} us")
        xi = torch.empty(nbytes // 2, dtype=torch.int8, device=dev)
        yi = torch.empty(nbytes // 2 // world, dtype=torch.int8, device=dev)
        t = graph_us(lambda: dist.all_gather_into_tensor(xi, yi), 20)
        p0(f"nccl all_gather int8  {nbytes >> 11:6d} KB: {t:8.2f} us")
    except Exception as e:
        p0("nccl rs/ag failed", repr(e)[:200])

try:
    import torch.distributed._symmetric_memory as symm

    try:
        symm.enable_symm_mem_for_group(dist.group.WORLD.group_name)
    except Exception as e:
        p0("enable_symm_mem_for_group:", repr(e)[:120])
    t = symm.empty(1 << 22, dtype=torch.float16, device=dev)
    h = symm.rendezvous(t, dist.group.WORLD)
    p0("symm handle:", type(h).__name__, [a for a in dir(h) if not a.startswith("_")])
    for a in ("rank", "world_size", "buffer_ptrs", "signal_pad_ptrs", "multicast_ptr", "buffer_size", "signal_pad_size",
              "buffer_ptrs_dev", "signal_pad_ptrs_dev"):
        try:
            print(f"[rank {rank}] {a} =", getattr(h, a), flush=True)
        except Exception as e:
            print(f"[rank {rank}] {a} failed: {e!r}"[:200], flush=True)
    try:
        tb = graph_us(lambda: h.barrier(channel=0), 50)
        p0(f"symm barrier: {tb:.2f} us")
    except Exception as e:
        p0("symm barrier graph failed", repr(e)[:300])
    # peer write through get_buffer
    try:
        peer = h.get_buffer((rank + 1) % world, (1 << 22,), torch.float16)
        src = torch.full((1 << 22,), float(rank + 1), dtype=torch.float16, device=dev)
        tcopy = graph_us(lambda: peer.copy_(src), 10)
        p0(f"peer copy 8 MB: {tcopy:.2f} us = {8.388608e6 / tcopy * 1e-3:.1f} GB/s")
        h.barrier(channel=0)
        torch.cuda.synchronize()
        print(f"[rank {rank}] my buffer after peer write: {float(t[0])} (expect {float((rank - 1) % world + 1)})", flush=True)
    except Exception as e:
        p0("peer copy failed", repr(e)[:300])
    try:
        t.zero_()
        h.barrier(channel=0)
        ar = torch.ops.symm_mem.multimem_all_reduce_(t[: 1 << 20].fill_(1.0), "sum", dist.group.WORLD.group_name)
        torch.cuda.synchronize()
        p0("multimem_all_reduce_ ok:", float(ar[0]))
        for n in (4 << 20, 512 << 10, 32 << 10):
            v = t[:n]
            tm = graph_us(lambda: torch.ops.symm_mem.multimem_all_reduce_(v, "sum", dist.group.WORLD.group_name), 20)
            p0(f"symm multimem_all_reduce_ fp16 {n * 2 >> 10:6d} KB: {tm:8.2f} us")
            tm = graph_us(lambda: torch.ops.symm_mem.two_shot_all_reduce_(v, "sum", dist.group.WORLD.group_name), 20)
            p0(f"symm two_shot_all_reduce_ fp16 {n * 2 >> 10:6d} KB: {tm:8.2f} us")
    except Exception as e:
        p0("multimem_all_reduce_ failed", repr(e)[:300])
except Exception as e:
    p0("symmetric memory unavailable:", repr(e)[:400])

torch.cuda.synchronize()
dist.barrier()
sys.stdout.flush()
os._exit(0)
