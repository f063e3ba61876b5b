"""GPU x2: GEMM + all-reduce in one kernel (tp.FusedRowParallelQuantLinear: multimem.red epilogue over NVSwitch) against
the NCCL path and against the fp32 sum of the per-rank GEMM outputs.  Needs two B200s of one NVSwitch domain: skipped
on a single-GPU box.

Run green on 2 x B200 in round 2 (profiles/r02/call_a/pytest_tp_fused.log)."""
import os
import socket
import sys

import numpy as np
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


def _worker(rank, world, port, gs, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dev = torch.device(f"cuda:{rank}")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import qqq_b200
    from oracle import qqq_oracle as O
    from qqq_b200 import tp
    from test_tp_gloo import _full_module

    K, N = 1024, 512
    ws = tp.AllReduceWorkspace(max_tokens=300, max_features=N, device=dev)
    res = []
    for it, M in enumerate((300, 7, 64, 300)):
        p = O.make_problem(M, K, N, gs, seed=30 + it)
        full = _full_module(p, K, N, gs)
        shard = tp.shard_quant_linear(full, rank, world, "row").to(dev)
        _, offs = tp.split_sizes(K, world, 128 if gs != -1 else 64)
        x_loc = torch.from_numpy(p["x"][:, offs[rank]:offs[rank + 1]].copy()).to(dev)
        # this rank's partial output through the plain kernel, summed over ranks in fp32: what the fused epilogue must
        # deliver up to the fp16 rounding of the adds (world = 2: a single rounding, order-independent)
        part = shard(x_loc).float()
        dist.all_reduce(part)
        y_nccl = tp.RowParallelQuantLinear(shard)(x_loc)
        l0 = qqq_b200.launch_count()
        y_fused = tp.FusedRowParallelQuantLinear(shard, ws)(x_loc)
        torch.cuda.synchronize()
        n_launch = qqq_b200.launch_count() - l0
        err_sum = float((y_fused.float() - part).abs().max())
        err_nccl = float((y_fused.float() - y_nccl.float()).abs().max())
        scale = float(part.abs().max())
        # exact mode: shared s1 + int32 partial sums -> the 1-GPU module's output bit for bit
        y_exact = tp.ExactRowParallelQuantLinear(shard)(x_loc)
        y_one = full.to(dev)(torch.from_numpy(p["x"]).to(dev))
        exact_ok = bool(torch.equal(y_exact.view(torch.int16), y_one.view(torch.int16)))
        res.append((M, err_sum, err_nccl, scale, n_launch, tuple(y_fused.shape), int(shard.workspace.abs().sum()), exact_ok))
    if rank == 0:
        out.put(res)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("gs", [-1, 128])
def test_fused_gemm_allreduce_matches_nccl_path(gs):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, gs, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    try:
        res = q.get(timeout=300)
        for pr in procs:
            pr.join(timeout=60)
            assert pr.exitcode == 0
    finally:
        for pr in procs:
            if pr.is_alive():
                pr.kill()
    for (M, err_sum, err_nccl, scale, n_launch, shape, ws_sum, exact_ok) in res:
        assert shape == (M, 512) and n_launch == 2 and ws_sum == 0
        assert exact_ok, f"exact row-parallel mode differs from the 1-GPU module at M={M}"
        # one fp16 rounding of a value of magnitude <= scale: half an ulp = scale * 2^-11
        assert err_sum <= scale * 2.0 ** -10 + 1e-6, (M, err_sum, scale)
        assert err_nccl <= scale * 2.0 ** -9 + 1e-6, (M, err_nccl, scale)
