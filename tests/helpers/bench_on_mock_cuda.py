"""Run bench.py's main() on a machine without a GPU (test infrastructure, used by tests/test_bench_hostlogic.py).

The two C-ABI entry points are replaced by the CPU oracle and the small CUDA surface bench.py touches (events, streams,
graphs, pinned memory) by inert stand-ins; the model is shrunk to 2 tiny layers.  What this exercises is the host logic of
bench.py: argument handling, the call chains, the guarded auxiliary sections and the JSON contract.  Timings printed by
such a run mean nothing."""
import sys, os, types, time, contextlib
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch, numpy as np
from oracle import qqq_oracle as O
from qqq_b200 import ops
import qqq_b200
cnt = {"c": 0}
def dq(x):
    cnt["c"] += 1
    q, s = O.dynamic_quant(x.detach().contiguous().numpy(), cuda_semantics=True); return torch.from_numpy(q), torch.from_numpy(s)
def gemm(A,B,C,D,s1,s2,s3,ws,*a,**k):
    cnt["c"] += 1
    D.copy_(torch.from_numpy(O.qqq_gemm_oracle(A.numpy(),B.numpy(),s1.numpy(),s2.numpy(),s3.numpy() if s3.numel() else None)))
def gemm_acc(A, B, C, D32, s3, workspace, max_par=16, sms=-1):
    cnt["c"] += 1
    W8 = O.weights_int8(B.numpy(), s3.numpy() if s3.numel() else None)
    D32.copy_(torch.from_numpy((A.numpy().astype(np.int64) @ W8.astype(np.int64)).astype(np.int32)))
ops.dynamic_quant = dq; ops.qqq_gemm = gemm; qqq_b200.qqq_gemm = gemm; ops.qqq_gemm_acc = gemm_acc
qqq_b200.launch_count = lambda: cnt["c"]
# --- fake CUDA surface ---
class Ev:
    def __init__(s, enable_timing=True): s.t = 0
    def record(s): s.t = time.perf_counter()
    def elapsed_time(s, o): return (o.t - s.t) * 1e3 + 1e-3
class G:
    def replay(s): s.fn and s.fn()
class Strm:
    def __init__(s, device=None, **k): pass
    def wait_stream(s, o): pass
torch.cuda.is_available = lambda: True
torch.cuda.set_device = lambda d: None
torch.cuda.synchronize = lambda *a: None
torch.cuda.Event = Ev
torch.cuda.Stream = Strm
torch.cuda.current_stream = lambda d=None: Strm()
torch.cuda.stream = lambda s: contextlib.nullcontext()
torch.cuda.get_device_capability = lambda *a: (10, 0)
torch.Tensor.pin_memory = lambda self: self
class FakeGraph:
    def __init__(s): s.fn=None
    def replay(s): s.fn()
class fakegraphctx:
    def __init__(s, g, pool=None): s.g=g
    def __enter__(s): return s
    def __exit__(s,*a): pass
torch.cuda.CUDAGraph = FakeGraph
torch.cuda.graph = fakegraphctx
# graph capture: run fn eagerly at capture and at every replay
import qqq_b200.graph as qg
class GC:
    def __init__(s, fn, x, warmup=2, pool=None): s.fn, s.x = fn, x.clone(); s.out = fn(s.x)
    def __call__(s, x=None):
        if x is not None: s.x.copy_(x)
        s.out = s.fn(s.x); return s.out
qg.GraphedCallable = GC
qg.capture = lambda fn, x, warmup=2: GC(fn, x)
qg.fork_join = lambda thunks: [t() for t in thunks]  # no streams on the mock surface
class PR:
    def __init__(s, fn, x, warmup=2): s.fn, s.x, s.i = fn, x.clone(), 0
    def step(s, x_host, out_host): s.x.copy_(x_host); out_host.copy_(s.fn(s.x)); s.i += 1
    def drain(s): pass
qg.PipelinedRunner = PR
src = open(os.path.join(ROOT, "bench.py")).read()
src = src.replace('dev = f"cuda:{local_rank}"', 'dev = "cpu"')
src = src.replace('dist.init_process_group("nccl", device_id=torch.device(dev))', 'dist.init_process_group("gloo")')
src = src.replace('ROOT = os.path.dirname(os.path.abspath(__file__))', 'ROOT = %r' % ROOT)
mod = types.ModuleType("bench_emu"); mod.__file__ = os.path.join(ROOT, "bench.py")
exec(compile(src, "bench.py", "exec"), mod.__dict__)

mod.MODEL.update(layers=2, hidden=256, inter=512, kv=256, seq=16, batch=1)
mod.FULL_MODEL.update(vocab_size=128, hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, max_position_embeddings=64, seq=12)
mod.LLAMA3_8B.update(layers=2, hidden=256, inter=512, kv=128, batch=4)
mod.LLAMA2_70B.update(layers=2, hidden=256, inter=512, kv=128, seq=16, batch=1)
_orig_sweep_tp = mod.gemm_sweep_tp
mod.gemm_sweep_tp = lambda dev, peaks, rank, world, ws, mor, K=8192, N=21760, Ms=None: _orig_sweep_tp(dev, peaks, rank, world, ws, mor, 256, 256, (4, 16))
mod.TP_PARITY_CASES = (("pc_small", 16, 256, 256, -1), ("g128_small", 6, 256, 128, 128))
def _gtu(launch_all, n_launches, reps=3, warm=1):
    launch_all(); return 1.0
mod.graph_time_us = _gtu
_orig_sweep = mod.gemm_sweep
mod.gemm_sweep = lambda dev, peaks, quick=False, K=8192, N=21760, Ms=None, ref_kernel=None: _orig_sweep(dev, peaks, quick, 256 if K == 8192 else 512, 256, (1, 16))
mod.load_reference_kernel = lambda: None
if "--gpus" in sys.argv and "--fused-allreduce" not in sys.argv and "nccl" not in sys.argv:
    # default N > 1 mode (scatter): symmetric memory and the two exchange kernels emulated over gloo
    import torch.distributed as dist
    from qqq_b200 import tp

    class _EmuBytes:
        def __init__(self, group=None):
            self.rank, self.world, self.group = dist.get_rank(), dist.get_world_size(), None

        def alloc(self, nbytes, device):
            self.buf = torch.zeros(nbytes, dtype=torch.uint8)
            _EmuBytes.last = self
            return self.buf, [(r + 1) << 40 for r in range(self.world)], 0

    def _find(addr_list, rank):  # which emulated buffer and byte offset an "address" of this rank means
        return addr_list[rank] - ((rank + 1) << 40)

    def _gemm_scatter(A, B, C, peer_partials, s1, s2, s3, workspace, prob_n, tp_rank, tp_world, tp_rows, max_par=16, sms=-1):
        cnt["c"] += 1
        be = _EmuBytes.last
        off = _find(peer_partials, tp_rank)
        part = torch.from_numpy(O.qqq_gemm_oracle(A.numpy(), B.numpy(), s1.numpy(), s2.numpy(), s3.numpy() if s3.numel() else None))
        padded = torch.zeros(tp_rows * tp_world, prob_n, dtype=torch.float16)
        padded[: part.shape[0]] = part
        allp = [torch.empty_like(padded) for _ in range(tp_world)]
        dist.all_gather(allp, padded)
        slots = be.buf[off:off + 2 * tp_world * tp_rows * prob_n].view(torch.float16).view(tp_world, tp_rows, prob_n)
        for s_ in range(tp_world):
            slots[s_] = allp[s_][tp_rank * tp_rows:(tp_rank + 1) * tp_rows]

    def _reduce_quant(partials_ptr, a8_dst, a8_mc, s1_dst, s1_mc, h_out, bias, flags_ptr, peer_flags, tp_rank, tp_world,
                      tp_rows, prob_m, prob_n, dev_):
        cnt["c"] += 1
        be = _EmuBytes.last
        off = partials_ptr - ((tp_rank + 1) << 40)
        slots = be.buf[off:off + 2 * tp_world * tp_rows * prob_n].view(torch.float16).view(tp_world, tp_rows, prob_n)
        acc = torch.zeros(tp_rows, prob_n)
        for s_ in range(tp_world):
            acc = acc + slots[s_].float()
        h = acc.half()
        if bias is not None:
            h = h + bias
        if h_out is not None:
            n = h_out.shape[0]
            h_out.copy_(h[:n])
        q, s1 = dq(h); cnt["c"] -= 1
        allq = [torch.empty_like(q) for _ in range(tp_world)]
        alls = [torch.empty_like(s1) for _ in range(tp_world)]
        dist.all_gather(allq, q); dist.all_gather(alls, s1)
        mpad = tp_rows * tp_world
        oa, os1 = _find(a8_dst, tp_rank), _find(s1_dst, tp_rank)
        be.buf[oa:oa + mpad * prob_n].view(torch.int8).view(mpad, prob_n).copy_(torch.cat(allq))
        be.buf[os1:os1 + 4 * mpad].view(torch.float32).view(mpad, 1).copy_(torch.cat(alls))

    tp._SymmBytesBackend = _EmuBytes
    ops.qqq_gemm_scatter = _gemm_scatter
    ops.tp_reduce_quant = _reduce_quant
    tp.ScatterWorkspace.timeouts = lambda self: 0
if "--fused-allreduce" in sys.argv:
    # NVLS multicast emulated over gloo: every rank's partial output is added into every rank's replica
    import torch.distributed as dist
    from qqq_b200 import tp

    class _Emu:
        by_addr = {}

        def __init__(self, group=None):
            pass

        def alloc(self, numel, device):
            t = torch.full((numel,), 555.0, dtype=torch.float16)
            addr = 4096 * (len(_Emu.by_addr) + 1)
            _Emu.by_addr[addr] = t
            return t, addr, (lambda: dist.barrier())

    def _gemm_reduce(A, B, C, mc, s1, s2, s3, workspace, prob_n, max_par=16, sms=-1):
        cnt["c"] += 1
        part = torch.from_numpy(O.qqq_gemm_oracle(A.numpy(), B.numpy(), s1.numpy(), s2.numpy(),
                                                  s3.numpy() if s3.numel() else None)).float()
        dist.all_reduce(part)
        view = _Emu.by_addr[mc][: part.numel()].view(part.shape)
        view += part.half()

    tp._SymmMemBackend = _Emu
    ops.qqq_gemm_reduce = _gemm_reduce
if "--break-sweep" in sys.argv:  # an auxiliary section that raises must be reported, not propagated
    sys.argv.remove("--break-sweep")
    mod.gemm_sweep = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("injected failure"))
sys.argv = ["bench.py", "--steps", "1", "--warmup", "1"] + sys.argv[1:]
rc = mod.main()
print("rc", rc)
