"""CUDA-graph capture of a chain of QuantLinear calls.

At decode (and for the 4096-wide linears of a 7B model even at prefill) one W4A8 GEMM is a 10-30 us kernel, so
the Python/ctypes launch path would bound throughput.  `capture()` records the launches once — the C ABI launches on
torch's current stream, so it is capturable as is; tensor maps are baked into the kernel parameters — and replays
them with one `cudaGraphLaunch`.
"""
from __future__ import annotations

import torch

from .qlinear import _ActQuantCache


_side_streams = {}


def fork_join(thunks):
    """Run independent launches (linears that consume the same activation: q/k/v, gate/up) on forked streams and join them
    again: `[t() for t in thunks]`, but thunks[1:] are issued on side streams that wait for everything issued so far, and the
    current stream waits for them before it goes on.  Captured into a CUDA graph this becomes parallel branches: the CTAs
    of the second GEMM move onto SMs as the first one's CTAs retire instead of waiting for its last drain and completion,
    and GEMMs smaller than the machine (tensor-parallel shards, decode) run side by side.  Each thunk must use its own
    scratch (every QuantLinear owns its reduce buffer and lock words; `model.share_scratch` gives that up)."""
    thunks = list(thunks)
    if len(thunks) <= 1 or not torch.cuda.is_available():
        return [t() for t in thunks]
    dev = torch.cuda.current_device()
    cur = torch.cuda.current_stream(dev)
    pool = _side_streams.setdefault(dev, [])
    while len(pool) < len(thunks) - 1:
        pool.append(torch.cuda.Stream(device=dev))
    fork = torch.cuda.Event()
    fork.record(cur)
    outs = [None] * len(thunks)
    joins = []
    for i in range(1, len(thunks)):
        s = pool[i - 1]
        s.wait_event(fork)
        with torch.cuda.stream(s):
            outs[i] = thunks[i]()
            e = torch.cuda.Event()
            e.record(s)
            joins.append(e)
    outs[0] = thunks[0]()
    for e in joins:
        cur.wait_event(e)
    return outs


class GraphedCallable:
    def __init__(self, fn, example_input: torch.Tensor, warmup: int = 2, pool=None):
        self.static_in = example_input.clone()
        side = torch.cuda.Stream(device=example_input.device)
        side.wait_stream(torch.cuda.current_stream(example_input.device))
        with torch.cuda.stream(side):
            for _ in range(warmup):  # first calls set function attributes / fill the tensor-map cache
                fn(self.static_in)
        torch.cuda.current_stream(example_input.device).wait_stream(side)
        torch.cuda.synchronize(example_input.device)
        # an activation-quant cache entry from the warm-up (or from eager code) must not be "hit" during capture: the
        # graph would then read int8 activations it never produces
        _ActQuantCache.clear()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, pool=pool):
            self.static_out = fn(self.static_in)
        _ActQuantCache.clear()

    def __call__(self, x: torch.Tensor = None) -> torch.Tensor:
        if x is not None and x.data_ptr() != self.static_in.data_ptr():
            self.static_in.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.static_out


def capture(fn, example_input: torch.Tensor, warmup: int = 2) -> GraphedCallable:
    """Capture `fn(example_input)` (any composition of QuantLinear / qqq_gemm / dynamic_quant calls)."""
    return GraphedCallable(fn, example_input, warmup)
