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


class PipelinedRunner:
    """Steps whose input arrives in pinned host memory and whose result goes back to host memory, software-pipelined over two
    captured graphs: the host-to-device copy of step i+1 and the device-to-host copy of step i-1 run on their own streams
    under the kernels of step i (one copy engine per direction), so that a serving loop pays the PCIe time of a step only
    where it exceeds the step's compute.  Every step still copies its own input in and its own result out."""

    def __init__(self, fn, example_input: torch.Tensor, warmup: int = 2):
        dev = example_input.device
        self.g = [GraphedCallable(fn, example_input, warmup), GraphedCallable(fn, example_input, warmup)]
        self.s_in, self.s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        mk = lambda: [torch.cuda.Event(), torch.cuda.Event()]  # noqa: E731
        self.ev_in, self.ev_done, self.ev_out = mk(), mk(), mk()
        self.i = 0
        self.dev = dev

    def step(self, x_host: torch.Tensor, out_host: torch.Tensor) -> None:
        """Enqueue one step: x_host (pinned) -> device -> graph -> out_host (pinned).  Asynchronous; `drain()` joins."""
        b = self.i & 1
        g = self.g[b]
        cur = torch.cuda.current_stream(self.dev)
        self.s_in.wait_event(self.ev_done[b])  # the step that last read this input buffer has finished
        with torch.cuda.stream(self.s_in):
            g.static_in.copy_(x_host, non_blocking=True)
            self.ev_in[b].record(self.s_in)
        cur.wait_event(self.ev_in[b])
        cur.wait_event(self.ev_out[b])  # this graph's previous result has left its static output
        g.graph.replay()
        self.ev_done[b].record(cur)
        self.s_out.wait_event(self.ev_done[b])
        with torch.cuda.stream(self.s_out):
            out_host.copy_(g.static_out, non_blocking=True)
            self.ev_out[b].record(self.s_out)
        self.i += 1

    def drain(self) -> None:
        """Make the current stream wait for every copy issued so far (call before timing / before reading results)."""
        cur = torch.cuda.current_stream(self.dev)
        for e in self.ev_out + self.ev_in:
            cur.wait_event(e)


def capture(fn, example_input: torch.Tensor, warmup: int = 2) -> GraphedCallable:
    """Capture `fn(example_input)` (any composition of QuantLinear / qqq_gemm / dynamic_quant calls)."""
    return GraphedCallable(fn, example_input, warmup)
