"""CUDA-graph capture of a chain of QuantLinear calls.

At decode (and for the 4096-wide linears of a 7B model even at prefill) one W4A8 GEMM is a 10-30 us kernel, so
the Python/ctypes launch path would bound throughput.  `capture()` records the launches once — the C ABI launches on
torch's current stream, so it is capturable as is; tensor maps are baked into the kernel parameters — and replays
them with one `cudaGraphLaunch`.
"""
from __future__ import annotations

import torch

from .qlinear import _ActQuantCache


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
