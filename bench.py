#!/usr/bin/env python
"""bench.py — the W4A8 hot path on BASELINE.json's configs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload of a "step" (config.workload): every quantized Linear of Llama-2-7B (32 layers x {q,k,v,o,gate,up,down},
per-channel W4A8), prefill seq=1024 batch=1  ->  BASELINE.json configs[1].  A step = for each of the 224 linears:
fused per-token activation quant + tcgen05 W4A8 GEMM (through QuantLinear.forward -> C ABI).  Everything between the
linears (attention, norms, SwiGLU) is outside the reference's hot path and is not executed; shapes chain as
h -> q,k,v ; o(q) -> h' ; gate,up(h') ; down(gate) -> h.  tokens/s = 1024 / step time.
The same JSON line carries the GEMM sweep of configs[4] (TFLOP/s and speed-up over fp16 cuBLAS at
M in {1,16,128,1024,4096}, K=8192, N=21760, per-channel and g=128) under "gemm_sweep".

N > 1: tensor parallel (strong scaling, same model): q,k,v,gate,up split N (no collective), o/down split K + one
NCCL all-reduce of the fp16 [1024,4096] output per layer.

--impl reference: the dequant-to-fp16 torch.matmul CPU path (the oracle port; the reference has no CPU implementation
of its own) on this box's host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

MODEL = dict(name="llama-2-7b", layers=32, hidden=4096, inter=11008, seq=1024, batch=1)
# (name, K, N, parallel mode)
LAYER_LINEARS = [("q", "hidden", "hidden", "column"), ("k", "hidden", "hidden", "column"),
                 ("v", "hidden", "hidden", "column"), ("o", "hidden", "hidden", "row"),
                 ("gate", "hidden", "inter", "column"), ("up", "hidden", "inter", "column"),
                 ("down", "inter", "hidden", "row")]


def env_int(name, default):
    return int(os.environ.get(name, default))


def model_flops(M):
    f = 0
    for (_, k, n, _) in LAYER_LINEARS:
        f += 2.0 * M * MODEL[k] * MODEL[n]
    return f * MODEL["layers"]


def model_gemm_bytes(M):
    """Algorithmic bytes of the 224 GEMMs: M*K + K*N/2 + 2*M*N + 4*M + 4*N each."""
    b = 0
    for (_, k, n, _) in LAYER_LINEARS:
        K, N = MODEL[k], MODEL[n]
        b += M * K + K * N / 2 + 2 * M * N + 4 * M + 4 * N
    return b * MODEL["layers"]


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"],
                    bf16_tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


# ----------------------------------------------------------------------------------------------------------
# clocks sampling (NVML) during the timed region
# ----------------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._th = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        if self.nv is not None:
            self._th = threading.Thread(target=self._run, daemon=True)
            self._th.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._th is not None:
            self._th.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ----------------------------------------------------------------------------------------------------------
# synthetic model
# ----------------------------------------------------------------------------------------------------------
def random_packed(K, N, gen, dev):
    # uniform nibbles; int32 words drawn directly (packing 200 M weights through pack() is test-only work)
    return torch.randint(-2**31, 2**31 - 1, (K // 16, 2 * N), dtype=torch.int32, device=dev, generator=gen)


def build_model(dev, rank, world, gen):
    """Random-init Llama-2-7B quantized linears (per-channel), already sharded for `world` ranks."""
    import qqq_b200
    from qqq_b200 import tp

    layers = []
    for _ in range(MODEL["layers"]):
        mods = {}
        for (name, k, n, mode) in LAYER_LINEARS:
            K, N = MODEL[k], MODEL[n]
            if world > 1:
                if mode == "column":
                    sizes, _ = tp.split_sizes(N, world, 64)
                    N = sizes[rank]
                else:
                    sizes, _ = tp.split_sizes(K, world, 64)
                    K = sizes[rank]
            ql = qqq_b200.QuantLinear(4, -1, K, N, bias=False)
            ql = ql.to(dev)
            ql.B = random_packed(K, N, gen, dev)
            # W8 = 16*w4 with w4 ~ U[-8,7] (std 4.6): unit-variance outputs for unit-variance inputs
            ql.s_channel = torch.full((1, N), 1.0 / (16 * 4.6 * (MODEL[k] ** 0.5)), dtype=torch.float32, device=dev)
            mods[name] = ql
        layers.append(mods)
    return layers


def _all_reduce_after(mod, y, world):
    """NCCL all-reduce after a row-parallel linear — unless the module reduces in its own epilogue (--fused-allreduce)."""
    import torch.distributed as dist
    from qqq_b200 import tp

    if world > 1 and not isinstance(mod, tp.FusedRowParallelQuantLinear):
        dist.all_reduce(y)


def forward_chain(layers, h, world):
    """The reference's module structure: 7 separate linears per layer (q, k, v re-quantise the same input)."""
    for m in layers:
        q = m["q"](h)
        m["k"](h)
        m["v"](h)
        o = m["o"](q)
        _all_reduce_after(m["o"], o, world)
        g = m["gate"](o)
        m["up"](o)
        d = m["down"](g)
        _all_reduce_after(m["down"], d, world)
        h = d
    return h


def merge_layers(layers):
    """q/k/v and gate/up merged by column concatenation of the packed tensors (qqq_b200.merge_quant_linears):
    bit-identical outputs, one activation quant + one GEMM per group of linears that share their input."""
    import qqq_b200

    merged = []
    for m in layers:
        merged.append(dict(qkv=qqq_b200.merge_quant_linears([m["q"], m["k"], m["v"]]), o=m["o"],
                           gate_up=qqq_b200.merge_quant_linears([m["gate"], m["up"]]), down=m["down"]))
    return merged


def forward_chain_merged(mlayers, h, world):
    for m in mlayers:
        qkv = m["qkv"](h)
        q = qkv[:, : m["qkv"].split_sizes[0]]  # column slice, consumed in place by the strided activation quant
        o = m["o"](q)
        _all_reduce_after(m["o"], o, world)
        gu = m["gate_up"](o)
        g = gu[:, : m["gate_up"].split_sizes[0]]
        d = m["down"](g)
        _all_reduce_after(m["down"], d, world)
        h = d
    return h


GEMM_NAMES = ("q", "k", "v", "o", "gate", "up", "down")


def gemm_only_chain(mlayers, qin):
    """The GEMM launches of a step alone, on pre-quantised inputs (for the roofline figure of the dominant kernel)."""
    import qqq_b200

    for m in mlayers:
        for name in GEMM_NAMES:
            ql = getattr(m[name], "shard", m[name])
            A8, s1, D = qin[(ql.infeatures, ql.outfeatures)]
            qqq_b200.qqq_gemm(A8, ql.B, ql.reduce_buffer, D, s1, ql.s_channel, ql.s_group, ql.workspace, -1, -1, -1, 16)


def timed(fn, steps, warmup, barrier):
    for _ in range(warmup):
        fn()
    barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    barrier()
    return e0.elapsed_time(e1) / steps  # ms per step


# ----------------------------------------------------------------------------------------------------------
# GEMM sweep (configs[4])
# ----------------------------------------------------------------------------------------------------------
def graph_time_us(launch_all, n_launches, reps=3, warm=1):
    """Time `launch_all()` (n_launches kernels back to back) replayed from a CUDA graph: device time per launch,
    free of Python/ctypes launch cost."""
    launch_all()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        launch_all()
    for _ in range(warm):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * n_launches)


def gemm_sweep(dev, peaks, quick=False, K=8192, N=21760, Ms=(1, 16, 128, 1024, 4096)):
    """BASELINE configs[4]: M in {1,16,128,1024,4096}, K=8192, N=21760, per-channel and g128, vs fp16 cuBLAS."""
    import qqq_b200

    ncopy = 4  # 4 x 89 MB packed weights > 126 MB L2: every launch streams its weights from HBM
    g = torch.Generator(device=dev).manual_seed(0)
    Bs = [random_packed(K, N, g, dev) for _ in range(ncopy)]
    Wh = [torch.randn(K, N, dtype=torch.float16, device=dev) * 0.02 for _ in range(ncopy)]
    s2 = torch.rand(1, N, device=dev) * 1e-3 + 5e-4
    s3g = (torch.rand(K // 128, N, device=dev) * 8 + 4).half()
    s3e = torch.zeros(0, dtype=torch.float16, device=dev)
    C = torch.zeros(16 * 64, N, dtype=torch.int32, device=dev)
    ws = torch.zeros(N // 128 * 16, dtype=torch.int32, device=dev)
    int8_peak = 2.0 * peaks["bf16_tflops"]  # burst figure: each GEMM is timed on its own
    out = []
    for M in (Ms if not quick else (16, 1024)):
        A = torch.randint(-127, 128, (M, K), dtype=torch.int8, device=dev)
        Ah = torch.randn(M, K, dtype=torch.float16, device=dev)
        s1 = torch.rand(M, 1, device=dev) * 1e-2 + 1e-3
        D = torch.empty(M, N, dtype=torch.float16, device=dev)
        Dh = torch.empty(M, N, dtype=torch.float16, device=dev)

        def run_h():
            for i in range(ncopy):
                torch.matmul(Ah, Wh[i], out=Dh)

        reps = 5 if M <= 1024 else 2
        t_h = graph_time_us(run_h, ncopy, reps)
        fl = 2.0 * M * K * N
        for mode, s3 in (("per-channel", s3e), ("g128", s3g)):
            def run_q():
                for i in range(ncopy):
                    qqq_b200.qqq_gemm(A, Bs[i], C, D, s1, s2, s3, ws, -1, -1, -1, 16)

            t = graph_time_us(run_q, ncopy, reps)
            by = M * K + K * N / 2 + 2 * M * N + 4 * M + 4 * N + (2 * (K // 128) * N if mode == "g128" else 0)
            hbm_frac = by / (t * 1e-6) / 1e9 / peaks["hbm_gbs"]
            tc_frac = fl / (t * 1e-6) / 1e12 / int8_peak
            out.append(dict(M=M, K=K, N=N, mode=mode, us=round(t, 2), tflops=round(fl / t * 1e-6, 1),
                            gbps=round(by / t * 1e-3, 1), fp16_cublas_us=round(t_h, 2),
                            speedup_vs_fp16=round(t_h / t, 3), bound="hbm" if hbm_frac > tc_frac else "tensor",
                            roofline_frac=round(max(hbm_frac, tc_frac), 3)))
    return out


# ----------------------------------------------------------------------------------------------------------
# CPU path (oracle port): dequant-to-fp16 torch.matmul on host cores
# ----------------------------------------------------------------------------------------------------------
def cpu_layer_setup(seed=0):
    """One decoder layer's 7 linears: packed weights -> fp16 (untimed, one-off), int8 activations + scales."""
    from oracle import qqq_oracle as O

    rng = np.random.default_rng(seed)
    M = MODEL["seq"] * MODEL["batch"]
    items = []
    for (name, k, n, _) in LAYER_LINEARS:
        K, N = MODEL[k], MODEL[n]
        B = rng.integers(-2**31, 2**31 - 1, size=(K // 16, 2 * N), dtype=np.int64).astype(np.int32)
        s2 = np.full((1, N), 1.0 / (16 * 4.6 * K ** 0.5), np.float32)
        W = torch.from_numpy(O.dequant_weights_fp16(B, s2, None))
        A8 = rng.integers(-127, 128, size=(M, K), dtype=np.int64).astype(np.int8)
        s1 = np.full((M, 1), 0.03, np.float32)
        items.append((name, A8, s1, W))
    return items


def cpu_layer_run(items):
    from oracle import qqq_oracle as O

    for (_, A8, s1, W) in items:
        O.dequant_matmul_cpu(A8, s1, W.numpy())


def cpu_baseline(steps=2, warmup=1):
    torch.set_num_threads(os.cpu_count())
    items = cpu_layer_setup()
    for _ in range(warmup):
        cpu_layer_run(items)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_layer_run(items)
    t = (time.perf_counter() - t0) / steps
    tok = MODEL["seq"] * MODEL["batch"] / (t * MODEL["layers"])
    return dict(value=round(tok, 3), unit="tokens/s", cores=os.cpu_count(), kind="port",
                sample=f"1 of {MODEL['layers']} decoder layers (7 linears, M=1024) dequant-to-fp16 torch.matmul, "
                       f"weights pre-dequantised; {steps} timed passes, extrapolated x{MODEL['layers']}"), t


# ----------------------------------------------------------------------------------------------------------
# configs[2]: Llama-3-8B per-group g=128, decode batch=32 seq=1 (reported beside the headline, not the headline)
# ----------------------------------------------------------------------------------------------------------
LLAMA3_8B = dict(name="llama-3-8b", layers=32, hidden=4096, inter=14336, kv=1024, batch=32)
LLAMA3_LINEARS = [("q", "hidden", "hidden"), ("k", "hidden", "kv"), ("v", "hidden", "kv"), ("o", "hidden", "hidden"),
                  ("gate", "hidden", "inter"), ("up", "hidden", "inter"), ("down", "inter", "hidden")]


def decode_g128(dev, peaks, steps, warmup):
    """All 224 quantized linears of Llama-3-8B (g128) on a batch of 32 single-token rows, replayed from one CUDA
    graph: per layer 7 x (act-quant + GEMM), and the merged q/k/v + gate/up variant.  HBM-bound: the figure of merit is
    algorithmic bytes / time against the measured HBM peak."""
    import qqq_b200
    from qqq_b200 import graph as qgraph

    cfg, M = LLAMA3_8B, LLAMA3_8B["batch"]
    gen = torch.Generator(device=dev).manual_seed(4321)
    layers, by = [], 0.0
    for _ in range(cfg["layers"]):
        mods = {}
        for (name, k, n) in LLAMA3_LINEARS:
            K, N = cfg[k], cfg[n]
            ql = qqq_b200.QuantLinear(4, 128, K, N, bias=False).to(dev)
            ql.B = random_packed(K, N, gen, dev)
            ql.s_group = (torch.rand(K // 128, N, device=dev, generator=gen) * 8 + 4).half()
            ql.s_channel = torch.full((1, N), 1.0 / (8 * 4.6 * (K ** 0.5)), dtype=torch.float32, device=dev)
            mods[name] = ql
            by += M * K + K * N / 2 + 2 * M * N + 4 * M + 4 * N + 2 * (K // 128) * N
        layers.append(mods)

    def chain(h):
        for m in layers:
            q = m["q"](h)
            m["k"](h)
            m["v"](h)
            o = m["o"](q)
            g = m["gate"](o)
            m["up"](o)
            h = m["down"](g)
        return h

    x = torch.randn(M, cfg["hidden"], device=dev, generator=gen).half()
    none = lambda: None  # noqa: E731
    g = qgraph.capture(chain, x)
    ms = timed(lambda: g(x), steps, warmup, none)
    out = dict(workload="llama-3-8b decode batch=32 seq=1: all 224 quantized linears, per-group g=128, 7 x (act-quant + GEMM) "
                        "per layer; random-init weights; 3.6 GB of packed weights + group scales stream from HBM every step",
               ms_per_step=round(ms, 4), value=round(M / (ms * 1e-3), 1), unit="tokens/s",
               gbps=round(by / (ms * 1e-3) / 1e9, 1), hbm_frac=round(by / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"], 3),
               algorithmic_bytes_per_step=by)
    del g
    mlayers = [dict(qkv=qqq_b200.merge_quant_linears([m["q"], m["k"], m["v"]]), o=m["o"],
                    gate_up=qqq_b200.merge_quant_linears([m["gate"], m["up"]]), down=m["down"]) for m in layers]

    def chain_merged(h):
        for m in mlayers:
            qkv = m["qkv"](h)
            o = m["o"](qkv[:, : m["qkv"].split_sizes[0]])
            gu = m["gate_up"](o)
            h = m["down"](gu[:, : m["gate_up"].split_sizes[0]])
        return h

    gm = qgraph.capture(chain_merged, x)
    ms_m = timed(lambda: gm(x), steps, warmup, none)
    out["merged"] = dict(ms_per_step=round(ms_m, 4), value=round(M / (ms_m * 1e-3), 1))
    return out


# ----------------------------------------------------------------------------------------------------------
# configs[1] as a whole model: stock HF LlamaForCausalLM with its decoder linears swapped for QuantLinear
# ----------------------------------------------------------------------------------------------------------
FULL_MODEL = dict(vocab_size=32000, hidden_size=4096, intermediate_size=11008, num_hidden_layers=32,
                  num_attention_heads=32, num_key_value_heads=32, max_position_embeddings=4096, seq=1024)


def full_forward(dev, steps, warmup):
    """Llama-2-7B prefill (seq 1024, batch 1) through `transformers.LlamaForCausalLM.forward`: embeddings, RMSNorm,
    rotary, SDPA attention, SwiGLU and lm_head are stock PyTorch (not on the reference's native hot path either), the 224
    decoder linears are QuantLinear (random packed weights).  Eager launches, no CUDA graph: what a user of the
    reference's `model(input_ids)` gets.  Reported beside the linears-only headline, with and without q/k/v + gate/up
    fusion (qqq_b200.model.fuse_qkv_gate_up)."""
    import transformers

    import qqq_b200
    from qqq_b200 import model as qm

    kw = {k: v for k, v in FULL_MODEL.items() if k != "seq"}
    cfg = transformers.LlamaConfig(tie_word_embeddings=False, **kw)
    m = qm.build_quantized_model(cfg, qm.quantization_config(-1), dtype=torch.float16, device=torch.device(dev)).eval()
    gen = torch.Generator(device=dev).manual_seed(99)
    for ql in qm.find_layers(m, [qqq_b200.QuantLinear]).values():
        ql.B = random_packed(ql.infeatures, ql.outfeatures, gen, dev)
        ql.s_channel = torch.full((1, ql.outfeatures), 1.0 / (16 * 4.6 * ql.infeatures ** 0.5), dtype=torch.float32, device=dev)
    qm.share_scratch(m)
    S = FULL_MODEL["seq"]
    ids = torch.randint(0, FULL_MODEL["vocab_size"], (1, S), device=dev, generator=gen)
    none = lambda: None  # noqa: E731

    def step():
        with torch.no_grad():
            return m(input_ids=ids, use_cache=False).logits

    l0 = qqq_b200.launch_count()
    out = step()
    n_launch = qqq_b200.launch_count() - l0
    finite = bool(torch.isfinite(out.float()).all().item())
    ms = timed(step, steps, warmup, none)
    res = dict(workload="transformers LlamaForCausalLM (llama-2-7b shape, random init) prefill seq=1024 batch=1, decoder "
                        "linears = QuantLinear per-channel W4A8, everything else stock PyTorch, eager",
               ms_per_step=round(ms, 4), value=round(S / (ms * 1e-3), 1), unit="tokens/s",
               qqq_launches_per_step=int(n_launch), logits_finite=finite)
    qm.fuse_qkv_gate_up(m)
    ms_f = timed(step, steps, warmup, none)
    res["fused_qkv_gate_up"] = dict(ms_per_step=round(ms_f, 4), value=round(S / (ms_f * 1e-3), 1))
    return res


# ----------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-merged", action="store_true", help="skip the merged QKV / gate-up variant (profiler runs)")
    ap.add_argument("--fused-allreduce", action="store_true",
                    help="N > 1: row-parallel linears reduce in their own epilogue (tp.FusedRowParallelQuantLinear)")
    ap.add_argument("--no-decode", action="store_true", help="skip the Llama-3-8B g128 decode section (configs[2])")
    ap.add_argument("--no-full", action="store_true", help="skip the whole-model (HF Llama forward) section")
    ap.add_argument("--aux-budget", type=float, default=300.0, help="seconds the auxiliary sections may take in total")
    args = ap.parse_args()
    rank, world = env_int("RANK", 0), env_int("WORLD_SIZE", 1)
    local_rank = env_int("LOCAL_RANK", 0)
    M = MODEL["seq"] * MODEL["batch"]
    cfg = dict(workload="llama-2-7b prefill seq=1024 batch=1: all 224 quantized linears (per-channel W4A8), called module "
                        "by module as the reference model does: per layer 7 x (per-token act-quant + W4A8 GEMM); "
                        "random-init weights",
               global_batch=MODEL["batch"], seq_len=MODEL["seq"], parallelism=f"tp{world}" if world > 1 else "single",
               l2="inputs larger than L2: 3.2 GB of packed weights stream from HBM every step")

    if args.impl == "reference":
        if rank != 0:
            return 0
        base, t_layer = cpu_baseline(steps=max(1, args.steps), warmup=max(1, min(args.warmup, 2)))
        ms = t_layer * MODEL["layers"] * 1e3
        line = dict(impl="reference", metric="llama2_7b_w4a8_prefill_linears_tokens_per_s", value=base["value"],
                    unit="tokens/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=round(ms, 2),
                    higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f16", data="synthetic",
                    config=cfg, cpu_baseline=base,
                    e2e=dict(value=base["value"], unit="tokens/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        print(json.dumps(line))
        return 0

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU path in qqq_b200)"
    import qqq_b200
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO", "TRACE") and "NCCL_DEBUG_FILE" not in os.environ:
            os.environ["NCCL_DEBUG"] = "WARN"  # NCCL would print its version banner to stdout ahead of the JSON line
        dist.init_process_group("nccl", device_id=torch.device(dev))
        barrier = lambda: dist.barrier()  # noqa: E731
    else:
        barrier = lambda: None  # noqa: E731
    peaks = load_peaks()
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    layers = build_model(dev, rank, world, gen)
    if args.fused_allreduce and world > 1:
        # o_proj / down_proj add their output tiles into a multicast buffer from the GEMM epilogue (multimem.red over
        # NVSwitch) instead of being followed by an NCCL all-reduce
        from qqq_b200 import tp

        ar_ws = tp.AllReduceWorkspace(M, MODEL["hidden"], device=torch.device(dev))
        for m in layers:
            m["o"] = tp.FusedRowParallelQuantLinear(m["o"], ar_ws)
            m["down"] = tp.FusedRowParallelQuantLinear(m["down"], ar_ws)
        cfg["parallelism"] += " (all-reduce fused into the GEMM epilogue, multimem.red)"
    x_host = (torch.randn(M, MODEL["hidden"], generator=torch.Generator().manual_seed(7))).half().pin_memory()
    out_host = torch.empty(M, MODEL["hidden"], dtype=torch.float16).pin_memory()
    x_dev = x_host.to(dev)

    # --- device-resident throughput (value): a step's launches are replayed from one CUDA graph ---
    from qqq_b200 import graph as qgraph

    def max_over_ranks(ms_local):
        t = torch.tensor([ms_local], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # headline: the 224 linears called module by module like the reference model (7 activation quants + 7 GEMMs per layer)
    l0 = qqq_b200.launch_count()
    forward_chain(layers, x_dev, world)
    launches_per_step = qqq_b200.launch_count() - l0
    graphed = qgraph.capture(lambda x: forward_chain(layers, x, world), x_dev)
    with ClockSampler(local_rank) as cs:
        ms = max_over_ranks(timed(lambda: graphed(x_dev), args.steps, args.warmup, barrier))
    value = M / (ms * 1e-3)
    # the same weights with q/k/v and gate/up merged: 4 activation quants + 4 GEMMs per layer
    ms_mrg = None
    if not args.no_merged:
        mlayers = merge_layers(layers)
        graphed_mrg = qgraph.capture(lambda x: forward_chain_merged(mlayers, x, world), x_dev)
        ms_mrg = max_over_ranks(timed(lambda: graphed_mrg(x_dev), args.steps, args.warmup, barrier))
        del graphed_mrg, mlayers
    # --- end to end: pinned host input -> H2D -> 224 linears -> D2H of the result, every step ---
    def e2e_step():
        h = graphed(x_host)  # pinned host -> static device input (H2D), graph replay
        out_host.copy_(h, non_blocking=True)  # D2H of the step's result

    # the same step through the eager public API (no graph), for the record
    def eager_step():
        x_dev.copy_(x_host, non_blocking=True)
        out_host.copy_(forward_chain(layers, x_dev, world), non_blocking=True)

    ms_eager = timed(eager_step, args.steps, args.warmup, barrier)

    ms_e2e = timed(e2e_step, args.steps, args.warmup, barrier)
    t = torch.tensor([ms_e2e], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e2e = float(t.item())

    # --- dominant kernel alone: the 224 GEMM launches on pre-quantised inputs ---
    qin = {}
    for name in GEMM_NAMES:
        ql = getattr(layers[0][name], "shard", layers[0][name])
        key = (ql.infeatures, ql.outfeatures)
        if key not in qin:
            A8 = torch.randint(-127, 128, (M, key[0]), dtype=torch.int8, device=dev)
            qin[key] = (A8, torch.full((M, 1), 0.03, device=dev), torch.empty(M, key[1], dtype=torch.float16, device=dev))
    def _gemm_chain(x):
        gemm_only_chain(layers, qin)
        return x

    graphed_gemm = qgraph.capture(_gemm_chain, x_dev)
    ms_gemm = timed(lambda: graphed_gemm(x_dev), args.steps, args.warmup, barrier)
    n_gemm = MODEL["layers"] * len(GEMM_NAMES)
    flops_rank = model_flops(M) / world
    int8_peak = 2.0 * peaks["bf16_tflops_sustained"]
    achieved = flops_rank / (ms_gemm * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01", "final", "ncu", "traffic.json")
    if world == 1 and os.path.exists(tpath):  # DRAM bytes per launch from the committed ncu --set full captures
        traffic = json.load(open(tpath)).get("llama2_7b_prefill_m1024", {}).get("avg_per_launch")
    roofline = dict(bound="tensor", kernel="qqq_gemm_kernel<per-channel> (tcgen05 kind::i8)", achieved=round(achieved, 1),
                    peak=round(int8_peak, 1), unit="TFLOP/s", frac=round(achieved / int8_peak, 4), traffic=traffic,
                    traffic_note="dram__bytes_read+write per launch, mean over the 7 GEMMs of a layer (ncu, "
                                 "profiles/r01/final/ncu/traffic.json); below the algorithmic bytes because A8 and D stay in L2",
                    peak_source=f"2 x bf16_tflops_sustained of MEASURED_PEAKS.json ({peaks['source']}); int8 dense = 2x bf16 "
                                "on sm_100a; UTCIMMA-only microbenchmark on this pool measured 4428 TOP/s burst "
                                "(profiles/r01/probe_umma_i8.log)",
                    frac_of_burst_umma_i8_peak=round(achieved / 4428.0, 4),  # the stricter, clock-unthrottled denominator
                    launches=n_gemm, avg_launch_us=round(ms_gemm * 1e3 / n_gemm, 2),
                    algorithmic_flops_per_step=flops_rank, algorithmic_bytes_per_step=model_gemm_bytes(M) / world)

    line = None
    if rank == 0:
        line = dict(metric="llama2_7b_w4a8_prefill_linears_tokens_per_s", value=round(value, 1), unit="tokens/s",
                    n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=round(ms, 4), higher_is_better=True,
                    scaling="strong", vs_baseline=None, dtype="int8", data="synthetic", config=cfg,
                    clocks=cs.summary(),
                    e2e=dict(value=round(M / (ms_e2e * 1e-3), 1), unit="tokens/s", ms_per_step=round(ms_e2e, 4),
                             h2d_bytes_per_step=x_host.numel() * 2, d2h_bytes_per_step=out_host.numel() * 2,
                             api="qqq_b200.graph.capture(QuantLinear chain); eager QuantLinear.forward loop: "
                                 f"{ms_eager:.3f} ms/step"),
                    gpu_launches=int(launches_per_step * args.steps), gpu_launches_per_step=int(launches_per_step),
                    launch_mode="cuda-graph replay of the per-step launches",
                    merged=None if ms_mrg is None else dict(
                        ms_per_step=round(ms_mrg, 4), value=round(M / (ms_mrg * 1e-3), 1),
                        note="same weights with q/k/v and gate/up merged by concatenating their packed tensors "
                             "(qqq_b200.merge_quant_linears, bit-identical outputs): 4 act-quants + 4 GEMMs per layer"),
                    roofline=roofline, tflops_linears=round(flops_rank * world / (ms * 1e-3) / 1e12, 1))
        # Auxiliary sections: each guarded, all under one watchdog, so that neither an exception nor a stall in them
        # can cost the headline line above.
        aux = []
        if world == 1 and not args.no_cpu:  # first: the contract's cpu_baseline must not fall to the watchdog
            aux.append(("cpu_baseline", lambda: cpu_baseline()[0]))
        if world == 1 and not args.no_sweep:
            aux.append(("gemm_sweep", lambda: gemm_sweep(dev, peaks)))
        def shared_act_quant():
            # the headline's 7-module structure with the opt-in activation-quant cache: q/k/v and gate/up quantise their
            # shared input once (bit-identical outputs; 4 act-quants + 7 GEMMs per layer)
            qqq_b200.set_act_quant_cache(True)
            try:
                l1 = qqq_b200.launch_count()
                forward_chain(layers, x_dev, world)
                n_l = qqq_b200.launch_count() - l1
                gr = qgraph.capture(lambda x: forward_chain(layers, x, world), x_dev)
                ms_c = timed(lambda: gr(x_dev), args.steps, args.warmup, barrier)
            finally:
                qqq_b200.set_act_quant_cache(False)
            return dict(ms_per_step=round(ms_c, 4), value=round(M / (ms_c * 1e-3), 1), gpu_launches_per_step=int(n_l),
                        note="qqq_b200.set_act_quant_cache(True): same modules and call order as the headline")

        if world == 1 and not args.no_merged:
            aux.append(("shared_act_quant", shared_act_quant))
        if world == 1 and not args.no_decode:
            aux.append(("decode_g128", lambda: decode_g128(dev, peaks, args.steps, args.warmup)))

        if world == 1 and not args.no_full:
            aux.append(("full_forward", lambda: full_forward(dev, args.steps, args.warmup)))
        if world == 1 and not args.no_sweep:  # the sweep shape once transposed (K=21760, N=8192); last
            aux.append(("gemm_sweep_transposed", lambda: gemm_sweep(dev, peaks, K=21760, N=8192, Ms=(16, 1024))))

        def emit_and_exit():
            line["aux_timeout_s"] = args.aux_budget
            try:
                print(json.dumps(line))
            except Exception:
                print(json.dumps({k: v for k, v in list(line.items()) if k not in dict(aux)}))
            sys.stdout.flush()
            os._exit(0)

        wd = threading.Timer(args.aux_budget, emit_and_exit)
        wd.daemon = True
        wd.start()
        for key, fn in aux:
            try:
                line[key] = fn()
            except Exception as e:  # reported, never fatal
                line[key] = {"error": repr(e)[:300]}
        wd.cancel()
        print(json.dumps(line))
    if world > 1:
        # CUDA graphs captured above hold references into the NCCL communicator; tearing the process group down
        # underneath them can hang, so synchronise, flush and leave without running destructors.
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    return 0


if __name__ == "__main__":
    sys.exit(main())
