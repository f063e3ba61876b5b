"""bench.py's host logic on a CPU-only machine: main() runs end to end on a mocked CUDA surface with the oracle behind
the C ABI (tests/helpers/bench_on_mock_cuda.py) and must print ONE JSON line that honours the driver's contract —
including when an auxiliary section fails."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HELPER = os.path.join(ROOT, "tests", "helpers", "bench_on_mock_cuda.py")


def _run(*extra):
    for attempt in (0, 1):  # a loaded CI host may kill or starve one subprocess: one retry, then the failure is real
        out = subprocess.run([sys.executable, HELPER, *extra], capture_output=True, text=True, timeout=600, cwd=ROOT)
        lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
        if out.returncode == 0 and len(lines) == 1:
            return json.loads(lines[0])
    assert out.returncode == 0, out.stderr[-2000:]
    assert len(lines) == 1, f"bench.py must print exactly one JSON line, got {len(lines)}: {out.stdout[-500:]}"


@pytest.fixture(scope="module")
def line():
    return _run()


def test_contract_keys(line):
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in line, k
    assert line["n_gpus"] == 1 and line["steps"] == 1 and line["warmup"] == 1 and line["higher_is_better"] is True
    assert line["vs_baseline"] is None and line["data"] == "synthetic" and "workload" in line["config"]
    assert set(line["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    assert set(line["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert set(line["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"}
    assert set(line["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}


def test_launch_accounting_matches_the_call_structure(line):
    # 2 tiny layers x (7 GEMMs + 4 activation quants: q/k/v and gate/up share theirs); the reference's literal structure
    # (every linear quantises for itself) is the `per_module_quant` section
    assert line["gpu_launches_per_step"] == 14 + 8 and line["gpu_launches"] == 22 * line["steps"]
    assert line["per_module_quant"]["gpu_launches_per_step"] == 28


def test_auxiliary_sections_present(line):
    assert isinstance(line["gemm_sweep"], list) and {r["mode"] for r in line["gemm_sweep"]} == {"per-channel", "g128"}
    assert isinstance(line["gemm_sweep_transposed"], list)
    assert line["decode_g128"]["unit"] == "tokens/s" and "merged" in line["decode_g128"]
    assert line["merged"]["value"] > 0
    ff = line["full_forward"]
    assert ff["qqq_launches_per_step"] == 28 and ff["logits_finite"] is True and ff["fused_qkv_gate_up"]["value"] > 0


def test_a_failing_auxiliary_section_does_not_cost_the_line():
    line = _run("--break-sweep")
    assert "error" in line["gemm_sweep"] and "error" in line["gemm_sweep_transposed"]
    assert line["value"] > 0 and "decode_g128" in line and "cpu_baseline" in line


def test_flags_skip_sections():
    line = _run("--no-sweep", "--no-cpu", "--no-merged", "--no-decode", "--no-full")
    for k in ("gemm_sweep", "gemm_sweep_transposed", "cpu_baseline", "decode_g128", "per_module_quant", "full_forward"):
        assert k not in line
    assert line["merged"] is None


def _torchrun(script, *args, nproc=2, port=29541):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr",
           "127.0.0.1", "--master-port", str(port), script, *args]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    return [json.loads(ln) for ln in out.stdout.splitlines() if ln.startswith("{")]


def test_two_ranks_tensor_parallel_prints_one_line_from_rank0():
    """The driver's N > 1 launch (torchrun, one rank per GPU): gloo stands in for NCCL, shards come from tp.split_sizes."""
    lines = _torchrun(HELPER, "--gpus", "2")
    assert len(lines) == 1
    line = lines[0]
    assert line["n_gpus"] == 2 and line["config"]["parallelism"].startswith("tp2") and line["scaling"] == "strong"
    # default exchange: fused into the kernels (scatter GEMM + reduce/quant/gather), emulated over gloo here
    assert line["tp_mode"] == "scatter" and line["tp_parity"]["green"] is True and line["exchange_timeouts"] == 0
    # per rank and layer: 7 GEMMs + 2 local activation quants (inputs of o / down) + 2 reduce-quant kernels; + 1 quant of x
    assert line["gpu_launches_per_step"] == 2 * 11 + 1
    assert line["e2e"]["d2h_bytes_per_step"] * 2 == line["e2e"]["h2d_bytes_per_step"]  # this rank's half of the rows
    assert "gemm_sweep" not in line and "cpu_baseline" not in line  # rank 0 at N = 1 only
    rows = line["gemm_sweep_tp"]
    assert {r["split"] for r in rows} == {"N", "K"} and {r["mode"] for r in rows} == {"per-channel", "g128"}
    assert all("fused_exchange_us" in r and "nccl_allreduce_us" in r for r in rows if r["split"] == "K")
    assert line["llama2_70b_tp"]["value"] > 0 and line["llama2_70b_tp"]["exchange_timeouts"] == 0


def test_two_ranks_nccl_mode():
    lines = _torchrun(HELPER, "--gpus", "2", "--tp-mode", "nccl", "--no-70b", "--no-tp-sweep", port=29549)
    assert len(lines) == 1 and lines[0]["tp_mode"] == "nccl" and lines[0]["tp_parity"]["green"] is True
    assert lines[0]["gpu_launches_per_step"] == 22 and lines[0]["merged"]["value"] > 0


def test_reference_arm_under_torchrun_only_rank0_speaks():
    lines = _torchrun(os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1",
                      port=29543)
    assert len(lines) == 1 and lines[0]["impl"] == "reference"
    assert lines[0]["e2e"]["h2d_bytes_per_step"] == 0 and lines[0]["cpu_baseline"]["kind"] == "port"


def test_two_ranks_with_the_all_reduce_fused_into_the_gemm():
    """--fused-allreduce wiring (row-parallel linears reduce in their own epilogue; multicast emulated over gloo): same
    launch count, no NCCL all-reduce in the chain, and the same numbers as the unfused run up to fp16 summation order."""
    fused = _torchrun(HELPER, "--gpus", "2", "--fused-allreduce", "--no-70b", "--no-tp-sweep", port=29547)
    assert len(fused) == 1
    assert "multimem.red" in fused[0]["config"]["parallelism"] and fused[0]["tp_mode"] == "reduce"
    assert fused[0]["gpu_launches_per_step"] == 22
