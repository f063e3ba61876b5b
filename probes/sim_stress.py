"""Long run of the intra-CTA protocol model (tests/pipeline_sim.py) on a library's planner: more shapes, SM counts, CTAs and
seeds than the CPU suite affords (dev tooling; 10-20 min on one host core).

    python probes/sim_stress.py qqq_b200/libqqq_b200.so [seeds]      # pairs are modelled with both CTAs on one clock
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from pipeline_sim import CtaSim, PairSim, segments  # noqa: E402
from test_pipeline_sim import SHAPES  # noqa: E402
from test_schedule import KEYS  # noqa: E402

lib = ctypes.CDLL(os.path.abspath(sys.argv[1]))
seeds = int(sys.argv[2]) if len(sys.argv) > 2 else 4
lib.qqq_b200_plan.argtypes = [ctypes.c_int] * 6 + [ctypes.POINTER(ctypes.c_int)]
lib.qqq_b200_plan.restype = ctypes.c_int


def plan(M, N, K, gs, sms):
    out = (ctypes.c_int * 20)()
    assert lib.qqq_b200_plan(M, N, K, gs, sms, 16, out) == 0
    return dict(zip(KEYS, out))


extra = [(512, 4096, 4096, -1), (768, 11008, 4096, -1), (2048, 4096, 4096, 128), (333, 1024, 2048, -1), (96, 8192, 8192, -1),
         (48, 4096, 4096, 128), (1024, 128, 8192, -1), (1024, 3584, 8192, -1), (4096, 4096, 4096, -1), (4096, 8192, 21760, 128)]
n = n_pair = 0
for (M, N, K, gs) in SHAPES + extra:
    for sms in (148, 132, 37, 3):
        p = plan(M, N, K, gs, sms)
        grid = p["grid"] >> p["pair"]
        for cta in sorted({0, 1 % grid, grid // 3, grid // 2, grid - 1}):
            if not segments(p, cta):
                continue
            for seed in range(seeds):
                if p["pair"]:
                    PairSim(p, cta, M, seed=seed)
                    n_pair += 1
                else:
                    CtaSim(p, cta, M, seed=seed).run()
                n += 1
print(f"simulated {n} CTA runs ({n_pair} of them CTA pairs): all finished, all parity waits exact, "
      "all chunks drained once")
