"""Long run of the intra-CTA protocol model (tests/pipeline_sim.py) on the drain-helper experiment build's planner: more
shapes, SM counts, CTAs and seeds than the CPU suite affords (dev tooling; ~20 min)."""
import sys, ctypes
import os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'tests'))
from pipeline_sim import CtaSim, segments
from test_schedule import KEYS
from test_pipeline_sim import SHAPES
# usage: python probes/sim_stress_helpers.py VARIANT.so   (VARIANT.so built by probes/build_variant.py ... -DQQQ_DRAIN_HELPERS)
lib = ctypes.CDLL(sys.argv[1]); lib.qqq_b200_plan.argtypes=[ctypes.c_int]*6+[ctypes.POINTER(ctypes.c_int)]; lib.qqq_b200_plan.restype=ctypes.c_int
def plan(M,N,K,gs,sms):
    out=(ctypes.c_int*20)(); assert lib.qqq_b200_plan(M,N,K,gs,sms,16,out)==0; return dict(zip(KEYS,out))
n=0
extra=[(512,4096,4096,-1),(768,11008,4096,-1),(2048,4096,4096,128),(333,1024,2048,-1),(96,8192,8192,-1),(48,4096,4096,128),(1024,128,8192,-1),(1024,3584,8192,-1)]
for (M,N,K,gs) in SHAPES+extra:
    for sms in (148,132,37,3,1):
        p=plan(M,N,K,gs,sms)
        grid=p['grid']>>p['pair']
        for cta in sorted(set([0,1%grid,grid//3,grid//2,grid-1])):
            if not segments(p,cta): continue
            for seed in range(6):
                CtaSim(p,cta,M,seed=seed,helpers=True,twin=bool(p['pair'])).run(); n+=1
print("simulated", n, "CTA runs with helpers: all finished, all parity waits exact, all chunks drained once")
