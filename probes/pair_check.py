"""CTA-pair mode bring-up (dev tooling): QQQ_B200_PAIR=1 python probes/pair_check.py M K N gs  -> bit-exact check vs torch._int_mm"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from test_gemm_parity import check_vs_int_mm
M, K, N, gs = (int(a) for a in sys.argv[1:5])
check_vs_int_mm(M, K, N, gs)
torch.cuda.synchronize()
print(f"pair={os.environ.get('QQQ_B200_PAIR','0')} {(M,K,N,gs)}: bit-exact", flush=True)
