"""Helpers for the -m gpu tests: move an oracle problem to the device and call the product through its
public Python boundary (which goes through the C ABI in include/qqq_b200.h)."""
import numpy as np
import torch

import qqq_b200


def to_dev(p, dev="cuda:0"):
    return {k: torch.from_numpy(np.ascontiguousarray(p[k])).to(dev) for k in ("A8", "B", "s1", "s2", "s3")}


def run_gemm(p, N, max_par=16, dev="cuda:0", scratch=None, sms=-1):
    t = to_dev(p, dev)
    M = t["A8"].shape[0]
    if scratch is None:
        C = torch.zeros((max_par * 64, N), dtype=torch.int32, device=dev)
        ws = torch.zeros(max(N // 128 * max_par, 1), dtype=torch.int32, device=dev)
    else:
        C, ws = scratch
    D = torch.full((M, N), float("nan"), dtype=torch.float16, device=dev)
    qqq_b200.qqq_gemm(t["A8"], t["B"], C, D, t["s1"], t["s2"], t["s3"], ws, -1, -1, sms, max_par)
    torch.cuda.synchronize()
    return D.cpu().numpy(), C, ws


def bits(a):
    return np.ascontiguousarray(a).view(np.uint16)
