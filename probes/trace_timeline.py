"""In-kernel timeline of one CTA (dev tooling): needs probes/libqqq_b200_trace.so (built with -DQQQ_TRACE).
usage: QQQ_B200_LIB=probes/libqqq_b200_trace.so python probes/trace_timeline.py M gs [K N]"""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("QQQ_B200_LIB", os.path.join(ROOT, "probes", "libqqq_b200_trace.so"))
import qqq_b200
from qqq_b200 import _lib

M = int(sys.argv[1]); gs = int(sys.argv[2])
K = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
N = int(sys.argv[4]) if len(sys.argv) > 4 else 21760
dev = "cuda:0"
lib = _lib.load()
buf = torch.zeros(20 * 2048, dtype=torch.int64, device=dev)
lib.qqq_trace_set.argtypes = [ctypes.c_void_p]
assert lib.qqq_trace_set(buf.data_ptr()) == 0
g = torch.Generator(device=dev).manual_seed(0)
Bs = [torch.randint(-2**31, 2**31 - 1, (K // 16, 2 * N), dtype=torch.int32, device=dev, generator=g) for _ in range(3)]
s2 = torch.rand(1, N, device=dev) * 1e-3 + 5e-4
s3 = (torch.rand(K // 128, N, device=dev) * 8 + 4).half() if gs == 128 else torch.zeros(0, dtype=torch.float16, device=dev)
C = torch.zeros(16 * 64, N, dtype=torch.int32, device=dev)
ws = torch.zeros(max(N // 128 * 16, 16), dtype=torch.int32, device=dev)
A = torch.randint(-127, 128, (M, K), dtype=torch.int8, device=dev)
s1 = torch.rand(M, 1, device=dev) * 1e-2 + 1e-3
D = torch.empty(M, N, dtype=torch.float16, device=dev)
for i in range(4):
    buf.zero_()
    qqq_b200.qqq_gemm(A, Bs[i % 3], C, D, s1, s2, s3, ws, -1, -1, -1, 16)
torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(20, 2048)
t0 = t[12, 0]
def rel(role):
    r = t[role].astype(np.int64)
    n = int((r != 0).sum())
    return (r[:n] - t0) if n else np.zeros(0, np.int64)
P0, P1, M5, M6, E7, E8, E11 = (rel(i) for i in (0, 1, 5, 6, 7, 8, 11))
end = t[12, 1] - t0
print(f"== M={M} gs={gs} K={K} N={N}: CTA lifetime after setup {end} cycles; units {len(P0)} (mma units {len(M5)})")
def stats(name, a):
    if len(a): print(f"  {name:52s} n={len(a):4d} mean={a.mean():8.1f} p50={np.median(a):8.1f} max={a.max():8d}")
stats("weights producer: issue duration (TR1-TR0)", P1 - P0)
stats("weights producer: interval between stage grants", np.diff(P0))
stats("mma: interval between unit issues", np.diff(M5))
stats("mma: issue block duration (TR6-TR5)", M6 - M5)
T16, T17, T18 = rel(16), rel(17), rel(18)
if len(T16) and len(T16) == len(M5) and len(M6) == len(M5):
    stats("mma: wait for tokens after previous issue block (TR16[i]-TR6[i-1])", (T16[1:] - M6[:-1]))
    stats("mma: extra wait for unpacked weights (TR17-TR16)", T17 - T16)
    n = min(len(T18), len(T16))
    stats("tokens: TMA issue -> seen by the MMA warp (TR16-TR18)", T16[:n] - T18[:n])
stats("epilogue: drain (TR8-TR7)", E8 - E7)
stats("epilogue: total incl fixup (TR11-TR7)", E11 - E7)
E13, E14 = rel(13), rel(14)
if len(E13) and len(E13) == len(E14):
    stats("epilogue warp 0: chunk convert+store (TR14-TR13)", E14 - E13)
    if len(E13) > 1:
        stats("epilogue warp 0: wait for next chunk's tcgen05.ld (TR13[i+1]-TR14[i])", E13[1:] - E14[:-1])
E9, E15 = rel(9), rel(15)
if len(E9) and len(E9) == len(E13) and len(E15) == len(E13):
    stats("epilogue warp 0: next LDTM issue + convert + 16 STS (TR9-TR13)", E9 - E13)
    stats("epilogue warp 0: __syncwarp (TR15-TR9)", E15 - E9)
    stats("epilogue warp 0: 2 LDS.128 + 2 STG.128 + __syncwarp (TR14-TR15)", E14 - E15)
idx = np.nonzero(t[2])[0]
u2 = t[2, idx].astype(np.int64) - t0; u3 = t[3, idx].astype(np.int64) - t0; u10 = t[10, idx].astype(np.int64) - t0; u4 = t[4, idx].astype(np.int64) - t0
stats("unpack(q0 warps): LDS issue + wait aempty (TR3-TR2)", u3 - u2)
stats("unpack: ALU + STTM issue (TR10-TR3)", u10 - u3)
stats("unpack: wait::st + arrive (TR4-TR10)", u4 - u10)
stats("unpack: interval between sub-blocks (all groups)", np.diff(u2))
print("  first 12 sub-blocks: [itn] wfull_seen slot_ok st_issued arrived")
for j, it in enumerate(idx[:12]):
    print(f"   [{it:3d}] {u2[j]:8d} {u3[j]:8d} {u10[j]:8d} {u4[j]:8d}")
print("  weights producer grants (first 12):", P0[:12].tolist())
print("  mma unit issue times (first 12):", M5[:12].tolist())
print("  epilogue segments start/drained/end:", list(zip(E7.tolist(), E8.tolist(), E11.tolist()))[:6])
