"""Shapes that dominate the exact-GP and SGPR paths, timed in isolation (CUDA events, best of 3)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpjax_b200 import ops

dev = "cuda"
from gpjax_b200._lib import lib
VAR = int(sys.argv[1]) if len(sys.argv) > 1 else 0
lib().gpb_debug_set_gemm_variant(VAR)
print("=== gemm variant", VAR)

def t(fn, rep=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(rep):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e-3)
    return best

n = 8192
A = torch.randn(n, n, dtype=torch.float64, device=dev); B = torch.randn(n, n, dtype=torch.float64, device=dev)
C = torch.empty(n, n, dtype=torch.float64, device=dev)
print("cuBLAS 8192^3      %.2f TF/s" % (2 * n**3 / t(lambda: torch.mm(A, B.T, out=C)) / 1e12))
for al, bl in ((0, 0), (0, 1), (1, 0), (1, 1)):
    print("ours 8192^3 l%d%d    %.2f TF/s" % (al, bl, 2 * n**3 / t(lambda: ops.gemm(A, B, C, a_layout=al, b_layout=bl)) / 1e12))
del A, B, C
m = 44544
for k in (256, 512, 1024):
    P = torch.randn(m, k, dtype=torch.float64, device=dev)
    Cb = torch.zeros(m, m, dtype=torch.float64, device=dev)
    tt = t(lambda: ops.gemm(P, P, Cb, alpha=-1.0, beta=1.0, mask=1))
    print("syrk lower %d^2 K=%d: %.2f ms %.2f TF/s (n^2 k)" % (m, k, tt * 1e3, m * m * k / tt / 1e12))
    tt = t(lambda: ops.gemm(P, P, Cb, alpha=-1.0, beta=1.0))
    print("rank-k full %d^2 K=%d: %.2f ms %.2f TF/s" % (m, k, tt * 1e3, 2 * m * m * k / tt / 1e12))
    del P, Cb
# trtri-like rectangle: C[20480 x 24064] += A[20480 x 512] * B[24064 x 512]^T, beta=1
Ar = torch.randn(20480, 512, dtype=torch.float64, device=dev); Br = torch.randn(24064, 512, dtype=torch.float64, device=dev)
Cr = torch.zeros(20480, 24064, dtype=torch.float64, device=dev)
tt = t(lambda: ops.gemm(Ar, Br, Cr, beta=1.0))
print("rect 20480x24064 K=512 beta=1: %.2f TF/s" % (2 * 20480 * 24064 * 512 / tt / 1e12))
# SGPR shapes: At = K^T Linv^T (65536 x 2048 x 2048, beta 0) and SYRK TN (2050 x 2050, K = 65536)
Kt = torch.randn(65536, 2050, dtype=torch.float64, device=dev); Li = torch.randn(2048, 2048, dtype=torch.float64, device=dev)
At = torch.empty(65536, 2050, dtype=torch.float64, device=dev)
tt = t(lambda: ops.gemm(Kt[:, :2048], Li, At[:, :2048]))
print("sgpr whiten 65536x2048x2048: %.2f TF/s (full count)" % (2 * 65536 * 2048 * 2048 / tt / 1e12))
Pm = torch.zeros(2050, 2050, dtype=torch.float64, device=dev)
tt = t(lambda: ops.gemm(At, At, Pm, beta=1.0, a_layout=1, b_layout=1, mask=1))
print("sgpr syrk TN 2050^2 K=65536 lower: %.2f ms %.2f TF/s (n^2 k)" % (tt * 1e3, 2050 * 2050 * 65536 / tt / 1e12))
