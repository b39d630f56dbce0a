"""First-contact measurements on the B200: FP64 ceilings (cuBLAS DGEMM, cuSOLVER potrf via torch)
next to our DMMA GEMM / blocked Cholesky / MLL fwd+bwd.  Writes gpurun_out/first_look.json."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpjax_b200 import ops  # noqa: E402

out = {}
dev = "cuda"


def timeit(fn, warm=1, rep=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(rep):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    return min(ts), float(np.median(ts))


print(torch.cuda.get_device_name(0), torch.cuda.mem_get_info())
# --- ceilings -------------------------------------------------------------------------------
n = 8192
A = torch.randn(n, n, dtype=torch.float64, device=dev)
B = torch.randn(n, n, dtype=torch.float64, device=dev)
tb, tm = timeit(lambda: A @ B.T, 2, 5)
out["cublas_dgemm_8192_tflops"] = 2 * n**3 / tb / 1e12
print("cuBLAS DGEMM 8192^3: %.2f TF/s (best) %.2f (median)" % (2 * n**3 / tb / 1e12, 2 * n**3 / tm / 1e12))
C = torch.empty(n, n, dtype=torch.float64, device=dev)
tb, tm = timeit(lambda: ops.gemm(A, B, C), 2, 5)
out["ours_dgemm_8192_tflops"] = 2 * n**3 / tb / 1e12
print("ours   DGEMM 8192^3 NT: %.2f TF/s (best) %.2f (median)" % (2 * n**3 / tb / 1e12, 2 * n**3 / tm / 1e12))
err = float((C - A @ B.T).abs().max())
print("   max abs err vs cuBLAS", err)
for al, bl in [(0, 1), (1, 0), (1, 1)]:
    tb, _ = timeit(lambda: ops.gemm(A, B, C, a_layout=al, b_layout=bl), 1, 3)
    print("ours   DGEMM 8192^3 layouts", al, bl, ": %.2f TF/s" % (2 * n**3 / tb / 1e12))
    out[f"ours_dgemm_8192_l{al}{bl}_tflops"] = 2 * n**3 / tb / 1e12
# rank-256 update shape (the Cholesky trailing update): M=N=32768, K=256, beta=1
m = 32768
P = torch.randn(m, 256, dtype=torch.float64, device=dev)
Cb = torch.zeros(m, m, dtype=torch.float64, device=dev)
tb, tm = timeit(lambda: ops.gemm(P, P, Cb, alpha=-1.0, beta=1.0), 1, 3)
out["ours_rank256_full_tflops"] = 2 * m * m * 256 / tb / 1e12
print("ours rank-256 update %d^2 full: %.2f TF/s" % (m, 2 * m * m * 256 / tb / 1e12))
tb, tm = timeit(lambda: ops.gemm(P, P, Cb, alpha=-1.0, beta=1.0, mask=1), 1, 3)
out["ours_rank256_lower_tflops"] = m * m * 256 / tb / 1e12
print("ours rank-256 update %d^2 lower: %.2f TF/s (counting n^2 k)" % (m, m * m * 256 / tb / 1e12))
tb, tm = timeit(lambda: torch.addmm(Cb, P, P.T, beta=1.0, alpha=-1.0, out=Cb), 1, 3)
print("cuBLAS rank-256 update full: %.2f TF/s" % (2 * m * m * 256 / tb / 1e12))
out["cublas_rank256_full_tflops"] = 2 * m * m * 256 / tb / 1e12
del A, B, C, P, Cb
torch.cuda.empty_cache()

# --- Cholesky ---------------------------------------------------------------------------------
for n in (8192, 20000, 50000):
    rng = np.random.default_rng(0)
    X = torch.as_tensor(rng.uniform(-2, 2, (n, 8)), device=dev)
    ell = torch.as_tensor(np.linspace(0.8, 1.6, 8), device=dev)
    var = torch.tensor(1.0, dtype=torch.float64, device=dev)
    sn = torch.tensor(0.3, dtype=torch.float64, device=dev)
    S = torch.empty(n, n, dtype=torch.float64, device=dev)
    tg, _ = timeit(lambda: ops.gram_forward(0, X, X, ell, var, 1e-6, sn, False, S), 1, 3)
    print(f"N={n}: gram full {tg*1e3:.2f} ms ({n*n*8/tg/1e9:.0f} GB/s write)")
    out[f"gram_full_{n}_ms"] = tg * 1e3
    tg, _ = timeit(lambda: ops.gram_forward(0, X, X, ell, var, 1e-6, sn, True, S), 1, 3)
    print(f"N={n}: gram lower {tg*1e3:.2f} ms")
    out[f"gram_lower_{n}_ms"] = tg * 1e3
    ws = ops.FactorWorkspace(n, 8, potri=True, device=dev)

    def ours():
        ops.gram_forward(0, X, X, ell, var, 1e-6, sn, True, S)
        ops.potrf_lower_(S, ws, zero_upper=False)

    tb, _ = timeit(ours, 1, 2)
    tb -= tg
    out[f"ours_potrf_{n}_s"] = tb
    print(f"N={n}: ours potrf {tb:.3f} s = {n**3/3/tb/1e12:.2f} TF/s")
    if n <= 20000:
        ops.gram_forward(0, X, X, ell, var, 1e-6, sn, False, S)
        S2 = S.clone()
        t0, _ = timeit(lambda: torch.linalg.cholesky(S2), 1, 2)
        out[f"cusolver_potrf_{n}_s"] = t0
        print(f"N={n}: torch.linalg.cholesky {t0:.3f} s = {n**3/3/t0/1e12:.2f} TF/s")
        Lref = torch.linalg.cholesky(S2)
        ours()
        d = float((torch.tril(S) - Lref).abs().max())
        print("   max |L - Lref| =", d)
        del S2, Lref
    del S, ws
    torch.cuda.empty_cache()

# --- full MLL value + grad -----------------------------------------------------------------------
for n in (20000, 50000):
    rng = np.random.default_rng(n)
    Xn = rng.uniform(-2, 2, (n, 8))
    yn = np.sin(Xn[:, :1]) + 0.1 * rng.standard_normal((n, 1))
    X, y = torch.as_tensor(Xn, device=dev), torch.as_tensor(yn, device=dev)
    ell = torch.as_tensor(np.linspace(0.8, 1.6, 8), device=dev).requires_grad_(True)
    var = torch.tensor(1.0, dtype=torch.float64, device=dev, requires_grad=True)
    sn = torch.tensor(0.3, dtype=torch.float64, device=dev, requires_grad=True)
    c = torch.tensor(0.0, dtype=torch.float64, device=dev, requires_grad=True)
    kind = 2 if n == 20000 else 0

    def step():
        for p in (ell, var, sn, c):
            p.grad = None
        v = ops.conjugate_mll_fused(kind, X, y, ell, var, sn, c, 1e-6)
        v.backward()
        return v

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    step()
    torch.cuda.synchronize()
    ev[0].record()
    v = ops.conjugate_mll_fused(kind, X, y, ell, var, sn, c, 1e-6)
    ev[1].record()
    v.backward()
    ev[2].record()
    torch.cuda.synchronize()
    tf, tbk = ev[0].elapsed_time(ev[1]) * 1e-3, ev[1].elapsed_time(ev[2]) * 1e-3
    print(f"N={n} kind={kind}: MLL fwd {tf:.3f} s, bwd {tbk:.3f} s, total {tf+tbk:.3f} s -> {n**3/(tf+tbk)/1e12:.2f} TF/s;"
          f" value {v.item():.10f} g_ell[0] {ell.grad[0].item():.8e} g_var {var.grad.item():.8e}")
    out[f"mll_{n}"] = dict(fwd_s=tf, bwd_s=tbk, value=v.item(), tflops=n**3 / (tf + tbk) / 1e12)
    ops.release_buffers()
    torch.cuda.empty_cache()

os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/first_look.json", "w"), indent=1)
print(json.dumps(out))

# --- SGPR collapsed_elbo value + grad --------------------------------------------------------------
from gpjax_b200.sgpr_ops import collapsed_elbo_fused  # noqa: E402
import gpjax_b200.sgpr_ops as sgpr_ops  # noqa: E402

for n, block in ((1_000_000, 32768), (1_000_000, 65536)):
    m, d = 2048, 8
    rng = np.random.default_rng(4)
    X = torch.as_tensor(rng.uniform(-2, 2, (n, d)), device=dev)
    y = torch.sin(X[:, :1]) + 0.1 * torch.randn(n, 1, dtype=torch.float64, device=dev)
    Z = torch.as_tensor(np.random.default_rng(5).uniform(-2, 2, (m, d)), device=dev).requires_grad_(True)
    ell = torch.as_tensor(np.linspace(0.8, 1.6, d), device=dev).requires_grad_(True)
    var = torch.tensor(1.0, dtype=torch.float64, device=dev, requires_grad=True)
    sn = torch.tensor(0.3, dtype=torch.float64, device=dev, requires_grad=True)
    c = torch.tensor(0.0, dtype=torch.float64, device=dev, requires_grad=True)

    def fwd():
        return collapsed_elbo_fused(0, X, y, Z, ell, var, sn, c, 1e-6, block)

    fwd().backward()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    v = fwd()
    ev[1].record()
    v.backward()
    ev[2].record()
    torch.cuda.synchronize()
    tf, tbk = ev[0].elapsed_time(ev[1]) * 1e-3, ev[1].elapsed_time(ev[2]) * 1e-3
    print(f"SGPR N={n} M={m} block={block}: fwd {tf:.3f} s bwd {tbk:.3f} s -> {n/(tf+tbk)/1e6:.3f} Mpoints/s, "
          f"{4*n*m*m/(tf+tbk)/1e12:.2f} TF/s (4NM^2); elbo {v.item():.8f}")
    out[f"sgpr_{n}_{block}"] = dict(fwd_s=tf, bwd_s=tbk, mpoints_s=n / (tf + tbk) / 1e6)
    sgpr_ops.release_buffers()
    del X, y
    torch.cuda.empty_cache()
json.dump(out, open("gpurun_out/first_look.json", "w"), indent=1)
