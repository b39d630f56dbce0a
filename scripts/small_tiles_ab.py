"""A/B of the quarter-tile dispatch for few-tile DMMA GEMMs (GPB_GEMM_SMALL_TILES, read once per process): per-evaluation wall time of
conjugate_mll value + gradient at small N, and the value / gradient themselves (the two settings must agree to rounding)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpjax_b200 import ops
dev = "cuda"
for n, d in ((1000, 1), (2000, 8), (5000, 8), (20000, 8)):
    rng = np.random.default_rng(123)
    X = torch.as_tensor(rng.uniform(-2, 2, (n, d)), device=dev); y = torch.sin(X[:, :1]) + 0.1 * torch.as_tensor(rng.standard_normal((n, 1)), device=dev)
    ell = torch.ones(d, dtype=torch.float64, device=dev).requires_grad_(True)
    var = torch.tensor(1.0, dtype=torch.float64, device=dev, requires_grad=True); sn = torch.tensor(0.3, dtype=torch.float64, device=dev, requires_grad=True)
    def step():
        for p in (ell, var, sn): p.grad = None
        v = ops.conjugate_mll_fused(0, X, y, ell, var, sn, None, 1e-6); v.backward(); return v
    reps = 50 if n <= 5000 else 5
    for _ in range(3): v = step()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): step()
    torch.cuda.synchronize(); t = (time.perf_counter() - t0) / reps
    print(f"small_tiles={os.environ.get('GPB_GEMM_SMALL_TILES', '1')} N={n} D={d}: {t*1e3:.3f} ms per MLL value+grad; value {v.item():.12e} "
          f"g_ell0 {ell.grad[0].item():.12e} g_sn {sn.grad.item():.12e}", flush=True)
