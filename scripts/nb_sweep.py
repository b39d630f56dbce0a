"""conjugate_mll value+grad time vs N for the block size the library was built with (GPB_NB=... python -m gpjax_b200.build --force)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpjax_b200 import ops
from gpjax_b200._lib import lib
dev = "cuda"
print("NB =", lib().gpb_block_size(), "(small orders),", lib().gpb_block_size_for(50000), "(N = 50,000)")
for n in [int(a) for a in sys.argv[1:]] or (2000, 5000, 10000, 20000, 30000):
    d = 8
    rng = np.random.default_rng(123)
    X = torch.as_tensor(rng.uniform(-2, 2, (n, d)), device=dev); y = torch.sin(X[:, :1]) + 0.1 * torch.randn(n, 1, dtype=torch.float64, device=dev)
    ell = torch.linspace(0.8, 1.6, d, dtype=torch.float64, device=dev).requires_grad_(True)
    var = torch.tensor(1.0, dtype=torch.float64, device=dev, requires_grad=True); sn = torch.tensor(0.3, dtype=torch.float64, device=dev, requires_grad=True)
    def step():
        for p in (ell, var, sn): p.grad = None
        v = ops.conjugate_mll_fused(0, X, y, ell, var, sn, None, 1e-6); v.backward(); return v
    for _ in range(2): step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3 if n >= 20000 else 10
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): step()
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / reps
    print(f"N={n}: {t:.2f} ms  ({n**3 / t / 1e9:.2f} TF/s effective)", flush=True)
    ops.release_buffers() if hasattr(ops, "release_buffers") else None
