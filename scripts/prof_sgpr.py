"""One SGPR collapsed_elbo value+grad at N=1M, M=2048 (for ncu launch lists)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpjax_b200.sgpr_ops import collapsed_elbo_fused

dev = "cuda"
n, m, d = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000, 2048, 8
rng = np.random.default_rng(4)
X = torch.as_tensor(rng.uniform(-2, 2, (n, d)), device=dev)
y = torch.sin(X[:, :1]) + 0.1 * torch.randn(n, 1, dtype=torch.float64, device=dev)
Z = torch.as_tensor(np.random.default_rng(5).uniform(-2, 2, (m, d)), device=dev).requires_grad_(True)
ell = torch.as_tensor(np.linspace(0.8, 1.6, d), device=dev).requires_grad_(True)
var = torch.tensor(1.0, dtype=torch.float64, device=dev, requires_grad=True)
sn = torch.tensor(0.3, dtype=torch.float64, device=dev, requires_grad=True)
c = torch.tensor(0.0, dtype=torch.float64, device=dev, requires_grad=True)
stats = sys.argv[2] if len(sys.argv) > 2 else "auto"  # "raw" = the route the benchmark settles on (cond(Kzz) ~ 5e2)
v = collapsed_elbo_fused(0, X, y, Z, ell, var, sn, c, 1e-6, 65536, None, stats)
v.backward()
torch.cuda.synchronize()
print(v.item())
