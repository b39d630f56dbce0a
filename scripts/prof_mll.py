"""One MLL value+grad (for ncu launch lists) and a few isolated GEMM shapes (for ncu --set full)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpjax_b200 import ops

mode = sys.argv[1]
dev = "cuda"
if mode == "mll":
    n = int(sys.argv[2])
    rng = np.random.default_rng(n)
    Xn = rng.uniform(-2, 2, (n, 8))
    yn = np.sin(Xn[:, :1]) + 0.1 * rng.standard_normal((n, 1))
    X, y = torch.as_tensor(Xn, device=dev), torch.as_tensor(yn, device=dev)
    ell = torch.as_tensor(np.linspace(0.8, 1.6, 8), device=dev).requires_grad_(True)
    var = torch.tensor(1.0, dtype=torch.float64, device=dev, requires_grad=True)
    sn = torch.tensor(0.3, dtype=torch.float64, device=dev, requires_grad=True)
    c = torch.tensor(0.0, dtype=torch.float64, device=dev, requires_grad=True)
    v = ops.conjugate_mll_fused(0, X, y, ell, var, sn, c, 1e-6)
    v.backward()
    torch.cuda.synchronize()
    print(v.item())
elif mode == "gemm":
    m = 16384
    P = torch.randn(m, 256, dtype=torch.float64, device=dev)
    Cb = torch.zeros(m, m, dtype=torch.float64, device=dev)
    for _ in range(2):
        ops.gemm(P, P, Cb, alpha=-1.0, beta=1.0)          # rank-256 full
    for _ in range(2):
        ops.gemm(P, P, Cb, alpha=-1.0, beta=1.0, mask=1)  # rank-256 lower
    A = torch.randn(4096, 4096, dtype=torch.float64, device=dev)
    C = torch.empty(4096, 4096, dtype=torch.float64, device=dev)
    for _ in range(2):
        ops.gemm(A, A, C)
    torch.cuda.synchronize()
elif mode == "ozaki":
    # one lower-masked rank-1024 update of the potrf shape through the int8 digit-plane kernel (for ncu --set full)
    m, k, s = 16384, 1024, 7
    X = torch.randn(m, k, dtype=torch.float64, device=dev) * 0.05
    Cb = torch.zeros(m, m, dtype=torch.float64, device=dev)
    Q, sc = ops.ozaki_slice(X, s)
    for _ in range(2):
        ops.ozaki_gemm_(Cb, Q, sc, Q, sc, k, s, alpha=-1.0, mask_lower=True)
    torch.cuda.synchronize()
