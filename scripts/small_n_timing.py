"""Per-iteration cost of conjugate_mll value+grad at small N (launch-bound regime), BASELINE config 1 shape."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpjax_b200 import ops
import gpjax_b200 as gpx
dev = "cuda"
for n, d in ((1000, 1), (2000, 8), (5000, 8)):
    rng = np.random.default_rng(123)
    X = torch.as_tensor(rng.uniform(-2, 2, (n, d)), device=dev); y = torch.sin(X[:, :1]) + 0.1 * torch.randn(n, 1, dtype=torch.float64, device=dev)
    ell = torch.ones(d, dtype=torch.float64, device=dev).requires_grad_(True)
    var = torch.tensor(1.0, dtype=torch.float64, device=dev, requires_grad=True); sn = torch.tensor(0.3, dtype=torch.float64, device=dev, requires_grad=True)
    def step():
        for p in (ell, var, sn): p.grad = None
        v = ops.conjugate_mll_fused(0, X, y, ell, var, sn, None, 1e-6); v.backward(); return v
    for _ in range(5): step()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(50): step()
    torch.cuda.synchronize(); t = (time.perf_counter() - t0) / 50
    print(f"N={n} D={d}: {t*1e3:.2f} ms per MLL value+grad (wall, incl. Python/ctypes)")
D = gpx.Dataset(X=X, y=y)
post = gpx.gps.Prior(mean_function=gpx.mean_functions.Zero(), kernel=gpx.kernels.RBF(lengthscale=[1.0]*8)) * gpx.likelihoods.Gaussian(num_datapoints=D.n)
t0 = time.perf_counter(); opt, hist = gpx.fit(model=post, objective=lambda p, d: -gpx.objectives.conjugate_mll(p, d), train_data=D, optim=gpx.optim.adam(0.01), num_iters=50, verbose=False); torch.cuda.synchronize()
print(f"gpx.fit 50 iters at N=5000: {(time.perf_counter()-t0)/50*1e3:.2f} ms/iter; loss {hist[0].item():.3f} -> {hist[-1].item():.3f}")
# fit(cuda_graph=True): the same loop with the step captured once and replayed (device time per iteration from CUDA events around the
# whole run, the first 3 ordinary iterations and the capture included in the graphed figure)
for n, d in ((1000, 1), (2000, 8), (5000, 8)):
    rng = np.random.default_rng(5)
    Xn = torch.as_tensor(rng.uniform(-2, 2, (n, d)), device=dev); yn = torch.sin(Xn[:, :1]) + 0.1 * torch.randn(n, 1, dtype=torch.float64, device=dev)
    Dn = gpx.Dataset(X=Xn, y=yn)
    out = {}
    for graphed in (False, True):
        post = gpx.gps.Prior(mean_function=gpx.mean_functions.Zero(), kernel=gpx.kernels.RBF(lengthscale=[1.0] * d)) * gpx.likelihoods.Gaussian(num_datapoints=n)
        iters = 203
        for rep in range(2):  # the first repetition pays the one-off costs (workspace allocation, attribute opt-ins)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            _, hist = gpx.fit(model=post, objective=lambda p, d_: -gpx.objectives.conjugate_mll(p, d_), train_data=Dn, optim=gpx.optim.adam(0.01), num_iters=iters, verbose=False, cuda_graph=graphed)
            torch.cuda.synchronize(); out[graphed] = ((time.perf_counter() - t0) / iters, hist[-1].item())
    print(f"gpx.fit N={n} D={d}, {iters} iters: ordinary {out[False][0]*1e3:.3f} ms/iter, cuda_graph {out[True][0]*1e3:.3f} ms/iter "
          f"(final loss {out[False][1]:.6f} / {out[True][1]:.6f})")
