"""Per-parameter errors of the SVGP ELBO gradient at the benchmarked M = 4096 against the oracle's autodiff (diagnostic)."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as o
from gpjax_b200.svgp_ops import svgp_elbo_fused
from gpjax_b200 import ops
n, m, d = 8192, 4096, 16
rng = np.random.default_rng(5)
X = rng.uniform(-2.0, 2.0, (n, d)); y = np.sin(X[:, :1]) + 0.1 * rng.standard_normal((n, 1)); Z = rng.uniform(-2.0, 2.0, (m, d))
mu = rng.standard_normal((m, 1)) * 0.3; W = np.tril(np.random.default_rng(6).standard_normal((m, m)) * 0.002) + 0.6 * np.eye(m)
ell = np.linspace(0.8, 1.6, d) * 2.0
ref, gref = o.svgp_elbo_value_and_grad_autodiff("matern32", X, y, Z, ell, 1.0, 0.3, 0.1, mu, W, 5e7)
dev = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda")
out = {"cond_kzz": float(np.linalg.cond(o.gram("matern32", Z, ell, 1.0) + 1e-6 * np.eye(m))), "ref": ref}
for mode in (-1, 0):
    ops.set_ozaki_slices(mode)
    for route in ("whitened", "raw"):
        p = {k: dev(v).requires_grad_(True) for k, v in dict(Z=Z, ell=ell, var=1.0, sn=0.3, c=0.1, mu=mu, W=W).items()}
        val = svgp_elbo_fused(1, dev(X), dev(y), p["Z"], p["ell"], p["var"], p["sn"], p["c"], p["mu"], p["W"], 5e7, 1e-6, 4096, None, route)
        val.backward()
        got = dict(inducing_inputs=p["Z"].grad, lengthscale=p["ell"].grad, variance=p["var"].grad, obs_stddev=p["sn"].grad,
                   mean_const=p["c"].grad, variational_mean=p["mu"].grad.reshape(-1), variational_root_covariance=p["W"].grad)
        e = {"value": abs(val.item() - ref) / abs(ref)}
        for k, b in gref.items():
            a, b = got[k].cpu().numpy().reshape(np.shape(b)), np.asarray(b)
            e[k] = float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
        out[f"ozaki={mode},{route}"] = e
print(json.dumps(out, indent=1))
