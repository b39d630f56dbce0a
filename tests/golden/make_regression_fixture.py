"""Regenerates tests/golden/regression_example.npz.

The data set of the reference's examples/regression.py:59-70 (key = jr.key(123), n = 100) is re-drawn with the
restated threefry generator (oracle/jax_prng.py, pre-0.5 "original" stream -- the one the stored goldens of
tests/integration_tests.py:99-107 were produced with), the MLL of a conjugate GP with RBF kernel and trainable
constant mean is minimised with SciPy BFGS on the CPU oracle, and the data, the optimum and the reference's own
stored golden numbers are saved.  jax is not needed.  Run from the repository root:  python tests/golden/make_regression_fixture.py
"""
import os
import sys

import numpy as np
from scipy.optimize import minimize
from scipy.special import erfinv

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as o  # noqa: E402
from oracle import jax_prng as jr  # noqa: E402


def uniform(bits, lo, hi):
    fb = (bits >> np.uint64(12)) | np.float64(1.0).view(np.uint64)
    return np.maximum(lo, (fb.view(np.float64) - 1.0) * (hi - lo) + lo)


key = jr.key(123)
key, sub = jr.split_original(key)
n = 100
x = uniform(jr.random_bits64_original(key, n), -3.0, 3.0).reshape(-1, 1)
lo = np.nextafter(np.float64(-1.0), 0.0)
noise = np.sqrt(2) * erfinv(uniform(jr.random_bits64_original(sub, n), lo, 1.0))
y = np.sin(4 * x) + np.cos(2 * x) + noise.reshape(-1, 1) * 0.3


def fun(u):
    ell, var, sn = o.softplus(u[:3])
    v, g = o.conjugate_mll_value_and_grad_autodiff("rbf", x, y, ell, var, sn, u[3])
    gc = np.array([g["lengthscale"], g["variance"], g["obs_stddev"]]) / (1 + np.exp(-u[:3]))
    return -v, -np.concatenate([gc, [g["mean_const"]]])


u0 = np.concatenate([o.softplus_inv(np.ones(3)), [0.0]])
res = minimize(fun, u0, jac=True, options={"maxiter": 500, "gtol": 1e-9})
ell, var, sn = o.softplus(res.x[:3])
xt = np.linspace(-3.5, 3.5, 500).reshape(-1, 1)
mean, cov = o.conjugate_predict("rbf", x, y, xt, ell, var, sn, res.x[3])
np.savez(os.path.join(os.path.dirname(os.path.abspath(__file__)), "regression_example.npz"),
         x=x, y=y, lengthscale=ell, variance=var, obs_stddev=sn, mean_const=res.x[3], neg_mll_at_optimum=res.fun,
         neg_mll_at_init=fun(u0)[0], xtest=xt, predictive_mean=mean, predictive_std=np.sqrt(np.diag(cov) + sn**2),
         reference_golden_history_last=55.07405622, reference_golden_predictive_mean_sum=36.24383416,
         reference_golden_predictive_std_sum=197.04727051)
print("optimum", res.fun, "stored golden 55.07405622; sums", mean.sum(), np.sqrt(np.diag(cov) + sn**2).sum())
