#!/usr/bin/env python
"""ATTEMPT (negative result, kept as evidence) to reach the reference's stored golden for examples/collapsed_vi.py
(tests/integration_tests.py:110-118: history[-1] = 1924.7634809, sum(predictive_mean) = -8.39869652, sum(predictive_std) =
255.74838027) with the CPU oracle, the way tests/test_oracle_goldens.py::test_regression_example_golden reaches the one for
examples/regression.py: data from the restated threefry PRNG (both the original and the partitionable stream), Z = linspace(-3, 3, 50),
500 AdamW(1e-2) steps on the softplus-unconstrained parameters (examples/collapsed_vi.py:121-133, gpjax/fit.py:133-170).

Result (8 vCPU, ~5 s per run): every variant -- either PRNG stream, Adam or AdamW, noise parameterised as a standard deviation or as
a variance (older GPJax), inducing inputs / mean constant trainable or fixed -- ends at -ELBO = 1858.66 .. 1858.81 (original stream)
or 1850.9 .. 1851.1 (partitionable stream), i.e. the optimisation is converged and insensitive to those choices, and 66 away from
the stored 1924.76; the predictive sums come out as -5.34 / 249.08 against the stored -8.40 / 255.75 (an average predictive std of
0.498 against 0.512: the stored run saw noisier data).  The stored value therefore belongs to a different revision of the example's
DATA, not to a different optimiser state, and cannot be reached from this revision of the reference; the reference's own check
cannot notice (`Result._compare` catches its AssertionError and prints it, tests/integration_tests.py:56-63).  The ELBO oracle stays
pinned through the reference-held identity ELBO(Z = X) = MLL (tests/test_objectives.py:170-199 of the reference,
test_elbo_equals_mll_when_z_is_x here), a 40-digit adjudicator and the long-double route.

    python tests/golden/attempt_collapsed_vi_golden.py
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle as o
from oracle import jax_prng as jr
from scipy.special import erfinv

def uniform_from_bits(bits, lo, hi):
    fb = (bits >> np.uint64(12)) | np.float64(1.0).view(np.uint64)
    return np.maximum(lo, (fb.view(np.float64) - 1.0) * (hi - lo) + lo)

def data(original):
    k = jr.key(42)
    n = 2500
    if original:
        k, sub = jr.split_original(k)
        bx, bn = jr.random_bits64_original(k, n), jr.random_bits64_original(sub, n)
    else:
        k, sub = jr.split(k)
        bx, bn = jr.random_bits64(k, (n,)), jr.random_bits64(sub, (n,))
    x = uniform_from_bits(bx, -3.0, 3.0).reshape(-1, 1)
    f = lambda x: np.sin(2 * x) + x * np.cos(5 * x)
    lo = np.nextafter(np.float64(-1.0), 0.0)
    y = f(x) + (np.sqrt(2) * erfinv(uniform_from_bits(bn, lo, 1.0))).reshape(-1, 1) * 0.5
    return x, y

def run(original, train_mean, iters=500, lr=1e-2, wd=1e-4):
    x, y = data(original)
    z = np.linspace(-3.0, 3.0, 50).reshape(-1, 1)
    u = np.concatenate([o.softplus_inv(np.ones(3)), [0.0], z.ravel()])
    m = np.zeros_like(u); v = np.zeros_like(u)
    hist = []
    for t in range(1, iters + 1):
        ell, var, sn = o.softplus(u[:3])
        val, g = o.collapsed_elbo_value_and_grad_autodiff("rbf", x, y, u[4:].reshape(-1, 1), ell, var, sn, u[3])
        gu = np.concatenate([np.array([np.sum(g["lengthscale"]), g["variance"], g["obs_stddev"]]) / (1 + np.exp(-u[:3])),
                             [g["mean_const"] if train_mean else 0.0], np.asarray(g["inducing_inputs"]).ravel()])
        gu = -gu
        hist.append(-val)
        m = 0.9 * m + 0.1 * gu; v = 0.999 * v + 0.001 * gu * gu
        mh = m / (1 - 0.9 ** t); vh = v / (1 - 0.999 ** t)
        upd = mh / (np.sqrt(vh) + 1e-8) + wd * u
        if not train_mean: upd[3] = 0.0
        u = u - lr * upd
    return hist, u, x, y

for original in (True, False):
    for train_mean in (True, False):
        t0 = time.time()
        hist, u, x, y = run(original, train_mean)
        print('original stream' if original else 'partitionable', 'train_mean' if train_mean else 'fixed_mean', 'history[0]=%.6f history[-1]=%.7f (golden 1924.7634809)  %.0f s' % (hist[0], hist[-1], time.time() - t0), flush=True)

print("---- variants (original stream)")
def run2(original=True, wd=1e-4, noise_as_variance=False, train_z=True, train_mean=True, iters=500, lr=1e-2):
    x, y = data(original)
    z = np.linspace(-3.0, 3.0, 50).reshape(-1, 1)
    u = np.concatenate([o.softplus_inv(np.ones(3)), [0.0], z.ravel()])
    m = np.zeros_like(u); v = np.zeros_like(u); hist = []
    for t in range(1, iters + 1):
        ell, var, s3 = o.softplus(u[:3])
        sn = np.sqrt(s3) if noise_as_variance else s3
        val, g = o.collapsed_elbo_value_and_grad_autodiff("rbf", x, y, u[4:].reshape(-1, 1), ell, var, sn, u[3])
        gsn = g["obs_stddev"] * (0.5 / sn if noise_as_variance else 1.0)
        gu = -np.concatenate([np.array([np.sum(g["lengthscale"]), g["variance"], gsn]) / (1 + np.exp(-u[:3])),
                              [g["mean_const"] if train_mean else 0.0], np.asarray(g["inducing_inputs"]).ravel() * (1.0 if train_z else 0.0)])
        hist.append(-val)
        m = 0.9 * m + 0.1 * gu; v = 0.999 * v + 0.001 * gu * gu
        upd = (m / (1 - 0.9 ** t)) / (np.sqrt(v / (1 - 0.999 ** t)) + 1e-8) + wd * u
        if not train_mean: upd[3] = 0
        if not train_z: upd[4:] = 0
        u = u - lr * upd
    return hist
for kw in [dict(wd=0.0), dict(noise_as_variance=True), dict(noise_as_variance=True, wd=0.0), dict(train_z=False), dict(train_z=False, noise_as_variance=True),
           dict(train_z=False, wd=0.0, noise_as_variance=True), dict(noise_as_variance=True, train_mean=False), dict(original=False, noise_as_variance=True)]:
    h = run2(**kw); print(kw, 'history[-1]=%.7f' % h[-1], flush=True)

print("---- predictive check")
x, y = data(True)
hist, u, x, y = run(True, True)
ell, var, sn = o.softplus(u[:3])
xt = np.linspace(-3.1, 3.1, 500).reshape(-1, 1)
mean, cov = o.collapsed_predict("rbf", x, y, xt, u[4:].reshape(-1,1), ell, var, sn, u[3])
print('sum mean %.6f (golden -8.39869652)  sum std %.6f (golden 255.74838027)' % (mean.sum(), np.sqrt(np.diag(cov) + sn**2).sum()), 'sn', sn, 'ell', ell, 'var', var)
x2, y2 = data(False)
print('partitionable stream: y sum', y2.sum(), ' original: y sum', y.sum())
