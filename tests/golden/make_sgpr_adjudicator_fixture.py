#!/usr/bin/env python
"""40-digit adjudicator for the collapsed ELBO and ITS GRADIENT in the ill-conditioned regime of the reference's own example
(examples/collapsed_vi.py: 1-D inputs, equally spaced inducing points, RBF -> cond(Kzz + jitter I) ~ 1e7).

Two float64 evaluation orders of the bound (the oracle's reverse-mode autodiff, its two-pass closed form, the CUDA path) drift
apart like cond * eps there; the SGPR tests therefore scale their gradient tolerance with the condition number.  This script
settles WHICH side is right: the value follows gpjax/objectives.py:342-416 in mpmath at 40 significant digits and the gradient is
a central difference of that value with step 1e-12 (truncation ~1e-24, round-off ~1e-40 * cond / 1e-12: both negligible).

    python tests/golden/make_sgpr_adjudicator_fixture.py        # ~1 min, writes tests/golden/sgpr_adjudicator.json
"""
import json
import os

import mpmath as mp
import numpy as np

mp.mp.dps = 40
HERE = os.path.dirname(os.path.abspath(__file__))


def make_inputs():
    rng = np.random.default_rng(123)
    n, m = 150, 24
    X = np.sort(rng.uniform(-3.0, 3.0, (n, 1)), axis=0)
    y = np.sin(2.0 * X) + 0.2 * rng.standard_normal((n, 1))
    Z = np.linspace(-3.0, 3.0, m).reshape(-1, 1)
    return X, y, Z


def fwd_solve(L, B):
    """L^-1 B by forward substitution (L lower triangular), working precision."""
    n, k = L.rows, B.cols
    X = mp.matrix(n, k)
    for j in range(k):
        for i in range(n):
            s = B[i, j]
            for t in range(i):
                s -= L[i, t] * X[t, j]
            X[i, j] = s / L[i, i]
    return X


def elbo(X, y, Z, ell, var, sn, c, jitter):
    """objectives.py:342-416; all arguments mp numbers / lists of mp numbers."""
    n, m = len(X), len(Z)
    k = lambda a, b: var * mp.exp(-mp.mpf("0.5") * ((a - b) / ell) ** 2)
    Kzz = mp.matrix(m, m)
    for i in range(m):
        for j in range(m):
            Kzz[i, j] = k(Z[i], Z[j]) + (jitter if i == j else 0)
    Kzx = mp.matrix(m, n)
    for i in range(m):
        for j in range(n):
            Kzx[i, j] = k(Z[i], X[j])
    Lz = mp.cholesky(Kzz)
    A = fwd_solve(Lz, Kzx) / sn
    AAT = A * A.T
    B = AAT + mp.eye(m)
    L = mp.cholesky(B)
    diff = mp.matrix([yi - c for yi in y])
    v = fwd_solve(L, A * diff)
    noise = sn * sn
    quad = (sum(diff[i] * diff[i] for i in range(n)) - sum(v[i] * v[i] for i in range(m))) / noise
    log_det_B = 2 * sum(mp.log(L[i, i]) for i in range(m))
    two_log_prob = -n * mp.log(2 * mp.pi * noise) - log_det_B - quad
    two_trace = n * var / noise - sum(AAT[i, i] for i in range(m))
    return (two_log_prob - two_trace) / 2


def main():
    X, y, Z = make_inputs()
    hyper = dict(lengthscale=1.0, variance=1.0, obs_stddev=0.2, mean_const=0.1, jitter=1e-6)
    Xm, ym, Zm = [mp.mpf(float(v)) for v in X[:, 0]], [mp.mpf(float(v)) for v in y[:, 0]], [mp.mpf(float(v)) for v in Z[:, 0]]
    base = [mp.mpf(hyper[k]) for k in ("lengthscale", "variance", "obs_stddev", "mean_const")]
    jit = mp.mpf(hyper["jitter"])
    f = lambda th, Zz: elbo(Xm, ym, Zz, th[0], th[1], th[2], th[3], jit)
    val = f(base, Zm)
    h = mp.mpf("1e-12")
    grad = {}
    for i, name in enumerate(("lengthscale", "variance", "obs_stddev", "mean_const")):
        up, dn = list(base), list(base)
        up[i] += h
        dn[i] -= h
        grad[name] = float((f(up, Zm) - f(dn, Zm)) / (2 * h))
    gz = []
    for a in range(len(Zm)):
        up, dn = list(Zm), list(Zm)
        up[a] += h
        dn[a] -= h
        gz.append(float((f(base, up) - f(base, dn)) / (2 * h)))
    grad["inducing_inputs"] = gz
    Kzz = np.exp(-0.5 * (Z - Z.T) ** 2) + 1e-6 * np.eye(len(Z))
    out = dict(n=len(Xm), m=len(Zm), seed=123, hyper=hyper, cond_kzz=float(np.linalg.cond(Kzz)), digits=mp.mp.dps,
               fd_step=1e-12, x_checksum=float(X.sum()), y_checksum=float(y.sum()), value=float(val), grad=grad)
    with open(os.path.join(HERE, "sgpr_adjudicator.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps({k: out[k] for k in ("n", "m", "cond_kzz", "value")}))


if __name__ == "__main__":
    main()
