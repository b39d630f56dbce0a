#!/usr/bin/env python
"""Generates tests/golden/conditioning_sweep.json: the CPU oracle's conjugate_mll value + gradient (reference formulation:
LU slogdet / solve, gpjax/objectives.py:93-107 + gpjax/linalg/operations.py:109-111,163-165) over a conditioning sweep at
N = 8192, D = 8, so that the GPU test (tests/test_gpu_conditioning.py) does not spend ~45 s of host time per cell on the GPU box.

    python tests/golden/make_conditioning_fixture.py          # ~10 min on 16 cores

Inputs are NOT stored: the test regenerates them from the same NumPy PCG64 seed (`make_inputs`).  Per cell the fixture holds
cond_2(Sigma) (eigvalsh), the LU value, the Cholesky value (the oracle's own float64 noise floor) and the closed-form gradients.
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle as o  # noqa: E402

N, D, SEED = 8192, 8, 8192
KERNELS = ("rbf", "matern52")
LENGTHSCALES = {"ard0.8-1.6": np.linspace(0.8, 1.6, D), "3.0": np.full(D, 3.0)}
OBS_STDDEVS = (0.3, 0.03, 0.003)


def make_inputs(n=N, d=D, seed=SEED):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-2.0, 2.0, (n, d))
    y = np.sin(X[:, :1]) + 0.1 * rng.standard_normal((n, 1))
    return X, y


def main():
    X, y = make_inputs()
    cells = []
    for name in KERNELS:
        for tag, ell in LENGTHSCALES.items():
            for sn in OBS_STDDEVS:
                t0 = time.time()
                S = o.gp_oracle._sigma(name, X, ell, 1.0, sn, 1e-6)
                ev = np.linalg.eigvalsh(S)
                g = o.conjugate_mll_grad_closed_form(name, X, y, ell, 1.0, sn, 0.0)
                cells.append(dict(kernel=name, lengthscale=tag, obs_stddev=sn, cond=float(ev[-1] / ev[0]),
                                  value_lu=o.conjugate_mll(name, X, y, ell, 1.0, sn, 0.0),
                                  value_chol=o.conjugate_mll_chol(name, X, y, ell, 1.0, sn, 0.0),
                                  grad=dict(lengthscale=[float(v) for v in g["lengthscale"]], variance=g["variance"],
                                            obs_stddev=g["obs_stddev"], mean_const=g["mean_const"])))
                print(name, tag, sn, f"cond {cells[-1]['cond']:.2e}", f"{time.time() - t0:.0f} s", flush=True)
    out = dict(n=N, d=D, seed=SEED, variance=1.0, mean_const=0.0, jitter=1e-6,
               x_checksum=float(X.sum()), y_checksum=float(y.sum()), cells=cells)
    with open(os.path.join(HERE, "conditioning_sweep.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
