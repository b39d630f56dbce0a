#!/usr/bin/env python
"""Conditioning sweep of conjugate_mll (value + gradient) against the CPU oracle: how far may the int8 digit-plane
trailing updates go before they leave the 1e-8 contract?

    python scripts/cond_sweep.py [N] > gpurun_out/cond_sweep.json

For every (kernel, obs_stddev, lengthscale) cell it records cond_2(Sigma) (eigvalsh), the oracle's own noise floor
(LU value vs Cholesky value -- two float64 routes to the same number) and, for each arithmetic of the CUDA path
(`planes` = 0: FP64 DMMA everywhere, 5 / 6 / 7: int8 radix-256 digit planes, "auto": the library default), the relative error of the
value and of every gradient.  Measurement script: its output under profiles/ is what tests/test_gpu_conditioning.py
and DESIGN section 5 quote.
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as o  # noqa: E402
from gpjax_b200 import ops  # noqa: E402


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda")


def gpu_eval(kind, X, y, ell, var, sn, planes):
    ops.set_ozaki_slices(planes)
    p = [dev(ell).requires_grad_(True), dev(var).requires_grad_(True), dev(sn).requires_grad_(True),
         dev(0.0).requires_grad_(True)]
    val = ops.conjugate_mll_fused(kind, dev(X), dev(y), p[0], p[1], p[2], p[3], 1e-6)
    val.backward()
    return val.item(), dict(lengthscale=p[0].grad.cpu().numpy(), variance=p[1].grad.item(), obs_stddev=p[2].grad.item(),
                            mean_const=p[3].grad.item())


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    plane_list = [0, 5, 6, 7]  # radix-256 digit planes: 40 / 48 / 56 bits below the row maximum
    d = 8
    rng = np.random.default_rng(8192)
    X = rng.uniform(-2.0, 2.0, (n, d))
    y = np.sin(X[:, :1]) + 0.1 * rng.standard_normal((n, 1))
    rows = []
    for kind, name in ((0, "rbf"), (2, "matern52")):
        for ell_tag, ell in (("ard0.8-1.6", np.linspace(0.8, 1.6, d)), ("3.0", np.full(d, 3.0))):
            for sn in (0.3, 0.03, 0.003):
                t0 = time.time()
                S = o.gp_oracle._sigma(name, X, ell, 1.0, sn, 1e-6)
                ev = np.linalg.eigvalsh(S)
                cond = float(ev[-1] / ev[0])
                vref = o.conjugate_mll(name, X, y, ell, 1.0, sn, 0.0)
                vchol = o.conjugate_mll_chol(name, X, y, ell, 1.0, sn, 0.0)
                gref = o.conjugate_mll_grad_closed_form(name, X, y, ell, 1.0, sn, 0.0)
                row = dict(kernel=name, lengthscale=ell_tag, obs_stddev=sn, n=n, cond=cond, value=vref,
                           oracle_lu_vs_chol=abs(vref - vchol) / abs(vref), oracle_s=time.time() - t0, paths={})
                base = None
                for planes in plane_list:
                    val, g = gpu_eval(kind, X, y, ell, 1.0, sn, planes)
                    errs = dict(value=abs(val - vref) / abs(vref))
                    for k in ("lengthscale", "variance", "obs_stddev", "mean_const"):
                        a, b = np.asarray(g[k]), np.asarray(gref[k])
                        errs[k] = float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
                    if planes == 0:
                        base = (val, g)
                    else:  # against the DMMA path: the pure arithmetic difference, no oracle noise
                        errs["value_vs_dmma"] = abs(val - base[0]) / abs(base[0])
                        errs["grad_vs_dmma"] = max(
                            float(np.max(np.abs(np.asarray(g[k]) - np.asarray(base[1][k]))) /
                                  max(np.max(np.abs(np.asarray(base[1][k]))), 1e-300))
                            for k in ("lengthscale", "variance", "obs_stddev", "mean_const"))
                    row["paths"][str(planes)] = errs
                rows.append(row)
                print(json.dumps(row), flush=True)
                ops.release_buffers()
    keys = ("value", "lengthscale", "variance", "obs_stddev", "mean_const")
    worst = {str(p): max(max(r["paths"][str(p)][k] for k in keys) for r in rows) for p in plane_list}
    print(json.dumps({"summary_worst_rel_error": worst}), flush=True)


if __name__ == "__main__":
    main()
