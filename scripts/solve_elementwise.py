#!/usr/bin/env python
"""Element-wise accuracy of the blocked triangular solves after a factorisation on each arithmetic (planes 0 / 6 / 7 radix-256), at
the sizes of tests/test_gpu_primitives.py::test_potrf_trsv_trsm_logdet_potri that cross the int8 threshold."""
import json
import os
import sys

import numpy as np
import scipy.linalg as sla
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as o  # noqa: E402
from gpjax_b200 import ops  # noqa: E402

dev = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda")
rel = lambda a, r: float(np.max(np.abs(a - r) / np.maximum(np.abs(r), 1e-300)))
for n in (3200, 5000):
    rng = np.random.default_rng(n)
    X = rng.uniform(-2, 2, (n, 4))
    S = o.gram("rbf", X, np.linspace(0.8, 1.4, 4), 1.0) + 0.09 * np.eye(n)
    Lref = np.linalg.cholesky(S)
    b = np.random.default_rng(n).standard_normal(n)
    xr = sla.solve_triangular(Lref, b, lower=True)
    xtr = sla.solve_triangular(Lref.T, b, lower=False)
    for planes in (0, 6, 7):
        ops.set_ozaki_slices(planes)
        A = dev(S)
        ws = ops.FactorWorkspace(n, 1, potri=True, device="cuda")
        ops.potrf_lower_(A, ws, zero_upper=True)
        L = A.cpu().numpy()
        x = ops.trsv_lower_(A, dev(b), ws).cpu().numpy()
        xt = ops.trsv_lower_(A, dev(b), ws, trans=True).cpu().numpy()
        print(json.dumps(dict(n=n, planes=planes, L_max_abs=float(np.max(np.abs(L - Lref)) / np.abs(Lref).max()),
                              trsv_rel=rel(x, xr), trsv_t_rel=rel(xt, xtr),
                              trsv_norm=float(np.max(np.abs(x - xr)) / np.max(np.abs(xr))))), flush=True)
