#!/usr/bin/env python
"""Errors of conjugate_mll (value + gradients) against the committed oracle fixture tests/golden/conditioning_sweep.json for every
cell of the conditioning sweep and every arithmetic (planes 0 = FP64 DMMA, 6, 7, auto).  Unlike scripts/cond_sweep.py it does not
recompute the oracle (45 s per cell), so it is cheap enough to run once per build-time switch:

    GPB_OZ_PANELS=0 python scripts/cond_sweep_fixture.py > gpurun_out/sweep_panels0.jsonl
    GPB_OZ_PANELS=1 python scripts/cond_sweep_fixture.py > gpurun_out/sweep_panels1.jsonl
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_conditioning_fixture import make_inputs  # noqa: E402

from gpjax_b200 import ops  # noqa: E402

FIX = json.load(open(os.path.join(ROOT, "tests", "golden", "conditioning_sweep.json")))
KEYS = ("lengthscale", "variance", "obs_stddev", "mean_const")
KIND = {"rbf": 0, "matern52": 2}


def main():
    X, y = make_inputs(FIX["n"], FIX["d"], FIX["seed"])
    X, y = torch.as_tensor(X, device="cuda"), torch.as_tensor(y, device="cuda")
    dev = lambda a: torch.as_tensor(np.asarray(a, np.float64), device="cuda")
    for cell in FIX["cells"]:
        ell = np.linspace(0.8, 1.6, FIX["d"]) if cell["lengthscale"] == "ard0.8-1.6" else np.full(FIX["d"], float(cell["lengthscale"]))
        row = dict(kernel=cell["kernel"], lengthscale=cell["lengthscale"], obs_stddev=cell["obs_stddev"], cond=cell["cond"],
                   panels=os.environ.get("GPB_OZ_PANELS", "default"), paths={})
        for planes in (0, 6, 7, ops.OZAKI_AUTO):
            ops.set_ozaki_slices(planes)
            p = [dev(ell).requires_grad_(True), dev(FIX["variance"]).requires_grad_(True), dev(cell["obs_stddev"]).requires_grad_(True),
                 dev(FIX["mean_const"]).requires_grad_(True)]
            val = ops.conjugate_mll_fused(KIND[cell["kernel"]], X, y, p[0], p[1], p[2], p[3], FIX["jitter"])
            val.backward()
            g = dict(zip(KEYS, (p[0].grad.cpu().numpy(), p[1].grad.item(), p[2].grad.item(), p[3].grad.item())))
            errs = {"value": abs(val.item() - cell["value_lu"]) / abs(cell["value_lu"])}
            for k in KEYS:
                a, b = np.asarray(g[k]), np.asarray(cell["grad"][k])
                errs[k] = float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
            row["paths"]["auto" if planes == ops.OZAKI_AUTO else str(planes)] = errs
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
