#!/usr/bin/env python
"""CUDA collapsed_elbo value / gradient against the 40-digit adjudicator fixture (tests/golden/sgpr_adjudicator.json), both
statistics routes, next to the oracle's two float64 routes: the record behind DESIGN section 5's accuracy statement."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import oracle as o
from make_sgpr_adjudicator_fixture import make_inputs
from gpjax_b200.sgpr_ops import collapsed_elbo_fused

F = json.load(open(os.path.join(ROOT, "tests", "golden", "sgpr_adjudicator.json")))
X, y, Z = make_inputs(); h = F["hyper"]
dev = lambda a: torch.as_tensor(np.asarray(a, np.float64), device="cuda")
rel = lambda a, b: float(np.max(np.abs(np.asarray(a).reshape(-1) - np.asarray(b).reshape(-1))) / np.max(np.abs(b)))
out = {"cond_kzz": F["cond_kzz"], "paths": {}}
args = ("rbf", X, y, Z, np.array([h["lengthscale"]]), h["variance"], h["obs_stddev"], h["mean_const"])
va, ga = o.collapsed_elbo_value_and_grad_autodiff(*args)
gc = o.collapsed_elbo_grad_closed_form(*args); gc = gc[1] if isinstance(gc, tuple) else gc
out["paths"]["oracle_autodiff_reference_order"] = {"value": abs(va - F["value"]) / abs(F["value"]), **{k: rel(ga[k], b) for k, b in F["grad"].items()}}
out["paths"]["oracle_two_pass_closed_form"] = {k: rel(gc[k], b) for k, b in F["grad"].items()}
for route in ("whitened", "raw"):
    p = [dev(Z).requires_grad_(True), dev(np.array([h["lengthscale"]])).requires_grad_(True), dev(h["variance"]).requires_grad_(True),
         dev(h["obs_stddev"]).requires_grad_(True), dev(h["mean_const"]).requires_grad_(True)]
    v = collapsed_elbo_fused(0, dev(X), dev(y), p[0], p[1], p[2], p[3], p[4], h["jitter"], 64, None, route)
    v.backward()
    got = dict(inducing_inputs=p[0].grad, lengthscale=p[1].grad, variance=p[2].grad, obs_stddev=p[3].grad, mean_const=p[4].grad)
    out["paths"][f"cuda_{route}"] = {"value": abs(v.item() - F["value"]) / abs(F["value"]),
                                     **{k: rel(got[k].cpu().numpy(), b) for k, b in F["grad"].items()}}
print(json.dumps(out, indent=1))
