"""Bitwise run-to-run reproducibility of conjugate_mll value + gradient on one device (and across devices when there are two)."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpjax_b200 import ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4500
rng = np.random.default_rng(77)
X = rng.uniform(-2, 2, (n, 8)); y = np.sin(X[:, :1]) + 0.1 * rng.standard_normal((n, 1))
ell = np.linspace(0.8, 1.6, 8)
def run(idx):
    with torch.cuda.device(idx):
        d = f"cuda:{idx}"
        t = lambda a: torch.as_tensor(np.asarray(a, np.float64), device=d)
        p = [t(ell).requires_grad_(True), t(1.0).requires_grad_(True), t(0.3).requires_grad_(True)]
        v = ops.conjugate_mll_fused(2, t(X), t(y), p[0], p[1], p[2], None, 1e-6)
        v.backward()
        torch.cuda.synchronize()
        return v.item(), np.concatenate([q.grad.cpu().numpy().reshape(-1) for q in p])
res = [run(0) for _ in range(4)]
out = {"n": n, "red": os.environ.get("GPB_OZ_RED", "1"), "values_equal": all(r[0] == res[0][0] for r in res),
       "grads_equal": all(np.array_equal(r[1], res[0][1]) for r in res),
       "max_grad_rel_diff": float(max(np.max(np.abs(r[1] - res[0][1]) / np.abs(res[0][1])) for r in res))}
if torch.cuda.device_count() > 1:
    r1 = run(1)
    out["dev1_value_equal"] = r1[0] == res[0][0]
    out["dev1_grad_rel_diff"] = float(np.max(np.abs(r1[1] - res[0][1]) / np.abs(res[0][1])))
print(json.dumps(out))
