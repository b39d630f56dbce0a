"""Timing of the hand-written tcgen05 kind::i8 path (csrc/ozaki_i8.cu) against the FP64 DMMA path.  Analysis script.

    python scripts/ozaki_bench.py [N_for_potrf] > gpurun_out/ozaki_bench.json
"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from gpjax_b200 import ops  # noqa: E402


def ev(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3


def main():
    out = {"raw": [], "update": [], "potrf": []}
    for (m, n, k) in ((8192, 8192, 8192), (16384, 16384, 4096), (32768, 32768, 1024)):
        A = torch.randint(-128, 128, (m, k), dtype=torch.int8, device="cuda")
        B = torch.randint(-128, 128, (n, k), dtype=torch.int8, device="cuda")
        t = ev(lambda: ops.igemm_i8(A, B))
        t_lib = ev(lambda: torch._int_mm(A, B.t()))
        out["raw"].append({"m": m, "n": n, "k": k, "ours_Pop_s": 2.0 * m * n * k / t / 1e15, "cublaslt_Pop_s": 2.0 * m * n * k / t_lib / 1e15})
        del A, B
    m, k = 32768, 1024
    X = torch.randn(m, k, dtype=torch.float64, device="cuda") * 0.05
    C = torch.zeros(m, m, dtype=torch.float64, device="cuda")
    t_d = ev(lambda: ops.gemm(X, X, C, alpha=-1.0, beta=1.0, mask=1), reps=2)
    for s in (4, 5, 6, 7):
        t_s = ev(lambda: ops.ozaki_slice(X, s), reps=3)
        Q, sc = ops.ozaki_slice(X, s)
        t_g = ev(lambda: ops.ozaki_gemm_(C, Q, sc, Q, sc, k, s, alpha=-1.0, mask_lower=True), reps=2)
        flop = m * m * k  # lower half of 2 m^2 k
        out["update"].append({"m": m, "k": k, "slices": s, "slice_s": t_s, "ozaki_gemm_s": t_g, "dmma_s": t_d,
                              "fp64_equiv_TF_s": flop / t_g / 1e12, "dmma_TF_s": flop / t_d / 1e12,
                              "int8_Pop_s": flop * s * (s + 1) / 2 / t_g / 1e15})
        del Q, sc
    del X, C
    import os
    if os.environ.get("OZ_BENCH_QUICK"):
        json.dump(out, sys.stdout, indent=1)
        return
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
    rng = np.random.default_rng(0)
    Xd = torch.as_tensor(rng.uniform(-2, 2, (n, 8)), device="cuda")
    ell = torch.as_tensor(np.linspace(0.8, 1.6, 8), device="cuda")
    one = torch.tensor(1.0, dtype=torch.float64, device="cuda")
    ws = ops.FactorWorkspace(n, 8, potri=False, device="cuda")
    A = torch.empty(n, n, dtype=torch.float64, device="cuda")
    L0 = None
    for s in (0, 6, 7):
        ops.set_ozaki_slices(s)

        def run():
            ops.gram_forward(0, Xd, Xd, ell, one, diag_add=1e-6 + 0.09, lower_only=True, out=A)
            ops.potrf_lower_(A, ws, zero_upper=False)

        t = ev(run, reps=2)
        t_gram = ev(lambda: ops.gram_forward(0, Xd, Xd, ell, one, diag_add=1e-6 + 0.09, lower_only=True, out=A), reps=2)
        run()
        torch.cuda.synchronize()
        row = {"N": n, "slices": s, "potrf_s": t - t_gram, "TF_s_fp64_equiv": n**3 / 3 / (t - t_gram) / 1e12}
        if s == 0:
            L0 = torch.tril(A).clone() if n <= 40000 else None
        elif L0 is not None:
            row["max_rel_diff_vs_dmma"] = float((torch.tril(A) - L0).abs().max() / L0.abs().max())
        out["potrf"].append(row)
    ops.set_ozaki_slices(0)
    del A, ws, L0
    torch.cuda.empty_cache()
    # whole conjugate_mll value + gradient step (potrf + trtri + lauum + streamed backward)
    y = torch.as_tensor(np.sin(rng.uniform(-2, 2, (n, 1))), device="cuda")
    base = None
    out["mll_step"] = []
    for s in (0, 6, 7):
        ops.set_ozaki_slices(s)
        p = [ell.clone().requires_grad_(True), one.clone().requires_grad_(True),
             torch.tensor(0.3, dtype=torch.float64, device="cuda", requires_grad=True),
             torch.tensor(0.0, dtype=torch.float64, device="cuda", requires_grad=True)]

        def step():
            for q in p:
                q.grad = None
            v = ops.conjugate_mll_fused(0, Xd, y, p[0], p[1], p[2], p[3], 1e-6)
            v.backward()
            return v

        t = ev(step, reps=2)
        v = step()
        torch.cuda.synchronize()
        g = torch.cat([q.grad.reshape(-1) for q in p])
        row = {"N": n, "slices": s, "step_s": t, "TF_s_fp64_equiv": n**3 / t / 1e12, "value": v.item()}
        if s == 0:
            base = (v.item(), g.clone())
        else:
            row["value_rel_diff_vs_dmma"] = abs(v.item() - base[0]) / abs(base[0])
            row["grad_rel_diff_vs_dmma"] = float((g - base[1]).abs().max() / base[1].abs().max())
        out["mll_step"].append(row)
    ops.set_ozaki_slices(0)
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
