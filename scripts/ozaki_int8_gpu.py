"""GPU feasibility measurement for DESIGN section 11.1 (NOT a product path; analysis only).

Question: how fast does the int8 tensor pipe of THIS B200 run, and what FP64-equivalent rate does an
Ozaki-scheme rank-k update reach on it?  The int8 products here go through the library (`torch._int_mm`,
i.e. cuBLASLt IGEMM) purely to measure the pipe before a hand-written tcgen05 `kind::i8` kernel is attempted;
nothing in `gpjax_b200/` imports this file.

Slicing follows scripts/ozaki_prototype.py (7 magnitude bits per signed-int8 slice after power-of-two row
scaling).  Slice pairs of equal total order p+q=t share one scale, so they are accumulated *inside* one integer
GEMM by concatenating along K:  [A_0|A_1|..|A_t] . [B_t|..|B_0]^T  (exact in int32 while (t+1) k 64^2 < 2^31).
That turns s(s+1)/2 products into s GEMMs with growing K and one int32 -> fp64 recombination each.

    python scripts/ozaki_int8_gpu.py > gpurun_out/ozaki_int8.json
"""
import json
import sys
import time

import torch

BETA = 7


def ev_time(fn, reps=5, warm=2):
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


def split_cat(A, s, reverse):
    """A [m,k] fp64 -> int8 [m, s*k] (slice p at columns p*k.., or reversed order) and row exponents."""
    m, k = A.shape
    mx = A.abs().amax(dim=1)
    e = torch.where(mx > 0, torch.ceil(torch.log2(torch.where(mx > 0, mx, torch.ones_like(mx)))) + 1, torch.zeros_like(mx))
    R = A * torch.exp2(-e)[:, None]
    out = torch.empty(m, s * k, dtype=torch.int8, device=A.device)
    for p in range(s):
        R = R * 2.0**BETA
        S = torch.round(R)
        q = (s - 1 - p) if reverse else p
        out[:, q * k:(q + 1) * k] = S.to(torch.int8)
        R = R - S
    return out, e


def ozaki_abt(A, B, s, contiguous_views):
    Ac, ea = split_cat(A, s, reverse=False)
    Bc, eb = split_cat(B, s, reverse=True)
    k = A.shape[1]
    acc = torch.zeros(A.shape[0], B.shape[0], dtype=torch.float64, device=A.device)
    t_gemm = 0.0
    for t in range(s):
        a = Ac[:, :(t + 1) * k]
        b = Bc[:, (s - 1 - t) * k:]
        if contiguous_views:
            a, b = a.contiguous(), b.contiguous()
        torch.cuda.synchronize()
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        P = torch._int_mm(a, b.t())
        t1.record()
        torch.cuda.synchronize()
        t_gemm += t0.elapsed_time(t1) * 1e-3
        acc += P.to(torch.float64) * 2.0 ** (-BETA * (t + 2))
        del P
    return acc * torch.exp2(ea)[:, None] * torch.exp2(eb)[None, :], t_gemm


def main():
    out = {"device": torch.cuda.get_device_name(0), "int_mm": [], "ozaki": []}
    dev = "cuda"
    # 1. raw int8 pipe rate through the library
    for (m, n, k) in ((8192, 8192, 8192), (16384, 16384, 16384), (32768, 32768, 1024), (32768, 32768, 4096), (32768, 32768, 7168)):
        try:
            a = torch.randint(-64, 64, (m, k), dtype=torch.int8, device=dev)
            b = torch.randint(-64, 64, (n, k), dtype=torch.int8, device=dev)
            t = ev_time(lambda: torch._int_mm(a, b.t()))
            out["int_mm"].append({"m": m, "n": n, "k": k, "s": t, "Pop_s": 2.0 * m * n * k / t / 1e15})
            del a, b
        except Exception as ex:  # noqa: BLE001
            out["int_mm"].append({"m": m, "n": n, "k": k, "error": repr(ex)[:300]})
    # fp64 / bf16 library rates for the same box, same clocks
    try:
        a = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
        t = ev_time(lambda: a @ a.t(), reps=3)
        out["dgemm_8192_TF_s"] = 2.0 * 8192**3 / t / 1e12
        a = a.to(torch.bfloat16)
        t = ev_time(lambda: a @ a.t(), reps=10)
        out["bf16_8192_TF_s"] = 2.0 * 8192**3 / t / 1e12
        del a
    except Exception as ex:  # noqa: BLE001
        out["lib_rates_error"] = repr(ex)[:300]
    # 2. Ozaki-scheme rank-k update at the shape of the potrf trailing update (m x m, k = NB = 1024)
    torch.manual_seed(0)
    for (m, k) in ((16384, 1024), (32768, 1024)):
        # operands shaped like a scaled Cholesky panel: entries of widely varying magnitude per row
        A = torch.randn(m, k, dtype=torch.float64, device=dev) * torch.exp(torch.randn(m, 1, dtype=torch.float64, device=dev))
        A = A * torch.exp(2.0 * torch.randn(1, k, dtype=torch.float64, device=dev))  # column spread (hurts row scaling)
        ref = A @ A.t()
        t_dgemm = ev_time(lambda: A @ A.t(), reps=2, warm=1)
        scale = (A.norm(dim=1)[:, None] * A.norm(dim=1)[None, :])  # Cauchy-Schwarz bound: natural error scale
        for s in (5, 6, 7, 8):
            for contig in (False, True):
                try:
                    C, t_g = ozaki_abt(A, A, s, contig)
                    err = ((C - ref).abs() / scale).max().item()
                    out["ozaki"].append({"m": m, "k": k, "slices": s, "contiguous_operands": contig,
                                          "int8_gemm_s": t_g, "fp64_equiv_TF_s_gemm_only": 2.0 * m * m * k / t_g / 1e12,
                                          "dgemm_s": t_dgemm, "dgemm_TF_s": 2.0 * m * m * k / t_dgemm / 1e12,
                                          "max_err_over_rownorm_product": err})
                    del C
                except Exception as ex:  # noqa: BLE001
                    out["ozaki"].append({"m": m, "k": k, "slices": s, "contiguous_operands": contig, "error": repr(ex)[:300]})
        del A, ref, scale
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    t0 = time.time()
    main()
    print(f"# wall {time.time() - t0:.1f} s", file=sys.stderr)
