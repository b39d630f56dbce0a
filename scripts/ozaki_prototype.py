"""CPU feasibility study for DESIGN section 11.1: blocked Cholesky whose rank-NB trailing updates run as an
Ozaki-scheme product (operands split into `s` slices of `beta`-bit integers after power-of-two row scaling; every
slice-pair product is an exact integer GEMM -- int8 x int8 -> int32 on tcgen05, emulated here with int64 matmul).
Reports the error of the exact-GP MLL ingredients against the native float64 factorisation.

    python scripts/ozaki_prototype.py [N] [NB]
"""
import sys

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import oracle as o  # noqa: E402  (test/analysis infrastructure only)

BETA = 7  # magnitude bits of a signed int8 slice


def split(A, s):
    """A[m,k] -> (slices [s][m,k] of integers in [-127,127], row exponents e[m]) with A ~= 2^e * sum_p S_p 2^(-BETA (p+1))."""
    mx = np.max(np.abs(A), axis=1)
    e = np.where(mx > 0, np.ceil(np.log2(np.where(mx > 0, mx, 1.0))) + 1, 0.0)  # |A| * 2^-e < 1/2
    R = A * np.exp2(-e)[:, None]
    out = []
    for _ in range(s):
        R = R * 2.0**BETA
        S = np.round(R)  # |S| <= 64 after the first step, <= 64 afterwards because |R - S| <= 1/2
        out.append(S.astype(np.int64))
        R = R - S
    return out, e


def ozaki_abt(A, B, s):
    """A @ B.T with all slice pairs of total order p + q < s (the s (s+1) / 2 most significant products)."""
    Sa, ea = split(A, s)
    Sb, eb = split(B, s)
    acc = np.zeros((A.shape[0], B.shape[0]))
    for p in range(s):
        for q in range(s - p):
            P = Sa[p] @ Sb[q].T  # exact: |entries| <= k * 64 * 64 < 2^31 for k < 2^19
            acc += P.astype(np.float64) * 2.0 ** (-BETA * (p + q + 2))
    return acc * np.exp2(ea)[:, None] * np.exp2(eb)[None, :]


def blocked_cholesky(S, nb, gemm):
    A = S.copy()
    n = A.shape[0]
    for k in range(0, n, nb):
        e = min(k + nb, n)
        A[k:e, k:e] = np.linalg.cholesky(A[k:e, k:e])
        if e < n:
            A[e:, k:e] = np.linalg.solve(A[k:e, k:e], A[e:, k:e].T).T
            A[e:, e:] -= gemm(A[e:, k:e], A[e:, k:e])
    return np.tril(A)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1536
    nb = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    rng = np.random.default_rng(0)
    X = rng.uniform(-2, 2, (n, 8))
    y = np.sin(X[:, 0]) + 0.1 * rng.standard_normal(n)
    for tag, ell in (("config ell=linspace(0.8,1.6)", np.linspace(0.8, 1.6, 8)), ("smooth ell=3", np.full(8, 3.0))):
        S = o.gram("rbf", X, ell, 1.0) + (1e-6 + 0.09) * np.eye(n)
        Lref = blocked_cholesky(S, nb, lambda a, b: a @ b.T)
        ld_ref = 2 * np.log(np.diag(Lref)).sum()
        w = np.linalg.solve(Lref, y)
        q_ref = w @ w
        print(f"{tag}: N={n} NB={nb} cond={np.linalg.cond(S):.1e}")
        for s in (5, 6, 7, 8):
            L = blocked_cholesky(S, nb, lambda a, b: ozaki_abt(a, b, s))
            w = np.linalg.solve(L, y)
            mll_ref = -0.5 * (n * np.log(2 * np.pi) + ld_ref + q_ref)
            mll = -0.5 * (n * np.log(2 * np.pi) + 2 * np.log(np.diag(L)).sum() + w @ w)
            print(f"  s={s} ({s * (s + 1) // 2:2d} int8 GEMMs): max|L-Lref|/max|L| = {np.max(np.abs(L - Lref)) / np.max(np.abs(Lref)):.1e}"
                  f"  MLL rel err = {abs(mll - mll_ref) / abs(mll_ref):.1e}")


if __name__ == "__main__":
    main()
