"""Type-dispatched lower_cholesky / solve / logdet / diag -- gpjax/linalg/operations.py:22-237.

Dense and Triangular branches run on the hand-written CUDA path (blocked DMMA Cholesky, triangular
solves through explicit diagonal-block inverses, fused diagonal log-sum).  Differences from the
reference that callers should know:
  * Dense `solve` / `logdet` factor with Cholesky instead of LU (identical for the SPD matrices every
    call site on the hot path passes; a non-SPD matrix yields NaN exactly like jnp.linalg.cholesky).
  * The fused objectives (gpjax_b200.objectives) carry their own analytic backward and never go through these
    functions.  For user-composed expressions the Dense / Triangular branches are differentiable as well: when an
    operand requires grad they run through ops.CholeskyFunction / ops.TriangularSolveFunction (the reverse mode
    jax.grad derives through jnp.linalg.cholesky / solve_triangular), otherwise through the in-place fast path.
"""
from __future__ import annotations

import torch

from .. import ops
from .operators import Dense, Diagonal, Identity, LinearOperator, Triangular


def _factor(A: torch.Tensor):
    """Cholesky of a dense tensor -> (L storage with zero upper, workspace).  As jnp.linalg.cholesky, the factor is that of
    (A + A^T) / 2: a Dense that is not exactly symmetric gives the same L as the reference."""
    n = A.shape[0]
    L = A.detach().clone().contiguous()
    ws = ops.FactorWorkspace(n, 1, potri=False, device=A.device)
    ops.potrf_lower_(L, ws, zero_upper=True, symmetrize=True)
    return L, ws


def lower_cholesky(A: LinearOperator) -> LinearOperator:
    if isinstance(A, Identity):
        return A
    if isinstance(A, Diagonal):
        return Diagonal(torch.sqrt(A.diagonal))
    if isinstance(A, Triangular):
        if A.lower:
            return A
        return lower_cholesky(Dense(A.to_dense()))  # operations.py:40-43: cholesky of the (symmetrised) dense upper factor
    if isinstance(A, Dense):
        if A.array.requires_grad and torch.is_grad_enabled():
            return Triangular(ops.CholeskyFunction.apply(A.array.contiguous()), lower=True)
        L, ws = _factor(A.array)
        out = Triangular(L, lower=True)
        out._ws = ws
        return out
    return lower_cholesky(Dense(A.to_dense()))


def _tri_storage(A: Triangular):
    """(lower-triangular storage, workspace, trans) for a Triangular operator."""
    if A._base is not None and A._base.lower:  # transposed view of a lower factor: no copy
        base, trans, store = A._base, True, A._base.array
    elif A.lower:
        base, trans, store = A, False, A.array
    else:  # genuinely upper storage: U = L^T with L = U^T (one materialised transpose)
        base, trans, store = A, True, A.array.T
    if not store.is_contiguous():
        store = store.contiguous()
    ws = getattr(base, "_ws", None)
    if ws is None or ws.n < store.shape[0]:
        ws = ops.FactorWorkspace(store.shape[0], 1, potri=False, device=store.device)
        ops.diag_inverses(store, ws)
        base._ws = ws
    return store, ws, trans


def solve(A: LinearOperator, b: torch.Tensor) -> torch.Tensor:
    """operations.py:73-120 incl. the 1-D promote/squeeze rule."""
    was_1d = b.ndim == 1
    if isinstance(A, Identity):
        return b
    if isinstance(A, Diagonal):
        return b / (A.diagonal if was_1d else A.diagonal[:, None])
    if isinstance(A, Triangular):
        store, ws, trans = _tri_storage(A)
        if torch.is_grad_enabled() and (store.requires_grad or b.requires_grad):
            if not was_1d and b.shape[1] > ws.n:
                ws = ops.FactorWorkspace(max(store.shape[0], b.shape[1]), 1, device=store.device)
                ops.diag_inverses(store.detach(), ws)
            return ops.TriangularSolveFunction.apply(store, b, trans, ws)
        if was_1d:
            return ops.trsv_lower_(store, b.detach().clone().contiguous(), ws, trans=trans)
        x = b.detach().clone().contiguous()
        if x.shape[1] > ws.n:
            ws = ops.FactorWorkspace(max(store.shape[0], x.shape[1]), 1, device=store.device)
            ops.diag_inverses(store, ws)
        return ops.trsm_lower_left_(store, x, ws, trans=trans)
    # Dense (SPD): Cholesky, then two triangular solves
    L = lower_cholesky(A if isinstance(A, Dense) else Dense(A.to_dense()))
    return solve(L.T, solve(L, b))


def logdet(A: LinearOperator) -> torch.Tensor:
    """operations.py:123-181.  Triangular: sum(log(diag)) -- NO factor 2 (operations.py:142-144)."""
    if isinstance(A, Identity):
        return torch.zeros((), dtype=torch.float64, device=A._device)
    if isinstance(A, Diagonal):
        return torch.sum(torch.log(A.diagonal))
    if isinstance(A, Triangular):
        if A.array.requires_grad and torch.is_grad_enabled():
            return torch.sum(torch.log(torch.diagonal(A.array)))  # N-element glue, differentiable
        arr = A.array if A.array.stride(1) == 1 else A.array.T  # the diagonal is transpose-invariant
        return ops.sum_log_diag(arr)
    L = lower_cholesky(A if isinstance(A, Dense) else Dense(A.to_dense()))
    return 2.0 * logdet(L)


def diag(A: LinearOperator) -> torch.Tensor:
    if isinstance(A, Diagonal):
        return A.diagonal
    if isinstance(A, Identity):
        return torch.ones(A.shape[0], dtype=torch.float64, device=A._device)
    return torch.diagonal(A.to_dense() if not isinstance(A, (Dense, Triangular)) else A.array).clone()
