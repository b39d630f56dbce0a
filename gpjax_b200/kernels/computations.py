"""Kernel computation engines -- gpjax/kernels/computations/{base,dense}.py.

`DenseKernelComputation` is the plugin seam (compute_engine=) of every kernel.  The reference engine
evaluates `vmap(vmap(kernel))`; this one pattern-matches the kernel type and launches the fused
sm_100a Gram tile kernel.  There is deliberately no generic/CPU fallback: an unsupported kernel
raises NotImplementedError.
"""
from __future__ import annotations

import torch

from .. import ops
from ..linalg import Dense, Diagonal, psd


class AbstractKernelComputation:
    """computations/base.py:36-109."""

    def gram(self, kernel, x):
        return psd(Dense(self.cross_covariance(kernel, x, x)))

    def cross_covariance(self, kernel, x, y):
        return self._cross_covariance(kernel, x, y)

    def _cross_covariance(self, kernel, x, y):  # pragma: no cover - abstract
        raise NotImplementedError

    def diagonal(self, kernel, inputs):
        return self._diagonal(kernel, inputs)

    def _diagonal(self, kernel, inputs):
        raise NotImplementedError


def _prep(kernel, x: torch.Tensor) -> torch.Tensor:
    if not isinstance(x, torch.Tensor):
        raise TypeError("inputs must be torch tensors")
    if x.dim() != 2:
        raise ValueError(f"inputs must be 2-dimensional (N, D); got shape {tuple(x.shape)}")
    xs = kernel.slice_input(x)
    return xs if xs.is_contiguous() else xs.contiguous()


class DenseKernelComputation(AbstractKernelComputation):
    """B200 engine for RBF / Matern32 / Matern52 (computations/dense.py:27-36)."""

    def _kind(self, kernel) -> int:
        kind = getattr(kernel, "_b200_kind", None)
        if kind is None:
            raise NotImplementedError(
                f"{type(kernel).__name__} has no fused sm_100a Gram epilogue (supported: RBF, Matern32, Matern52); "
                "gpjax_b200 has no generic fallback path"
            )
        return kind

    def _cross_covariance(self, kernel, x, y):
        kind = self._kind(kernel)
        xs, ys = _prep(kernel, x), _prep(kernel, y)
        return ops.GramFunction.apply(kind, xs, ys, kernel.lengthscale.value, kernel.variance.value, False)

    def _diagonal(self, kernel, inputs):
        # k(x, x) = variance for every stationary kernel here (exp(0) = 1; Matern tau clamp ~ 1e-18)
        self._kind(kernel)
        n = inputs.shape[0]
        return psd(Diagonal(kernel.variance.value.reshape(()) * torch.ones(n, dtype=torch.float64, device=inputs.device)))
