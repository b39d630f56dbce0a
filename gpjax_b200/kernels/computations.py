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


def _scalars(kernel) -> torch.Tensor:
    return kernel.kernel_scalars() if hasattr(kernel, "kernel_scalars") else kernel.variance.value


class DenseKernelComputation(AbstractKernelComputation):
    """B200 engine (computations/dense.py:27-36): one fused Gram launch per stationary kernel (RBF, Matern12/32/52,
    RationalQuadratic, PoweredExponential, Periodic, White); sum / product / constant kernels are assembled from the
    launches of their parts (kernels/base.py:246-339).  Anything else raises -- there is no generic fallback."""

    def _kind(self, kernel) -> int:
        kind = getattr(kernel, "_b200_kind", None)
        if kind is None:
            raise NotImplementedError(
                f"{type(kernel).__name__} has no fused sm_100a Gram epilogue (supported: RBF, Matern12/32/52, "
                "RationalQuadratic, PoweredExponential, Periodic, White and sums / products of them); "
                "gpjax_b200 has no generic fallback path"
            )
        return kind

    def _cross_covariance(self, kernel, x, y):
        parts = getattr(kernel, "kernels", None)
        if parts is not None:  # CombinationKernel: reduce the parts' matrices with its operator
            out = None
            for k in parts:
                # every part evaluates its own matrix with its own engine semantics but always densely,
                # as CombinationKernel.__call__ does pair by pair (kernels/base.py:297-310)
                m = DenseKernelComputation._cross_covariance(self, k, x, y)
                out = m if out is None else (out + m if kernel.operator_name == "sum" else out * m)
            return out
        if getattr(kernel, "_is_constant_kernel", False):
            c = kernel.constant.value.reshape(1, 1).to(x.device)
            return c.expand(x.shape[0], y.shape[0]).contiguous()
        kind = self._kind(kernel)
        xs, ys = _prep(kernel, x), _prep(kernel, y)
        return ops.GramFunction.apply(kind, xs, ys, kernel.lengthscale.value, _scalars(kernel), False)

    def _diagonal(self, kernel, inputs):
        parts = getattr(kernel, "kernels", None)
        n = inputs.shape[0]
        ones = torch.ones(n, dtype=torch.float64, device=inputs.device)
        if parts is not None:
            out = None
            for k in parts:
                d = DenseKernelComputation._diagonal(self, k, inputs).diagonal
                out = d if out is None else (out + d if kernel.operator_name == "sum" else out * d)
            return psd(Diagonal(out))
        if getattr(kernel, "_is_constant_kernel", False):
            return psd(Diagonal(kernel.constant.value.reshape(()).to(inputs.device) * ones))
        kind = self._kind(kernel)
        var = kernel.variance.value.reshape(())
        if kind == 5:
            # the reference's distance clamp (stationary/utils.py:67) leaves tau = 1e-18 on the diagonal:
            # k(x, x) = variance * exp(-(1e-18)^power), visibly below variance for small powers
            power = _scalars(kernel)[1]
            var = var * torch.exp(-torch.pow(torch.tensor(1e-18, dtype=torch.float64, device=var.device), power))
        # every other kernel here has k(x, x) = variance exactly (exp(0) = 1; Matern tau clamp ~ 1e-18)
        return psd(Diagonal(var * ones))


class ConstantDiagonalKernelComputation(DenseKernelComputation):
    """computations/constant_diagonal.py:36-57 -- the White kernel's engine: gram = k(x0, x0) on the diagonal
    (a Diagonal operator, never an N x N matrix); cross-covariances stay dense launches."""

    def gram(self, kernel, x):
        n = x.shape[0]
        value = kernel.variance.value.reshape(())
        return psd(Diagonal(value * torch.ones(n, dtype=torch.float64, device=x.device)))
