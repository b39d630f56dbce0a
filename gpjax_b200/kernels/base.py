"""AbstractKernel -- gpjax/kernels/base.py:40-149 (engine delegation + active_dims slicing)."""
from __future__ import annotations

import typing as tp

import torch

from ..parameters import Module
from .computations import AbstractKernelComputation, DenseKernelComputation


def _check_active_dims(active_dims):
    if not isinstance(active_dims, (list, slice)):
        raise TypeError(f"Expected active_dims to be a list or slice. Got {active_dims} instead.")


def _check_n_dims(n_dims):
    if n_dims is not None:
        if not isinstance(n_dims, int):
            raise TypeError(f"Expected n_dims to be an integer. Got {n_dims} instead.")
        if n_dims <= 0:
            raise ValueError(f"Expected n_dims to be strictly positive. Got {n_dims} instead.")


def _check_dims_compat(active_dims, n_dims):
    if isinstance(active_dims, list) and n_dims is None:
        n_dims = len(active_dims)
    if isinstance(active_dims, list) and n_dims is not None and len(active_dims) != n_dims:
        raise ValueError(
            "Expected the length of active_dims to be equal to the specified n_dims. "
            f"Got active_dims: {active_dims} and n_dims: {n_dims}."
        )
    return active_dims, n_dims


class AbstractKernel(Module):
    name: str = "AbstractKernel"

    def __init__(self, active_dims=None, n_dims: tp.Optional[int] = None,
                 compute_engine: AbstractKernelComputation = None):
        active_dims = active_dims or slice(None)
        _check_active_dims(active_dims)
        _check_n_dims(n_dims)
        self.active_dims, self.n_dims = _check_dims_compat(active_dims, n_dims)
        self.compute_engine = compute_engine if compute_engine is not None else DenseKernelComputation()

    def __call__(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        """Scalar k(x, y) for a single pair: a 1 x 1 launch of the same fused tile kernel."""
        return self.compute_engine.cross_covariance(self, x.reshape(1, -1), y.reshape(1, -1)).reshape(())

    def cross_covariance(self, x, y):
        return self.compute_engine.cross_covariance(self, x, y)

    def gram(self, x):
        return self.compute_engine.gram(self, x)

    def diagonal(self, x):
        return self.compute_engine.diagonal(self, x)

    def slice_input(self, x: torch.Tensor) -> torch.Tensor:
        return x[..., self.active_dims] if self.active_dims is not None else x


    # ---- kernel algebra (kernels/base.py:150-205) ------------------------------------------------------------
    def __add__(self, other):
        return SumKernel(kernels=[self, other if isinstance(other, AbstractKernel) else Constant(constant=other)])

    def __radd__(self, other):
        return self.__add__(other)

    def __mul__(self, other):
        return ProductKernel(kernels=[self, other if isinstance(other, AbstractKernel) else Constant(constant=other)])


class Constant(AbstractKernel):
    """k(x, y) = constant  (kernels/base.py:223-243)."""

    name = "Constant"
    _is_constant_kernel = True

    def __init__(self, active_dims=None, constant=0.0, compute_engine: AbstractKernelComputation = None):
        from ..parameters import Parameter, Real

        self.constant = constant if isinstance(constant, Parameter) else Real(constant)
        super().__init__(active_dims=active_dims, compute_engine=compute_engine)


class CombinationKernel(AbstractKernel):
    """Sum or product of kernels, nested instances of the same combination flattened (kernels/base.py:246-310)."""

    name = "Combination"

    def __init__(self, kernels, operator: str, compute_engine: AbstractKernelComputation = None):
        if operator not in ("sum", "prod"):
            raise ValueError("operator must be 'sum' or 'prod'")
        flat = []
        for kernel in kernels:
            if not isinstance(kernel, AbstractKernel):
                raise TypeError("can only combine Kernel instances")
            if isinstance(kernel, CombinationKernel) and kernel.operator_name == operator:
                flat.extend(kernel.kernels)
            else:
                flat.append(kernel)
        self.kernels = flat
        self.operator_name = operator
        super().__init__(compute_engine=compute_engine)


def SumKernel(kernels, compute_engine: AbstractKernelComputation = None) -> CombinationKernel:
    """kernels/base.py:338."""
    return CombinationKernel(kernels, "sum", compute_engine)


def ProductKernel(kernels, compute_engine: AbstractKernelComputation = None) -> CombinationKernel:
    """kernels/base.py:339."""
    return CombinationKernel(kernels, "prod", compute_engine)
