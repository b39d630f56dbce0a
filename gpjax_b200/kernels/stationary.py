"""Stationary kernels -- gpjax/kernels/stationary/{base,rbf,matern12,matern32,matern52,rational_quadratic,
powered_exponential,periodic,white}.py."""
from __future__ import annotations

import numbers
import typing as tp

import numpy as np
import torch

from ..parameters import NonNegativeReal, Parameter, PositiveReal, as_tensor
from .base import AbstractKernel
from .computations import AbstractKernelComputation, ConstantDiagonalKernelComputation


def _shape_of(lengthscale) -> tuple:
    if isinstance(lengthscale, Parameter):
        return tuple(lengthscale.value.shape)
    if isinstance(lengthscale, torch.Tensor):
        return tuple(lengthscale.shape)
    return np.shape(np.asarray(lengthscale))


def _check_lengthscale(lengthscale: tp.Any):
    """stationary/base.py:150-168."""
    if isinstance(lengthscale, Parameter):
        return _check_lengthscale(lengthscale.value)
    if not isinstance(lengthscale, (numbers.Real, np.ndarray, torch.Tensor, list, tuple)) or isinstance(lengthscale, bool):
        raise TypeError(f"Expected `lengthscale` to be a array-like. Got {lengthscale}.")
    if isinstance(lengthscale, (np.ndarray, torch.Tensor, list)):
        ls_shape = _shape_of(lengthscale)
        if len(ls_shape) > 1:
            raise ValueError(
                f"Expected `lengthscale` to be a scalar or 1D array. Got `lengthscale` with shape {ls_shape}."
            )


def _check_lengthscale_dims_compat(lengthscale, n_dims):
    """stationary/base.py:120-147."""
    ls_shape = _shape_of(lengthscale)
    if ls_shape == ():
        return n_dims
    if n_dims is None:
        return ls_shape[0]
    if ls_shape != (n_dims,):
        raise ValueError(
            "Expected `lengthscale` to be compatible with the number "
            f"of input dimensions. Got `lengthscale` with shape {ls_shape}, "
            f"but the number of input dimensions is {n_dims}."
        )
    return n_dims


class StationaryKernel(AbstractKernel):
    """stationary/base.py:42-106."""

    _b200_kind: tp.Optional[int] = None
    _b200_shape: tp.Optional[str] = None  # attribute holding the extra scalar of kinds 4..6

    def __init__(self, active_dims=None, lengthscale=1.0, variance=1.0, n_dims: tp.Optional[int] = None,
                 compute_engine: AbstractKernelComputation = None):
        super().__init__(active_dims, n_dims, compute_engine)
        _check_lengthscale(lengthscale)
        self.n_dims = _check_lengthscale_dims_compat(lengthscale, self.n_dims)
        self.lengthscale = lengthscale if isinstance(lengthscale, Parameter) else PositiveReal(lengthscale)
        self.variance = variance if isinstance(variance, Parameter) else NonNegativeReal(variance)

    def kernel_scalars(self) -> torch.Tensor:
        """[variance] or [variance, shape] -- the packing the C ABI takes (include/gpjax_b200.h)."""
        var = self.variance.value
        if self._b200_shape is None:
            return var
        shape = getattr(self, self._b200_shape)
        shape = shape.value if isinstance(shape, Parameter) else as_tensor(shape, var.device)
        return torch.stack([var.reshape(()), shape.reshape(()).to(var.device)])


class RBF(StationaryKernel):
    """k = s2 exp(-|x-y|^2 / (2 l^2))  (stationary/rbf.py:40-44)."""

    name = "RBF"
    _b200_kind = 0


class Matern32(StationaryKernel):
    """k = s2 (1 + sqrt3 tau) exp(-sqrt3 tau)  (stationary/matern32.py:41-54)."""

    name = "Matérn32"
    _b200_kind = 1


class Matern52(StationaryKernel):
    """k = s2 (1 + sqrt5 tau + 5/3 tau^2) exp(-sqrt5 tau)  (stationary/matern52.py:42-53)."""

    name = "Matérn52"
    _b200_kind = 2


class Matern12(StationaryKernel):
    """k = s2 exp(-tau)  (stationary/matern12.py:44-48)."""

    name = "Matérn12"
    _b200_kind = 3


class RationalQuadratic(StationaryKernel):
    """k = s2 (1 + |x-y|^2 / (2 alpha l^2))^(-alpha)  (stationary/rational_quadratic.py:44-83).  As in the reference,
    `alpha` is stored as given: pass a Parameter (e.g. PositiveReal) to make it trainable."""

    name = "Rational Quadratic"
    _b200_kind = 4
    _b200_shape = "alpha"

    def __init__(self, active_dims=None, lengthscale=1.0, variance=1.0, alpha=1.0, n_dims=None, compute_engine=None):
        self.alpha = alpha
        super().__init__(active_dims, lengthscale, variance, n_dims, compute_engine)


class PoweredExponential(StationaryKernel):
    """k = s2 exp(-tau^kappa)  (stationary/powered_exponential.py:48-89); `power` stored as given (the reference
    documents a SigmoidBounded parameter for a trainable power)."""

    name = "Powered Exponential"
    _b200_kind = 5
    _b200_shape = "power"

    def __init__(self, active_dims=None, lengthscale=1.0, variance=1.0, power=1.0, n_dims=None, compute_engine=None):
        self.power = power
        super().__init__(active_dims, lengthscale, variance, n_dims, compute_engine)


class Periodic(StationaryKernel):
    """k = s2 exp(-1/2 sum_d (sin(pi (x_d - y_d) / p) / l_d)^2)  (stationary/periodic.py:46-88)."""

    name = "Periodic"
    _b200_kind = 6
    _b200_shape = "period"

    def __init__(self, active_dims=None, lengthscale=1.0, variance=1.0, period=1.0, n_dims=None, compute_engine=None):
        self.period = period
        super().__init__(active_dims, lengthscale, variance, n_dims, compute_engine)


class White(StationaryKernel):
    """k = s2 * all(x == y)  (stationary/white.py:33-64); Gram through the constant-diagonal engine."""

    name = "White"
    _b200_kind = 7

    def __init__(self, active_dims=None, variance=1.0, n_dims=None, compute_engine=None):
        super().__init__(active_dims, 1.0, variance, n_dims,
                         compute_engine if compute_engine is not None else ConstantDiagonalKernelComputation())
