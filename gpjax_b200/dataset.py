"""gpjax/dataset.py:26-126 mirrored on torch tensors."""
from __future__ import annotations

import warnings
from dataclasses import dataclass
from typing import Optional

import torch


@dataclass
class Dataset:
    X: Optional[torch.Tensor] = None
    y: Optional[torch.Tensor] = None

    def __post_init__(self) -> None:
        import numpy as np

        for name in ("X", "y"):
            v = getattr(self, name)
            if v is not None and not isinstance(v, torch.Tensor):
                setattr(self, name, torch.as_tensor(np.asarray(v)))
        _check_shape(self.X, self.y)
        _check_precision(self.X, self.y)

    def __repr__(self) -> str:
        return f"Dataset(Number of observations: {self.n:=} - Input dimension: {self.in_dim})"

    def is_supervised(self) -> bool:
        return self.X is not None and self.y is not None

    def is_unsupervised(self) -> bool:
        return self.X is None and self.y is not None

    def __add__(self, other: "Dataset") -> "Dataset":
        X = torch.cat((self.X, other.X)) if self.X is not None and other.X is not None else None
        y = torch.cat((self.y, other.y)) if self.y is not None and other.y is not None else None
        return Dataset(X=X, y=y)

    @property
    def n(self) -> int:
        return self.X.shape[0]

    @property
    def in_dim(self) -> int:
        return self.X.shape[1]

    def to(self, device, non_blocking: bool = False) -> "Dataset":
        return Dataset(X=None if self.X is None else self.X.to(device, non_blocking=non_blocking),
                       y=None if self.y is None else self.y.to(device, non_blocking=non_blocking))


def _check_shape(X, y) -> None:
    if X is not None and y is not None and X.shape[0] != y.shape[0]:
        raise ValueError(
            "Inputs, X, and outputs, y, must have the same number of rows."
            f" Got X.shape={tuple(X.shape)} and y.shape={tuple(y.shape)}."
        )
    if X is not None and X.ndim != 2:
        raise ValueError(f"Inputs, X, must be a 2-dimensional array. Got X.ndim={X.ndim}.")
    if y is not None and y.ndim != 2:
        raise ValueError(f"Outputs, y, must be a 2-dimensional array. Got y.ndim={y.ndim}.")


def _check_precision(X, y) -> None:
    if X is not None and X.dtype != torch.float64:
        warnings.warn(f"X is not of type float64. Got X.dtype={X.dtype}. This may lead to numerical instability. ",
                      stacklevel=2)
    if y is not None and y.dtype != torch.float64:
        warnings.warn(f"y is not of type float64.Got y.dtype={y.dtype}. This may lead to numerical instability.",
                      stacklevel=2)


__all__ = ["Dataset"]
