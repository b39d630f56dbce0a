"""Structured-matrix wrappers -- gpjax/linalg/operators.py:20-218 (Dense, Diagonal, Identity, Triangular)."""
from __future__ import annotations

import torch


class LinearOperator:
    def __init__(self):
        self.annotations = set()

    @property
    def shape(self):
        raise NotImplementedError

    def to_dense(self) -> torch.Tensor:
        raise NotImplementedError

    @property
    def T(self):
        return Dense(self.to_dense().T.contiguous())

    def __matmul__(self, other):
        o = other.to_dense() if isinstance(other, LinearOperator) else other
        from .. import ops

        d = self.to_dense()
        if o.dim() == 1:
            return ops.gemm(d.contiguous(), o.reshape(1, -1).contiguous()).reshape(-1)
        return ops.gemm(d.contiguous(), o.contiguous(), b_layout=1)


class Dense(LinearOperator):
    def __init__(self, array: torch.Tensor):
        super().__init__()
        self.array = array

    @property
    def shape(self):
        return tuple(self.array.shape)

    @property
    def dtype(self):
        return self.array.dtype

    def to_dense(self):
        return self.array


class Diagonal(LinearOperator):
    def __init__(self, diag: torch.Tensor):
        super().__init__()
        self.diagonal = diag

    @property
    def shape(self):
        n = self.diagonal.shape[0]
        return (n, n)

    def to_dense(self):
        return torch.diag(self.diagonal)

    @property
    def T(self):
        return self


class Identity(LinearOperator):
    def __init__(self, shape, dtype=torch.float64, device=None):
        super().__init__()
        n = shape if isinstance(shape, int) else shape[0]
        self._n = int(n)
        self._dtype = dtype
        self._device = device

    @property
    def shape(self):
        return (self._n, self._n)

    def to_dense(self):
        return torch.eye(self._n, dtype=self._dtype, device=self._device)

    @property
    def T(self):
        return self


class Triangular(LinearOperator):
    """operators.py:194-218.  `.array` holds the raw storage; `.T` flips `lower` WITHOUT copying: the
    transposed operator shares storage and remembers it (solve(Lz.T, .) becomes a trans=1 solve)."""

    def __init__(self, array: torch.Tensor, lower: bool = True, _base=None):
        super().__init__()
        self.array = array
        self.lower = lower
        self._base = _base  # (lower-triangular storage, factor workspace) when this is a transposed view
        self._ws = None

    @property
    def shape(self):
        return tuple(self.array.shape)

    def to_dense(self):
        return torch.tril(self.array) if self.lower else torch.triu(self.array)

    @property
    def T(self):
        if self._base is not None:  # transposing a transposed view gives the original back
            return self._base
        return Triangular(self.array.T, lower=not self.lower, _base=self)
