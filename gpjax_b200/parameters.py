"""Parameter containers and bijections -- host-side mirror of gpjax/parameters.py.

Values are float64 torch tensors (device memory is torch's job here, nothing more).  The softplus
bijection and its chain rule stay in host code exactly as in the reference (fit.py:136-143,
parameters.py:140-146); the CUDA path always sees constrained values.
"""
from __future__ import annotations

import math
import numbers
import typing as tp

import torch


def default_device() -> torch.device:
    return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")


def as_tensor(value, device=None) -> torch.Tensor:
    """float64 tensor on the compute device; accepts numbers, lists, numpy arrays, tensors."""
    if isinstance(value, Parameter):
        value = value.value
    if isinstance(value, torch.Tensor):
        t = value.to(dtype=torch.float64)
        return t if device is None else t.to(device)
    import numpy as np

    return torch.as_tensor(np.asarray(value, dtype=np.float64), device=device or default_device())


def _check_is_arraylike(value) -> None:
    import numpy as np

    if not isinstance(value, (numbers.Number, list, np.ndarray, torch.Tensor, np.generic)):
        raise TypeError(f"Expected parameter value to be an array-like type. Got {value}.")


class Parameter:
    """gpjax/parameters.py:69-80.  `.value` is the constrained value; `.tag` selects the bijection."""

    def __init__(self, value, tag: str, device=None):
        _check_is_arraylike(value)
        self.value = as_tensor(value, device)
        self.tag = tag

    def replace(self, value) -> "Parameter":
        new = object.__new__(type(self))
        new.value, new.tag = value, self.tag
        return new

    @property
    def shape(self):
        return self.value.shape

    def to(self, device) -> "Parameter":
        self.value = self.value.to(device)
        return self

    def __repr__(self):
        return f"{type(self).__name__}(value={self.value.detach().cpu().numpy()!r}, tag={self.tag!r})"


class NonNegativeReal(Parameter):
    def __init__(self, value, tag: str = "non_negative", device=None):
        super().__init__(value, tag, device)
        if not bool((self.value >= 0).all()):
            raise ValueError(f"value needs to be non-negative, got {self.value}")


class PositiveReal(Parameter):
    def __init__(self, value, tag: str = "positive", device=None):
        super().__init__(value, tag, device)
        if not bool((self.value > 0).all()):
            raise ValueError(f"value needs to be positive, got {self.value}")


class Real(Parameter):
    def __init__(self, value, tag: str = "real", device=None):
        super().__init__(value, tag, device)


class SigmoidBounded(Parameter):
    """Parameter bounded between 0 and 1 (gpjax/parameters.py:106-122)."""

    def __init__(self, value, tag: str = "sigmoid", device=None):
        super().__init__(value, tag, device)
        if not bool(((self.value >= 0) & (self.value <= 1)).all()):
            raise ValueError(f"value needs to be bounded between 0.0 and 1.0, got {self.value}")


class LowerTriangular(Parameter):
    """Lower-triangular matrix parameter (gpjax/parameters.py:125-137): square and zero above the diagonal."""

    def __init__(self, value, tag: str = "lower_triangular", device=None):
        super().__init__(value, tag, device)
        v = self.value
        if v.dim() != 2 or v.shape[0] != v.shape[1]:
            raise ValueError(f"value needs to be a square matrix, got {v}")
        if not bool((torch.tril(v) == v).all()):
            raise ValueError(f"value needs to be a lower triangular matrix, got {v}")


# ---- bijections (numpyro SoftplusTransform / IdentityTransform as used by parameters.py:140-146) ----
class Bijection:
    def __call__(self, u: torch.Tensor) -> torch.Tensor:  # unconstrained -> constrained
        raise NotImplementedError

    def inv(self, y: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError


class SoftplusTransform(Bijection):
    def __call__(self, u):
        return torch.nn.functional.softplus(u, beta=1.0, threshold=1e30) if u.numel() else u

    def inv(self, y):
        return y + torch.log(-torch.expm1(-y))


class IdentityTransform(Bijection):
    def __call__(self, u):
        return u

    def inv(self, y):
        return y


class SigmoidTransform(Bijection):
    def __call__(self, u):
        return torch.sigmoid(u)

    def inv(self, y):
        return torch.log(y) - torch.log1p(-y)


class FillTriangularTransform(Bijection):
    """Vector of n(n+1)/2 entries <-> lower-triangular n x n matrix (gpjax/numpyro_extras.py:12-106);
    row-major order of the lower triangle."""

    def __call__(self, u):
        k = u.shape[-1]
        n = int((math.isqrt(8 * k + 1) - 1) // 2)
        if n * (n + 1) // 2 != k:
            raise ValueError(f"a vector of length {k} does not fill a lower triangle")
        idx = torch.tril_indices(n, n, device=u.device)
        out = torch.zeros((n, n), dtype=u.dtype, device=u.device)
        out[idx[0], idx[1]] = u
        return out

    def inv(self, y):
        n = y.shape[-1]
        idx = torch.tril_indices(n, n, device=y.device)
        return y[idx[0], idx[1]]


DEFAULT_BIJECTION: tp.Dict[str, Bijection] = {
    "positive": SoftplusTransform(),
    "non_negative": SoftplusTransform(),
    "real": IdentityTransform(),
    "sigmoid": SigmoidTransform(),
    "lower_triangular": FillTriangularTransform(),
}


class Module:
    """Minimal stand-in for flax.nnx.Module: attribute containers whose Parameter leaves can be
    enumerated (what nnx.split does for fit.py:133)."""

    def named_parameters(self, prefix: str = "", _seen=None):
        _seen = set() if _seen is None else _seen
        for name, attr in vars(self).items():
            path = f"{prefix}{name}"
            if isinstance(attr, Parameter):
                if id(attr) not in _seen:
                    _seen.add(id(attr))
                    yield path, attr
            elif isinstance(attr, Module):
                yield from attr.named_parameters(path + ".", _seen)
            elif isinstance(attr, (list, tuple)):
                for i, a in enumerate(attr):
                    if isinstance(a, Module):
                        yield from a.named_parameters(f"{path}[{i}].", _seen)
                    elif isinstance(a, Parameter) and id(a) not in _seen:
                        _seen.add(id(a))
                        yield f"{path}[{i}]", a

    def to(self, device):
        for _, p in self.named_parameters():
            p.to(device)
        return self


def transform(params: tp.Dict[str, Parameter], params_bijection: tp.Dict[str, Bijection], inverse: bool = False):
    """gpjax/parameters.py:16-66 on a {path: Parameter} mapping."""
    out = {}
    for k, p in params.items():
        bij = params_bijection.get(p.tag, IdentityTransform())
        out[k] = p.replace(bij.inv(p.value) if inverse else bij(p.value))
    return out


LOG_2PI = math.log(2.0 * math.pi)
