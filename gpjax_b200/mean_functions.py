"""gpjax/mean_functions.py:123-164 (Constant / Zero)."""
from __future__ import annotations

import torch

from .parameters import Module, Parameter, as_tensor


class AbstractMeanFunction(Module):
    def __call__(self, x):
        raise NotImplementedError


class Constant(AbstractMeanFunction):
    """Returns ones((N, 1)) * constant; trainable only when given as a Parameter (mean_functions.py:135-138)."""

    def __init__(self, constant=0.0):
        self.constant = constant if isinstance(constant, Parameter) else as_tensor(constant)

    def constant_tensor(self) -> torch.Tensor:
        return self.constant.value if isinstance(self.constant, Parameter) else self.constant

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        c = self.constant_tensor().to(x.device)
        return torch.ones((x.shape[0], 1), dtype=torch.float64, device=x.device) * c


class Zero(Constant):
    def __init__(self):
        super().__init__(constant=0.0)
