"""CollapsedVariationalGaussian container -- gpjax/variational_families.py:107-131,767-784."""
from __future__ import annotations

from .gps import AbstractPosterior
from .likelihoods import Gaussian
from .parameters import Module, Real


class AbstractVariationalFamily(Module):
    def __init__(self, posterior: AbstractPosterior):
        self.posterior = posterior


class AbstractVariationalGaussian(AbstractVariationalFamily):
    def __init__(self, posterior: AbstractPosterior, inducing_inputs, jitter: float = 1e-6):
        if not isinstance(inducing_inputs, Real):
            inducing_inputs = Real(inducing_inputs)
        self.inducing_inputs = inducing_inputs
        self.jitter = jitter
        super().__init__(posterior)

    @property
    def num_inducing(self) -> int:
        return self.inducing_inputs.value.shape[0]


class CollapsedVariationalGaussian(AbstractVariationalGaussian):
    """Titsias (2009) collapsed bound; holds only the inducing inputs (trainable) and the jitter."""

    def __init__(self, posterior: AbstractPosterior, inducing_inputs, jitter: float = 1e-6):
        super().__init__(posterior, inducing_inputs, jitter)
        if not isinstance(posterior.likelihood, Gaussian):
            raise TypeError("Likelihood must be Gaussian.")
