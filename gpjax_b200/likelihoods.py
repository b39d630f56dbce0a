"""gpjax/likelihoods.py:129-186 (Gaussian only: the conjugate path)."""
from __future__ import annotations

from .parameters import Module, NonNegativeReal


class AbstractLikelihood(Module):
    def __init__(self, num_datapoints: int):
        self.num_datapoints = num_datapoints


class Gaussian(AbstractLikelihood):
    def __init__(self, num_datapoints: int, obs_stddev=1.0):
        if not isinstance(obs_stddev, NonNegativeReal):
            obs_stddev = NonNegativeReal(obs_stddev)
        self.obs_stddev = obs_stddev
        super().__init__(num_datapoints)

    def predict(self, dist):
        """Adds obs_stddev^2 to the diagonal of the latent covariance (likelihoods.py:165-186)."""
        from .distributions import GaussianDistribution
        from .linalg import Dense

        cov = dist.covariance().clone()
        cov.diagonal().add_(self.obs_stddev.value.to(cov.device) ** 2)
        return GaussianDistribution(dist.loc, Dense(cov))

    __call__ = predict
