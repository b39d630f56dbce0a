"""Prior / ConjugatePosterior containers and predict -- gpjax/gps.py:64-78,443-526,724-743."""
from __future__ import annotations

import torch

from . import ops
from .dataset import Dataset
from .distributions import GaussianDistribution
from .likelihoods import AbstractLikelihood, Gaussian
from .linalg import Dense, psd
from .mean_functions import Constant
from .parameters import Module


class AbstractPrior(Module):
    def __init__(self, kernel, mean_function, jitter: float = 1e-6):
        self.kernel = kernel
        self.mean_function = mean_function
        self.jitter = jitter

    def __call__(self, test_inputs):
        return self.predict(test_inputs)


class Prior(AbstractPrior):
    def __mul__(self, other: AbstractLikelihood):
        return construct_posterior(prior=self, likelihood=other)

    def __rmul__(self, other: AbstractLikelihood):
        return self.__mul__(other)

    @torch.no_grad()  # forward-only: the predictive moments are not part of the training graph
    def predict(self, test_inputs: torch.Tensor) -> GaussianDistribution:
        """gps.py:224-254: N(m(t), K(t,t) + jitter I)."""
        mean = self.mean_function(test_inputs)
        K = self.kernel.gram(test_inputs).to_dense().detach().clone()
        K.diagonal().add_(self.jitter)
        return GaussianDistribution(torch.atleast_1d(mean.squeeze()), psd(Dense(K)))


class AbstractPosterior(Module):
    def __init__(self, prior: AbstractPrior, likelihood: AbstractLikelihood, jitter: float = 1e-6):
        self.prior = prior
        self.likelihood = likelihood
        self.jitter = jitter

    def __call__(self, test_inputs, train_data):
        return self.predict(test_inputs, train_data)


class ConjugatePosterior(AbstractPosterior):
    @torch.no_grad()  # forward-only: the predictive moments are not part of the training graph
    def predict(self, test_inputs: torch.Tensor, train_data: Dataset) -> GaussianDistribution:
        """gps.py:495-526.  Sigma = K + posterior.jitter I + s^2 I is built and factored by the fused
        Gram + blocked DMMA Cholesky; L^-1 Kxt is a blocked DMMA triangular solve; the Schur
        complement is a DMMA SYRK.  Output covariance gets prior.jitter on its diagonal (gps.py:523)."""
        kern = self.prior.kernel
        x = kern.slice_input(train_data.X).contiguous()
        t = kern.slice_input(test_inputs).contiguous()
        y = train_data.y
        n, T = x.shape[0], t.shape[0]
        sn = self.likelihood.obs_stddev.value.reshape(1)
        mx = self.prior.mean_function(train_data.X)
        mean_t = self.prior.mean_function(test_inputs)
        from .objectives import _is_fused

        fused = _is_fused(kern)
        if fused:
            kind, ell, var = kern._b200_kind, kern.lengthscale.value, kern.kernel_scalars()
            Sigma = ops.gram_forward(kind, x, x, ell, var, diag_add=self.jitter, diag_add_sq=sn, lower_only=True)
            Kxt = ops.gram_forward(kind, x, t, ell, var)                  # [n, T]
        else:  # sum / product kernels: matrices assembled from the parts' Gram launches
            with torch.no_grad():
                Sigma = kern.gram(train_data.X).to_dense().clone()
                torch.diagonal(Sigma).add_(float(self.jitter) + sn.reshape(()) ** 2)
                Kxt = kern.cross_covariance(train_data.X, test_inputs).contiguous()
        ws = ops.FactorWorkspace(max(n, T), 1, potri=False, device=x.device)
        ops.potrf_lower_(Sigma, ws, zero_upper=False)
        V = ops.trsm_lower_left_(Sigma, Kxt, ws)                          # L^-1 Kxt (in place)
        w = ops.trsv_lower_(Sigma, (y - mx).reshape(-1).contiguous(), ws)  # L^-1 (y - m)
        mean = mean_t.reshape(-1) + ops.gemm(V, w.reshape(1, -1), a_layout=1).reshape(-1)
        if fused:
            cov = ops.gram_forward(kind, t, t, ell, var, diag_add=self.prior.jitter)
        else:
            with torch.no_grad():
                cov = kern.gram(test_inputs).to_dense().clone()
                torch.diagonal(cov).add_(float(self.prior.jitter))
        ops.gemm(V, V, cov, alpha=-1.0, beta=1.0, a_layout=1, b_layout=1)  # Ktt - V^T V
        return GaussianDistribution(torch.atleast_1d(mean), psd(Dense(cov)))

    def predict_mean_and_variance(self, test_inputs: torch.Tensor, train_data: Dataset):
        """Config-3 variant: mean[T] and diag(cov)[T] without forming the T x T covariance twice."""
        d = self.predict(test_inputs, train_data)
        return d.mean(), d.variance()


def construct_posterior(prior, likelihood):
    if isinstance(likelihood, Gaussian):
        return ConjugatePosterior(prior=prior, likelihood=likelihood)
    raise NotImplementedError("only the conjugate (Gaussian-likelihood) posterior is on the B200 hot path")


__all__ = ["AbstractPrior", "Prior", "AbstractPosterior", "ConjugatePosterior", "construct_posterior", "Constant"]
