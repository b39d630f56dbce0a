"""CollapsedVariationalGaussian container -- gpjax/variational_families.py:107-131,767-784."""
from __future__ import annotations

from .gps import AbstractPosterior
from .likelihoods import Gaussian
import torch

from .parameters import LowerTriangular, Module, Real, default_device


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


class VariationalGaussian(AbstractVariationalGaussian):
    """q(u) = N(mu, S), S = sqrt sqrt^T (gpjax/variational_families.py:134-285).  `prior_kl` and the
    per-point moments of `predict` are consumed by `gpjax_b200.objectives.elbo`, which evaluates them through
    the fused statistics path; `predict` here returns the full predictive Gaussian at test inputs."""

    def __init__(self, posterior: AbstractPosterior, inducing_inputs, variational_mean=None,
                 variational_root_covariance=None, jitter: float = 1e-6):
        super().__init__(posterior, inducing_inputs, jitter)
        m = self.num_inducing
        dev = self.inducing_inputs.value.device
        if variational_mean is None:
            variational_mean = torch.zeros((m, 1), dtype=torch.float64, device=dev)
        if variational_root_covariance is None:
            variational_root_covariance = torch.eye(m, dtype=torch.float64, device=dev)
        self.variational_mean = variational_mean if isinstance(variational_mean, Real) else Real(variational_mean)
        self.variational_root_covariance = (variational_root_covariance
                                            if isinstance(variational_root_covariance, LowerTriangular)
                                            else LowerTriangular(variational_root_covariance))

    def predict(self, test_inputs):
        """N(mu_t + Ktz Kzz^-1 (mu - mu_z), Ktt - Ktz Kzz^-1 Kzt + Ktz Kzz^-1 S Kzz^-1 Kzt + jitter I)
        (variational_families.py:234-285) on the CUDA path: fused Gram tiles, blocked DMMA Cholesky,
        triangular solves through explicit block inverses, DMMA GEMMs for the Schur terms."""
        from . import ops
        from .distributions import GaussianDistribution
        from .linalg import Dense

        kern = self.posterior.prior.kernel
        kind = kern.compute_engine._kind(kern)
        z = kern.slice_input(self.inducing_inputs.value).contiguous()
        t = kern.slice_input(test_inputs).contiguous()
        ell, var = kern.lengthscale.value, kern.variance.value
        mean_fn = self.posterior.prior.mean_function
        m, T = z.shape[0], t.shape[0]
        Lz = ops.gram_forward(kind, z, z, ell, var, diag_add=self.jitter, lower_only=True)
        ws = ops.FactorWorkspace(max(m, T), 1, potri=False, device=z.device)
        ops.potrf_lower_(Lz, ws, zero_upper=False)
        Kzt = ops.gram_forward(kind, z, t, ell, var)                        # [m, T]
        A = ops.trsm_lower_left_(Lz, Kzt, ws)                               # Lz^-1 Kzt
        KiK = ops.trsm_lower_left_(Lz, A.clone(), ws, trans=True)           # Kzz^-1 Kzt
        W = torch.tril(self.variational_root_covariance.value).contiguous()
        R = ops.gemm(KiK, W, a_layout=1, b_layout=1)                        # (Kzz^-1 Kzt)^T W   [T, m]
        mu_tilde = (self.variational_mean.value.reshape(-1) - mean_fn(self.inducing_inputs.value).reshape(-1)).contiguous()
        mean = mean_fn(test_inputs).reshape(-1) + ops.gemm(KiK, mu_tilde.reshape(1, -1), a_layout=1).reshape(-1)
        cov = ops.gram_forward(kind, t, t, ell, var, diag_add=self.jitter)
        ops.gemm(A, A, cov, alpha=-1.0, beta=1.0, a_layout=1, b_layout=1)   # - A^T A
        ops.gemm(R, R, cov, alpha=1.0, beta=1.0)                            # + R R^T
        return GaussianDistribution(torch.atleast_1d(mean), Dense(cov))
