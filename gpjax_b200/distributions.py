"""GaussianDistribution.log_prob -- gpjax/distributions.py:45-55,115-134."""
from __future__ import annotations

import torch

from .linalg import Dense, LinearOperator, logdet, solve
from .parameters import LOG_2PI


class GaussianDistribution:
    def __init__(self, loc: torch.Tensor, scale: LinearOperator):
        self.loc = loc
        self.scale = scale

    def mean(self) -> torch.Tensor:
        return self.loc

    def covariance(self) -> torch.Tensor:
        return self.scale.to_dense()

    def variance(self) -> torch.Tensor:
        return torch.diagonal(self.covariance()).clone()

    def stddev(self) -> torch.Tensor:
        return torch.sqrt(self.variance())

    def log_prob(self, y: torch.Tensor) -> torch.Tensor:
        """-1/2 [ n log 2pi + logdet(Sigma) + d^T solve(Sigma, d) ].  Differentiable w.r.t. a dense Sigma, the location
        and y through ops.GaussianLogProbFunction (dSigma = 1/2 (alpha alpha^T - Sigma^-1) from TRTRI + LAUUM)."""
        from . import ops

        mu, sigma = self.loc, self.scale
        n = mu.shape[-1]
        diff = (y - mu).contiguous()
        if isinstance(sigma, Dense) and torch.is_grad_enabled() and (sigma.array.requires_grad or diff.requires_grad):
            return ops.GaussianLogProbFunction.apply(sigma.array.contiguous(), diff.reshape(-1))
        if isinstance(sigma, Dense):  # one Cholesky serves both the log-determinant and the solve
            from .linalg import lower_cholesky

            L = lower_cholesky(sigma)
            w = solve(L, diff)
            quad = ops.gemm(w.reshape(1, -1), w.reshape(1, -1)).reshape(())
            return -0.5 * (n * LOG_2PI + 2.0 * logdet(L) + quad)
        sol = solve(sigma, diff)
        quad = ops.gemm(diff.reshape(1, -1), sol.reshape(1, -1).contiguous()).reshape(())
        return -0.5 * (n * LOG_2PI + logdet(sigma) + quad)
