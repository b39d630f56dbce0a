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

    @torch.no_grad()  # forward-only: the predictive moments are not part of the training graph
    def predict(self, test_inputs, train_data, *, block_rows: int = 32768, group=None):
        """Predictive Gaussian of the collapsed bound (variational_families.py:786-870).  The training data
        enter through the streamed, row-additive statistics of the ELBO path (all-reduced over `group`
        when `train_data` is this rank's shard), never through an N x M matrix:
            mean = mu_t + (Lz^-1 Kzt)^T B^-1 (Lz^-1 Kzx d) / s,      B = I + Lz^-1 Kzx Kxz Lz^-T / s
            cov  = Ktt - At^T At + (L^-1 At)^T (L^-1 At) + jitter I, At = Lz^-1 Kzt, L L^T = B."""
        from . import _abi, ops
        from ._lib import lib
        from .distributions import GaussianDistribution
        from .linalg import Dense
        from .objectives import _mean_constant
        from .ops import _ell_args, _p, _scalar, _stream
        from .sgpr_ops import _all_reduce, _state

        post = self.posterior
        kern = post.prior.kernel
        from .objectives import _has_fused_sparse_path

        if not _has_fused_sparse_path(kern):
            return self._predict_composable(test_inputs, train_data)
        kind = kern.compute_engine._kind(kern)
        x = kern.slice_input(train_data.X).contiguous()
        t = kern.slice_input(test_inputs).contiguous()
        z = kern.slice_input(self.inducing_inputs.value).contiguous()
        y = train_data.y.reshape(-1).contiguous()
        n_loc, D = x.shape
        M, T = z.shape[0], t.shape[0]
        ell_v, iso = _ell_args(kern.lengthscale.value, D)
        var = ops._kscalars(kind, kern.kernel_scalars())
        sn = _scalar(post.likelihood.obs_stddev.value, "obs_stddev")
        mean = _mean_constant(post.prior.mean_function)
        mean = None if mean is None else _scalar(mean.to(x.device), "mean constant")
        block_rows = int(min(block_rows, max(n_loc, 1)))
        st = _state(M, D, block_rows, z.device)
        L_ = lib()
        P = torch.empty(L_.gpb_sgpr_stats_count(M), dtype=torch.float64, device=z.device)
        rc = L_.gpb_sgpr_stats(_stream(), kind, n_loc, M, D, _p(x), x.stride(0), _p(y), _p(z), z.stride(0), _p(ell_v), iso,
                               _p(var), _p(sn), _p(mean), float(self.jitter), block_rows, _p(st.ws), st.nbytes, _p(P))
        _abi.check(rc, "gpb_sgpr_stats")
        st.generation = ops.next_generation()  # the workspace was overwritten: a pending backward must replay its forward
        _all_reduce(P, group)
        P = P.reshape(M + 2, M + 2)
        s = sn.reshape(()) ** 2
        Phi = torch.tril(P[:M, :M])
        Bmat = (Phi + torch.tril(Phi, -1).T) / s + torch.eye(M, dtype=torch.float64, device=z.device)  # M x M glue
        psi = P[M, :M].contiguous()
        wsB = ops.FactorWorkspace(max(M, T), 1, device=z.device)
        ops.potrf_lower_(Bmat, wsB, zero_upper=False)                              # L L^T = I + A A^T
        v = ops.trsv_lower_(Bmat, ops.trsv_lower_(Bmat, psi.clone(), wsB), wsB, trans=True)   # B^-1 (Lz^-1 Kzx d)
        Lz = ops.gram_forward(kind, z, z, ell_v, var, diag_add=self.jitter, lower_only=True)
        wsZ = ops.FactorWorkspace(max(M, T), 1, device=z.device)
        ops.potrf_lower_(Lz, wsZ, zero_upper=False)
        At = ops.trsm_lower_left_(Lz, ops.gram_forward(kind, z, t, ell_v, var), wsZ)          # Lz^-1 Kzt  [M, T]
        mean_fn = post.prior.mean_function
        mu = mean_fn(test_inputs).reshape(-1) + ops.gemm(At, (v / s).reshape(1, -1).contiguous(), a_layout=1).reshape(-1)
        cov = ops.gram_forward(kind, t, t, ell_v, var, diag_add=self.jitter)
        ops.gemm(At, At, cov, alpha=-1.0, beta=1.0, a_layout=1, b_layout=1)
        LAt = ops.trsm_lower_left_(Bmat, At.clone(), wsB)                                      # L^-1 At
        ops.gemm(LAt, LAt, cov, alpha=1.0, beta=1.0, a_layout=1, b_layout=1)
        return GaussianDistribution(torch.atleast_1d(mu), Dense(cov))


    def _predict_composable(self, test_inputs, train_data):
        """variational_families.py:766-870 operation by operation (sum / product kernels, PoweredExponential): dense in N x M."""
        from . import ops
        from .distributions import GaussianDistribution
        from .linalg import Dense, lower_cholesky, psd, solve

        post = self.posterior
        kern, mean_fn = post.prior.kernel, post.prior.mean_function
        x, y, t = train_data.X, train_data.y, test_inputs
        z = self.inducing_inputs.value
        n, m, T = x.shape[0], z.shape[0], t.shape[0]
        sn = post.likelihood.obs_stddev.value.reshape(()).to(x.device)
        eye = torch.eye(m, dtype=torch.float64, device=x.device)
        Lz = lower_cholesky(psd(Dense(kern.gram(z).to_dense() + float(self.jitter) * eye)))
        Kzx = kern.cross_covariance(z, x)
        A = solve(Lz, Kzx) / sn                                            # Lz^-1 Kzx / sigma
        L = lower_cholesky(Dense(eye + ops.matmul_nt(A, A)))               # L L^T = I + A A^T
        diff = (y.reshape(n, 1) - mean_fn(x).reshape(n, 1))
        Ad = ops.matmul_nt(A, diff.reshape(1, n)) * sn                     # Lz^-1 Kzx (y - mu)
        v = solve(L.T, solve(L, Ad))                                       # B^-1 Lz^-1 Kzx diff
        At = solve(Lz, kern.cross_covariance(z, t))                        # Lz^-1 Kzt   [m, T]
        mean = mean_fn(t).reshape(-1) + ops.matmul_nt(At, (v / (sn * sn)).reshape(1, m), a_layout=1).reshape(-1)
        LAt = solve(L, At)
        cov = kern.gram(t).to_dense() - ops.matmul_nt(At, At, a_layout=1, b_layout=1) \
            + ops.matmul_nt(LAt, LAt, a_layout=1, b_layout=1) + float(self.jitter) * torch.eye(T, dtype=torch.float64, device=x.device)
        return GaussianDistribution(torch.atleast_1d(mean), Dense(cov))


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

    @torch.no_grad()  # forward-only: the predictive moments are not part of the training graph
    def predict(self, test_inputs):
        """N(mu_t + Ktz Kzz^-1 (mu - mu_z), Ktt - Ktz Kzz^-1 Kzt + Ktz Kzz^-1 S Kzz^-1 Kzt + jitter I)
        (variational_families.py:234-285) on the CUDA path: fused Gram tiles, blocked DMMA Cholesky,
        triangular solves through explicit block inverses, DMMA GEMMs for the Schur terms."""
        from . import ops
        from .distributions import GaussianDistribution
        from .linalg import Dense

        kern = self.posterior.prior.kernel
        from .objectives import _is_fused

        if not _is_fused(kern):
            return self._predict_composable(test_inputs)
        kind = kern.compute_engine._kind(kern)
        z = kern.slice_input(self.inducing_inputs.value).contiguous()
        t = kern.slice_input(test_inputs).contiguous()
        ell, var = kern.lengthscale.value, kern.kernel_scalars()
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

    def _predict_composable(self, test_inputs):
        """variational_families.py:234-285 operation by operation, for kernels assembled from several Gram launches."""
        from . import ops
        from .distributions import GaussianDistribution
        from .linalg import Dense, lower_cholesky, psd, solve

        kern, mean_fn = self.posterior.prior.kernel, self.posterior.prior.mean_function
        z, t = self.inducing_inputs.value, test_inputs
        m, T = z.shape[0], t.shape[0]
        dev = z.device
        Lz = lower_cholesky(psd(Dense(kern.gram(z).to_dense() + float(self.jitter) * torch.eye(m, dtype=torch.float64, device=dev))))
        A = solve(Lz, kern.cross_covariance(z, t))                          # Lz^-1 Kzt
        KiK = solve(Lz.T, A)                                                # Kzz^-1 Kzt
        W = torch.tril(self.variational_root_covariance.value).contiguous()
        R = ops.matmul_nt(KiK, W, a_layout=1, b_layout=1)                   # (Kzz^-1 Kzt)^T W
        mu_t = (self.variational_mean.value.reshape(-1) - mean_fn(z).reshape(-1)).contiguous()
        mean = mean_fn(t).reshape(-1) + ops.matmul_nt(KiK, mu_t.reshape(1, m), a_layout=1).reshape(-1)
        cov = kern.gram(t).to_dense() - ops.matmul_nt(A, A, a_layout=1, b_layout=1) + ops.matmul_nt(R, R) \
            + float(self.jitter) * torch.eye(T, dtype=torch.float64, device=dev)
        return GaussianDistribution(torch.atleast_1d(mean), Dense(cov))
