"""conjugate_mll and collapsed_elbo -- gpjax/objectives.py:36-107 and :321-416.

Same call signature as the reference (`Objective = Callable[[Module, Dataset], scalar]`,
objectives.py:33); the returned scalar is a device tensor attached to the autograd graph through a
fused custom backward, so user lambdas such as ``lambda p, d: -conjugate_mll(p, d)``
(examples/regression.py:200) compose unchanged.
"""
from __future__ import annotations

import torch

from . import ops, sgpr_ops, svgp_ops
from .dataset import Dataset
from .mean_functions import Constant, Zero
from .parameters import Parameter


def _mean_constant(mean_function):
    """Zero -> None; Constant -> its scalar tensor (objectives read the mean only as a vector m(x))."""
    if not isinstance(mean_function, Constant):
        raise NotImplementedError(
            "the fused objectives support Zero / Constant mean functions (SURVEY section 2, component 15)"
        )
    if isinstance(mean_function, Zero):
        return None
    c = mean_function.constant
    return c.value if isinstance(c, Parameter) else c


def _kernel_args(kernel):
    """(kind, lengthscale, [variance(, shape)]) of a kernel with a fused epilogue."""
    if getattr(kernel, "_b200_kind", None) is None:
        raise NotImplementedError(
            f"{type(kernel).__name__}: the fused objectives take a single stationary kernel with a fused sm_100a epilogue"
        )
    return kernel._b200_kind, kernel.lengthscale.value, kernel.kernel_scalars()


def _dist_world() -> int:
    import torch.distributed as dist

    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def _is_fused(kernel) -> bool:
    """True when `kernel.gram` is exactly one launch of the fused dense epilogue.  A kernel whose engine is not the plain
    DenseKernelComputation (White's constant-diagonal engine, a user-supplied engine) goes through `kernel.gram` instead,
    so that the engine's semantics are kept (computations/constant_diagonal.py:39-43 ignores coincident rows)."""
    from .kernels.computations import DenseKernelComputation

    return getattr(kernel, "_b200_kind", None) is not None and type(kernel.compute_engine) is DenseKernelComputation


def _dense_sigma(posterior, data: Dataset):
    """(Sigma, y - m(x)) with Sigma = Kxx + (jitter + obs_stddev^2) I assembled from differentiable Gram launches
    (objectives.py:96-103) -- the composable route for sum / product kernels."""
    x, y = data.X, data.y
    if y.shape[-1] != 1 and y.dim() > 1:
        raise ValueError("single-output objectives: y must have shape [N, 1]")
    prior = posterior.prior
    K = prior.kernel.gram(x).to_dense()
    Sigma = K.clone() if K.requires_grad or not K.is_contiguous() else K
    sn = posterior.likelihood.obs_stddev.value.reshape(()).to(Sigma.device)
    torch.diagonal(Sigma).add_(sn * sn + float(prior.jitter))
    mean = _mean_constant(prior.mean_function)
    d = y.reshape(-1).to(Sigma.device)
    if mean is not None:
        d = d - mean.reshape(()).to(Sigma.device)
    return Sigma, d


def conjugate_mll(posterior, data: Dataset) -> torch.Tensor:
    """log p(y | X, theta) of a conjugate GP: Sigma = Kxx + prior.jitter I + obs_stddev^2 I
    (objectives.py:96-103), value through the fused Gram -> Cholesky -> solve/logdet pipeline."""
    x, y = data.X, data.y
    kernel = posterior.prior.kernel
    if not _is_fused(kernel):  # combination kernels: dense Sigma from the parts' Gram launches
        Sigma, d = _dense_sigma(posterior, data)
        return ops.GaussianLogProbFunction.apply(Sigma, d)
    kind, ell, var = _kernel_args(kernel)
    xs = kernel.slice_input(x)
    xs = xs if xs.is_contiguous() else xs.contiguous()
    mean = _mean_constant(posterior.prior.mean_function)
    if mean is not None:
        mean = mean.to(xs.device)
    return ops.conjugate_mll_fused(kind, xs, y, ell, var, posterior.likelihood.obs_stddev.value, mean,
                                   float(posterior.prior.jitter))


def conjugate_loocv(posterior, data: Dataset) -> torch.Tensor:
    """Leave-one-out log predictive probability of a conjugate GP (objectives.py:110-178): with P = Sigma^-1 and
    alpha = P (y - m), sum_i log N(y_i; y_i - alpha_i / P_ii, 1 / P_ii).  Sigma^-1 comes from the blocked
    POTRF + TRTRI + LAUUM (the reference calls jnp.linalg.inv, objectives.py:171)."""
    Sigma, d = _dense_sigma(posterior, data)
    return ops.LoocvFunction.apply(Sigma, d)


def _has_fused_sparse_path(kernel) -> bool:
    """The streamed SGPR / SVGP kernels (csrc/sgpr.cpp) evaluate K_b tiles of ONE stationary kernel whose clamp-free diagonal is
    its variance.  Sum / product kernels, PoweredExponential (clamped diagonal) and user engines take the composable route."""
    return _is_fused(kernel) and kernel._b200_kind in sgpr_ops.FUSED_SPARSE_KINDS


def _collapsed_elbo_composable(q, data: Dataset) -> torch.Tensor:
    """objectives.py:342-416 evaluated operation by operation for kernels without a streamed sparse path (sums / products of
    kernels, PoweredExponential): every matrix product, factorisation and solve is a launch of this library's CUDA kernels
    (differentiable Gram tiles per part, DMMA GEMM, blocked Cholesky, triangular solves), the graph is torch autograd's.  Dense in
    N x M -- sized for one GPU's memory, not streamed, not sharded."""
    from .linalg import Dense, Triangular, lower_cholesky, psd, solve

    x, y = data.X, data.y
    n = x.shape[0]
    post = q.posterior
    kernel, mean_function = post.prior.kernel, post.prior.mean_function
    m = q.num_inducing
    sn = post.likelihood.obs_stddev.value.reshape(()).to(x.device)
    noise = sn * sn
    z = q.inducing_inputs.value
    eye = torch.eye(m, dtype=torch.float64, device=x.device)
    Kzz = kernel.gram(z).to_dense() + float(q.jitter) * eye
    Kzx = kernel.cross_covariance(z, x)                                   # [m, n]
    kxx = kernel.diagonal(x).diagonal                                     # k(x_i, x_i)
    diff = y.reshape(n, 1) - mean_function(x).reshape(n, 1)
    Lz = lower_cholesky(psd(Dense(Kzz)))
    A = solve(Lz, Kzx) / sn                                               # Lz^-1 Kzx / sigma
    AAT = ops.matmul_nt(A, A)                                             # [m, m]
    L = lower_cholesky(Dense(eye + AAT))
    log_det_B = 2.0 * torch.sum(torch.log(torch.diagonal(L.to_dense())))
    Ad = ops.matmul_nt(A, diff.reshape(1, n))                             # A (y - mu)   [m, 1]
    c = solve(L, Ad)
    quad = (torch.sum(diff * diff) - torch.sum(c * c)) / noise
    two_log_prob = -n * torch.log(2.0 * torch.pi * noise) - log_det_B - quad
    two_trace = torch.sum(kxx) / noise - torch.trace(AAT)
    return ((two_log_prob - two_trace) / 2.0).reshape(())


def _elbo_composable(q, data: Dataset) -> torch.Tensor:
    """objectives.py:241-318 with the Gaussian likelihood's analytical integrator (integrators.py:151-158) and the moments of
    variational_families.py:234-285 at the batch points, operation by operation (see _collapsed_elbo_composable)."""
    from .linalg import Dense, lower_cholesky, psd, solve

    x, y = data.X, data.y
    n = x.shape[0]
    post = q.posterior
    kernel, mean_function = post.prior.kernel, post.prior.mean_function
    m = q.num_inducing
    sn = post.likelihood.obs_stddev.value.reshape(()).to(x.device)
    noise = sn * sn
    z = q.inducing_inputs.value
    eye = torch.eye(m, dtype=torch.float64, device=x.device)
    Lz = lower_cholesky(psd(Dense(kernel.gram(z).to_dense() + float(q.jitter) * eye)))
    W = torch.tril(q.variational_root_covariance.value)
    mu_t = (q.variational_mean.value.reshape(m, 1) - mean_function(z).reshape(m, 1))
    # KL[N(mu, W W^T) || N(mu_z, Kzz)]  (variational_families.py:169-210)
    LiW = solve(Lz, W)
    Lim = solve(Lz, mu_t)
    Lzd = Lz.to_dense()
    kl = 0.5 * (torch.sum(LiW * LiW) + torch.sum(Lim * Lim) - m
                + 2.0 * torch.sum(torch.log(torch.diagonal(Lzd))) - 2.0 * torch.sum(torch.log(torch.abs(torch.diagonal(W)))))
    # moments of q(f(x_i))
    Kzx = kernel.cross_covariance(z, x)                                   # [m, n]
    A = solve(Lz, Kzx)                                                    # Lz^-1 Kzx
    KiK = solve(Lz.T, A)                                                  # Kzz^-1 Kzx
    R = ops.matmul_nt(KiK, W, a_layout=1, b_layout=1)                     # (Kzz^-1 Kzx)^T W   [n, m]
    mean = mean_function(x).reshape(n, 1) + ops.matmul_nt(KiK, mu_t.reshape(1, m), a_layout=1)
    var = kernel.diagonal(x).diagonal - torch.sum(A * A, dim=0) + torch.sum(R * R, dim=1) + float(q.jitter)
    err = y.reshape(n, 1) - mean
    expectation = -0.5 * (torch.log(torch.tensor(2.0 * torch.pi, dtype=torch.float64, device=x.device)) + torch.log(noise)
                          + (err.reshape(-1) ** 2 + var) / noise)
    return (torch.sum(expectation) * float(post.likelihood.num_datapoints) / n - kl).reshape(())


def collapsed_elbo(variational_family, data: Dataset, *, block_rows: int = sgpr_ops.DEFAULT_BLOCK_ROWS,
                   group=None, statistics: str = "auto") -> torch.Tensor:
    """Collapsed (Titsias) evidence lower bound (objectives.py:342-416).

    `data` holds THIS rank's rows; when torch.distributed is initialised the row-additive statistics
    and the gradient are all-reduced over `group`, so every rank returns the full-data ELBO.
    `statistics` selects how pass 1 forms them ("auto" | "whitened" | "raw", see sgpr_ops.collapsed_elbo_fused):
    the reference's whiten-first order always, or the cheaper raw-product route while Kzz is well conditioned."""
    x, y = data.X, data.y
    post = variational_family.posterior
    kernel = post.prior.kernel
    if not _has_fused_sparse_path(kernel):
        if group is not None or _dist_world() > 1:
            raise NotImplementedError("the composable collapsed_elbo (sum / product kernels) is single-GPU: it is not row-sharded")
        return _collapsed_elbo_composable(variational_family, data)
    kind, ell, var = _kernel_args(kernel)
    xs = kernel.slice_input(x)
    xs = xs if xs.is_contiguous() else xs.contiguous()
    z = kernel.slice_input(variational_family.inducing_inputs.value)
    mean = _mean_constant(post.prior.mean_function)
    if mean is not None:
        mean = mean.to(xs.device)
    return sgpr_ops.collapsed_elbo_fused(kind, xs, y, z, ell, var, post.likelihood.obs_stddev.value, mean,
                                         float(variational_family.jitter), block_rows, group, statistics)


def elbo(variational_family, data: Dataset, *, block_rows: int = sgpr_ops.DEFAULT_BLOCK_ROWS, group=None,
         statistics: str = "auto") -> torch.Tensor:
    """Evidence lower bound of a VariationalGaussian (objectives.py:241-273):
    sum_b E_q[log p(y_b | f(x_b))] * num_datapoints / batch - KL[q(u) || p(u)], Gaussian likelihood (analytical
    integrator, integrators.py:151-158).  `data` is THIS rank's minibatch; with torch.distributed initialised the
    statistics and gradients are all-reduced over `group` (effective batch = sum of the rank batches)."""
    from .likelihoods import Gaussian

    q = variational_family
    post = q.posterior
    if not isinstance(post.likelihood, Gaussian):
        raise NotImplementedError("the fused ELBO covers the Gaussian likelihood (analytical integrator) only")
    kernel = post.prior.kernel
    if not _has_fused_sparse_path(kernel):
        if group is not None or _dist_world() > 1:
            raise NotImplementedError("the composable elbo (sum / product kernels) is single-GPU: it is not data-parallel")
        return _elbo_composable(q, data)
    kind, ell, var = _kernel_args(kernel)
    xs = kernel.slice_input(data.X)
    xs = xs if xs.is_contiguous() else xs.contiguous()
    z = kernel.slice_input(q.inducing_inputs.value)
    mean = _mean_constant(post.prior.mean_function)
    if mean is not None:
        mean = mean.to(xs.device)
    return svgp_ops.svgp_elbo_fused(kind, xs, data.y, z, ell, var, post.likelihood.obs_stddev.value, mean,
                                    q.variational_mean.value, q.variational_root_covariance.value,
                                    float(post.likelihood.num_datapoints), float(q.jitter), block_rows, group, statistics)


__all__ = ["conjugate_mll", "conjugate_loocv", "collapsed_elbo", "elbo"]
