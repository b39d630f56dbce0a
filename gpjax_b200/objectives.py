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
            f"{type(kernel).__name__}: the fused sparse objectives take a single stationary kernel with a fused "
            "sm_100a epilogue (sum / product kernels are supported by conjugate_mll and conjugate_loocv only)"
        )
    return kernel._b200_kind, kernel.lengthscale.value, kernel.kernel_scalars()


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
