"""NumPy / torch-CPU float64 restatement of the GPJax hot path (TEST INFRASTRUCTURE).

Every function cites the reference file:line (relative to the gpjax 0.13.2 tree) whose
operation order it follows.  Values are NumPy; gradients come from two independent routes
that the tests cross-check:

* ``*_value_and_grad_autodiff`` -- a literal torch-CPU float64 restatement differentiated by
  reverse-mode autograd (the analogue of ``jax.value_and_grad`` in ``gpjax/fit.py:160``),
* ``*_grad_closed_form``        -- the analytic expressions the CUDA path implements.

Nothing in ``gpjax_b200/`` imports this module.
"""
from __future__ import annotations

import math

import numpy as np
import scipy.linalg as sla

KINDS = {"rbf": 0, "matern32": 1, "matern52": 2, "matern12": 3,
         "rational_quadratic": 4, "powered_exponential": 5, "periodic": 6, "white": 7}
# Kinds 4..6 have one extra scalar (alpha / power / period).  Throughout this module their `variance` argument is
# the pair (variance, shape) -- the same packing the C ABI uses -- and gradients come back as a pair as well.
SHAPE_KINDS = (4, 5, 6)


def _split_scal(kind: int, variance):
    if kind in SHAPE_KINDS:
        return variance[0], variance[1]
    return variance, None
KIND_NAMES = {v: k for k, v in KINDS.items()}

__all__ = [
    "KINDS",
    "KIND_NAMES",
    "squared_distance",
    "euclidean_distance",
    "kernel_pair",
    "cross_covariance",
    "gram",
    "add_jitter",
    "softplus",
    "softplus_inv",
    "gaussian_log_prob_lu",
    "conjugate_mll",
    "conjugate_mll_chol",
    "conjugate_mll_grad_closed_form",
    "conjugate_mll_value_and_grad_autodiff",
    "collapsed_elbo",
    "collapsed_elbo_streamed",
    "collapsed_elbo_grad_closed_form",
    "collapsed_elbo_value_and_grad_autodiff",
    "conjugate_predict",
    "gram_longdouble",
    "conjugate_mll_longdouble",
    "collapsed_elbo_longdouble",
    "collapsed_elbo_grad_longdouble_fd",
    "reference_cpu_mll_value_and_grad",
    "reference_cpu_elbo_value_and_grad",
    "svgp_elbo",
    "svgp_elbo_value_and_grad_autodiff",
    "svgp_predict",
    "collapsed_predict",
    "conjugate_loocv",
    "conjugate_loocv_value_and_grad_autodiff",
    "SHAPE_KINDS",
    "kernel_diagonal",
]


def _kind_id(kind):
    """Kernel selector: a name / integer id of a stationary kernel, or a COMBINATION spec
    ``("sum" | "prod", [(kind, lengthscale, variance), ...])`` (gpjax/kernels/base.py:246-339: the parts' values are summed /
    multiplied pair by pair); for a combination the ``lengthscale`` / ``variance`` arguments of the callers are ignored."""
    if isinstance(kind, tuple):
        return kind
    if isinstance(kind, str):
        return KINDS[kind.lower()]
    return int(kind)


def _combine(kind, fn):
    op, parts = kind
    out = None
    for (k, ell, var) in parts:
        m = fn(k, ell, var)
        out = m if out is None else (out + m if op == "sum" else out * m)
    return out


def kernel_diagonal(kind, x, lengthscale, variance) -> np.ndarray:
    """``vmap(kernel, in_axes=(0, 0))(x, x)`` (objectives.py:356) INCLUDING the 1e-36 distance clamp of utils.py:67, i.e.
    variance * exp(-(1e-18)^power) for PoweredExponential."""
    kind = _kind_id(kind)
    x = np.atleast_2d(np.asarray(x, np.float64))
    if isinstance(kind, tuple):
        return _combine(kind, lambda k, ell, var: kernel_diagonal(k, x, ell, var))
    return np.array([kernel_pair(kind, xi, xi, lengthscale, variance) for xi in x], np.float64)


# ----------------------------------------------------------------------------------------
# a1: gpjax/kernels/stationary/utils.py:42-67
# ----------------------------------------------------------------------------------------
def squared_distance(x: np.ndarray, y: np.ndarray) -> np.floating:
    """``jnp.sum((x - y) ** 2)`` -- gpjax/kernels/stationary/utils.py:53."""
    return np.sum((np.asarray(x, np.float64) - np.asarray(y, np.float64)) ** 2)


def euclidean_distance(x: np.ndarray, y: np.ndarray) -> np.floating:
    """``sqrt(max(r2, 1e-36))`` -- gpjax/kernels/stationary/utils.py:67."""
    return np.sqrt(np.maximum(squared_distance(x, y), 1e-36))


def _profile(kind: int, r2: np.ndarray, variance) -> np.ndarray:
    """Kernel profile on an array of squared scaled distances.

    rbf.py:43, matern32.py:46-53 (tau via utils.py:67), matern52.py:45-52.
    """
    variance, shape = _split_scal(kind, variance)
    if kind == 0 or kind == 6:  # periodic.py:88: r2 is then sum_d (sin(pi (x_d - y_d) / p) / l_d)^2
        return variance * np.exp(-0.5 * r2)
    if kind == 4:  # rational_quadratic.py:80-82
        return variance * (1.0 + 0.5 * r2 / shape) ** (-shape)
    if kind == 7:  # white.py:63: all(x == y) * variance  (lengthscale is 1, so r2 == 0 <=> equal)
        return variance * (np.asarray(r2) == 0.0)
    tau = np.sqrt(np.maximum(r2, 1e-36))
    if kind == 5:  # powered_exponential.py:88
        return variance * np.exp(-(tau**shape))
    if kind == 3:  # matern12.py:44-48
        return variance * np.exp(-tau)
    if kind == 1:
        return variance * (1.0 + np.sqrt(3.0) * tau) * np.exp(-np.sqrt(3.0) * tau)
    if kind == 2:
        return (
            variance
            * (1.0 + np.sqrt(5.0) * tau + 5.0 / 3.0 * np.square(tau))
            * np.exp(-np.sqrt(5.0) * tau)
        )
    raise ValueError(f"unknown kernel kind {kind}")


def kernel_pair(kind, x, y, lengthscale, variance) -> np.floating:
    """Scalar ``kernel(x, y)``: scale both inputs by 1/l first (rbf.py:41-42), then a1."""
    kind = _kind_id(kind)
    if kind == 6:  # periodic.py:81-88
        period = variance[1]
        sine_squared = (np.sin(np.pi * (np.asarray(x, np.float64) - np.asarray(y, np.float64)) / period) / lengthscale) ** 2
        return variance[0] * np.exp(-0.5 * np.sum(sine_squared, axis=0))
    xs = np.asarray(x, np.float64) / lengthscale
    ys = np.asarray(y, np.float64) / lengthscale
    return _profile(kind, squared_distance(xs, ys), variance)


def _r2_matrix(xs: np.ndarray, zs: np.ndarray) -> np.ndarray:
    """Direct-difference squared distances (never the |a|^2+|b|^2-2ab expansion)."""
    n, d = xs.shape
    r2 = np.zeros((n, zs.shape[0]), np.float64)
    for k in range(d):
        diff = xs[:, k : k + 1] - zs[:, k][None, :]
        r2 += diff * diff
    return r2


def cross_covariance(kind, x, z, lengthscale, variance) -> np.ndarray:
    """``vmap(vmap(kernel))`` -- gpjax/kernels/computations/dense.py:32-36."""
    kind = _kind_id(kind)
    x = np.atleast_2d(np.asarray(x, np.float64))
    z = np.atleast_2d(np.asarray(z, np.float64))
    if isinstance(kind, tuple):
        return _combine(kind, lambda k, ell, var: cross_covariance(k, x, z, ell, var))
    if kind == 6:
        ell = np.broadcast_to(np.asarray(lengthscale, np.float64), (x.shape[1],))
        r2 = np.zeros((x.shape[0], z.shape[0]), np.float64)
        for k in range(x.shape[1]):
            sk = np.sin(np.pi * (x[:, k : k + 1] - z[:, k][None, :]) / variance[1]) / ell[k]
            r2 += sk * sk
        return _profile(kind, r2, variance)
    xs = x / lengthscale
    zs = z / lengthscale
    return _profile(kind, _r2_matrix(xs, zs), variance)


def gram(kind, x, lengthscale, variance) -> np.ndarray:
    """``cross_covariance(x, x)`` -- gpjax/kernels/computations/base.py:56-72."""
    return cross_covariance(kind, x, x, lengthscale, variance)


def add_jitter(matrix: np.ndarray, jitter: float) -> np.ndarray:
    """gpjax/linalg/utils.py:39-65 (same error behaviour)."""
    if matrix.ndim != 2:
        raise ValueError(f"Expected 2D matrix, got {matrix.ndim}D array")
    if matrix.shape[0] != matrix.shape[1]:
        raise ValueError(f"Expected square matrix, got shape {matrix.shape}")
    return matrix + np.eye(matrix.shape[0]) * jitter


# bijections: gpjax/parameters.py:140-146 (numpyro SoftplusTransform)
def softplus(u):
    u = np.asarray(u, np.float64)
    return np.logaddexp(u, 0.0)


def softplus_inv(y):
    y = np.asarray(y, np.float64)
    return y + np.log(-np.expm1(-y))


# ----------------------------------------------------------------------------------------
# a9/a10: conjugate_mll, LU formulation exactly as the reference dispatches it
# ----------------------------------------------------------------------------------------
def gaussian_log_prob_lu(mu: np.ndarray, sigma: np.ndarray, y: np.ndarray) -> float:
    """gpjax/distributions.py:124-134 with Dense dispatch:
    logdet -> slogdet (LU, linalg/operations.py:163-165), solve -> LU (operations.py:109-111)."""
    n = mu.shape[-1]
    diff = y - mu
    logdet = np.linalg.slogdet(sigma)[1]
    quad = diff @ np.linalg.solve(sigma, diff)
    return float(-0.5 * (n * np.log(2.0 * np.pi) + logdet + quad))


def _sigma(kind, X, lengthscale, variance, obs_stddev, jitter):
    """objectives.py:96-102: K + jitter*I, then + eye*obs_noise (two separate adds)."""
    obs_noise = obs_stddev**2
    Kxx = gram(kind, X, lengthscale, variance)
    Kxx = add_jitter(Kxx, jitter)
    return Kxx + np.eye(Kxx.shape[0]) * obs_noise


def conjugate_mll(kind, X, y, lengthscale, variance, obs_stddev, mean_const=0.0, jitter=1e-6):
    """gpjax/objectives.py:93-107 (LU path, reference order)."""
    X = np.asarray(X, np.float64)
    y = np.asarray(y, np.float64).reshape(-1)
    mx = np.ones(X.shape[0]) * mean_const
    return gaussian_log_prob_lu(mx, _sigma(kind, X, lengthscale, variance, obs_stddev, jitter), y)


def conjugate_mll_chol(kind, X, y, lengthscale, variance, obs_stddev, mean_const=0.0, jitter=1e-6):
    """Cholesky twin of :func:`conjugate_mll` (what the CUDA path does)."""
    X = np.asarray(X, np.float64)
    y = np.asarray(y, np.float64).reshape(-1)
    n = X.shape[0]
    d = y - mean_const
    L = np.linalg.cholesky(_sigma(kind, X, lengthscale, variance, obs_stddev, jitter))
    w = sla.solve_triangular(L, d, lower=True)
    return float(-0.5 * (n * np.log(2.0 * np.pi) + 2.0 * np.sum(np.log(np.diag(L))) + w @ w))


def _dK_dr2(kind: int, r2: np.ndarray, K: np.ndarray, variance) -> np.ndarray:
    """dK/d(r^2); zero where the 1e-36 clamp is active."""
    if kind == 0:
        return -0.5 * K
    tau = np.sqrt(np.maximum(r2, 1e-36))
    live = r2 > 1e-36
    if kind == 3:
        g = -0.5 * K / tau
    elif kind == 1:
        g = -1.5 * variance * np.exp(-np.sqrt(3.0) * tau)
    else:
        g = -(5.0 / 6.0) * variance * (1.0 + np.sqrt(5.0) * tau) * np.exp(-np.sqrt(5.0) * tau)
    return np.where(live, g, 0.0)


def conjugate_mll_grad_closed_form(
    kind, X, y, lengthscale, variance, obs_stddev, mean_const=0.0, jitter=1e-6
):
    """Analytic gradient of conjugate_mll in constrained space.

    alpha = Sigma^-1 d, W = 1/2 (alpha alpha^T - Sigma^-1);
    d/dvar = <W,K>/var; d/dobs_stddev = 2 obs_stddev tr W; d/dc = sum alpha;
    d/dl_d = <W, dK/dr2 * (-2 Delta_d^2 / l_d^3)>.   Returns dict of numpy values.
    """
    kind = _kind_id(kind)
    X = np.asarray(X, np.float64)
    y = np.asarray(y, np.float64).reshape(-1)
    n, D = X.shape
    ell = np.asarray(lengthscale, np.float64)
    iso = ell.ndim == 0
    ell_v = np.full(D, float(ell)) if iso else ell
    K = gram(kind, X, ell_v, variance)
    Sigma = add_jitter(K, jitter) + np.eye(n) * obs_stddev**2
    Sinv = np.linalg.inv(Sigma)
    d = y - mean_const
    alpha = Sinv @ d
    W = 0.5 * (np.outer(alpha, alpha) - Sinv)
    xs = X / ell_v
    r2 = _r2_matrix(xs, xs)
    G = W * _dK_dr2(kind, r2, K, variance)
    g_ell = np.zeros(D)
    for k in range(D):
        diff = X[:, k : k + 1] - X[:, k][None, :]
        g_ell[k] = np.sum(G * (-2.0 * diff * diff / ell_v[k] ** 3))
    out = {
        "lengthscale": float(g_ell.sum()) if iso else g_ell,
        "variance": float(np.sum(W * K) / variance),
        "obs_stddev": float(2.0 * obs_stddev * np.trace(W)),
        "mean_const": float(alpha.sum()),
    }
    return out


# ----------------------------------------------------------------------------------------
# torch-CPU literal restatement, differentiated by autograd (jax.value_and_grad analogue)
# ----------------------------------------------------------------------------------------
def _grad_of(t):
    """numpy gradient of a leaf (zeros when the value does not depend on it, e.g. White's lengthscale)."""
    g = t.grad if t.grad is not None else t.detach() * 0.0
    return g.numpy().copy() if g.ndim else float(g)


def _t_profile(torch, kind, r2, variance):
    variance, shape = _split_scal(kind, variance)
    if kind == 0 or kind == 6:
        return variance * torch.exp(-0.5 * r2)
    if kind == 4:
        return variance * (1.0 + 0.5 * r2 / shape) ** (-shape)
    if kind == 7:
        return variance * (r2 == 0.0).to(torch.float64)
    tau = torch.sqrt(torch.clamp_min(r2, 1e-36))
    if kind == 5:
        return variance * torch.exp(-(tau**shape))
    if kind == 3:
        return variance * torch.exp(-tau)
    if kind == 1:
        s3 = math.sqrt(3.0)
        return variance * (1.0 + s3 * tau) * torch.exp(-s3 * tau)
    s5 = math.sqrt(5.0)
    return variance * (1.0 + s5 * tau + 5.0 / 3.0 * tau * tau) * torch.exp(-s5 * tau)


def _t_cross(torch, kind, x, z, ell, variance):
    if kind == 6:
        sine = torch.sin(math.pi * (x[:, None, :] - z[None, :, :]) / variance[1]) / ell
        return _t_profile(torch, kind, (sine * sine).sum(-1), variance)
    xs = x / ell
    zs = z / ell
    diff = xs[:, None, :] - zs[None, :, :]
    return _t_profile(torch, kind, (diff * diff).sum(-1), variance)


def conjugate_mll_value_and_grad_autodiff(
    kind, X, y, lengthscale, variance, obs_stddev, mean_const=0.0, jitter=1e-6, threads=None
):
    """Reverse-mode gradient of the literal LU-path restatement (objectives.py:93-107,
    distributions.py:124-134, operations.py:109-111,163-165).  O(N^2 D) memory: small N only
    unless D is small.  Returns (value, grads dict)."""
    import torch

    if threads:
        torch.set_num_threads(threads)
    kind = _kind_id(kind)
    t = lambda a: torch.tensor(np.asarray(a, np.float64), dtype=torch.float64, requires_grad=True)
    Xt = torch.tensor(np.asarray(X, np.float64))
    yt = torch.tensor(np.asarray(y, np.float64).reshape(-1))
    ell, var, sn, c = t(lengthscale), t(variance), t(obs_stddev), t(mean_const)
    n = Xt.shape[0]
    obs_noise = sn**2
    mx = torch.ones(n, dtype=torch.float64) * c
    Kxx = _t_cross(torch, kind, Xt, Xt, ell, var)
    eye = torch.eye(n, dtype=torch.float64)
    Sigma = (Kxx + eye * jitter) + eye * obs_noise
    diff = yt - mx
    logdet = torch.linalg.slogdet(Sigma)[1]
    quad = diff @ torch.linalg.solve(Sigma, diff)
    val = -0.5 * (n * math.log(2.0 * math.pi) + logdet + quad)
    val.backward()
    g = {
        "lengthscale": _grad_of(ell),
        "variance": _grad_of(var),
        "obs_stddev": float(sn.grad),
        "mean_const": float(c.grad),
    }
    return float(val.detach()), g


# ----------------------------------------------------------------------------------------
# a11: collapsed_elbo -- gpjax/objectives.py:342-416
# ----------------------------------------------------------------------------------------
def collapsed_elbo(kind, X, y, Z, lengthscale, variance, obs_stddev, mean_const=0.0, jitter=1e-6):
    """Literal order of gpjax/objectives.py:342-416 (A materialised; small N only)."""
    kind = _kind_id(kind)
    X = np.asarray(X, np.float64)
    y = np.asarray(y, np.float64).reshape(-1, 1)
    Z = np.asarray(Z, np.float64)
    n = X.shape[0]
    m = Z.shape[0]
    noise = obs_stddev**2
    Kzz = add_jitter(gram(kind, Z, lengthscale, variance), jitter)  # :352-354
    Kzx = cross_covariance(kind, Z, X, lengthscale, variance)  # :355
    Kxx_diag = kernel_diagonal(kind, X, lengthscale, variance)  # :356
    mux = np.ones((n, 1)) * mean_const  # :357
    Lz = np.linalg.cholesky(Kzz)  # :359
    A = sla.solve_triangular(Lz, Kzx, lower=True) / np.sqrt(noise)  # :387
    AAT = A @ A.T  # :390
    B = np.eye(m) + AAT  # :393
    L = np.linalg.cholesky(B)  # :396
    log_det_B = 2.0 * np.sum(np.log(np.diag(L)))  # :399
    diff = y - mux  # :401
    L_inv_A_diff = sla.solve_triangular(L, A @ diff, lower=True)  # :404
    quad = (np.sum(diff**2) - np.sum(L_inv_A_diff**2)) / noise  # :407
    two_log_prob = -n * np.log(2.0 * np.pi * noise) - log_det_B - quad  # :410
    two_trace = np.sum(Kxx_diag) / noise - np.trace(AAT)  # :413
    return float((two_log_prob - two_trace) / 2.0)  # :416


def _sgpr_stats_block(kind, Xb, db, Z, Lz, ell, variance, noise):
    Kzx = cross_covariance(kind, Z, Xb, ell, variance)
    A = sla.solve_triangular(Lz, Kzx, lower=True) / np.sqrt(noise)
    return A @ A.T, A @ db, A.sum(axis=1)


def collapsed_elbo_streamed(
    kind, X, y, Z, lengthscale, variance, obs_stddev, mean_const=0.0, jitter=1e-6,
    block=4096, return_stats=False,
):
    """Same value through row-additive statistics Phi, psi;
    this is the formulation that is sharded over GPUs, and the CPU baseline for config 4."""
    kind = _kind_id(kind)
    X = np.asarray(X, np.float64)
    y = np.asarray(y, np.float64).reshape(-1)
    Z = np.asarray(Z, np.float64)
    n, m = X.shape[0], Z.shape[0]
    noise = obs_stddev**2
    Lz = np.linalg.cholesky(add_jitter(gram(kind, Z, lengthscale, variance), jitter))
    Phi = np.zeros((m, m))
    psi = np.zeros(m)
    a1 = np.zeros(m)
    dd = 0.0
    sd = 0.0
    for s in range(0, n, block):
        db = y[s : s + block] - mean_const
        P, q, a = _sgpr_stats_block(kind, X[s : s + block], db, Z, Lz, lengthscale, variance, noise)
        Phi += P
        psi += q
        a1 += a
        dd += float(db @ db)
        sd += float(db.sum())
    val = _sgpr_finish(Phi, psi, dd, n, variance, noise)
    if return_stats:
        return val, dict(Phi=Phi, psi=psi, a1=a1, dd=dd, sd=sd, n=n, Lz=Lz)
    return val


def _sgpr_finish(Phi, psi, dd, n, variance, noise):
    m = Phi.shape[0]
    L = np.linalg.cholesky(np.eye(m) + Phi)
    w = sla.solve_triangular(L, psi, lower=True)
    quad = (dd - w @ w) / noise
    two_log_prob = -n * np.log(2.0 * np.pi * noise) - 2.0 * np.sum(np.log(np.diag(L))) - quad
    two_trace = n * (variance[0] if np.ndim(variance) else variance) / noise - np.trace(Phi)
    return float((two_log_prob - two_trace) / 2.0)


def collapsed_elbo_value_and_grad_autodiff(
    kind, X, y, Z, lengthscale, variance, obs_stddev, mean_const=0.0, jitter=1e-6, threads=None
):
    """Reverse-mode gradient of the literal restatement of objectives.py:342-416.
    Returns (value, grads) with grads for lengthscale, variance, obs_stddev, mean_const,
    inducing_inputs."""
    import torch

    if threads:
        torch.set_num_threads(threads)
    kind = _kind_id(kind)
    t = lambda a: torch.tensor(np.asarray(a, np.float64), dtype=torch.float64, requires_grad=True)
    Xt = torch.tensor(np.asarray(X, np.float64))
    yt = torch.tensor(np.asarray(y, np.float64).reshape(-1, 1))
    ell, var, sn, c, Zt = t(lengthscale), t(variance), t(obs_stddev), t(mean_const), t(Z)
    n, m = Xt.shape[0], Zt.shape[0]
    noise = sn**2
    eye = torch.eye(m, dtype=torch.float64)
    Kzz = _t_cross(torch, kind, Zt, Zt, ell, var) + eye * jitter
    Kzx = _t_cross(torch, kind, Zt, Xt, ell, var)
    Kxx_diag = _t_profile(torch, kind, torch.zeros(n, dtype=torch.float64), var)
    mux = torch.ones((n, 1), dtype=torch.float64) * c
    Lz = torch.linalg.cholesky(Kzz)
    A = torch.linalg.solve_triangular(Lz, Kzx, upper=False) / torch.sqrt(noise)
    AAT = A @ A.T
    L = torch.linalg.cholesky(eye + AAT)
    log_det_B = 2.0 * torch.sum(torch.log(torch.diagonal(L)))
    diff = yt - mux
    L_inv_A_diff = torch.linalg.solve_triangular(L, A @ diff, upper=False)
    quad = (torch.sum(diff**2) - torch.sum(L_inv_A_diff**2)) / noise
    two_log_prob = -n * torch.log(2.0 * math.pi * noise) - log_det_B - quad
    two_trace = torch.sum(Kxx_diag) / noise - torch.trace(AAT)
    val = (two_log_prob - two_trace) / 2.0
    val.backward()
    g = {
        "lengthscale": _grad_of(ell),
        "variance": _grad_of(var),
        "obs_stddev": float(sn.grad),
        "mean_const": float(c.grad),
        "inducing_inputs": Zt.grad.numpy().copy(),
    }
    return float(val.detach()), g


def collapsed_elbo_grad_closed_form(
    kind, X, y, Z, lengthscale, variance, obs_stddev, mean_const=0.0, jitter=1e-6, block=4096
):
    """Two-pass analytic gradient -- what the CUDA path implements.

    Pass 1 accumulates Phi, psi, a1, dd, sd; the M x M finish yields adjoints
    C = (2/s) Lz^-T dPhi Lz^-1, c = Lz^-T dpsi / sqrt(s), dKzz = Lz^-T (dPhi - Phi/2)... ;
    pass 2 recomputes K_b = k(Z, X_b) and contracts dK_b = C K_b + c d_b^T with dK/dtheta.
    """
    kind = _kind_id(kind)
    X = np.asarray(X, np.float64)
    y = np.asarray(y, np.float64).reshape(-1)
    Z = np.asarray(Z, np.float64)
    n, D = X.shape
    m = Z.shape[0]
    ell = np.asarray(lengthscale, np.float64)
    iso = ell.ndim == 0
    ell_v = np.full(D, float(ell)) if iso else ell
    s = obs_stddev**2
    val, st = collapsed_elbo_streamed(
        kind, X, y, Z, ell_v, variance, obs_stddev, mean_const, jitter, block, return_stats=True
    )
    Phi, psi, a1, dd, sd, Lz = st["Phi"], st["psi"], st["a1"], st["dd"], st["sd"], st["Lz"]
    Bm = np.eye(m) + Phi
    Binv = np.linalg.inv(Bm)
    v = Binv @ psi
    dPhi = 0.5 * (np.eye(m) - Binv - np.outer(v, v) / s)
    dpsi = v / s
    Lzinv = sla.solve_triangular(Lz, np.eye(m), lower=True)
    C = (2.0 / s) * Lzinv.T @ dPhi @ Lzinv
    C = 0.5 * (C + C.T)
    cvec = Lzinv.T @ dpsi / np.sqrt(s)
    # dELBO/dKzz from the square-root-invariant form (Q = Kzz + P/s, P = Kzx Kxz, b = Kzx d):
    #   logdet(I+Phi) = logdet Q - logdet Kzz,  psi^T B^-1 psi = b^T Q^-1 b / s,  tr Phi = tr(Kzz^-1 P)/s
    #   => dELBO/dKzz = Lz^-T (dPhi - Phi/2) Lz^-1.
    dKzz = Lzinv.T @ (dPhi - 0.5 * Phi) @ Lzinv
    dKzz = 0.5 * (dKzz + dKzz.T)
    # -- contract with dk/dtheta -------------------------------------------------------
    g_ell = np.zeros(D)
    g_Z = np.zeros((m, D))
    g_var = 0.0
    zs = Z / ell_v

    def contract(dK, Xr, Kmat, r2, z_is_both):
        nonlocal g_ell, g_Z, g_var
        Gm = dK * _dK_dr2(kind, r2, Kmat, variance)
        g_var += np.sum(dK * Kmat) / variance
        for k in range(D):
            diff = Z[:, k : k + 1] - Xr[:, k][None, :]
            g_ell[k] += np.sum(Gm * (-2.0 * diff * diff / ell_v[k] ** 3))
            gz = 2.0 * Gm * diff / ell_v[k] ** 2
            g_Z[:, k] += gz.sum(axis=1)
            if z_is_both:
                g_Z[:, k] -= gz.sum(axis=0)

    for s0 in range(0, n, block):
        Xb = X[s0 : s0 + block]
        db = y[s0 : s0 + block] - mean_const
        r2 = _r2_matrix(zs, Xb / ell_v)
        Kb = _profile(kind, r2, variance)
        dKb = C @ Kb + np.outer(cvec, db)
        contract(dKb, Xb, Kb, r2, False)
    r2zz = _r2_matrix(zs, zs)
    Kzz0 = _profile(kind, r2zz, variance)
    contract(dKzz, Z, Kzz0, r2zz, True)
    g_var += -n / (2.0 * s)
    quad_s = dd - psi @ v
    g_s = (
        -n / (2.0 * s)
        + quad_s / (2.0 * s * s)
        + n * variance / (2.0 * s * s)
        - (2.0 * np.sum(dPhi * Phi) + dpsi @ psi) / (2.0 * s)
    )
    g_c = -(dpsi @ a1) * 1.0 + sd / s
    # a1 = sum_rows A = Lz^-1 Kzx 1 / sqrt(s); psi = A d  => dpsi/dc = -a1
    return val, {
        "lengthscale": float(g_ell.sum()) if iso else g_ell,
        "variance": float(g_var),
        "obs_stddev": float(2.0 * obs_stddev * g_s),
        "mean_const": float(g_c),
        "inducing_inputs": g_Z,
    }


# ----------------------------------------------------------------------------------------
# a15: ConjugatePosterior.predict -- gpjax/gps.py:495-526
# ----------------------------------------------------------------------------------------
def conjugate_predict(
    kind, X, y, T, lengthscale, variance, obs_stddev, mean_const=0.0, jitter=1e-6,
    prior_jitter=1e-6,
):
    """Returns (mean[T], cov[T,T]) of the latent function, gps.py:495-526."""
    kind = _kind_id(kind)
    X = np.asarray(X, np.float64)
    y = np.asarray(y, np.float64).reshape(-1)
    T = np.asarray(T, np.float64)
    n = X.shape[0]
    Sigma = gram(kind, X, lengthscale, variance) + np.eye(n) * jitter  # :505-506
    Sigma = Sigma + np.eye(n) * obs_stddev**2  # :507-509
    L = np.linalg.cholesky(Sigma)  # :511
    Ktt = gram(kind, T, lengthscale, variance)  # :514
    Kxt = cross_covariance(kind, X, T, lengthscale, variance)  # :515
    V = sla.solve_triangular(L, Kxt, lower=True)  # :517
    w = sla.solve_triangular(L, y - mean_const, lower=True)  # :518
    mean = mean_const + V.T @ w  # :520
    cov = Ktt - V.T @ V + np.eye(T.shape[0]) * prior_jitter  # :522-523
    return mean, cov


# ----------------------------------------------------------------------------------------
# 80-bit adjudicators (small N only)
# ----------------------------------------------------------------------------------------
def gram_longdouble(kind, x, z, lengthscale, variance) -> np.ndarray:
    kind = _kind_id(kind)
    ld = np.longdouble
    x = np.atleast_2d(np.asarray(x, np.float64)).astype(ld)
    z = np.atleast_2d(np.asarray(z, np.float64)).astype(ld)
    ell = np.asarray(lengthscale, np.float64).astype(ld)
    xs, zs = x / ell, z / ell
    r2 = np.zeros((x.shape[0], z.shape[0]), ld)
    for k in range(x.shape[1]):
        diff = xs[:, k : k + 1] - zs[:, k][None, :]
        r2 += diff * diff
    var = ld(variance)
    if kind == 0:
        return var * np.exp(ld(-0.5) * r2)
    tau = np.sqrt(np.maximum(r2, ld(1e-36)))
    if kind == 3:
        return var * np.exp(-tau)
    if kind == 1:
        s3 = np.sqrt(ld(3.0))
        return var * (1 + s3 * tau) * np.exp(-s3 * tau)
    s5 = np.sqrt(ld(5.0))
    return var * (1 + s5 * tau + ld(5.0) / ld(3.0) * tau * tau) * np.exp(-s5 * tau)


def conjugate_mll_longdouble(kind, X, y, lengthscale, variance, obs_stddev, mean_const=0.0, jitter=1e-6):
    """Unblocked 80-bit Cholesky MLL (N <= ~400) to adjudicate LU-vs-Cholesky disagreements."""
    ld = np.longdouble
    X = np.asarray(X, np.float64)
    n = X.shape[0]
    S = gram_longdouble(kind, X, X, lengthscale, variance)
    S = S + np.eye(n, dtype=ld) * ld(jitter)
    S = S + np.eye(n, dtype=ld) * (ld(obs_stddev) ** 2)
    d = np.asarray(y, np.float64).reshape(-1).astype(ld) - ld(mean_const)
    L = np.zeros_like(S)
    for j in range(n):
        v = S[j:, j] - L[j:, :j] @ L[j, :j]
        L[j, j] = np.sqrt(v[0])
        L[j + 1 :, j] = v[1:] / L[j, j]
    w = np.zeros(n, ld)
    for i in range(n):
        w[i] = (d[i] - L[i, :i] @ w[:i]) / L[i, i]
    val = ld(-0.5) * (n * np.log(2 * ld(np.pi)) + 2 * np.sum(np.log(np.diag(L))) + w @ w)
    return val


def _chol_longdouble(S):
    n = S.shape[0]
    L = np.zeros_like(S)
    for j in range(n):
        v = S[j:, j] - L[j:, :j] @ L[j, :j]
        L[j, j] = np.sqrt(v[0])
        L[j + 1 :, j] = v[1:] / L[j, j]
    return L


def _trsm_longdouble(L, B):
    """L^-1 B by forward substitution, 80-bit."""
    X = np.zeros_like(B)
    for i in range(L.shape[0]):
        X[i] = (B[i] - L[i, :i] @ X[:i]) / L[i, i]
    return X


def collapsed_elbo_longdouble(kind, X, y, Z, lengthscale, variance, obs_stddev, mean_const=0.0, jitter=1e-6):
    """gpjax/objectives.py:342-416 in 80-bit arithmetic (unblocked Cholesky / substitution; M <= ~100, N <= ~1000): the
    adjudicator for disagreements between two float64 evaluation orders of the collapsed bound."""
    ld = np.longdouble
    X = np.asarray(X, np.float64)
    Z = np.asarray(Z, np.float64)
    n, m = X.shape[0], Z.shape[0]
    noise = ld(obs_stddev) ** 2
    Kzz = gram_longdouble(kind, Z, Z, lengthscale, variance) + np.eye(m, dtype=ld) * ld(jitter)
    Kzx = gram_longdouble(kind, Z, X, lengthscale, variance)
    diff = np.asarray(y, np.float64).reshape(-1).astype(ld) - ld(mean_const)
    Lz = _chol_longdouble(Kzz)
    A = _trsm_longdouble(Lz, Kzx) / np.sqrt(noise)
    AAT = A @ A.T
    L = _chol_longdouble(np.eye(m, dtype=ld) + AAT)
    c = _trsm_longdouble(L, (A @ diff).reshape(-1, 1)).reshape(-1)
    quad = (diff @ diff - c @ c) / noise
    two_log_prob = -n * np.log(2 * ld(np.pi) * noise) - 2 * np.sum(np.log(np.diag(L))) - quad
    two_trace = n * ld(variance) / noise - np.trace(AAT)  # k(x, x) = variance for the kernels gram_longdouble covers
    return (two_log_prob - two_trace) / 2


def collapsed_elbo_grad_longdouble_fd(kind, X, y, Z, lengthscale, variance, obs_stddev, mean_const=0.0, jitter=1e-6, h=1e-6):
    """Central differences of the 80-bit value (step h: truncation ~h^2, round-off ~1e-19 / h -> ~1e-12 relative): gradient with
    respect to lengthscale (vector), variance, obs_stddev, mean_const and the inducing inputs, as a dict like the autodiff oracle's."""
    ld = np.longdouble
    ell = np.atleast_1d(np.asarray(lengthscale, np.float64)).astype(ld)
    Zl = np.asarray(Z, np.float64).astype(ld)

    def f(ell_, var_, sn_, c_, Z_):
        return collapsed_elbo_longdouble_raw(kind, X, y, Z_, ell_, var_, sn_, c_, jitter)

    base = (ell, ld(variance), ld(obs_stddev), ld(mean_const), Zl)
    out = {}
    g = np.zeros(ell.shape[0])
    for i in range(ell.shape[0]):
        e = np.zeros_like(ell); e[i] = ld(h)
        g[i] = float((f(ell + e, *base[1:]) - f(ell - e, *base[1:])) / (2 * ld(h)))
    out["lengthscale"] = g
    out["variance"] = float((f(ell, base[1] + ld(h), *base[2:]) - f(ell, base[1] - ld(h), *base[2:])) / (2 * ld(h)))
    out["obs_stddev"] = float((f(ell, base[1], base[2] + ld(h), base[3], Zl) - f(ell, base[1], base[2] - ld(h), base[3], Zl)) / (2 * ld(h)))
    out["mean_const"] = float((f(ell, base[1], base[2], base[3] + ld(h), Zl) - f(ell, base[1], base[2], base[3] - ld(h), Zl)) / (2 * ld(h)))
    gz = np.zeros(Zl.shape)
    for a in range(Zl.shape[0]):
        for b in range(Zl.shape[1]):
            E = np.zeros_like(Zl); E[a, b] = ld(h)
            gz[a, b] = float((f(ell, base[1], base[2], base[3], Zl + E) - f(ell, base[1], base[2], base[3], Zl - E)) / (2 * ld(h)))
    out["inducing_inputs"] = gz
    return out


def collapsed_elbo_longdouble_raw(kind, X, y, Z, ell, variance, obs_stddev, mean_const, jitter):
    """collapsed_elbo_longdouble with 80-bit PARAMETERS (the finite differences perturb them below float64 resolution)."""
    ld = np.longdouble
    kind = _kind_id(kind)
    Xl = np.asarray(X, np.float64).astype(ld)
    n, m = Xl.shape[0], Z.shape[0]

    def gram_ld(a, b):
        xs, zs = a / ell, b / ell
        r2 = np.zeros((a.shape[0], b.shape[0]), ld)
        for k in range(a.shape[1]):
            dd = xs[:, k : k + 1] - zs[:, k][None, :]
            r2 += dd * dd
        if kind == 0:
            return variance * np.exp(ld(-0.5) * r2)
        tau = np.sqrt(np.maximum(r2, ld(1e-36)))
        if kind == 3:
            return variance * np.exp(-tau)
        if kind == 1:
            s3 = np.sqrt(ld(3.0))
            return variance * (1 + s3 * tau) * np.exp(-s3 * tau)
        s5 = np.sqrt(ld(5.0))
        return variance * (1 + s5 * tau + ld(5.0) / ld(3.0) * tau * tau) * np.exp(-s5 * tau)

    noise = obs_stddev * obs_stddev
    Kzz = gram_ld(Z, Z) + np.eye(m, dtype=ld) * ld(jitter)
    Kzx = gram_ld(Z, Xl)
    diff = np.asarray(y, np.float64).reshape(-1).astype(ld) - mean_const
    Lz = _chol_longdouble(Kzz)
    A = _trsm_longdouble(Lz, Kzx) / np.sqrt(noise)
    AAT = A @ A.T
    L = _chol_longdouble(np.eye(m, dtype=ld) + AAT)
    c = _trsm_longdouble(L, (A @ diff).reshape(-1, 1)).reshape(-1)
    quad = (diff @ diff - c @ c) / noise
    two_log_prob = -n * np.log(2 * ld(np.pi) * noise) - 2 * np.sum(np.log(np.diag(L))) - quad
    two_trace = n * variance / noise - np.trace(AAT)
    return (two_log_prob - two_trace) / 2


# ----------------------------------------------------------------------------------------
# Timed CPU stand-ins for the reference's jax[cpu] path (bench.py cpu_baseline / --impl reference)
# ----------------------------------------------------------------------------------------
def reference_cpu_mll_value_and_grad(kind, X, y, lengthscale, variance, obs_stddev, mean_const=0.0, jitter=1e-6):
    """conjugate_mll value + gradient in the reference's operation order on LAPACK:
    LU slogdet + LU solve for the value (distributions.py:132-134 -> operations.py:109-111,163-165)
    and the full inverse that reverse-mode of slogdet/solve materialises for the gradient
    (SURVEY section 3.1).  O(N^2) memory beyond Sigma and its inverse; D-loop kept out of N^2 D storage."""
    kind = _kind_id(kind)
    X = np.asarray(X, np.float64)
    y = np.asarray(y, np.float64).reshape(-1)
    n, D = X.shape
    ell = np.asarray(lengthscale, np.float64)
    ell_v = np.full(D, float(ell)) if ell.ndim == 0 else ell
    xs = X / ell_v
    r2 = _r2_matrix(xs, xs)
    K = _profile(kind, r2, variance)
    Sigma = add_jitter(K, jitter) + np.eye(n) * obs_stddev**2
    d = y - mean_const
    logdet = np.linalg.slogdet(Sigma)[1]            # LU #1
    alpha = np.linalg.solve(Sigma, d)               # LU #2
    value = float(-0.5 * (n * np.log(2.0 * np.pi) + logdet + d @ alpha))
    Sinv = np.linalg.inv(Sigma)                     # LU #3 + N-RHS solve
    W = 0.5 * (np.outer(alpha, alpha) - Sinv)
    G = W * _dK_dr2(kind, r2, K, variance)
    g_ell = np.empty(D)
    for k in range(D):
        diff = xs[:, k : k + 1] - xs[:, k][None, :]
        g_ell[k] = -2.0 / ell_v[k] * np.sum(G * diff * diff)
    grads = {
        "lengthscale": float(g_ell.sum()) if ell.ndim == 0 else g_ell,
        "variance": float(np.sum(W * K) / variance),
        "obs_stddev": float(2.0 * obs_stddev * np.trace(W)),
        "mean_const": float(alpha.sum()),
    }
    return value, grads


def reference_cpu_elbo_value_and_grad(kind, X, y, Z, lengthscale, variance, obs_stddev, mean_const=0.0, jitter=1e-6,
                                      block=8192):
    """collapsed_elbo value + gradient, blocked over rows (A never materialised for all N), LAPACK/BLAS."""
    return collapsed_elbo_grad_closed_form(kind, X, y, Z, lengthscale, variance, obs_stddev, mean_const, jitter, block)


# ----------------------------------------------------------------------------------------
# section 8f rank 1: SVGP elbo -- gpjax/objectives.py:241-315, variational_families.py:169-285,
# distributions.py:188-228, integrators.py:151-158
# ----------------------------------------------------------------------------------------
def svgp_elbo(kind, X, y, Z, lengthscale, variance, obs_stddev, mean_const, var_mean, var_sqrt, num_datapoints,
              jitter=1e-6):
    """Literal NumPy restatement: prior_kl (KL[N(mu,S)||N(mu_z,Kzz)]) + per-point predictive moments
    (VariationalGaussian.predict evaluated one point at a time, incl. its add_jitter on the 1x1 covariance)
    + the analytical Gaussian expected log-likelihood, scaled by num_datapoints / batch."""
    kind = _kind_id(kind)
    X = np.asarray(X, np.float64)
    y = np.asarray(y, np.float64).reshape(-1)
    Z = np.asarray(Z, np.float64)
    mu = np.asarray(var_mean, np.float64).reshape(-1)
    W = np.tril(np.asarray(var_sqrt, np.float64))
    m, b = Z.shape[0], X.shape[0]
    s = obs_stddev**2
    Kzz = add_jitter(gram(kind, Z, lengthscale, variance), jitter)
    # prior_kl: variational_families.py:169-213 -> distributions.py:188-228
    S = W @ W.T
    sqrt_p = np.linalg.cholesky(Kzz)
    sqrt_q = np.linalg.cholesky(S)
    diff = mean_const - mu
    trace = np.sum(np.square(sla.solve_triangular(sqrt_p, sqrt_q, lower=True)))
    mahal = np.sum(np.square(sla.solve_triangular(sqrt_p, diff, lower=True)))
    kl = 0.5 * (mahal - m - np.linalg.slogdet(S)[1] + np.linalg.slogdet(Kzz)[1] + trace)
    # predict: variational_families.py:234-285 (vectorised over the batch, diagonal only)
    Lz = sqrt_p
    Kzt = cross_covariance(kind, Z, X, lengthscale, variance)
    Lz_inv_Kzt = sla.solve_triangular(Lz, Kzt, lower=True)
    Kzz_inv_Kzt = sla.solve_triangular(Lz.T, Lz_inv_Kzt, lower=False)
    Ktz_Kzz_inv_sqrt = Kzz_inv_Kzt.T @ W
    mean = mean_const + Kzz_inv_Kzt.T @ (mu - mean_const)
    var = kernel_diagonal(kind, X, lengthscale, variance) - np.sum(Lz_inv_Kzt**2, axis=0) + np.sum(Ktz_Kzz_inv_sqrt**2, axis=1) + jitter
    # integrators.py:151-158
    ell = -0.5 * np.sum(np.log(2.0 * np.pi) + np.log(s) + ((y - mean) ** 2 + var) / s)
    return float(ell * num_datapoints / b - kl)


def svgp_elbo_value_and_grad_autodiff(kind, X, y, Z, lengthscale, variance, obs_stddev, mean_const, var_mean, var_sqrt,
                                      num_datapoints, jitter=1e-6):
    """torch-CPU reverse mode of the same restatement.  The gradient w.r.t. var_sqrt is reported on the
    lower triangle (the parameter is LowerTriangular)."""
    import torch

    kind = _kind_id(kind)
    t = lambda a: torch.tensor(np.asarray(a, np.float64), dtype=torch.float64, requires_grad=True)
    Xt = torch.tensor(np.asarray(X, np.float64))
    yt = torch.tensor(np.asarray(y, np.float64).reshape(-1))
    ell, var, sn, c, Zt = t(lengthscale), t(variance), t(obs_stddev), t(mean_const), t(Z)
    mu, Wp = t(np.asarray(var_mean).reshape(-1)), t(var_sqrt)
    W = torch.tril(Wp)
    m, b = Zt.shape[0], Xt.shape[0]
    s = sn**2
    eye = torch.eye(m, dtype=torch.float64)
    Kzz = _t_cross(torch, kind, Zt, Zt, ell, var) + eye * jitter
    S = W @ W.T
    Lz = torch.linalg.cholesky(Kzz)
    Lq = torch.linalg.cholesky(S)
    diff = c - mu
    trace = torch.sum(torch.linalg.solve_triangular(Lz, Lq, upper=False) ** 2)
    mahal = torch.sum(torch.linalg.solve_triangular(Lz, diff[:, None], upper=False) ** 2)
    kl = 0.5 * (mahal - m - torch.linalg.slogdet(S)[1] + torch.linalg.slogdet(Kzz)[1] + trace)
    Kzt = _t_cross(torch, kind, Zt, Xt, ell, var)
    A = torch.linalg.solve_triangular(Lz, Kzt, upper=False)
    KiK = torch.linalg.solve_triangular(Lz.T, A, upper=True)
    R = KiK.T @ W
    mean = c + KiK.T @ (mu - c)
    vpt = var - torch.sum(A**2, dim=0) + torch.sum(R**2, dim=1) + jitter
    ellv = -0.5 * torch.sum(math.log(2.0 * math.pi) + torch.log(s) + ((yt - mean) ** 2 + vpt) / s)
    val = ellv * num_datapoints / b - kl
    val.backward()
    g = {
        "lengthscale": _grad_of(ell),
        "variance": _grad_of(var), "obs_stddev": float(sn.grad), "mean_const": float(c.grad),
        "inducing_inputs": Zt.grad.numpy().copy(), "variational_mean": mu.grad.numpy().copy(),
        "variational_root_covariance": np.tril(Wp.grad.numpy()),
    }
    return float(val.detach()), g


def svgp_predict(kind, T, Z, lengthscale, variance, mean_const, var_mean, var_sqrt, jitter=1e-6):
    """VariationalGaussian.predict (variational_families.py:234-285): mean[T], cov[T,T]."""
    kind = _kind_id(kind)
    T, Z = np.asarray(T, np.float64), np.asarray(Z, np.float64)
    mu = np.asarray(var_mean, np.float64).reshape(-1)
    W = np.tril(np.asarray(var_sqrt, np.float64))
    Lz = np.linalg.cholesky(add_jitter(gram(kind, Z, lengthscale, variance), jitter))
    Ktt = gram(kind, T, lengthscale, variance)
    Kzt = cross_covariance(kind, Z, T, lengthscale, variance)
    A = sla.solve_triangular(Lz, Kzt, lower=True)
    KiK = sla.solve_triangular(Lz.T, A, lower=False)
    R = KiK.T @ W
    mean = mean_const + KiK.T @ (mu - mean_const)
    cov = Ktt - A.T @ A + R @ R.T
    return mean, add_jitter(cov, jitter)


def collapsed_predict(kind, X, y, T, Z, lengthscale, variance, obs_stddev, mean_const=0.0, jitter=1e-6):
    """CollapsedVariationalGaussian.predict (variational_families.py:786-870): mean[T], cov[T,T]."""
    kind = _kind_id(kind)
    X, T, Z = np.asarray(X, np.float64), np.asarray(T, np.float64), np.asarray(Z, np.float64)
    y = np.asarray(y, np.float64).reshape(-1)
    m = Z.shape[0]
    noise = obs_stddev**2
    Kzx = cross_covariance(kind, Z, X, lengthscale, variance)
    Lz = np.linalg.cholesky(add_jitter(gram(kind, Z, lengthscale, variance), jitter))
    Lz_inv_Kzx = sla.solve_triangular(Lz, Kzx, lower=True)
    A = Lz_inv_Kzx / obs_stddev
    L = np.linalg.cholesky(np.eye(m) + A @ A.T)
    diff = y - mean_const
    Lz_inv_Kzx_diff = sla.cho_solve((L, True), Lz_inv_Kzx @ diff)
    Kzz_inv_Kzx_diff = sla.solve_triangular(Lz.T, Lz_inv_Kzx_diff, lower=False)
    Ktt = gram(kind, T, lengthscale, variance)
    Kzt = cross_covariance(kind, Z, T, lengthscale, variance)
    Lz_inv_Kzt = sla.solve_triangular(Lz, Kzt, lower=True)
    L_inv_Lz_inv_Kzt = sla.solve_triangular(L, Lz_inv_Kzt, lower=True)
    mean = mean_const + (Kzt.T / noise) @ Kzz_inv_Kzx_diff
    cov = Ktt - Lz_inv_Kzt.T @ Lz_inv_Kzt + L_inv_Lz_inv_Kzt.T @ L_inv_Lz_inv_Kzt
    return mean, add_jitter(cov, jitter)


# ----------------------------------------------------------------------------------------
# f-4: conjugate_loocv -- gpjax/objectives.py:161-178
# ----------------------------------------------------------------------------------------
def conjugate_loocv(kind, X, y, lengthscale, variance, obs_stddev, mean_const=0.0, jitter=1e-6, sigma=None):
    """Literal order of objectives.py:161-178 (LU solve + explicit inverse).  `sigma` overrides the covariance
    (used for combination kernels assembled by the caller)."""
    X = np.asarray(X, np.float64)
    y = np.asarray(y, np.float64).reshape(-1, 1)
    n = X.shape[0]
    mx = np.ones((n, 1)) * mean_const
    if sigma is None:
        sigma = gram(kind, X, lengthscale, variance) + np.eye(n) * (obs_stddev**2 + jitter)  # :166-168
    sigma_inv_y = np.linalg.solve(sigma, y - mx)  # :171
    sigma_inv_diag = np.diag(np.linalg.inv(sigma))[:, None]  # :172-173
    loocv_means = mx + (y - mx) - sigma_inv_y / sigma_inv_diag  # :175
    loocv_stds = np.sqrt(1.0 / sigma_inv_diag)  # :176
    z = (y - loocv_means) / loocv_stds
    return float(np.sum(-0.5 * z * z - np.log(loocv_stds) - 0.5 * np.log(2.0 * np.pi)))  # :177-178


def conjugate_loocv_value_and_grad_autodiff(kind, X, y, lengthscale, variance, obs_stddev, mean_const=0.0, jitter=1e-6):
    """Reverse-mode gradient of the literal restatement of objectives.py:161-178."""
    import torch

    kind = _kind_id(kind)
    t = lambda a: torch.tensor(np.asarray(a, np.float64), dtype=torch.float64, requires_grad=True)
    Xt = torch.tensor(np.asarray(X, np.float64))
    yt = torch.tensor(np.asarray(y, np.float64).reshape(-1, 1))
    ell, var, sn, c = t(lengthscale), t(variance), t(obs_stddev), t(mean_const)
    n = Xt.shape[0]
    mx = torch.ones((n, 1), dtype=torch.float64) * c
    sigma = _t_cross(torch, kind, Xt, Xt, ell, var) + torch.eye(n, dtype=torch.float64) * (sn**2 + jitter)
    sigma_inv_y = torch.linalg.solve(sigma, yt - mx)
    sigma_inv_diag = torch.diagonal(torch.linalg.inv(sigma))[:, None]
    means = mx + (yt - mx) - sigma_inv_y / sigma_inv_diag
    stds = torch.sqrt(1.0 / sigma_inv_diag)
    z = (yt - means) / stds
    val = torch.sum(-0.5 * z * z - torch.log(stds) - 0.5 * math.log(2.0 * math.pi))
    val.backward()
    g = {"lengthscale": _grad_of(ell), "variance": _grad_of(var), "obs_stddev": float(sn.grad), "mean_const": float(c.grad)}
    return float(val.detach()), g
