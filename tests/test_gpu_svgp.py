"""SVGP minibatch ELBO (gpjax/objectives.py:241-315) on the GPU vs the oracle's reverse-mode autodiff of the
literal restatement; VariationalGaussian.predict vs the oracle; minibatch fit through the public API."""
import numpy as np
import pytest
import torch

import oracle as o

pytestmark = pytest.mark.gpu
KINDS = [(0, "rbf"), (1, "matern32"), (2, "matern52"), (3, "matern12")]
TOL = 1e-8


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda")


def make(n, m, d, seed):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-2.0, 2.0, (n, d))
    y = np.sin(X[:, :1]) + 0.1 * rng.standard_normal((n, 1))
    Z = rng.uniform(-2.0, 2.0, (m, d))
    mu = rng.standard_normal((m, 1)) * 0.3
    W = np.tril(rng.standard_normal((m, m)) * 0.05) + 0.6 * np.eye(m)
    return X, y, Z, mu, W


@pytest.mark.parametrize("kind,name", KINDS)
@pytest.mark.parametrize("n,m,d,iso,block", [(50, 5, 1, True, 16), (500, 40, 3, False, 128), (2000, 300, 8, False, 700),
                                             (4096, 600, 16, False, 4096)])
def test_svgp_elbo_vs_autodiff_oracle(kind, name, n, m, d, iso, block):
    from gpjax_b200.svgp_ops import svgp_elbo_fused

    X, y, Z, mu, W = make(n, m, d, n + m)
    ell = np.array(0.9) if iso else np.linspace(0.8, 1.6, d)
    ref, gref = o.svgp_elbo_value_and_grad_autodiff(name, X, y, Z, ell, 1.2, 0.5, 0.1, mu, W, 1e6)
    p = {k: dev(v).requires_grad_(True) for k, v in dict(Z=Z, ell=ell, var=1.2, sn=0.5, c=0.1, mu=mu, W=W).items()}
    val = svgp_elbo_fused(kind, dev(X), dev(y), p["Z"], p["ell"], p["var"], p["sn"], p["c"], p["mu"], p["W"], 1e6, 1e-6, block)
    val.backward()
    assert abs(val.item() - ref) <= TOL * abs(ref)
    got = dict(inducing_inputs=p["Z"].grad, lengthscale=p["ell"].grad, variance=p["var"].grad, obs_stddev=p["sn"].grad,
               mean_const=p["c"].grad, variational_mean=p["mu"].grad.reshape(-1), variational_root_covariance=p["W"].grad)
    cond = float(np.linalg.cond(o.gram(name, Z, ell, 1.2) + 1e-6 * np.eye(m)))
    tol_g = TOL * max(1.0, cond / 1e4)  # see test_gpu_sgpr.check
    for k, b in gref.items():
        a, b = got[k].cpu().numpy().reshape(np.shape(b)), np.asarray(b)
        assert np.max(np.abs(a - b)) <= tol_g * max(np.max(np.abs(b)), 1e-6 * abs(ref)), (k, cond)


def test_svgp_elbo_at_the_benchmarked_inducing_count():
    """BASELINE config 5's M = 4096 (D = 16, Matern32) on a 2 x 4096-row minibatch: the statistics SYRK and the pass-2 product run
    on the int8 pipe (7 planes), the replicated M x M finish is the 4096-point one the benchmark times.  Value and every gradient
    against the oracle's autodiff (~25 s of host time)."""
    from gpjax_b200.svgp_ops import svgp_elbo_fused

    n, m, d = 8192, 4096, 16
    X, y, Z, mu, _ = make(n, m, d, 5)
    # a well-conditioned root covariance: a random unit-lower-triangular factor with O(1) row sums is exponentially ill-conditioned in
    # M (with make()'s 0.05 off-diagonals the W^-T term of dELBO/dW differs by 7e-7 between ANY two float64 evaluations at M = 4096)
    W = np.tril(np.random.default_rng(6).standard_normal((m, m)) * 0.002) + 0.6 * np.eye(m)
    ell = np.linspace(0.8, 1.6, d) * 2.0
    ref, gref = o.svgp_elbo_value_and_grad_autodiff("matern32", X, y, Z, ell, 1.0, 0.3, 0.1, mu, W, 5e7)
    p = {k: dev(v).requires_grad_(True) for k, v in dict(Z=Z, ell=ell, var=1.0, sn=0.3, c=0.1, mu=mu, W=W).items()}
    val = svgp_elbo_fused(1, dev(X), dev(y), p["Z"], p["ell"], p["var"], p["sn"], p["c"], p["mu"], p["W"], 5e7, 1e-6, 4096)
    val.backward()
    assert abs(val.item() - ref) <= TOL * abs(ref)
    got = dict(inducing_inputs=p["Z"].grad, lengthscale=p["ell"].grad, variance=p["var"].grad, obs_stddev=p["sn"].grad,
               mean_const=p["c"].grad, variational_mean=p["mu"].grad.reshape(-1), variational_root_covariance=p["W"].grad)
    cond = float(np.linalg.cond(o.gram("matern32", Z, ell, 1.0) + 1e-6 * np.eye(m)))
    tol_g = TOL * max(1.0, cond / 1e4)
    for k, b in gref.items():
        a, b = got[k].cpu().numpy().reshape(np.shape(b)), np.asarray(b)
        assert np.max(np.abs(a - b)) <= tol_g * max(np.max(np.abs(b)), 1e-6 * abs(ref)), (k, cond)


def test_svgp_api_fit_minibatch_and_predict():
    import gpjax_b200 as gpx

    X, y, Z, _, _ = make(3000, 30, 1, 7)
    D = gpx.Dataset(X=dev(X), y=dev(y))
    post = gpx.gps.Prior(mean_function=gpx.mean_functions.Constant(), kernel=gpx.kernels.Matern32()) * \
        gpx.likelihoods.Gaussian(num_datapoints=D.n)
    q = gpx.variational_families.VariationalGaussian(posterior=post, inducing_inputs=dev(Z))
    e0 = gpx.objectives.elbo(q, D)
    ref = o.svgp_elbo("matern32", X, y, Z, 1.0, 1.0, 1.0, 0.0, np.zeros(30), np.eye(30), 3000)
    assert abs(e0.item() - ref) <= TOL * abs(ref)
    neg = lambda p, d: -gpx.objectives.elbo(p, d)
    opt, hist = gpx.fit(model=q, objective=neg, train_data=D, optim=gpx.optim.adam(0.02), num_iters=40, batch_size=256,
                        key=3, verbose=False)     # tests/test_fit.py:298-339 of the reference: minibatched ELBO
    assert hist.shape == (40,) and float(hist[-5:].mean()) < float(hist[:5].mean())
    W = opt.variational_root_covariance.value
    assert torch.equal(W, torch.tril(W)) and not torch.equal(W, torch.eye(30, dtype=torch.float64, device="cuda"))
    # predict against the oracle at the optimised parameters
    T = np.linspace(-2, 2, 57).reshape(-1, 1)
    dist = opt.predict(dev(T))
    k = opt.posterior.prior.kernel
    mean, cov = o.svgp_predict("matern32", T, opt.inducing_inputs.value.cpu().numpy(), k.lengthscale.value.item(),
                               k.variance.value.item(), opt.posterior.prior.mean_function.constant.item()
                               if hasattr(opt.posterior.prior.mean_function.constant, "item")
                               else opt.posterior.prior.mean_function.constant.value.item(),
                               opt.variational_mean.value.cpu().numpy(), W.cpu().numpy())
    assert np.max(np.abs(dist.mean().cpu().numpy() - mean)) <= 1e-8 * max(1.0, np.abs(mean).max())
    assert np.max(np.abs(dist.covariance().cpu().numpy() - cov)) <= 1e-8
