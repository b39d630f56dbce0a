"""The reference-facing Python surface on the GPU: same names, argument meaning and return types as gpjax,
results checked against the oracle.  Modelled on the reference's tests/test_kernels/test_stationary.py:186-226,
tests/test_kernels/test_computation.py:43-81, tests/test_linalg.py:237-396, tests/test_objectives.py:47-85,170-199,
tests/test_gps.py:108-146, tests/test_fit.py:193-257."""
import numpy as np
import pytest
import scipy.linalg as sla
import torch

import oracle as o

pytestmark = pytest.mark.gpu
NAMES = {"RBF": "rbf", "Matern32": "matern32", "Matern52": "matern52", "Matern12": "matern12"}


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda")


def build_data(n, d, seed):  # tests/test_objectives.py:25-44
    rng = np.random.default_rng(seed)
    X = rng.uniform(-2.0, 2.0, (n, d))
    y = np.sin(X[:, :1]) + 0.1 * rng.standard_normal((n, 1))
    return X, y


@pytest.mark.parametrize("kname", list(NAMES))
@pytest.mark.parametrize("n,d,ell", [(1, 1, 0.1), (2, 1, 0.1), (5, 2, [0.1, 0.2]), (300, 3, [0.5, 1.0, 2.0])])
def test_gram_and_cross_covariance(kname, n, d, ell):
    import gpjax_b200 as gpx
    from gpjax_b200.linalg import PSD, Dense, LinearOperator

    k = getattr(gpx.kernels, kname)(lengthscale=ell, variance=0.1)
    x = np.linspace(0.0, 1.0, n * d).reshape(n, d)
    K = k.gram(dev(x))
    assert isinstance(K, LinearOperator) and isinstance(K, Dense) and PSD in K.annotations and K.shape == (n, n)
    Kd = K.to_dense().cpu().numpy()
    assert np.all(np.linalg.eigvalsh(Kd + 1e-6 * np.eye(n)) > 0)
    ref = o.gram(NAMES[kname], x, np.asarray(ell), 0.1)
    assert np.max(np.abs(Kd - ref) / np.abs(ref)) <= 1e-12
    b = np.linspace(-1.0, 2.0, 7 * d).reshape(7, d)
    Kxb = k.cross_covariance(dev(x), dev(b))
    assert isinstance(Kxb, torch.Tensor) and Kxb.shape == (n, 7)
    assert np.max(np.abs(Kxb.cpu().numpy() - o.cross_covariance(NAMES[kname], x, b, np.asarray(ell), 0.1))) <= 1e-13
    assert abs(k(dev(x[0]), dev(b[0])).item() - o.kernel_pair(NAMES[kname], x[0], b[0], np.asarray(ell), 0.1)) <= 1e-14
    assert torch.allclose(k.diagonal(dev(x)).diagonal, torch.full((n,), 0.1, dtype=torch.float64, device="cuda"))


def test_active_dims_slicing():
    import gpjax_b200 as gpx

    rng = np.random.default_rng(0)
    x = rng.uniform(-1, 1, (40, 4))
    k = gpx.kernels.RBF(active_dims=[0, 2], lengthscale=[0.5, 0.7])
    K = k.gram(dev(x)).to_dense().cpu().numpy()
    assert np.max(np.abs(K - o.gram("rbf", x[:, [0, 2]], np.array([0.5, 0.7]), 1.0))) <= 1e-13


@pytest.mark.parametrize("n", [1, 5, 200, 700])
def test_linalg_dispatch(n):
    from gpjax_b200.linalg import Dense, Diagonal, Identity, Triangular, logdet, lower_cholesky, psd, solve

    rng = np.random.default_rng(n)
    X = rng.uniform(-2, 2, (n, 3))
    S = o.gram("matern52", X, np.array([0.9, 1.0, 1.1]), 1.0) + 0.1 * np.eye(n)
    A = psd(Dense(dev(S)))
    L = lower_cholesky(A)
    assert isinstance(L, Triangular) and L.lower
    Ld = L.to_dense().cpu().numpy()
    assert np.allclose(Ld @ Ld.T, S, atol=1e-10)            # tests/test_linalg.py:237-252
    assert lower_cholesky(L) is L
    b, B = rng.standard_normal(n), rng.standard_normal((n, 3))
    assert np.allclose(S @ solve(A, dev(b)).cpu().numpy(), b, atol=1e-8)         # Dense, vector (squeeze rule)
    assert solve(A, dev(b)).shape == (n,) and solve(A, dev(B)).shape == (n, 3)
    assert np.allclose(S @ solve(A, dev(B)).cpu().numpy(), B, atol=1e-8)         # Dense, matrix
    assert np.allclose(Ld @ solve(L, dev(B)).cpu().numpy(), B, atol=1e-9)        # Triangular lower
    assert np.allclose(Ld.T @ solve(L.T, dev(B)).cpu().numpy(), B, atol=1e-9)    # Triangular.T (upper solve)
    U = Triangular(dev(np.ascontiguousarray(Ld.T)), lower=False)                 # genuinely upper storage
    assert np.allclose(Ld.T @ solve(U, dev(b)).cpu().numpy(), b, atol=1e-9)
    assert abs(logdet(A).item() - np.linalg.slogdet(S)[1]) <= 1e-9 * max(1, n)    # tests/test_linalg.py:364-396
    assert abs(logdet(L).item() - np.sum(np.log(np.diag(Ld)))) <= 1e-10 * max(1, n)  # no factor 2
    d = dev(rng.uniform(1, 2, n))
    assert torch.allclose(solve(Diagonal(d), dev(b)), dev(b) / d)
    assert torch.equal(solve(Identity(n, device="cuda"), dev(b)), dev(b))
    assert abs(logdet(Diagonal(d)).item() - float(torch.log(d).sum())) < 1e-12


@pytest.mark.parametrize("n", [3, 700, 2500])
def test_lower_cholesky_symmetrises_a_dense_input_like_jnp_cholesky(n):
    """jnp.linalg.cholesky(symmetrize_input=True) -- what gpjax/linalg/operations.py:54-55 calls -- factors (A + A^T) / 2; an
    upper Triangular goes through the same dense route (operations.py:40-43)."""
    from gpjax_b200.linalg import Dense, Triangular, lower_cholesky

    rng = np.random.default_rng(n)
    X = rng.uniform(-2, 2, (n, 2))
    S = o.gram("rbf", X, np.array([0.9, 1.1]), 1.0) + 0.3 * np.eye(n)
    A0 = S + np.triu(1e-3 * rng.standard_normal((n, n)), 1)
    L = lower_cholesky(Dense(dev(A0))).to_dense().cpu().numpy()
    ref = np.linalg.cholesky(0.5 * (A0 + A0.T))
    assert np.max(np.abs(L - ref)) <= 1e-11 * np.abs(ref).max()
    assert np.max(np.abs(L - np.linalg.cholesky(S))) > 1e-7
    Ad = dev(A0).requires_grad_(True)  # differentiable route: same factor, symmetric gradient
    Lg = lower_cholesky(Dense(Ad)).to_dense()
    assert np.max(np.abs(Lg.detach().cpu().numpy() - ref)) <= 1e-11 * np.abs(ref).max()
    Lg.sum().backward()
    assert torch.allclose(Ad.grad, Ad.grad.T)
    U = Triangular(dev(np.triu(S)), lower=False)
    Lu = lower_cholesky(U).to_dense().cpu().numpy()
    Ud = np.triu(S)
    assert np.max(np.abs(Lu - np.linalg.cholesky(0.5 * (Ud + Ud.T)))) <= 1e-11


@pytest.mark.parametrize("kname", list(NAMES))
@pytest.mark.parametrize("n,d", [(1, 1), (2, 2), (10, 3), (200, 2)])
def test_conjugate_mll_api(kname, n, d):
    import gpjax_b200 as gpx
    from gpjax_b200.parameters import Real

    X, y = build_data(n, d, n + d)
    D = gpx.Dataset(X=dev(X), y=dev(y))
    prior = gpx.gps.Prior(mean_function=gpx.mean_functions.Constant(Real(0.3)), kernel=getattr(gpx.kernels, kname)())
    post = prior * gpx.likelihoods.Gaussian(num_datapoints=n)
    v = gpx.objectives.conjugate_mll(post, D)
    assert isinstance(v, torch.Tensor) and v.shape == ()
    ref = o.conjugate_mll(NAMES[kname], X, y, 1.0, 1.0, 1.0, 0.3)
    assert abs(v.item() - ref) <= 1e-8 * abs(ref)
    # GaussianDistribution.log_prob through the generic linalg dispatch agrees (tests/test_gaussian_distribution.py)
    from gpjax_b200.distributions import GaussianDistribution
    from gpjax_b200.linalg import Dense, psd

    S = o.gram(NAMES[kname], X, 1.0, 1.0) + (1e-6 + 1.0) * np.eye(n)
    lp = GaussianDistribution(dev(np.full(n, 0.3)), psd(Dense(dev(S)))).log_prob(dev(y.reshape(-1)))
    assert abs(lp.item() - ref) <= 1e-8 * abs(ref)


def test_collapsed_elbo_api_and_identity_with_mll():
    import gpjax_b200 as gpx

    for n in (10, 20):  # tests/test_objectives.py:170-199
        X, y = build_data(n, 2, n)
        D = gpx.Dataset(X=dev(X), y=dev(y))
        post = gpx.gps.Prior(mean_function=gpx.mean_functions.Constant(), kernel=gpx.kernels.RBF()) * \
            gpx.likelihoods.Gaussian(num_datapoints=n)
        q = gpx.variational_families.CollapsedVariationalGaussian(posterior=post, inducing_inputs=dev(X))
        e = gpx.objectives.collapsed_elbo(q, D)
        m = gpx.objectives.conjugate_mll(post, D)
        assert e.shape == () and abs(e.item() - m.item()) <= 1e-5 * abs(m.item())
        assert abs(e.item() - o.collapsed_elbo("rbf", X, y, X, 1.0, 1.0, 1.0, 0.0)) <= 1e-8 * abs(e.item())


@pytest.mark.parametrize("kname", ["RBF", "Matern52"])
def test_conjugate_predict(kname):
    import gpjax_b200 as gpx

    X, y = build_data(150, 2, 3)
    T = np.random.default_rng(1).uniform(-2, 2, (33, 2))
    post = gpx.gps.Prior(mean_function=gpx.mean_functions.Zero(), kernel=getattr(gpx.kernels, kname)(lengthscale=[0.7, 1.2])) * \
        gpx.likelihoods.Gaussian(num_datapoints=150, obs_stddev=0.2)
    dist = post.predict(dev(T), gpx.Dataset(X=dev(X), y=dev(y)))
    mean, cov = o.conjugate_predict(NAMES[kname], X, y, T, np.array([0.7, 1.2]), 1.0, 0.2, 0.0)
    assert dist.mean().shape == (33,) and dist.covariance().shape == (33, 33)   # tests/test_gps.py:108-146
    assert np.max(np.abs(dist.mean().cpu().numpy() - mean)) <= 1e-8 * np.abs(mean).max()
    assert np.max(np.abs(dist.covariance().cpu().numpy() - cov)) <= 1e-8
    pred = post.likelihood(dist)
    assert np.allclose(pred.variance().cpu().numpy(), np.diag(cov) + 0.04, atol=1e-8)


def test_fit_regression_config1():
    """BASELINE config 1 shape: N=1000, D=1, RBF, conjugate_mll + gradient via gpx.fit; the objective must
    decrease and the end point must match the oracle evaluated at the same hyper-parameters."""
    import gpjax_b200 as gpx

    X, y = build_data(1000, 1, 123)
    D = gpx.Dataset(X=dev(X), y=dev(y))
    post = gpx.gps.Prior(mean_function=gpx.mean_functions.Zero(), kernel=gpx.kernels.RBF()) * \
        gpx.likelihoods.Gaussian(num_datapoints=D.n)
    neg = lambda p, d: -gpx.objectives.conjugate_mll(p, d)
    opt, hist = gpx.fit(model=post, objective=neg, train_data=D, optim=gpx.optim.adam(0.05), num_iters=30, verbose=False)
    assert isinstance(opt, gpx.gps.ConjugatePosterior) and hist.shape == (30,) and hist[-1] < hist[0]
    ell, var, sn = (opt.prior.kernel.lengthscale.value.item(), opt.prior.kernel.variance.value.item(),
                    opt.likelihood.obs_stddev.value.item())
    ref = -o.conjugate_mll("rbf", X, y, ell, var, sn, 0.0)
    assert abs(neg(opt, D).item() - ref) <= 1e-8 * abs(ref)
    # first history entry = objective at the initial parameters
    assert abs(hist[0].item() + o.conjugate_mll("rbf", X, y, 1.0, 1.0, 1.0, 0.0)) <= 1e-8 * abs(hist[0].item())


def test_fit_sgpr_trains_inducing_inputs():
    import gpjax_b200 as gpx
    from gpjax_b200.parameters import PositiveReal

    X, y = build_data(600, 1, 42)
    D = gpx.Dataset(X=dev(X), y=dev(y))
    post = gpx.gps.Prior(mean_function=gpx.mean_functions.Constant(), kernel=gpx.kernels.RBF()) * \
        gpx.likelihoods.Gaussian(num_datapoints=D.n)
    z0 = np.linspace(-2, 2, 12).reshape(-1, 1)
    q = gpx.variational_families.CollapsedVariationalGaussian(posterior=post, inducing_inputs=dev(z0))
    neg = lambda p, d: -gpx.objectives.collapsed_elbo(p, d)
    opt, hist = gpx.fit(model=q, objective=neg, train_data=D, optim=gpx.optim.adam(0.02), num_iters=20, verbose=False)
    assert hist[-1] < hist[0]
    assert not torch.allclose(opt.inducing_inputs.value, dev(z0))
    frozen, _ = gpx.fit(model=q, objective=neg, train_data=D, optim=gpx.optim.adam(0.02), num_iters=3,
                        trainable=PositiveReal, verbose=False)
    assert torch.equal(frozen.inducing_inputs.value, dev(z0))


def test_collapsed_predict_and_fit_lbfgs():
    """section 8f rank 2 (CollapsedVariationalGaussian.predict, variational_families.py:786-870) and fit_lbfgs
    (fit.py:259-361; tests/test_fit.py:226-257 of the reference: loss decreases, model type preserved)."""
    import gpjax_b200 as gpx

    X, y = build_data(900, 2, 11)
    T = np.random.default_rng(2).uniform(-2, 2, (41, 2))
    Z = np.random.default_rng(3).uniform(-2, 2, (25, 2))
    D = gpx.Dataset(X=dev(X), y=dev(y))
    post = gpx.gps.Prior(mean_function=gpx.mean_functions.Constant(0.3), kernel=gpx.kernels.Matern52(lengthscale=[0.9, 1.3])) * \
        gpx.likelihoods.Gaussian(num_datapoints=D.n, obs_stddev=0.4)
    q = gpx.variational_families.CollapsedVariationalGaussian(posterior=post, inducing_inputs=dev(Z))
    dist = q.predict(dev(T), D, block_rows=256)
    mean, cov = o.collapsed_predict("matern52", X, y, T, Z, np.array([0.9, 1.3]), 1.0, 0.4, 0.3)
    assert np.max(np.abs(dist.mean().cpu().numpy() - mean)) <= 1e-8 * np.abs(mean).max()
    assert np.max(np.abs(dist.covariance().cpu().numpy() - cov)) <= 1e-8
    neg = lambda p, d: -gpx.objectives.collapsed_elbo(p, d)
    before = neg(q, D).item()
    opt, final = gpx.fit_lbfgs(model=q, objective=neg, train_data=D, max_iters=15)
    assert isinstance(opt, gpx.variational_families.CollapsedVariationalGaussian)
    assert final.item() < before and abs(neg(opt, D).item() - final.item()) <= 1e-8 * abs(final.item())


def test_reference_regression_golden_on_gpu():
    """The CUDA path on the data of the reference's examples/regression.py (tests/golden/regression_example.npz):
    the objective at the stored optimum equals the reference's stored golden 55.07405622 (tests/integration_tests.py:
    99-107, tolerance there: 1.0), gpx.fit_scipy from the example's initial parameters gets there, and the predictive
    moments match the fixture."""
    import os

    import gpjax_b200 as gpx
    from gpjax_b200.parameters import Real

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "regression_example.npz"))
    D = gpx.Dataset(X=dev(g["x"]), y=dev(g["y"]))
    neg = lambda p, d: -gpx.objectives.conjugate_mll(p, d)

    def posterior(ell, var, sn, c):
        prior = gpx.gps.Prior(mean_function=gpx.mean_functions.Constant(Real(c)), kernel=gpx.kernels.RBF(lengthscale=ell, variance=var))
        return prior * gpx.likelihoods.Gaussian(num_datapoints=D.n, obs_stddev=sn)

    at_opt = neg(posterior(float(g["lengthscale"]), float(g["variance"]), float(g["obs_stddev"]), float(g["mean_const"])), D).item()
    assert abs(at_opt - float(g["neg_mll_at_optimum"])) <= 1e-8 * abs(at_opt)
    assert abs(at_opt - float(g["reference_golden_history_last"])) < 1e-4
    init = posterior(1.0, 1.0, 1.0, 0.0)
    assert abs(neg(init, D).item() - float(g["neg_mll_at_init"])) <= 1e-8 * float(g["neg_mll_at_init"])
    opt, hist = gpx.fit_scipy(model=init, objective=neg, train_data=D, verbose=False)
    assert abs(hist[-1].item() - float(g["reference_golden_history_last"])) < 1e-3
    pred = opt.likelihood(opt.predict(dev(g["xtest"]), D))
    assert abs(pred.mean().sum().item() - float(g["reference_golden_predictive_mean_sum"])) < 1.0
    assert abs(pred.stddev().sum().item() - float(g["reference_golden_predictive_std_sum"])) < 1.0


@pytest.mark.parametrize("n", [5, 300, 1300])
def test_linalg_free_functions_are_differentiable(n):
    """What jax.grad derives through lower_cholesky / solve / logdet (linalg/operations.py:22-181) and
    GaussianDistribution.log_prob (distributions.py:115-134): compared with torch-CPU autograd of the same expressions."""
    from gpjax_b200.distributions import GaussianDistribution
    from gpjax_b200.linalg import Dense, logdet, lower_cholesky, psd, solve

    rng = np.random.default_rng(n)
    Xn = rng.uniform(-2, 2, (n, 3))
    S0 = o.gram("rbf", Xn, np.array([0.8, 1.0, 1.3]), 1.0) + 0.3 * np.eye(n)
    b0, B0, y0, m0 = rng.standard_normal(n), rng.standard_normal((n, 4)), rng.standard_normal(n), rng.standard_normal(n)

    def expr(S, b, B, y, m, mine):
        if mine:
            L = lower_cholesky(psd(Dense(S)))
            w = solve(L, b)                       # 1-D right-hand side, squeeze rule
            V = solve(L.T, solve(L, B))           # Sigma^-1 B through the transposed view
            ld = logdet(L)
            lp = GaussianDistribution(m, psd(Dense(S))).log_prob(y)
            L = L.array
        else:
            L = torch.linalg.cholesky(S)
            w = torch.linalg.solve_triangular(L, b[:, None], upper=False)[:, 0]
            V = torch.cholesky_solve(B, L)
            ld = torch.log(torch.diagonal(L)).sum()
            d = y - m
            lp = -0.5 * (n * np.log(2 * np.pi) + torch.linalg.slogdet(S)[1] + d @ torch.linalg.solve(S, d))
        return (w * w).sum() + 0.3 * (V * V).sum() + 2.0 * ld + 0.7 * lp + (L * L).sum() * 0.01

    outs = []
    for mine in (True, False):
        mk = (lambda a: dev(a).requires_grad_(True)) if mine else (lambda a: torch.tensor(a, requires_grad=True))
        leaves = [mk(a) for a in (S0, b0, B0, y0, m0)]
        val = expr(*leaves, mine)
        val.backward()
        outs.append((val.item(), [t.grad.detach().cpu().numpy() for t in leaves]))
    (v1, g1), (v2, g2) = outs
    assert abs(v1 - v2) <= 1e-10 * abs(v2)
    g1[0], g2[0] = 0.5 * (g1[0] + g1[0].T), 0.5 * (g2[0] + g2[0].T)  # only the symmetric part of dSigma is defined
    for a, b in zip(g1, g2):
        assert np.max(np.abs(a - b)) <= 1e-9 * max(np.max(np.abs(b)), 1.0)
