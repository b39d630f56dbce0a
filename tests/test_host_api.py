"""Host-side mirror of the reference interface (no GPU needed): constructor validation, containers,
parameter routing in fit -- modelled on the reference's tests/test_kernels/test_stationary.py:133-183,
tests/test_fit.py:193-257,512-690, tests/test_linalg.py:491-558, tests/test_variational_families.py:236-262."""
import numpy as np
import pytest
import torch

import gpjax_b200 as gpx
from gpjax_b200.parameters import (DEFAULT_BIJECTION, NonNegativeReal, Parameter, PositiveReal, Real,
                                   SoftplusTransform, transform)

CPU = torch.device("cpu")
KERNELS = [gpx.kernels.RBF, gpx.kernels.Matern12, gpx.kernels.Matern32, gpx.kernels.Matern52]


@pytest.mark.parametrize("K", KERNELS)
def test_kernel_ctor_validation(K):
    with pytest.raises(ValueError):
        K(lengthscale=-1.0)
    with pytest.raises(ValueError):
        K(variance=-1.0)
    with pytest.raises(ValueError):
        K(lengthscale=np.ones((2, 2)))
    with pytest.raises(TypeError):
        K(lengthscale="one")
    with pytest.raises(ValueError):
        K(lengthscale=[1.0, 2.0], n_dims=3)
    k = K(lengthscale=[0.1, 0.2])
    assert k.n_dims == 2 and isinstance(k.lengthscale, PositiveReal) and isinstance(k.variance, NonNegativeReal)
    assert K(active_dims=[0, 2]).n_dims == 2
    assert K().name in ("RBF", "Matérn12", "Matérn32", "Matérn52")
    assert isinstance(K().compute_engine, gpx.kernels.DenseKernelComputation)


def test_engine_is_swappable_and_has_no_fallback():
    k = gpx.kernels.RBF()
    eng = gpx.kernels.DenseKernelComputation()
    k.compute_engine = eng
    assert k.compute_engine is eng

    class Periodic(gpx.kernels.StationaryKernel):
        name = "Periodic"

    with pytest.raises(NotImplementedError):
        Periodic().gram(torch.zeros((3, 1), dtype=torch.float64))


def test_compute_on_cpu_tensor_fails_loudly():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        gpx.kernels.RBF().gram(torch.zeros((3, 1), dtype=torch.float64))


def test_dataset_checks():
    X, y = torch.zeros((5, 2), dtype=torch.float64), torch.zeros((5, 1), dtype=torch.float64)
    D = gpx.Dataset(X=X, y=y)
    assert D.n == 5 and D.in_dim == 2 and D.is_supervised()
    assert (D + D).n == 10
    with pytest.raises(ValueError):
        gpx.Dataset(X=X, y=torch.zeros((4, 1), dtype=torch.float64))
    with pytest.raises(ValueError):
        gpx.Dataset(X=torch.zeros(5, dtype=torch.float64), y=y)
    with pytest.raises(ValueError):
        gpx.Dataset(X=X, y=torch.zeros(5, dtype=torch.float64))
    with pytest.warns(UserWarning):
        gpx.Dataset(X=X.float(), y=y)


def test_add_jitter_and_psd():
    from gpjax_b200.linalg import PSD, Dense, add_jitter, psd

    M = torch.eye(3, dtype=torch.float64)
    assert torch.equal(add_jitter(M, 0.5), 1.5 * torch.eye(3, dtype=torch.float64))
    with pytest.raises(ValueError):
        add_jitter(torch.zeros((2, 3), dtype=torch.float64))
    with pytest.raises(ValueError):
        add_jitter(torch.zeros(3, dtype=torch.float64))
    assert PSD in psd(Dense(M)).annotations


def test_containers_and_posterior_construction():
    prior = gpx.gps.Prior(mean_function=gpx.mean_functions.Zero(), kernel=gpx.kernels.RBF())
    assert prior.jitter == 1e-6
    lik = gpx.likelihoods.Gaussian(num_datapoints=10)
    assert float(lik.obs_stddev.value) == 1.0 and isinstance(lik.obs_stddev, NonNegativeReal)
    post = prior * lik
    assert isinstance(post, gpx.gps.ConjugatePosterior) and isinstance(lik * prior, gpx.gps.ConjugatePosterior)
    q = gpx.variational_families.CollapsedVariationalGaussian(posterior=post, inducing_inputs=np.zeros((7, 1)))
    assert q.num_inducing == 7 and isinstance(q.inducing_inputs, Real) and q.jitter == 1e-6

    class NotGaussian(gpx.likelihoods.AbstractLikelihood):
        pass

    bad = gpx.gps.AbstractPosterior(prior, NotGaussian(10))
    with pytest.raises(TypeError):
        gpx.variational_families.CollapsedVariationalGaussian(posterior=bad, inducing_inputs=np.zeros((7, 1)))
    m = gpx.mean_functions.Constant(Real(1.5))
    assert m(torch.zeros((4, 2), dtype=torch.float64)).shape == (4, 1)


def test_softplus_bijection_roundtrip():
    sp = SoftplusTransform()
    y = torch.tensor([1e-3, 0.3, 1.0, 25.0], dtype=torch.float64)
    assert torch.allclose(sp(sp.inv(y)), y, rtol=1e-14)
    p = {"a": PositiveReal(torch.tensor([1.0], dtype=torch.float64, device=CPU))}
    out = transform(p, DEFAULT_BIJECTION)
    assert abs(float(out["a"].value) - 1.3132617) < 1e-6  # docstring example of gpjax/parameters.py:33-35


def _toy_model():
    k = gpx.kernels.RBF(lengthscale=PositiveReal(torch.tensor(2.0, dtype=torch.float64, device=CPU)),
                        variance=NonNegativeReal(torch.tensor(3.0, dtype=torch.float64, device=CPU)))
    prior = gpx.gps.Prior(mean_function=gpx.mean_functions.Constant(Real(torch.tensor(1.0, dtype=torch.float64, device=CPU))),
                          kernel=k)
    lik = gpx.likelihoods.Gaussian(10, NonNegativeReal(torch.tensor(0.5, dtype=torch.float64, device=CPU)))
    return prior * lik


def _toy_objective(p, d):  # quadratic bowl in constrained space; exercises the bijection chain rule
    return ((p.prior.kernel.lengthscale.value - 1.0) ** 2 + (p.prior.kernel.variance.value - 1.0) ** 2
            + (p.likelihood.obs_stddev.value - 1.0) ** 2 + (p.prior.mean_function.constant.value - 0.0) ** 2).sum()


def test_fit_host_logic_and_trainable_filters():
    D = gpx.Dataset(X=torch.zeros((10, 1), dtype=torch.float64), y=torch.zeros((10, 1), dtype=torch.float64))
    post = _toy_model()
    opt, hist = gpx.fit(model=post, objective=_toy_objective, train_data=D, optim=gpx.optim.adam(0.1), num_iters=15,
                        verbose=False)
    assert isinstance(opt, gpx.gps.ConjugatePosterior) and hist.shape == (15,) and hist[-1] < hist[0]
    assert float(post.prior.kernel.lengthscale.value) == 2.0  # the input model is not mutated
    # freeze everything but PositiveReal (lengthscale): tests/test_fit.py:649-690
    opt, _ = gpx.fit(model=post, objective=_toy_objective, train_data=D, optim=gpx.optim.adam(0.1), num_iters=5,
                     trainable=PositiveReal, verbose=False)
    assert float(opt.prior.kernel.variance.value) == 3.0 and float(opt.likelihood.obs_stddev.value) == 0.5
    assert float(opt.prior.kernel.lengthscale.value) != 2.0
    # predicate filter: variance frozen (tests/test_fit.py:512-545)
    opt, _ = gpx.fit(model=post, objective=_toy_objective, train_data=D, optim=gpx.optim.sgd(0.1), num_iters=5,
                     trainable=lambda path, p: "variance" not in path, verbose=False)
    assert float(opt.prior.kernel.variance.value) == 3.0 and float(opt.prior.mean_function.constant.value) != 1.0
    # Zero mean is never trained (tests/test_fit.py:548-646)
    post.prior.mean_function = gpx.mean_functions.Zero()
    assert all("mean_function" not in n for n, _ in post.named_parameters())


def test_fit_argument_checks():
    D = gpx.Dataset(X=torch.zeros((4, 1), dtype=torch.float64), y=torch.zeros((4, 1), dtype=torch.float64))
    post, opt = _toy_model(), gpx.optim.adam(0.1)
    kw = dict(model=post, objective=_toy_objective, train_data=D, optim=opt, verbose=False)
    with pytest.raises(TypeError):
        gpx.fit(**{**kw, "model": object()})
    with pytest.raises(TypeError):
        gpx.fit(**{**kw, "train_data": (1, 2)})
    with pytest.raises(TypeError):
        gpx.fit(**{**kw, "optim": 3})
    with pytest.raises(ValueError):
        gpx.fit(**kw, num_iters=0)
    with pytest.raises(TypeError):
        gpx.fit(**kw, num_iters=1.5)
    with pytest.raises(ValueError):
        gpx.fit(**kw, batch_size=0)
    with pytest.raises(ValueError):
        gpx.fit(**kw, log_rate=0)
    with pytest.raises(TypeError):
        gpx.fit(**{**kw, "verbose": "yes"})


def test_fit_scipy_host_logic():
    D = gpx.Dataset(X=torch.zeros((4, 1), dtype=torch.float64), y=torch.zeros((4, 1), dtype=torch.float64))
    opt, hist = gpx.fit_scipy(model=_toy_model(), objective=_toy_objective, train_data=D, verbose=False)
    assert hist[-1] < 1e-8 and abs(float(opt.prior.kernel.lengthscale.value) - 1.0) < 1e-4


def test_adam_matches_optax_semantics():
    """One step of adam from zero state moves every coordinate by -lr * sign(g) (bias-corrected)."""
    opt = gpx.optim.adam(0.01)
    p = {"a": torch.tensor([1.0, -2.0], dtype=torch.float64)}
    g = {"a": torch.tensor([0.5, -3.0], dtype=torch.float64)}
    upd, st = opt.update(g, opt.init(p), p)
    assert torch.allclose(upd["a"], torch.tensor([-0.01, 0.01], dtype=torch.float64), atol=1e-9)
    assert st["count"] == 1


def test_fit_lbfgs_host_logic():
    D = gpx.Dataset(X=torch.zeros((4, 1), dtype=torch.float64), y=torch.zeros((4, 1), dtype=torch.float64))
    opt, final = gpx.fit_lbfgs(model=_toy_model(), objective=_toy_objective, train_data=D, max_iters=50)
    assert final.item() < 1e-10 and abs(float(opt.prior.kernel.lengthscale.value) - 1.0) < 1e-5
    with pytest.raises(ValueError):
        gpx.fit_lbfgs(model=_toy_model(), objective=_toy_objective, train_data=D, max_iters=0)


def test_lower_triangular_parameter_and_bijection():
    from gpjax_b200.parameters import FillTriangularTransform, LowerTriangular

    L = torch.tril(torch.arange(1.0, 10.0, dtype=torch.float64).reshape(3, 3))
    p = LowerTriangular(L)
    assert p.tag == "lower_triangular"
    with pytest.raises(ValueError):
        LowerTriangular(torch.ones((3, 3), dtype=torch.float64))
    with pytest.raises(ValueError):
        LowerTriangular(torch.ones((2, 3), dtype=torch.float64))
    bij = FillTriangularTransform()
    v = bij.inv(L)
    assert v.shape == (6,) and torch.equal(bij(v), L)
    q = gpx.variational_families.VariationalGaussian(
        posterior=_toy_model(), inducing_inputs=torch.zeros((4, 1), dtype=torch.float64))
    assert q.variational_mean.value.shape == (4, 1) and torch.equal(q.variational_root_covariance.value,
                                                                    torch.eye(4, dtype=torch.float64))
    assert sorted(n for n, _ in q.named_parameters())[:2] == ["inducing_inputs", "posterior.likelihood.obs_stddev"]


# ---- kernels beyond RBF / Matern and kernel algebra -------------------------------------------
def test_shape_parameter_kernels_and_packing():
    from gpjax_b200.parameters import SigmoidBounded

    rq = gpx.kernels.RationalQuadratic(lengthscale=[0.5, 0.7], variance=2.0, alpha=0.3)
    assert rq.alpha == 0.3 and rq.n_dims == 2  # stored as given (rational_quadratic.py:72): not trainable by default
    assert torch.equal(rq.kernel_scalars(), torch.tensor([2.0, 0.3], dtype=torch.float64))
    assert set(dict(rq.named_parameters())) == {"lengthscale", "variance"}
    pe = gpx.kernels.PoweredExponential(power=SigmoidBounded(0.4))
    assert "power" in dict(pe.named_parameters()) and pe.power.tag == "sigmoid"
    assert torch.equal(pe.kernel_scalars(), torch.tensor([1.0, 0.4], dtype=torch.float64))
    with pytest.raises(ValueError):
        SigmoidBounded(1.5)
    per = gpx.kernels.Periodic(period=PositiveReal(2.0))
    assert per.kernel_scalars().tolist() == [1.0, 2.0]
    w = gpx.kernels.White(variance=0.1)
    assert isinstance(w.compute_engine, gpx.kernels.ConstantDiagonalKernelComputation)
    assert w.kernel_scalars().numel() == 1 and w.lengthscale.value.item() == 1.0
    assert gpx.kernels.RBF().kernel_scalars().numel() == 1
    u = torch.tensor([-3.0, 0.0, 2.0], dtype=torch.float64)
    sg = DEFAULT_BIJECTION["sigmoid"]
    assert torch.allclose(sg.inv(sg(u)), u, atol=1e-12)


def test_kernel_algebra_builds_flattened_combinations():
    k1, k2, k3 = gpx.kernels.RBF(), gpx.kernels.Matern32(), gpx.kernels.White()
    s = k1 + k2 + k3
    assert isinstance(s, gpx.kernels.CombinationKernel) and s.operator_name == "sum" and s.kernels == [k1, k2, k3]
    p = k1 * k2 * k3
    assert p.operator_name == "prod" and p.kernels == [k1, k2, k3]
    m = k1 * k2 + k3
    assert m.operator_name == "sum" and len(m.kernels) == 2 and m.kernels[0].operator_name == "prod"
    c = k1 + 2.0  # scalars become Constant kernels (kernels/base.py:150-165)
    assert isinstance(c.kernels[1], gpx.kernels.Constant) and c.kernels[1].constant.value.item() == 2.0
    assert (3.0 + k1).operator_name == "sum"
    names = set(dict(m.named_parameters()))
    assert {"kernels[0].kernels[0].lengthscale", "kernels[0].kernels[1].variance", "kernels[1].variance"} <= names
    with pytest.raises(TypeError):
        gpx.kernels.ProductKernel(kernels=[k1, "rbf"])
    with pytest.raises(NotImplementedError):  # a combination has no single fused epilogue: the sparse objectives refuse
        from gpjax_b200.objectives import _kernel_args
        _kernel_args(s)


def test_lbfgs_minimize_zoom_linesearch_reaches_the_scipy_optimum():
    """optim.lbfgs_minimize restates the optimiser gpjax/fit.py:259-361 assembles from optax (L-BFGS memory 10 + strong-Wolfe zoom
    search from step 1, loop `n == 0 or (n < max_iters and |g| >= gtol)`).  Checked on Rosenbrock (curved valley: the line search
    must bracket and zoom), an ill-conditioned quadratic (the two-loop recursion must pick up curvature) and a softplus-transformed
    GP-like objective; SciPy's L-BFGS-B is the independent reference for the optimum."""
    from scipy.optimize import minimize

    from gpjax_b200.optim import lbfgs_minimize

    def rosen(x):
        f = float(np.sum(100.0 * (x[1:] - x[:-1] ** 2) ** 2 + (1 - x[:-1]) ** 2))
        g = np.zeros_like(x)
        g[:-1] = -400.0 * x[:-1] * (x[1:] - x[:-1] ** 2) - 2 * (1 - x[:-1])
        g[1:] += 200.0 * (x[1:] - x[:-1] ** 2)
        return f, g

    x, f, g, n = lbfgs_minimize(rosen, np.array([-1.2, 1.0, -0.5, 0.8]), max_iters=200, gtol=1e-8)
    assert np.allclose(x, 1.0, atol=1e-6) and f < 1e-14 and np.linalg.norm(g) < 1e-8 and n < 120

    A = np.diag(np.logspace(0, 3, 12))
    Q = np.linalg.qr(np.random.default_rng(0).standard_normal((12, 12)))[0]
    H = Q @ A @ Q.T
    b = np.arange(12.0)
    quad = lambda x: (float(0.5 * x @ H @ x - b @ x), H @ x - b)
    x, f, g, n = lbfgs_minimize(quad, np.zeros(12), max_iters=300, gtol=1e-7)
    # float64 stops resolving the sufficient-decrease test around |g| ~ 1e-6 here (f ~ -30): the search then returns step 0 and the loop ends
    assert np.linalg.norm(x - np.linalg.solve(H, b)) <= 1e-6 * np.linalg.norm(np.linalg.solve(H, b)) and np.linalg.norm(g) < 1e-5

    rng = np.random.default_rng(1)
    Xs = np.sort(rng.uniform(-3, 3, 60))
    ys = np.sin(Xs) + 0.1 * rng.standard_normal(60)

    def nll(u):  # exact GP negative MLL in softplus-unconstrained (lengthscale, variance, noise), numerical gradient-free check below
        sp = np.log1p(np.exp(u))
        ell, var, sn = sp
        r2 = (Xs[:, None] - Xs[None, :]) ** 2 / ell**2
        K = var * np.exp(-0.5 * r2)
        S = K + (sn**2 + 1e-6) * np.eye(60)
        L = np.linalg.cholesky(S)
        a = np.linalg.solve(S, ys)
        f = 0.5 * ys @ a + np.sum(np.log(np.diag(L))) + 30 * np.log(2 * np.pi)
        Si = np.linalg.inv(S)
        Wm = 0.5 * (Si - np.outer(a, a))
        dS = [K * r2 / ell, K / var, 2 * sn * np.eye(60)]
        g = np.array([np.sum(Wm * d) for d in dS]) * (1.0 / (1.0 + np.exp(-u)))
        return float(f), g

    u0 = np.log(np.expm1(np.array([1.0, 1.0, 1.0])))
    x, f, g, n = lbfgs_minimize(nll, u0, max_iters=100, gtol=1e-6)
    ref = minimize(nll, u0, jac=True, method="L-BFGS-B", options={"gtol": 1e-9, "ftol": 1e-15})
    assert abs(f - ref.fun) <= 1e-8 * abs(ref.fun) and np.linalg.norm(g) < 1e-5 and n <= 100
    # the loop condition of the reference: at least one iteration even when the gradient is already below gtol
    x1, f1, g1, n1 = lbfgs_minimize(quad, np.linalg.solve(H, b), max_iters=5, gtol=1e3)
    assert n1 == 1



def test_fit_cuda_graph_argument_checks_on_cpu():
    """fit(cuda_graph=True) refuses what it cannot replay before touching the GPU: minibatches and CPU data."""
    post, opt = _toy_model(), gpx.optim.adam(0.1)
    D = gpx.Dataset(X=torch.zeros((10, 1), dtype=torch.float64), y=torch.zeros((10, 1), dtype=torch.float64))
    with pytest.raises(NotImplementedError):
        gpx.fit(model=post, objective=_toy_objective, train_data=D, optim=opt, num_iters=4, batch_size=2, verbose=False,
                cuda_graph=True)
    with pytest.raises(RuntimeError):
        gpx.fit(model=post, objective=_toy_objective, train_data=D, optim=opt, num_iters=4, verbose=False, cuda_graph=True)


def test_fit_cuda_graph_state_pairing():
    """The replayed step copies the new optimiser state over the old one leaf by leaf; a host-side leaf that moves is refused."""
    from gpjax_b200.fit import _tree_clone, _tree_tensors

    opt = gpx.optim.adam(0.1)
    p = {"a": torch.zeros(3, dtype=torch.float64), "b": torch.ones((), dtype=torch.float64)}
    st0 = opt.init(p)
    _, st1 = opt.update({k: torch.ones_like(v) for k, v in p.items()}, st0, p)
    pairs = _tree_tensors(st0, st1)
    assert len(pairs) == 5 and all(a.shape == b.shape for a, b in pairs)  # count, mu x 2, nu x 2
    c = _tree_clone(st1)
    assert c["count"] is not st1["count"] and float(c["count"]) == 1.0 and torch.equal(c["mu"]["a"], st1["mu"]["a"])
    with pytest.raises(TypeError):
        _tree_tensors({"count": 0}, {"count": 1})
    assert _tree_tensors((), ()) == []
