"""SURVEY section 8f rows 3 and 4 on the GPU: RationalQuadratic / PoweredExponential / Periodic / White in the fused
Gram epilogue (values 1e-12, gradients 1e-8 against torch-CPU autodiff of the literal restatement), sum / product /
constant kernels assembled from fused launches, and conjugate_loocv.  Modelled on the reference's
tests/test_kernels/test_stationary.py:186-226, tests/test_kernels/test_base.py (combination kernels) and
tests/test_objectives.py:88-126 (loocv)."""
import math

import numpy as np
import pytest
import torch

import oracle as o
from oracle import gp_oracle as go

pytestmark = pytest.mark.gpu
# class name, kind id, oracle name, shape kwarg, shape value
EXT = [("RationalQuadratic", 4, "rational_quadratic", "alpha", 0.7), ("PoweredExponential", 5, "powered_exponential", "power", 0.6),
       ("Periodic", 6, "periodic", "period", 1.7), ("White", 7, "white", None, None)]


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda")


def data(n, d, seed, dup=True):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-2.0, 2.0, (n, d))
    if dup and n > 3:
        X[n // 2] = X[1]
    y = np.sin(X[:, :1]) + 0.1 * rng.standard_normal((n, 1))
    return X, y


def make(cls, shape_kw, shape, ell, var=1.3, trainable_shape=True):
    import gpjax_b200 as gpx
    from gpjax_b200.parameters import PositiveReal, SigmoidBounded

    K = getattr(gpx.kernels, cls)
    if shape_kw is None:
        return K(variance=var)
    sv = (SigmoidBounded(shape) if shape_kw == "power" else PositiveReal(shape)) if trainable_shape else shape
    return K(lengthscale=ell, variance=var, **{shape_kw: sv})


def scal(kind, var, shape):
    return np.array([var, shape]) if kind in (4, 5, 6) else var


@pytest.mark.parametrize("cls,kind,name,skw,sv", EXT)
@pytest.mark.parametrize("n,m,d,iso", [(1, 1, 1, True), (5, 3, 2, False), (300, 257, 3, False), (200, 130, 8, False),
                                       (150, 150, 33, True)])
def test_gram_values(cls, kind, name, skw, sv, n, m, d, iso):
    X, _ = data(n, d, n + d)
    Z, _ = data(m, d, m + 3 * d)
    if m > 2 and n > 2:
        Z[2] = X[1]
    ell = 0.9 if iso else np.linspace(0.7, 1.4, d)
    k = make(cls, skw, sv, ell, trainable_shape=False)
    ellv = 1.0 if kind == 7 else ell
    K = k.cross_covariance(dev(X), dev(Z)).cpu().numpy()
    ref = o.cross_covariance(name, X, Z, ellv, scal(kind, 1.3, sv))
    assert np.max(np.abs(K - ref)) <= 1e-12 * 1.3
    G = k.gram(dev(X))
    Gd = G.to_dense().cpu().numpy()
    if kind == 7:
        from gpjax_b200.linalg import Diagonal
        # constant-diagonal engine (computations/constant_diagonal.py:39-43): k(x0, x0) on the diagonal, whatever
        # rows coincide -- unlike the dense pairwise evaluation the cross-covariance above uses
        assert isinstance(G, Diagonal) and np.array_equal(Gd, 1.3 * np.eye(n))
    else:
        assert np.max(np.abs(Gd - o.gram(name, X, ellv, scal(kind, 1.3, sv)))) <= 1e-12 * 1.3
    dg = k.diagonal(dev(X)).diagonal.cpu().numpy()
    refd = np.array([float(o.kernel_pair(name, X[i], X[i], ellv, scal(kind, 1.3, sv))) for i in range(min(n, 4))])
    assert np.max(np.abs(dg[: len(refd)] - refd)) <= 1e-14


@pytest.mark.parametrize("cls,kind,name,skw,sv", EXT)
@pytest.mark.parametrize("iso", [False, True])
def test_gram_backward(cls, kind, name, skw, sv, iso):
    n, m, d = 130, 200, 3
    X, _ = data(n, d, 5)
    Z, _ = data(m, d, 6)
    Z[7] = X[3]
    ell = 0.9 if iso else np.array([0.7, 1.0, 1.3])
    k = make(cls, skw, sv, ell)
    Xd, Zd = dev(X).requires_grad_(True), dev(Z).requires_grad_(True)
    params = dict(k.named_parameters())
    for p in params.values():
        p.value.requires_grad_(True)
    W = np.random.default_rng(0).standard_normal((n, m))
    (dev(W) * k.cross_covariance(Xd, Zd)).sum().backward()
    ellv = 1.0 if kind == 7 else ell
    t = lambda a: torch.tensor(np.asarray(a, np.float64), requires_grad=True)
    Xt, Zt, et, vt = t(X), t(Z), t(ellv), t(scal(kind, 1.3, sv))
    (torch.tensor(W) * go._t_cross(torch, kind, Xt, Zt, et, vt)).sum().backward()

    def close(a, b, what):
        b = np.asarray(b, np.float64)
        a = a.detach().cpu().numpy().reshape(b.shape)
        assert np.max(np.abs(a - b)) <= 1e-9 * max(np.max(np.abs(b)), 1.0), what

    close(params["variance"].value.grad, vt.grad.numpy().reshape(-1)[0], "variance")
    if skw is not None:
        close(params[skw].value.grad, vt.grad.numpy()[1], skw)
    if kind != 7:
        close(params["lengthscale"].value.grad, et.grad.numpy(), "lengthscale")
        close(Xd.grad, Xt.grad.numpy(), "X")
        close(Zd.grad, Zt.grad.numpy(), "Z")


@pytest.mark.parametrize("cls,kind,name,skw,sv", EXT)
@pytest.mark.parametrize("n,d,iso", [(100, 3, False), (700, 2, True), (1500, 8, False)])
def test_conjugate_mll_fused(cls, kind, name, skw, sv, n, d, iso):
    import gpjax_b200 as gpx
    from gpjax_b200.parameters import Real

    X, y = data(n, d, n + d, dup=(kind != 7))
    ell = 0.9 if iso else np.linspace(0.8, 1.6, d)
    k = make(cls, skw, sv, ell)
    post = gpx.gps.Prior(mean_function=gpx.mean_functions.Constant(Real(0.2)), kernel=k) * gpx.likelihoods.Gaussian(
        num_datapoints=n, obs_stddev=0.4)
    params = dict(post.named_parameters())
    for p in params.values():
        p.value.requires_grad_(True)
    val = gpx.objectives.conjugate_mll(post, gpx.Dataset(X=dev(X), y=dev(y)))
    val.backward()
    ellv = 1.0 if kind == 7 else ell
    ref, gr = o.conjugate_mll_value_and_grad_autodiff(name, X, y, ellv, scal(kind, 1.3, sv), 0.4, 0.2)
    assert abs(val.item() - ref) <= 1e-8 * abs(ref)
    floor = 1e-6 * abs(ref)

    def close(pname, b):
        b = np.asarray(b, np.float64)
        a = params[pname].value.grad.cpu().numpy().reshape(b.shape)
        assert np.max(np.abs(a - b)) <= 1e-8 * max(np.max(np.abs(b)), floor), pname

    gv = np.atleast_1d(gr["variance"])
    close("prior.kernel.variance", gv[0])
    if skw is not None:
        close("prior.kernel." + skw, gv[1])
    if kind != 7:
        close("prior.kernel.lengthscale", gr["lengthscale"])
    close("likelihood.obs_stddev", gr["obs_stddev"])
    close("prior.mean_function.constant", gr["mean_const"])


@pytest.mark.parametrize("cls,kind,name,skw,sv", [e for e in EXT if e[1] in (4, 6)])
def test_collapsed_elbo_fused(cls, kind, name, skw, sv):
    import gpjax_b200 as gpx
    from gpjax_b200.parameters import Real

    n, m, d = 900, 40, 3
    X, y = data(n, d, 21, dup=False)
    Z = data(m, d, 22, dup=False)[0]
    ell = np.array([0.9, 1.2, 1.5])
    k = make(cls, skw, sv, ell)
    post = gpx.gps.Prior(mean_function=gpx.mean_functions.Constant(Real(0.2)), kernel=k) * gpx.likelihoods.Gaussian(
        num_datapoints=n, obs_stddev=0.4)
    q = gpx.variational_families.CollapsedVariationalGaussian(posterior=post, inducing_inputs=dev(Z))
    params = dict(q.named_parameters())
    for p in params.values():
        p.value.requires_grad_(True)
    val = gpx.objectives.collapsed_elbo(q, gpx.Dataset(X=dev(X), y=dev(y)), block_rows=256)
    val.backward()
    ref, gr = o.collapsed_elbo_value_and_grad_autodiff(name, X, y, Z, ell, np.array([1.3, sv]), 0.4, 0.2)
    assert abs(val.item() - ref) <= 1e-8 * abs(ref)
    floor = 1e-6 * abs(ref)
    got = {"lengthscale": "posterior.prior.kernel.lengthscale", "obs_stddev": "posterior.likelihood.obs_stddev",
           "mean_const": "posterior.prior.mean_function.constant", "inducing_inputs": "inducing_inputs"}
    for key, pname in got.items():
        b = np.asarray(gr[key], np.float64)
        a = params[pname].value.grad.cpu().numpy().reshape(b.shape)
        assert np.max(np.abs(a - b)) <= 1e-7 * max(np.max(np.abs(b)), floor), key
    gv = gr["variance"]
    assert abs(params["posterior.prior.kernel.variance"].value.grad.item() - gv[0]) <= 1e-7 * max(abs(gv[0]), floor)
    assert abs(params["posterior.prior.kernel." + skw].value.grad.item() - gv[1]) <= 1e-7 * max(abs(gv[1]), floor)
    # predictive moments at test inputs run through the same fused launches
    T = data(50, d, 23, dup=False)[0]
    pred = q.predict(dev(T), gpx.Dataset(X=dev(X), y=dev(y)))
    mref, cref = o.collapsed_predict(name, X, y, T, Z, ell, np.array([1.3, sv]), 0.4, 0.2)
    assert np.max(np.abs(pred.mean().cpu().numpy() - mref)) <= 1e-8 * max(np.max(np.abs(mref)), 1.0)
    assert np.max(np.abs(pred.covariance().cpu().numpy() - cref)) <= 1e-8 * max(np.max(np.abs(cref)), 1.0)


def test_powered_exponential_runs_the_composable_sparse_route():
    """The streamed SGPR statistics assume k(x, x) = variance; PoweredExponential's clamped diagonal is variance * exp(-(1e-18)^power)
    (stationary/utils.py:67), visibly smaller for power = 0.3, so it takes the composable route -- and matches the oracle, whose
    trace term evaluates the kernel at (x, x) pair by pair as objectives.py:356 does."""
    import gpjax_b200 as gpx

    n = 300
    X, y = data(n, 2, 1)
    k = gpx.kernels.PoweredExponential(lengthscale=[0.9, 1.2], variance=1.3, power=0.3)
    post = gpx.gps.Prior(mean_function=gpx.mean_functions.Zero(), kernel=k) * gpx.likelihoods.Gaussian(num_datapoints=n, obs_stddev=0.4)
    q = gpx.variational_families.CollapsedVariationalGaussian(posterior=post, inducing_inputs=dev(X[:20] + 0.01))
    val = gpx.objectives.collapsed_elbo(q, gpx.Dataset(X=dev(X), y=dev(y)))
    ref = o.collapsed_elbo("powered_exponential", X, y, X[:20] + 0.01, np.array([0.9, 1.2]), np.array([1.3, 0.3]), 0.4, 0.0)
    assert abs(val.item() - ref) <= 1e-8 * abs(ref)
    # had the trace term used k(x, x) = variance, the value would be off by n variance (1 - exp(-(1e-18)^0.3)) / (2 sigma^2) ~ 5e-3
    assert abs(n * 1.3 * (1.0 - math.exp(-(1e-18 ** 0.3))) / (2 * 0.16)) > 100 * 1e-8 * abs(ref)


# ---- combination kernels (kernels/base.py:150-339) ---------------------------------------------------------------
def _combo(op):
    import gpjax_b200 as gpx

    k1 = gpx.kernels.RBF(lengthscale=[0.8, 1.3], variance=1.1)
    k2 = gpx.kernels.Matern32(lengthscale=0.6, variance=0.7)
    k3 = gpx.kernels.White(variance=0.05)
    if op == "sum":
        return k1 + k2 + k3, [k1, k2, k3]
    if op == "prod":
        return k1 * k2, [k1, k2]
    return k1 * k2 + k3 + 0.3, [k1, k2, k3]  # mixed: (k1 * k2) + k3 + Constant(0.3)


def _combo_t(op, x, z, P):
    """torch-CPU restatement of CombinationKernel.__call__ (kernels/base.py:297-310): operator over the parts."""
    k1 = go._t_cross(torch, 0, x, z, P["l1"], P["v1"])
    k2 = go._t_cross(torch, 1, x, z, P["l2"], P["v2"])
    k3 = go._t_cross(torch, 7, x, z, torch.tensor(1.0, dtype=torch.float64), P["v3"])
    if op == "sum":
        return k1 + k2 + k3
    if op == "prod":
        return k1 * k2
    return k1 * k2 + k3 + 0.3


def _combo_params():
    t = lambda a: torch.tensor(np.asarray(a, np.float64), requires_grad=True)
    return {"l1": t([0.8, 1.3]), "v1": t(1.1), "l2": t(0.6), "v2": t(0.7), "v3": t(0.05)}


@pytest.mark.parametrize("op", ["sum", "prod", "mixed"])
def test_combination_kernels(op):
    import gpjax_b200 as gpx
    from gpjax_b200.parameters import Real

    n = 400
    X, y = data(n, 2, 31)
    k, parts = _combo(op)
    assert isinstance(k, gpx.kernels.CombinationKernel)
    if op == "sum":
        assert len(k.kernels) == 3  # nested sums are flattened (kernels/base.py:275-281)
    P = _combo_params()
    Kref = _combo_t(op, torch.tensor(X), torch.tensor(X[:50]), P).detach().numpy()
    K = k.cross_covariance(dev(X), dev(X[:50])).cpu().numpy()
    assert np.max(np.abs(K - Kref)) <= 1e-12 * np.max(np.abs(Kref))
    # conjugate_mll through the dense-Sigma route, value and every gradient against CPU autodiff
    post = gpx.gps.Prior(mean_function=gpx.mean_functions.Constant(Real(0.2)), kernel=k) * gpx.likelihoods.Gaussian(
        num_datapoints=n, obs_stddev=0.4)
    params = dict(post.named_parameters())
    for p in params.values():
        p.value.requires_grad_(True)
    D = gpx.Dataset(X=dev(X), y=dev(y))
    val = gpx.objectives.conjugate_mll(post, D)
    val.backward()
    sn, c = torch.tensor(0.4, dtype=torch.float64, requires_grad=True), torch.tensor(0.2, dtype=torch.float64, requires_grad=True)
    Xt, yt = torch.tensor(X), torch.tensor(y).reshape(-1)
    Sig = _combo_t(op, Xt, Xt, P) + torch.eye(n, dtype=torch.float64) * 1e-6 + torch.eye(n, dtype=torch.float64) * sn**2
    diff = yt - c
    ref = -0.5 * (n * math.log(2 * math.pi) + torch.linalg.slogdet(Sig)[1] + diff @ torch.linalg.solve(Sig, diff))
    ref.backward()
    assert abs(val.item() - ref.item()) <= 1e-8 * abs(ref.item())
    floor = 1e-6 * abs(ref.item())
    pairs = [(parts[0].lengthscale, P["l1"]), (parts[0].variance, P["v1"]), (parts[1].lengthscale, P["l2"]),
             (parts[1].variance, P["v2"]), (post.likelihood.obs_stddev, sn), (post.prior.mean_function.constant, c)]
    if op != "prod":
        pairs.append((parts[2].variance, P["v3"]))
    for mine, theirs in pairs:
        b = theirs.grad.numpy()
        a = mine.value.grad.cpu().numpy().reshape(b.shape)
        assert np.max(np.abs(a - b)) <= 1e-8 * max(np.max(np.abs(b)), floor)
    # predict runs for combination kernels too (gps.py:495-526)
    T = data(30, 2, 32)[0]
    pred = post.predict(dev(T), D)
    with torch.no_grad():
        Pd = {kk: v.detach() for kk, v in P.items()}
        Sg = (_combo_t(op, Xt, Xt, Pd) + torch.eye(n, dtype=torch.float64) * (1e-6 + 0.16)).numpy()
        Kxt = _combo_t(op, Xt, torch.tensor(T), Pd).numpy()
        Ktt = _combo_t(op, torch.tensor(T), torch.tensor(T), Pd).numpy()
    mref = 0.2 + Kxt.T @ np.linalg.solve(Sg, y.reshape(-1) - 0.2)
    cref = Ktt - Kxt.T @ np.linalg.solve(Sg, Kxt) + 1e-6 * np.eye(30)
    assert np.max(np.abs(pred.mean().cpu().numpy() - mref)) <= 1e-8 * max(np.max(np.abs(mref)), 1.0)
    assert np.max(np.abs(pred.covariance().cpu().numpy() - cref)) <= 1e-8 * max(np.max(np.abs(cref)), 1.0)


def test_combination_rejects_non_kernels():
    import gpjax_b200 as gpx

    with pytest.raises(TypeError):
        gpx.kernels.SumKernel(kernels=[gpx.kernels.RBF(), 3.0])


def _t_collapsed_elbo(Kzz, Kzx, kdiag, diff, sn, jitter):
    """objectives.py:342-416 in torch-CPU float64 on given matrices (autodiff reference for the combination kernels)."""
    m, n = Kzx.shape
    Lz = torch.linalg.cholesky(Kzz + jitter * torch.eye(m, dtype=torch.float64))
    A = torch.linalg.solve_triangular(Lz, Kzx, upper=False) / sn
    AAT = A @ A.T
    L = torch.linalg.cholesky(torch.eye(m, dtype=torch.float64) + AAT)
    c = torch.linalg.solve_triangular(L, (A @ diff).reshape(-1, 1), upper=False)
    quad = (torch.sum(diff**2) - torch.sum(c**2)) / sn**2
    two_log_prob = -n * torch.log(2 * math.pi * sn**2) - 2.0 * torch.sum(torch.log(torch.diagonal(L))) - quad
    return (two_log_prob - (torch.sum(kdiag) / sn**2 - torch.trace(AAT))) / 2.0


def _t_svgp_elbo(Kzz, Kzx, kdiag, y, mean_c, sn, mu, W, num_datapoints, jitter):
    """objectives.py:241-318 + variational_families.py:169-285 + integrators.py:151-158 in torch-CPU float64."""
    m, n = Kzx.shape
    Lz = torch.linalg.cholesky(Kzz + jitter * torch.eye(m, dtype=torch.float64))
    W = torch.tril(W)
    LiW = torch.linalg.solve_triangular(Lz, W, upper=False)
    Lim = torch.linalg.solve_triangular(Lz, (mu - mean_c).reshape(-1, 1), upper=False)
    kl = 0.5 * (torch.sum(LiW**2) + torch.sum(Lim**2) - m + 2 * torch.sum(torch.log(torch.diagonal(Lz)))
                - 2 * torch.sum(torch.log(torch.abs(torch.diagonal(W)))))
    A = torch.linalg.solve_triangular(Lz, Kzx, upper=False)
    KiK = torch.linalg.solve_triangular(Lz.T, A, upper=True)
    R = KiK.T @ W
    mean = mean_c + KiK.T @ (mu - mean_c)
    var = kdiag - torch.sum(A**2, 0) + torch.sum(R**2, 1) + jitter
    ell = -0.5 * torch.sum(math.log(2 * math.pi) + torch.log(sn**2) + ((y - mean) ** 2 + var) / sn**2)
    return ell * num_datapoints / n - kl


@pytest.mark.parametrize("op", ["sum", "prod", "mixed"])
def test_sparse_objectives_and_predictions_accept_combination_kernels(op):
    """collapsed_elbo / elbo take ANY kernel in the reference (objectives.py:352-356 only call kernel.gram / cross_covariance /
    __call__).  Sums and products have no single streamed epilogue here, so they run the composable route: value, every gradient
    (kernel parts, noise, mean, inducing inputs, variational parameters) and both predictive distributions against CPU autodiff."""
    import gpjax_b200 as gpx
    from gpjax_b200.parameters import Real

    n, m = 600, 40
    X, y = data(n, 2, 41)
    Z0 = np.ascontiguousarray(X[:m] + (0.0 if op == "sum" else 0.01))  # op == "sum": coincident points exercise the White part
    T = data(25, 2, 42)[0]
    Xt, yt, Tt = torch.tensor(X), torch.tensor(y).reshape(-1), torch.tensor(T)

    def build():
        k, parts = _combo(op)
        post = gpx.gps.Prior(mean_function=gpx.mean_functions.Constant(Real(0.2)), kernel=k) * gpx.likelihoods.Gaussian(
            num_datapoints=5 * n, obs_stddev=0.4)
        return k, parts, post

    def check(val, ref, pairs):
        assert abs(val.item() - ref.item()) <= 1e-8 * abs(ref.item()), (val.item(), ref.item())
        floor = 1e-6 * abs(ref.item())
        for mine, theirs in pairs:
            b = theirs.grad.numpy()
            a = mine.value.grad.cpu().numpy().reshape(b.shape)
            assert np.max(np.abs(a - b)) <= 1e-8 * max(np.max(np.abs(b)), floor), (a, b)

    D = gpx.Dataset(X=dev(X), y=dev(y))
    t = lambda a: torch.tensor(np.asarray(a, np.float64), requires_grad=True)
    # ---- collapsed_elbo + CollapsedVariationalGaussian.predict
    k, parts, post = build()
    q = gpx.variational_families.CollapsedVariationalGaussian(posterior=post, inducing_inputs=dev(Z0))
    for p in dict(q.named_parameters()).values():
        p.value.requires_grad_(True)
    val = gpx.objectives.collapsed_elbo(q, D)
    val.backward()
    P, Zt, sn, c = _combo_params(), t(Z0), t(0.4), t(0.2)
    ref = _t_collapsed_elbo(_combo_t(op, Zt, Zt, P), _combo_t(op, Zt, Xt, P), torch.diagonal(_combo_t(op, Xt, Xt, P)), yt - c, sn, 1e-6)
    ref.backward()
    pairs = [(parts[0].lengthscale, P["l1"]), (parts[0].variance, P["v1"]), (parts[1].lengthscale, P["l2"]),
             (parts[1].variance, P["v2"]), (post.likelihood.obs_stddev, sn), (post.prior.mean_function.constant, c),
             (q.inducing_inputs, Zt)]
    if op != "prod":
        pairs.append((parts[2].variance, P["v3"]))
    check(val, ref, pairs)
    pred = q.predict(dev(T), D)
    with torch.no_grad():
        Pd = {kk: v.detach() for kk, v in P.items()}
        Kzz = _combo_t(op, Zt.detach(), Zt.detach(), Pd).numpy() + 1e-6 * np.eye(m)
        Kzx = _combo_t(op, Zt.detach(), Xt, Pd).numpy()
        Kzt = _combo_t(op, Zt.detach(), Tt, Pd).numpy()
        Ktt = _combo_t(op, Tt, Tt, Pd).numpy()
    Lz = np.linalg.cholesky(Kzz)
    A = np.linalg.solve(Lz, Kzx) / 0.4
    Bm = np.eye(m) + A @ A.T
    At = np.linalg.solve(Lz, Kzt)
    v = np.linalg.solve(Bm, A @ (y.reshape(-1) - 0.2)) * 0.4
    mref = 0.2 + At.T @ v / 0.16
    cref = Ktt - At.T @ At + At.T @ np.linalg.solve(Bm, At) + 1e-6 * np.eye(25)
    assert np.max(np.abs(pred.mean().cpu().numpy() - mref)) <= 1e-8 * max(np.max(np.abs(mref)), 1.0)
    assert np.max(np.abs(pred.covariance().cpu().numpy() - cref)) <= 1e-8 * max(np.max(np.abs(cref)), 1.0)
    # ---- elbo (SVGP) + VariationalGaussian.predict
    k, parts, post = build()
    rng = np.random.default_rng(7)
    mu0 = 0.3 * rng.standard_normal((m, 1))
    W0 = np.tril(0.1 * rng.standard_normal((m, m))) + 0.8 * np.eye(m)
    q = gpx.variational_families.VariationalGaussian(posterior=post, inducing_inputs=dev(Z0), variational_mean=dev(mu0),
                                                     variational_root_covariance=dev(W0))
    for p in dict(q.named_parameters()).values():
        p.value.requires_grad_(True)
    val = gpx.objectives.elbo(q, D)
    val.backward()
    P, Zt, sn, c, mu, W = _combo_params(), t(Z0), t(0.4), t(0.2), t(mu0.reshape(-1)), t(W0)
    ref = _t_svgp_elbo(_combo_t(op, Zt, Zt, P), _combo_t(op, Zt, Xt, P), torch.diagonal(_combo_t(op, Xt, Xt, P)), yt, c, sn, mu, W,
                       5.0 * n, 1e-6)
    ref.backward()
    pairs = [(parts[0].lengthscale, P["l1"]), (parts[0].variance, P["v1"]), (parts[1].lengthscale, P["l2"]),
             (parts[1].variance, P["v2"]), (post.likelihood.obs_stddev, sn), (post.prior.mean_function.constant, c),
             (q.inducing_inputs, Zt), (q.variational_mean, mu)]
    check(val, ref, pairs)
    gW = q.variational_root_covariance.value.grad.cpu().numpy()
    assert np.max(np.abs(np.tril(gW) - np.tril(W.grad.numpy()))) <= 1e-8 * np.max(np.abs(W.grad.numpy()))
    pred = q.predict(dev(T))
    KiK = np.linalg.solve(Kzz, Kzt)
    mref = 0.2 + KiK.T @ (mu0.reshape(-1) - 0.2)
    cref = Ktt - At.T @ At + (KiK.T @ W0) @ (KiK.T @ W0).T + 1e-6 * np.eye(25)
    assert np.max(np.abs(pred.mean().cpu().numpy() - mref)) <= 1e-8 * max(np.max(np.abs(mref)), 1.0)
    assert np.max(np.abs(pred.covariance().cpu().numpy() - cref)) <= 1e-8 * max(np.max(np.abs(cref)), 1.0)


# ---- conjugate_loocv (objectives.py:110-178) -----------------------------------------------------------------------
@pytest.mark.parametrize("kname,name", [("RBF", "rbf"), ("Matern52", "matern52")])
@pytest.mark.parametrize("n,d", [(30, 1), (300, 2), (1100, 4)])
def test_conjugate_loocv(kname, name, n, d):
    import gpjax_b200 as gpx
    from gpjax_b200.parameters import Real

    X, y = data(n, d, n + 7 * d, dup=False)
    ell = np.linspace(0.8, 1.4, d)
    k = getattr(gpx.kernels, kname)(lengthscale=ell, variance=1.3)
    post = gpx.gps.Prior(mean_function=gpx.mean_functions.Constant(Real(0.2)), kernel=k) * gpx.likelihoods.Gaussian(
        num_datapoints=n, obs_stddev=0.4)
    params = dict(post.named_parameters())
    for p in params.values():
        p.value.requires_grad_(True)
    val = gpx.objectives.conjugate_loocv(post, gpx.Dataset(X=dev(X), y=dev(y)))
    (-val).backward()  # the way users hand it to fit (objectives.py:139-141)
    ref, gr = o.conjugate_loocv_value_and_grad_autodiff(name, X, y, ell, 1.3, 0.4, 0.2)
    assert abs(ref - o.conjugate_loocv(name, X, y, ell, 1.3, 0.4, 0.2)) <= 1e-9 * abs(ref)
    assert abs(val.item() - ref) <= 1e-8 * abs(ref)
    floor = 1e-6 * abs(ref)
    for pname, key in [("prior.kernel.lengthscale", "lengthscale"), ("prior.kernel.variance", "variance"),
                       ("likelihood.obs_stddev", "obs_stddev"), ("prior.mean_function.constant", "mean_const")]:
        b = -np.asarray(gr[key], np.float64)
        a = params[pname].value.grad.cpu().numpy().reshape(b.shape)
        assert np.max(np.abs(a - b)) <= 1e-7 * max(np.max(np.abs(b)), floor), key


def test_fit_with_periodic_plus_white_kernel():
    """End to end: a sum kernel trains through gpx.fit and the objective improves (tests/test_fit.py:193-215 pattern)."""
    import gpjax_b200 as gpx

    rng = np.random.default_rng(0)
    X = np.sort(rng.uniform(0, 6, (120, 1)), axis=0)
    y = np.sin(2 * np.pi * X / 1.5) + 0.1 * rng.standard_normal((120, 1))
    from gpjax_b200.parameters import PositiveReal

    k = gpx.kernels.Periodic(lengthscale=1.0, period=PositiveReal(1.4)) + gpx.kernels.White(variance=0.1)
    post = gpx.gps.Prior(mean_function=gpx.mean_functions.Zero(), kernel=k) * gpx.likelihoods.Gaussian(num_datapoints=120, obs_stddev=0.2)
    D = gpx.Dataset(X=dev(X), y=dev(y))
    neg = lambda p, d: -gpx.objectives.conjugate_mll(p, d)
    start = neg(post, D).item()
    opt, hist = gpx.fit(model=post, objective=neg, train_data=D, optim=gpx.optim.adam(0.02), num_iters=150, verbose=False)
    assert hist[-1].item() < start - 5.0
    period = dict(opt.named_parameters())["prior.kernel.kernels[0].period"].value.item()
    assert abs(period - 1.5) < 0.1


def test_white_kernel_objective_keeps_the_constant_diagonal_engine_semantics():
    """White.gram goes through ConstantDiagonalKernelComputation (white.py:47, constant_diagonal.py:39-43): variance on the
    diagonal even when rows coincide -- conjugate_mll must see that matrix, not the pairwise all-equal evaluation."""
    import gpjax_b200 as gpx

    X, y = data(60, 2, 5, dup=True)  # rows 1 and 30 coincide
    post = gpx.gps.Prior(mean_function=gpx.mean_functions.Zero(), kernel=gpx.kernels.White(variance=0.7)) * \
        gpx.likelihoods.Gaussian(num_datapoints=60, obs_stddev=0.4)
    val = gpx.objectives.conjugate_mll(post, gpx.Dataset(X=dev(X), y=dev(y))).item()
    s2 = 0.7 + 1e-6 + 0.16
    ref = -0.5 * (60 * np.log(2 * np.pi * s2) + float((y**2).sum()) / s2)
    assert abs(val - ref) <= 1e-10 * abs(ref)
    Kc = gpx.kernels.White(variance=0.7).cross_covariance(dev(X), dev(X)).cpu().numpy()
    assert Kc[1, 30] == 0.7 and Kc[30, 1] == 0.7  # the dense pairwise evaluation does see the coincidence
