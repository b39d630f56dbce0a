"""fit(cuda_graph=True): one optimisation step captured as a CUDA graph and replayed (the jit + lax.scan part of the reference's
loop, gpjax/fit.py:160-170).  The replayed run must reproduce the ordinary run -- same history, same end point -- and the end point
must match the oracle (tests/test_fit.py:193-257 of the reference checks that fit lowers the objective and returns the history)."""
import numpy as np
import pytest
import torch

import oracle as o

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda")


def build_data(n, d, seed):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-2.0, 2.0, (n, d))
    y = np.sin(X[:, :1]) + 0.1 * rng.standard_normal((n, 1))
    return X, y


def _posterior(gpx, n, d, constant_mean):
    mean = gpx.mean_functions.Constant() if constant_mean else gpx.mean_functions.Zero()
    kernel = gpx.kernels.RBF(lengthscale=[1.0] * d) if d > 1 else gpx.kernels.RBF()
    return gpx.gps.Prior(mean_function=mean, kernel=kernel) * gpx.likelihoods.Gaussian(num_datapoints=n)


def _neg_mll(p, d):
    import gpjax_b200 as gpx

    return -gpx.objectives.conjugate_mll(p, d)


# (600, 1): config 1 shape, one diagonal block.  (2500, 3): three blocks of 1024 -> the look-ahead helper streams fork and join
# inside the capture.
@pytest.mark.parametrize("n,d,constant_mean,iters", [(600, 1, False, 12), (2500, 3, True, 8)])
def test_graphed_fit_reproduces_the_ordinary_fit(n, d, constant_mean, iters):
    import gpjax_b200 as gpx

    X, y = build_data(n, d, 7)
    D = gpx.Dataset(X=dev(X), y=dev(y))
    runs = {}
    for graphed in (False, True):
        post = _posterior(gpx, n, d, constant_mean)
        opt, hist = gpx.fit(model=post, objective=_neg_mll, train_data=D, optim=gpx.optim.adam(0.05), num_iters=iters,
                            verbose=False, cuda_graph=graphed)
        runs[graphed] = (opt, hist.cpu().numpy())
    h0, h1 = runs[False][1], runs[True][1]
    assert h1.shape == (iters,) and np.all(np.isfinite(h1))
    assert h1[-1] < h1[0]
    assert np.max(np.abs(h1 - h0)) <= 1e-10 * np.max(np.abs(h0))
    p0 = dict(runs[False][0].named_parameters())
    for path, p in runs[True][0].named_parameters():
        assert torch.allclose(p.value, p0[path].value, rtol=1e-9, atol=1e-12), path
    # the last history entry is the objective at the parameters BEFORE the last update: evaluate the oracle at the end point and
    # compare with one more evaluation there
    opt = runs[True][0]
    k = opt.prior.kernel
    c = opt.prior.mean_function.constant if constant_mean else None
    mean = 0.0 if c is None else float(getattr(c, "value", c))
    ref = o.conjugate_mll("rbf", X, y, k.lengthscale.value.cpu().numpy().reshape(-1), k.variance.value.item(),
                          opt.likelihood.obs_stddev.value.item(), mean)
    got = gpx.objectives.conjugate_mll(opt, D).item()
    assert abs(got - ref) <= 1e-8 * abs(ref)


def test_graphed_fit_loocv_dense_route():
    """conjugate_loocv takes the composable dense route (differentiable Gram launch + GaussianLogProb-style factorisation): its
    per-evaluation workspaces are allocated inside the capture and live in the graph's pool."""
    import gpjax_b200 as gpx

    n = 400
    X, y = build_data(n, 2, 11)
    D = gpx.Dataset(X=dev(X), y=dev(y))
    obj = lambda p, d: -gpx.objectives.conjugate_loocv(p, d)
    hists = []
    for graphed in (False, True):
        post = _posterior(gpx, n, 2, False)
        _, hist = gpx.fit(model=post, objective=obj, train_data=D, optim=gpx.optim.adam(0.05), num_iters=8, verbose=False,
                          cuda_graph=graphed)
        hists.append(hist.cpu().numpy())
    assert np.all(np.isfinite(hists[1])) and hists[1][-1] < hists[1][0]
    assert np.max(np.abs(hists[1] - hists[0])) <= 1e-10 * np.max(np.abs(hists[0]))


def test_graphed_fit_refuses_minibatches():
    import gpjax_b200 as gpx

    X, y = build_data(64, 1, 3)
    D = gpx.Dataset(X=dev(X), y=dev(y))
    with pytest.raises(NotImplementedError):
        gpx.fit(model=_posterior(gpx, 64, 1, False), objective=_neg_mll, train_data=D, optim=gpx.optim.adam(0.05), num_iters=5,
                batch_size=16, verbose=False, cuda_graph=True)
