"""CPU checks for the kernels beyond RBF / Matern: RationalQuadratic, PoweredExponential,
Periodic, White -- oracle self-consistency (pairwise definition vs matrix form), and the product's orchestration +
C ABI on the host model of the device primitives, against torch-CPU autodiff of the literal restatement.
Kinds 4..6 pass `variance` as the pair [variance, shape] (include/gpjax_b200.h)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

import oracle as o
from gpjax_b200 import _abi
from oracle import gp_oracle as go

HERE = os.path.dirname(os.path.abspath(__file__))
# (kind id, oracle name, kernel scalars [variance(, shape)])
EXT = [(4, "rational_quadratic", [1.3, 0.7]), (5, "powered_exponential", [1.3, 0.6]), (6, "periodic", [1.3, 1.7]),
       (7, "white", [1.3])]


@pytest.fixture(scope="module")
def lib():
    hs = os.path.join(HERE, "hostsim")
    subprocess.run(["make", "-s", "-C", hs], check=True)
    return _abi.declare(C.CDLL(os.path.join(hs, "libgpjax_b200_hostsim.so")))


def p(a):
    return None if a is None else a.ctypes.data


def scal(kind, s):
    return np.array(s) if kind in (4, 5, 6) else s[0]


def data(n, d, seed, dup=False):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-2, 2, (n, d))
    if dup and n > 3:
        X[n // 2] = X[1]  # an exact duplicate row: exercises the 1e-36 clamp and White's equality test off the diagonal
    y = np.sin(X[:, 0]) + 0.1 * rng.standard_normal(n)
    return X, y


@pytest.mark.parametrize("kind,name,s", EXT)
def test_oracle_matrix_form_matches_pairwise_definition(kind, name, s):
    """cross_covariance (vectorised) == kernel_pair (the literal `__call__` of rational_quadratic.py:77-83,
    powered_exponential.py:85-89, periodic.py:81-88, white.py:63-64) pair by pair."""
    X, _ = data(7, 3, 1, dup=True)
    Z, _ = data(5, 3, 2)
    Z[2] = X[3]
    ell = np.array([0.7, 1.0, 1.3]) if kind != 7 else np.ones(3)
    K = o.cross_covariance(name, X, Z, ell, scal(kind, s))
    for i in range(7):
        for j in range(5):
            ref = float(o.kernel_pair(name, X[i], Z[j], ell, scal(kind, s)))
            assert abs(K[i, j] - ref) <= 1e-15 * max(abs(ref), 1.0)
    if kind == 7:
        assert K[3, 2] == s[0] and K[1, 2] == s[0] and np.count_nonzero(K) == 2  # rows 1 and 3 of X coincide
    # PoweredExponential on coincident points keeps the clamp's tau = 1e-18: k = var * exp(-(1e-18)^power)
    if kind == 5:
        assert abs(K[3, 2] - s[0] * np.exp(-(1e-18 ** s[1]))) < 1e-15


@pytest.mark.parametrize("kind,name,s", EXT)
@pytest.mark.parametrize("iso", [False, True])
def test_gram_and_gram_bwd(lib, kind, name, s, iso):
    X, _ = data(70, 3, 5, dup=True)
    Z, _ = data(150, 3, 6)
    Z[7] = X[3]
    D = 3
    ell = np.ones(1 if iso else D) if kind == 7 else (np.array([0.9]) if iso else np.array([0.7, 1.0, 1.3]))
    var = np.array(s)
    K = np.zeros((70, 150))
    assert lib.gpb_gram(None, kind, 70, 150, D, p(X), D, p(Z), D, p(ell), int(iso), p(var), 0.0, None, 0, p(K), 150) == 0
    ellv = ell[0] if iso else ell
    Kref = o.cross_covariance(name, X, Z, ellv, scal(kind, s))
    assert np.max(np.abs(K - Kref)) <= 1e-13
    rng = np.random.default_rng(0)
    Wt = rng.standard_normal((70, 150))
    t = lambda a: torch.tensor(np.asarray(a, np.float64), requires_grad=True)
    Xt, Zt, et, vt = t(X), t(Z), t(ellv), t(scal(kind, s))
    (torch.tensor(Wt) * go._t_cross(torch, kind, Xt, Zt, et, vt)).sum().backward()
    nb = lib.gpb_gram_bwd_workspace_bytes(70, 150, D)
    ws = np.zeros(nb // 8 + 1)
    g_ell, g_var, gX, gZ = np.zeros_like(ell), np.zeros(len(s)), np.zeros_like(X), np.zeros_like(Z)
    assert lib.gpb_gram_bwd(None, kind, 70, 150, D, p(X), D, p(Z), D, p(ell), int(iso), p(var), p(Wt), 150, 1.0, p(ws), nb,
                            p(g_ell), p(g_var), p(gX), D, p(gZ), D) == 0

    def close(a, b, what):
        b = np.asarray(b, np.float64)
        assert np.max(np.abs(np.asarray(a).reshape(b.shape) - b)) <= 1e-10 * max(np.max(np.abs(b)), 1.0), what

    if kind != 7:
        close(g_ell, et.grad.numpy(), "lengthscale")
        close(gX, Xt.grad.numpy(), "X")
        close(gZ, Zt.grad.numpy(), "Z")
    close(g_var, vt.grad.numpy(), "variance/shape")


@pytest.mark.parametrize("kind,name,s", EXT)
@pytest.mark.parametrize("N,D,iso", [(100, 3, False), (300, 2, True), (513, 1, True)])
def test_mll_forward_backward(lib, kind, name, s, N, D, iso):
    X, y = data(N, D, N + D, dup=(kind != 7))
    ell = np.ones(1 if iso else D) if kind == 7 else (np.array([0.9]) if iso else np.linspace(0.8, 1.6, D))
    var, sn, c = np.array(s), np.array([0.4]), np.array([0.2])
    nbytes = lib.gpb_mll_workspace_bytes(N, D)
    ws = np.zeros(nbytes // 8 + 8)
    Sig = np.full((N, N), np.nan)
    val, alpha, info = np.zeros(1), np.zeros(N), np.zeros(1, np.int32)
    assert lib.gpb_mll_forward(None, kind, N, D, p(X), D, p(y), p(ell), int(iso), p(var), p(sn), p(c), 1e-6, p(Sig), N,
                               p(ws), nbytes, p(val), p(alpha), p(info)) == 0
    assert info[0] == 0
    ellv = ell[0] if iso else ell
    ref, gr = o.conjugate_mll_value_and_grad_autodiff(name, X, y, ellv, scal(kind, s), sn[0], c[0])
    assert abs(val[0] - ref) <= 1e-10 * abs(ref)
    g_ell, g_var, g_sn, g_c = np.zeros(1 if iso else D), np.zeros(len(s)), np.zeros(1), np.zeros(1)
    assert lib.gpb_mll_backward(None, kind, N, D, p(X), D, p(ell), int(iso), p(var), p(sn), p(Sig), N, p(ws), nbytes,
                                p(alpha), None, p(g_ell), p(g_var), p(g_sn), p(g_c)) == 0
    tol = 1e-8
    if kind != 7:
        gl = np.atleast_1d(gr["lengthscale"])
        assert np.max(np.abs(g_ell - gl)) <= tol * max(np.max(np.abs(gl)), 1e-6 * abs(ref))
    gv = np.atleast_1d(gr["variance"])
    assert np.max(np.abs(g_var - gv)) <= tol * max(np.max(np.abs(gv)), 1e-6 * abs(ref))
    assert abs(g_sn[0] - gr["obs_stddev"]) <= tol * max(abs(gr["obs_stddev"]), 1e-6 * abs(ref))
    assert abs(g_c[0] - gr["mean_const"]) <= tol * max(abs(gr["mean_const"]), 1e-6 * abs(ref))


@pytest.mark.parametrize("kind,name,s", [e for e in EXT if e[0] in (4, 6)])
@pytest.mark.parametrize("shards", [1, 2])
def test_sgpr_value_and_gradient(lib, kind, name, s, shards):
    N, M, D, block = 260, 24, 2, 100
    X, y = data(N, D, 11)
    Z = np.ascontiguousarray(data(M, D, 12)[0])
    ell, var, sn, c = np.array([0.9, 1.4]), np.array(s), np.array([0.4]), np.array([0.2])
    nbytes, cnt = lib.gpb_sgpr_workspace_bytes(M, D, block), lib.gpb_sgpr_stats_count(M)
    bounds = np.linspace(0, N, shards + 1).astype(int)
    parts = [(np.ascontiguousarray(X[a:b]), np.ascontiguousarray(y[a:b])) for a, b in zip(bounds[:-1], bounds[1:])]
    wss, Pall = [np.zeros(nbytes // 8 + 8) for _ in parts], np.zeros(cnt)
    for (Xr, yr), w in zip(parts, wss):
        P = np.zeros(cnt)
        assert lib.gpb_sgpr_stats(None, kind, len(Xr), M, D, p(Xr), D, p(yr), p(Z), D, p(ell), 0, p(var), p(sn), p(c),
                                  1e-6, block, p(w), nbytes, p(P)) == 0
        Pall += P
    val, info = np.zeros(1), np.zeros(2, np.int32)
    for w in wss:
        assert lib.gpb_sgpr_finish(None, kind, M, D, p(Z), D, p(ell), 0, p(var), p(sn), block, p(w), nbytes, p(Pall),
                                   1, p(val), p(info)) == 0
    ref, gref = o.collapsed_elbo_value_and_grad_autodiff(name, X, y, Z, ell, np.array(s), sn[0], c[0])
    assert abs(val[0] - ref) <= 1e-9 * abs(ref)
    tot = np.zeros(M * D + D + len(s))
    for (Xr, yr), w in zip(parts, wss):
        f = np.full_like(tot, np.nan)
        assert lib.gpb_sgpr_grad_local(None, kind, len(Xr), M, D, p(Xr), D, p(yr), p(Z), D, p(ell), 0, p(var), p(sn), p(c),
                                       block, p(w), nbytes, p(f[:M * D]), p(f[M * D:M * D + D]), p(f[M * D + D:])) == 0
        tot += f
    gZ, gl, gv = tot[:M * D].copy(), tot[M * D:M * D + D].copy(), tot[M * D + D:].copy()
    gs, gc = np.zeros(1), np.zeros(1)
    assert lib.gpb_sgpr_grad_finish(None, kind, M, D, p(Z), D, p(ell), 0, p(var), p(sn), block, p(wss[0]), nbytes, None,
                                    p(gZ), p(gl), p(gv), p(gs), p(gc)) == 0
    got = dict(lengthscale=gl, variance=gv, obs_stddev=gs[0], mean_const=gc[0], inducing_inputs=gZ.reshape(M, D))
    for k in gref:
        a, b = np.asarray(got[k]).reshape(np.shape(gref[k])), np.asarray(gref[k])
        assert np.max(np.abs(a - b)) <= 1e-7 * max(np.max(np.abs(b)), 1e-8 * abs(ref)), k


def test_powered_exponential_is_refused_by_the_sparse_objectives(lib):
    """k(x, x) != variance under the reference's distance clamp for small powers; the statistics path does not
    carry that, so it refuses (GPB_ERR_UNSUPPORTED) rather than return a slightly different bound."""
    N, M, D = 50, 8, 2
    X, y = data(N, D, 3)
    Z = np.ascontiguousarray(X[:M])
    ell, var, sn = np.ones(D), np.array([1.0, 0.5]), np.array([0.3])
    nbytes = lib.gpb_sgpr_workspace_bytes(M, D, 32)
    ws, P = np.zeros(nbytes // 8 + 8), np.zeros(lib.gpb_sgpr_stats_count(M))
    rc = lib.gpb_sgpr_stats(None, 5, N, M, D, p(X), D, p(y), p(Z), D, p(ell), 0, p(var), p(sn), None, 1e-6, 32, p(ws),
                            nbytes, p(P))
    assert rc == -2


def test_unknown_kind_is_invalid(lib):
    X = np.zeros((4, 1))
    K = np.zeros((4, 4))
    one = np.ones(2)
    assert lib.gpb_gram(None, 8, 4, 4, 1, p(X), 1, p(X), 1, p(one), 1, p(one), 0.0, None, 0, p(K), 4) == -1
