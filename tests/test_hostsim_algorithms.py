"""CPU check of the product's blocked-algorithm orchestration + C ABI (gpjax_b200/csrc/algorithms.cpp,
sgpr.cpp, abi.cpp) linked against a plain C++ host model of the device primitives
(tests/hostsim/primitives_host.cpp).  The CUDA kernels themselves are checked by the -m gpu tests."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import scipy.linalg as sla

import oracle as o
from gpjax_b200 import _abi

HERE = os.path.dirname(os.path.abspath(__file__))
KINDS = [(0, "rbf"), (1, "matern32"), (2, "matern52"), (3, "matern12")]


@pytest.fixture(scope="module")
def lib():
    hs = os.path.join(HERE, "hostsim")
    subprocess.run(["make", "-s", "-C", hs], check=True)
    return _abi.declare(C.CDLL(os.path.join(hs, "libgpjax_b200_hostsim.so")))


def p(a):
    return None if a is None else a.ctypes.data


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def data(n, d, seed):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-2, 2, (n, d))
    y = np.sin(X[:, 0]) + 0.1 * rng.standard_normal(n)
    return X, y


@pytest.mark.parametrize("kind,name", KINDS)
@pytest.mark.parametrize("N,D,iso", [(1, 1, True), (100, 3, False), (256, 2, False), (300, 3, False), (513, 1, True),
                                     (700, 8, False)])
def test_mll_forward_backward(lib, kind, name, N, D, iso):
    X, y = data(N, D, N + D)
    ell = np.array([0.9]) if iso else np.linspace(0.8, 1.6, D)
    var, sn, c = np.array([1.3]), np.array([0.4]), np.array([0.2])
    nbytes = lib.gpb_mll_workspace_bytes(N, D)
    ws = np.zeros(nbytes // 8 + 8)
    Sig = np.full((N, N), np.nan)
    val, alpha, info = np.zeros(1), np.zeros(N), np.zeros(1, np.int32)
    rc = lib.gpb_mll_forward(None, kind, N, D, p(X), D, p(y), p(ell), int(iso), p(var), p(sn), p(c), 1e-6, p(Sig), N,
                             p(ws), nbytes, p(val), p(alpha), p(info))
    assert rc == 0 and info[0] == 0
    ellv = ell[0] if iso else ell
    ref = o.conjugate_mll(name, X, y, ellv, var[0], sn[0], c[0])
    assert abs(val[0] - ref) <= 1e-10 * abs(ref)
    g_ell, g_var, g_sn, g_c = np.zeros(1 if iso else D), np.zeros(1), np.zeros(1), np.zeros(1)
    gout = np.array([-2.0])
    rc = lib.gpb_mll_backward(None, kind, N, D, p(X), D, p(ell), int(iso), p(var), p(sn), p(Sig), N, p(ws), nbytes,
                              p(alpha), p(gout), p(g_ell), p(g_var), p(g_sn), p(g_c))
    assert rc == 0
    gr = o.conjugate_mll_grad_closed_form(name, X, y, ellv, var[0], sn[0], c[0])
    tol = 1e-8
    scale = max(np.max(np.abs(np.asarray(gr["lengthscale"]))), 1e-6)
    assert np.max(np.abs(g_ell / -2.0 - gr["lengthscale"])) <= tol * scale
    assert abs(g_var[0] / -2.0 - gr["variance"]) <= tol * max(abs(gr["variance"]), 1e-6 * abs(ref))
    assert abs(g_sn[0] / -2.0 - gr["obs_stddev"]) <= tol * max(abs(gr["obs_stddev"]), 1e-6 * abs(ref))
    assert abs(g_c[0] / -2.0 - gr["mean_const"]) <= tol * max(abs(gr["mean_const"]), 1e-6 * abs(ref))
    # L survived the backward pass in the lower triangle
    Lref = np.linalg.cholesky(o.gram(name, X, ellv, var[0]) + (1e-6 + sn[0] ** 2) * np.eye(N))
    assert np.max(np.abs(np.tril(Sig) - Lref)) <= 1e-11 * np.abs(Lref).max()


def test_mll_not_pd_sets_info_and_nan(lib):
    N, D = 300, 2
    X = np.zeros((N, D))
    y = np.zeros(N)
    ell, var, sn = np.ones(D), np.ones(1), np.zeros(1)
    nbytes = lib.gpb_mll_workspace_bytes(N, D)
    ws, Sig = np.zeros(nbytes // 8 + 8), np.zeros((N, N))
    val, alpha, info = np.zeros(1), np.zeros(N), np.zeros(1, np.int32)
    rc = lib.gpb_mll_forward(None, 0, N, D, p(X), D, p(y), p(ell), 0, p(var), p(sn), None, 0.0, p(Sig), N, p(ws),
                             nbytes, p(val), p(alpha), p(info))
    assert rc == 0 and info[0] == 2 and np.isnan(val[0])


@pytest.mark.parametrize("n", [1, 5, 128, 129, 256, 257, 600])
def test_factor_family(lib, n):
    rng = np.random.default_rng(n)
    X = rng.uniform(-2, 2, (n, 4))
    S = o.gram("matern52", X, np.linspace(0.8, 1.4, 4), 1.0) + 0.09 * np.eye(n)
    A = S.copy()
    nbytes = lib.gpb_factor_workspace_bytes(n, 1, 1)
    ws = np.zeros(nbytes // 8 + 8)
    info = np.zeros(1, np.int32)
    wsa = (p(ws), nbytes, n, 1, 1)
    assert lib.gpb_potrf_lower(None, n, p(A), n, 1, *wsa, p(info)) == 0
    Lref = np.linalg.cholesky(S)
    assert np.max(np.abs(A - Lref)) <= 1e-12 * np.abs(Lref).max()
    out = np.zeros(1)
    assert lib.gpb_sum_log_diag(None, n, p(A), n, p(out)) == 0
    assert abs(out[0] - np.log(np.diag(Lref)).sum()) <= 1e-12 * max(1, n)
    b = rng.standard_normal(n)
    for trans, Lm in ((0, Lref), (1, Lref.T)):
        x = b.copy()
        assert lib.gpb_trsv_lower(None, n, p(A), n, trans, p(x), *wsa) == 0
        assert rel(x, sla.solve_triangular(Lm, b, lower=not trans)) <= 1e-9
        T = min(n, 19)
        Bm = rng.standard_normal((n, T))
        Xm = Bm.copy()
        assert lib.gpb_trsm_lower_left(None, n, T, p(A), n, trans, p(Xm), T, *wsa) == 0
        assert np.max(np.abs(Lm @ Xm - Bm)) <= 1e-10 * n
    # diag_inverses on a user-supplied factor reproduces the same workspace content
    ws2 = np.zeros_like(ws)
    assert lib.gpb_diag_inverses(None, n, p(A), n, p(ws2), nbytes, n, 1, 1) == 0
    x1, x2 = b.copy(), b.copy()
    lib.gpb_trsv_lower(None, n, p(A), n, 0, p(x1), *wsa)
    lib.gpb_trsv_lower(None, n, p(A), n, 0, p(x2), p(ws2), nbytes, n, 1, 1)
    assert rel(x2, x1) <= 1e-12
    Sinv = np.zeros((n, n))
    assert lib.gpb_potri_lower(None, n, p(A), n, p(Sinv), n, *wsa) == 0
    assert np.max(np.abs(Sinv @ S - np.eye(n))) <= 1e-9


def test_workspace_too_small_is_reported(lib):
    n = 300
    A = np.eye(n)
    ws = np.zeros(16)
    info = np.zeros(1, np.int32)
    assert lib.gpb_potrf_lower(None, n, p(A), n, 1, p(ws), ws.nbytes, n, 1, 0, p(info)) == -4


@pytest.mark.parametrize("kind,name", KINDS)
def test_gram_and_gram_bwd_abi(lib, kind, name):
    rng = np.random.default_rng(2)
    N, M, D = 70, 150, 3
    X, Z = rng.uniform(-2, 2, (N, D)), rng.uniform(-2, 2, (M, D))
    ell, var = np.array([0.7, 1.0, 1.3]), np.array([1.2])
    K = np.zeros((N, M))
    assert lib.gpb_gram(None, kind, N, M, D, p(X), D, p(Z), D, p(ell), 0, p(var), 0.0, None, 0, p(K), M) == 0
    assert rel(K, o.cross_covariance(name, X, Z, ell, var[0])) <= 1e-13
    assert lib.gpb_gram(None, kind, N, M, 65, p(X), D, p(Z), D, p(ell), 0, p(var), 0.0, None, 0, p(K), M) == -2


def _sgpr_run(lib, kind, X, y, Z, ell, iso, var, sn, c, jitter, block_rows, shards=1, gout=1.0, raw=False, stats_out=None):
    """Drive the 6-step SGPR protocol of include/gpjax_b200.h on `shards` simulated ranks."""
    stats_fn = lib.gpb_sgpr_stats_raw if raw else lib.gpb_sgpr_stats
    N, D = X.shape
    M = Z.shape[0]
    ellv = np.atleast_1d(np.asarray(ell, np.float64)).copy()
    var_a, sn_a = np.array([var]), np.array([sn])
    c_a = None if c is None else np.array([c])
    nbytes = lib.gpb_sgpr_workspace_bytes(M, D, block_rows)
    cnt = lib.gpb_sgpr_stats_count(M)
    bounds = np.linspace(0, N, shards + 1).astype(int)
    wss = [np.zeros(nbytes // 8 + 8) for _ in range(shards)]
    Ps = []
    for r in range(shards):
        Xr, yr = np.ascontiguousarray(X[bounds[r]:bounds[r + 1]]), np.ascontiguousarray(y[bounds[r]:bounds[r + 1]])
        P = np.full(cnt, np.nan)
        rc = stats_fn(None, kind, len(Xr), M, D, p(Xr), D, p(yr), p(Z), D, p(ellv), int(iso), p(var_a),
                      p(sn_a), p(c_a), jitter, block_rows, p(wss[r]), nbytes, p(P))
        assert rc == 0
        Ps.append(P)
    Pall = np.sum(Ps, axis=0)  # the all-reduce
    if stats_out is not None:
        stats_out.append(Pall.reshape(M + 2, M + 2).copy())
    vals = []
    for r in range(shards):
        val, info = np.zeros(1), np.zeros(2, np.int32)
        # flag word: bit 0 = gradients, bit 1 (GPB_FINISH_DENSE_INT8) = well-conditioned Kzz, as the public route sets it with raw
        rc = lib.gpb_sgpr_finish(None, kind, M, D, p(Z), D, p(ellv), int(iso), p(var_a), p(sn_a), block_rows,
                                 p(wss[r]), nbytes, p(Pall), 1 | (2 if raw else 0), p(val), p(info))
        assert rc == 0 and not info.any()
        vals.append(val[0])
    assert len(set(vals)) == 1  # replicated finish is bit-identical on every rank
    gZ, gl, gv = [], [], []
    for r in range(shards):
        Xr, yr = np.ascontiguousarray(X[bounds[r]:bounds[r + 1]]), np.ascontiguousarray(y[bounds[r]:bounds[r + 1]])
        a, b, cc = np.zeros((M, D)), np.zeros(1 if iso else D), np.zeros(1)
        rc = lib.gpb_sgpr_grad_local(None, kind, len(Xr), M, D, p(Xr), D, p(yr), p(Z), D, p(ellv), int(iso), p(var_a),
                                     p(sn_a), p(c_a), block_rows, p(wss[r]), nbytes, p(a), p(b), p(cc))
        assert rc == 0
        gZ.append(a), gl.append(b), gv.append(cc)
    g_Z, g_ell, g_var = np.sum(gZ, axis=0), np.sum(gl, axis=0), np.sum(gv, axis=0)
    g_sn, g_c, go = np.zeros(1), np.zeros(1), np.array([gout])
    rc = lib.gpb_sgpr_grad_finish(None, kind, M, D, p(Z), D, p(ellv), int(iso), p(var_a), p(sn_a), block_rows,
                                  p(wss[0]), nbytes, p(go), p(g_Z), p(g_ell), p(g_var), p(g_sn), p(g_c))
    assert rc == 0
    return vals[0], dict(lengthscale=g_ell, variance=g_var[0], obs_stddev=g_sn[0], mean_const=g_c[0],
                         inducing_inputs=g_Z)


@pytest.mark.parametrize("kind,name", KINDS)
@pytest.mark.parametrize("N,M,D,iso,block,shards", [(60, 12, 3, False, 16, 1), (300, 40, 2, True, 128, 3),
                                                     (500, 130, 8, False, 200, 2), (257, 30, 1, True, 64, 1),
                                                     (300, 260, 3, False, 100, 2), (700, 20, 2, False, 300, 1)])
def test_sgpr_value_and_gradient(lib, kind, name, N, M, D, iso, block, shards):
    X, y = data(N, D, N + M)
    rng = np.random.default_rng(M)
    Z = np.ascontiguousarray(rng.uniform(-2, 2, (M, D)))
    ell = 0.9 if iso else np.linspace(0.8, 1.6, D)
    val, g = _sgpr_run(lib, kind, X, y, Z, ell, iso, 1.3, 0.4, 0.2, 1e-6, block, shards, gout=-1.0)
    ref, gref = o.collapsed_elbo_value_and_grad_autodiff(name, X, y, Z, ell, 1.3, 0.4, 0.2)
    assert abs(val - ref) <= 1e-9 * abs(ref)
    for k in gref:
        a, b = -np.asarray(g[k]).reshape(np.shape(gref[k])), np.asarray(gref[k])
        assert np.max(np.abs(a - b)) <= 1e-7 * max(np.max(np.abs(b)), 1e-8 * abs(ref)), k


@pytest.mark.parametrize("N,M,D,block,shards", [(500, 130, 8, 200, 2), (300, 40, 2, 128, 3), (257, 30, 1, 64, 1)])
def test_sgpr_raw_statistics_route(lib, N, M, D, block, shards):
    """gpb_sgpr_stats_raw (raw Kzx Kxz products, one whitening of the M x M sums) returns the SAME statistics as the
    whiten-first gpb_sgpr_stats up to rounding amplified by cond(Kzz): the lower triangle agrees within
    eps * sqrt(N) * cond, and value / gradients meet the tolerance of the default route while Kzz is well conditioned."""
    X, y = data(N, D, N + M)
    Z = np.ascontiguousarray(np.random.default_rng(M).uniform(-2, 2, (M, D)))
    ell = np.linspace(0.8, 1.6, D)
    cond = np.linalg.cond(o.gram("matern52", Z, ell, 1.3) + 1e-6 * np.eye(M))
    Sw, Sr = [], []
    vw, gw = _sgpr_run(lib, 2, X, y, Z, ell, False, 1.3, 0.4, 0.2, 1e-6, block, shards, stats_out=Sw)
    vr, gr = _sgpr_run(lib, 2, X, y, Z, ell, False, 1.3, 0.4, 0.2, 1e-6, block, shards, raw=True, stats_out=Sr)
    tl = np.tril_indices(M + 2)
    scale = np.max(np.abs(Sw[0][tl]))
    assert np.max(np.abs(Sw[0][tl] - Sr[0][tl])) <= 50 * np.finfo(float).eps * np.sqrt(N) * cond * scale
    ref, gref = o.collapsed_elbo_value_and_grad_autodiff("matern52", X, y, Z, ell, 1.3, 0.4, 0.2)
    amp = max(1.0, cond / 1e3)  # the raw route is meant for cond <= ~1e3 (sgpr_ops.RAW_STATISTICS_COND_LIMIT)
    assert abs(vr - ref) <= 1e-9 * amp * abs(ref)
    for k in gref:
        a, b = np.asarray(gr[k]).reshape(np.shape(gref[k])), np.asarray(gref[k])
        assert np.max(np.abs(a - b)) <= 1e-7 * amp * max(np.max(np.abs(b)), 1e-8 * abs(ref)), (k, cond)


def test_sgpr_zero_mean_and_identity_with_mll(lib):
    """tests/test_objectives.py:170-199 of the reference: ELBO(z = X) ~= MLL (rel 1e-6)."""
    X, y = data(20, 2, 3)
    val, _ = _sgpr_run(lib, 0, X, y, X.copy(), 1.0, True, 1.0, 1.0, None, 1e-6, 8)
    assert abs(val - o.conjugate_mll("rbf", X, y, 1.0, 1.0, 1.0, 0.0)) <= 1e-5 * abs(val)


@pytest.mark.parametrize("kind,name", KINDS)
@pytest.mark.parametrize("N,M,D,iso,block,shards", [(120, 9, 2, False, 50, 1), (400, 40, 3, True, 128, 2),
                                                     (600, 140, 4, False, 300, 3)])
def test_svgp_elbo_value_and_gradient(lib, kind, name, N, M, D, iso, block, shards):
    """SVGP elbo through the shared SGPR statistics protocol + gpb_svgp_finish/grad_finish vs the oracle's autodiff."""
    X, y = data(N, D, N + M)
    rng = np.random.default_rng(M)
    Z = np.ascontiguousarray(rng.uniform(-2, 2, (M, D)))
    mu = rng.standard_normal(M) * 0.3
    W = np.ascontiguousarray(np.tril(rng.standard_normal((M, M)) * 0.1) + 0.7 * np.eye(M))
    W += np.triu(np.full((M, M), 123.0), 1)  # garbage above the diagonal must be ignored
    ell = np.array([0.9]) if iso else np.linspace(0.8, 1.6, D)
    var_a, sn_a, c_a = np.array([1.3]), np.array([0.4]), np.array([0.2])
    ndata, jitter = 5000.0, 1e-6
    nbytes = lib.gpb_sgpr_workspace_bytes(M, D, block)
    cnt = lib.gpb_sgpr_stats_count(M)
    bounds = np.linspace(0, N, shards + 1).astype(int)
    wss, Ps = [np.zeros(nbytes // 8 + 8) for _ in range(shards)], []
    for r in range(shards):
        Xr, yr = np.ascontiguousarray(X[bounds[r]:bounds[r + 1]]), np.ascontiguousarray(y[bounds[r]:bounds[r + 1]])
        P = np.zeros(cnt)
        assert lib.gpb_sgpr_stats(None, kind, len(Xr), M, D, p(Xr), D, p(yr), p(Z), D, p(ell), int(iso), p(var_a), p(sn_a),
                                  p(c_a), jitter, block, p(wss[r]), nbytes, p(P)) == 0
        Ps.append(P)
    Pall = np.sum(Ps, axis=0)
    val, info = np.zeros(1), np.zeros(2, np.int32)
    for r in range(shards):
        assert lib.gpb_svgp_finish(None, kind, M, D, p(Z), D, p(ell), int(iso), p(var_a), p(sn_a), p(c_a), p(mu), p(W), M,
                                   ndata, jitter, block, p(wss[r]), nbytes, p(Pall), 1, p(val), p(info)) == 0
    ellv = ell[0] if iso else ell
    ref, gref = o.svgp_elbo_value_and_grad_autodiff(name, X, y, Z, ellv, 1.3, 0.4, 0.2, mu, np.tril(W), ndata, jitter)
    assert abs(val[0] - ref) <= 1e-9 * abs(ref)
    flat = np.zeros(M * D + (1 if iso else D) + 1)
    nl = 1 if iso else D
    tot = np.zeros_like(flat)
    for r in range(shards):
        Xr, yr = np.ascontiguousarray(X[bounds[r]:bounds[r + 1]]), np.ascontiguousarray(y[bounds[r]:bounds[r + 1]])
        f = np.zeros_like(flat)
        assert lib.gpb_sgpr_grad_local(None, kind, len(Xr), M, D, p(Xr), D, p(yr), p(Z), D, p(ell), int(iso), p(var_a),
                                       p(sn_a), p(c_a), block, p(wss[r]), nbytes, p(f[:M * D]), p(f[M * D:M * D + nl]),
                                       p(f[M * D + nl:])) == 0
        tot += f
    gZ, gl, gv = tot[:M * D].copy(), tot[M * D:M * D + nl].copy(), tot[M * D + nl:].copy()
    gs, gc, gmu, gW = np.zeros(1), np.zeros(1), np.zeros(M), np.full((M, M), np.nan)
    assert lib.gpb_svgp_grad_finish(None, kind, M, D, p(Z), D, p(ell), int(iso), p(var_a), p(sn_a), jitter, block,
                                    p(wss[0]), nbytes, None, p(W), M, p(gZ), p(gl), p(gv), p(gs), p(gc), p(gmu), p(gW), M) == 0
    got = dict(lengthscale=gl, variance=gv[0], obs_stddev=gs[0], mean_const=gc[0], inducing_inputs=gZ.reshape(M, D),
               variational_mean=gmu, variational_root_covariance=gW)
    for k in gref:
        a, b = np.asarray(got[k]).reshape(np.shape(gref[k])), np.asarray(gref[k])
        assert np.max(np.abs(a - b)) <= 1e-7 * max(np.max(np.abs(b)), 1e-8 * abs(ref)), k


def test_potrf_flag_word_symmetrises_the_input_like_jnp_cholesky(lib):
    """jnp.linalg.cholesky factors (A + A^T) / 2 (symmetrize_input=True; gpjax/linalg/operations.py:54-55 passes whatever dense
    array it holds).  Bit 1 of gpb_potrf_lower's flag word does the same; without it only the lower triangle is read."""
    N = 300
    X, _ = data(N, 2, N)
    S = o.gram("rbf", X, np.array([0.9, 1.1]), 1.0) + 0.3 * np.eye(N)
    E = 1e-3 * np.random.default_rng(1).standard_normal((N, N))
    A0 = S + np.triu(E, 1)  # not symmetric: the strict upper triangle is perturbed
    nbytes = lib.gpb_factor_workspace_bytes(N, 2, 0)
    info = np.zeros(1, np.int32)
    out = {}
    for flags in (1, 3):
        ws = np.zeros(nbytes // 8 + 8)
        A = A0.copy()
        assert lib.gpb_potrf_lower(None, N, p(A), N, flags, p(ws), nbytes, N, 2, 0, p(info)) == 0 and info[0] == 0
        assert np.all(np.triu(A, 1) == 0.0)
        out[flags] = A
    assert np.max(np.abs(out[1] - np.linalg.cholesky(S))) <= 1e-12
    assert np.max(np.abs(out[3] - np.linalg.cholesky(0.5 * (A0 + A0.T)))) <= 1e-12
    assert np.max(np.abs(out[1] - out[3])) > 1e-6


# ---- Ozaki (int8 digit plane) trailing updates: orchestration, workspace carving, double buffering ---------------------
@pytest.mark.parametrize("N,planes", [(700, 6), (1024, 7), (900, 4)])
def test_potrf_with_int8_digit_plane_updates(lib, N, planes):
    """Host model block = 256, Ozaki threshold = 256 rows (hostsim/Makefile): N=700 runs two Ozaki steps and one DMMA-model
    step, so digit buffers alternate exactly like the panels under lookahead."""
    X, _ = data(N, 3, N)
    S = o.gram("rbf", X, np.array([0.9, 1.1, 1.3]), 1.0) + (1e-6 + 0.09) * np.eye(N)
    nbytes = lib.gpb_factor_workspace_bytes(N, 3, 0)
    info = np.zeros(1, np.int32)
    out = {}
    try:
        for s in (0, planes):
            lib.gpb_set_ozaki_slices(s)
            assert lib.gpb_get_ozaki_slices() == s
            ws = np.zeros(nbytes // 8 + 8)
            A = S.copy()
            rc = lib.gpb_potrf_lower(None, N, p(A), N, 1, p(ws), nbytes, N, 3, 0, p(info))
            assert rc == 0 and info[0] == 0
            out[s] = np.tril(A)
    finally:
        lib.gpb_set_ozaki_slices(0)
    Lref = np.linalg.cholesky(S)
    assert np.max(np.abs(out[0] - Lref)) <= 1e-12 * np.abs(Lref).max()
    assert not np.array_equal(out[0], out[planes])  # the int8 model really ran
    # radix-256 digit planes: 32 / 48 / 56 bits below the row maximum.  Per block step TWO products are truncated at that level
    # since the panel product X = P inv(L_kk)^T runs on the int8 model too (GPB_OZ_PANELS, default on): measured 1.1e-12 at 6 planes
    # (5.0e-13 with the panel products on the FP64 model), 9.5e-15 at 7 planes = the FP64 model's own 7.5e-15 level
    tol = {4: 1e-7, 6: 2e-12, 7: 1e-12}[planes]
    assert np.max(np.abs(out[planes] - Lref)) <= tol * np.abs(Lref).max()


def test_ozaki_switch_rejects_unsupported_plane_counts(lib):
    try:
        for bad in (1, 3, 8, 9, -3):  # -1 is OZ_AUTO and valid; 4..7 planes of 8 bits
            lib.gpb_set_ozaki_slices(bad)
            assert lib.gpb_get_ozaki_slices() == 0
    finally:
        lib.gpb_set_ozaki_slices(0)


def test_ozaki_slice_and_gemm_host_model_bounds(lib):
    rng = np.random.default_rng(3)
    m, n, k, s = 40, 30, 128, 6
    A = rng.standard_normal((m, k)) * np.exp(rng.standard_normal((m, 1)))
    B = rng.standard_normal((n, k))
    # rows that stress the carry into the top digit: maxima just below a power of two, of both signs, next to tiny entries
    A[0, :4] = [np.nextafter(2.0, 0), -np.nextafter(2.0, 0), 1.9765, -1.9765]
    A[1, :] = -np.nextafter(4.0, 0)
    A[2, :4] = [127.4999 / 128, -127.5001 / 128, 2.0 ** -60, -(2.0 ** -30)]
    A[2, 4:] *= 1e-3
    Qa, Qb = np.zeros((m, s * k), np.int8), np.zeros((n, s * k), np.int8)
    sa, sb = np.zeros(m), np.zeros(n)
    assert lib.gpb_ozaki_slice(None, m, k, p(A), k, s, p(Qa), s * k, p(sa)) == 0
    assert lib.gpb_ozaki_slice(None, n, k, p(B), k, s, p(Qb), s * k, p(sb)) == 0
    assert Qa.min() >= -128 and Qa.max() <= 127 and np.abs(Qa.astype(int)).max() > 64  # the whole int8 range is used
    assert np.all(np.log2(sa) == np.round(np.log2(sa))) and np.all(np.abs(A).max(1) <= 0.494 * sa)
    w = 2.0 ** (-8.0 * (np.arange(s) + 1))
    rec = (Qa.reshape(m, s, k).astype(float) * w[None, :, None]).sum(1) * sa[:, None]
    assert np.max(np.abs(rec - A) / sa[:, None]) <= 2.0 ** (-8 * s - 1)  # ONE rounding to the last plane
    C = np.ones((m, n))
    assert lib.gpb_ozaki_gemm(None, m, n, k, s, p(Qa), s * k, p(sa), p(Qb), s * k, p(sb), -1.0, p(C), n, 0) == 0
    ref = 1.0 - A @ B.T
    bound = k * np.abs(A).max(1)[:, None] * np.abs(B).max(1)[None, :]
    assert np.max(np.abs(C - ref) / bound) <= 2.0 ** (-8 * s + 4)
    Ci = np.zeros((m, n), np.int32)
    assert lib.gpb_igemm_i8(None, m, n, k, p(Qa), s * k, p(Qb), s * k, p(Ci), n) == 0
    assert np.array_equal(Ci, Qa[:, :k].astype(np.int64) @ Qb[:, :k].astype(np.int64).T)


@pytest.mark.parametrize("N", [700, 1100])
def test_mll_value_and_gradient_with_int8_digit_plane_updates(lib, N):
    """potrf, trtri and lauum trailing updates all through the Ozaki model (ragged last block -> zero-padded digit planes)."""
    D = 3
    X, y = data(N, D, N)
    ell, var, sn, c = np.linspace(0.8, 1.6, D), np.array([1.3]), np.array([0.4]), np.array([0.2])
    nbytes = lib.gpb_mll_workspace_bytes(N, D)
    res = {}
    try:
        for s in (0, 6):
            lib.gpb_set_ozaki_slices(s)
            ws = np.zeros(nbytes // 8 + 8)
            Sig = np.full((N, N), np.nan)
            val, alpha, info = np.zeros(1), np.zeros(N), np.zeros(1, np.int32)
            assert lib.gpb_mll_forward(None, 0, N, D, p(X), D, p(y), p(ell), 0, p(var), p(sn), p(c), 1e-6, p(Sig), N, p(ws),
                                       nbytes, p(val), p(alpha), p(info)) == 0
            g_ell, g_var, g_sn, g_c = np.zeros(D), np.zeros(1), np.zeros(1), np.zeros(1)
            assert lib.gpb_mll_backward(None, 0, N, D, p(X), D, p(ell), 0, p(var), p(sn), p(Sig), N, p(ws), nbytes, p(alpha),
                                        None, p(g_ell), p(g_var), p(g_sn), p(g_c)) == 0
            res[s] = (val[0], g_ell.copy(), g_var[0], g_sn[0], g_c[0])
    finally:
        lib.gpb_set_ozaki_slices(0)
    ref = o.conjugate_mll("rbf", X, y, ell, var[0], sn[0], c[0])
    gr = o.conjugate_mll_grad_closed_form("rbf", X, y, ell, var[0], sn[0], c[0])
    v, ge, gv, gs, gc = res[6]
    assert abs(v - ref) <= 1e-10 * abs(ref)
    assert np.max(np.abs(ge - gr["lengthscale"])) <= 1e-8 * np.max(np.abs(gr["lengthscale"]))
    assert abs(gv - gr["variance"]) <= 1e-8 * max(abs(gr["variance"]), 1e-6 * abs(ref))
    assert abs(gs - gr["obs_stddev"]) <= 1e-8 * max(abs(gr["obs_stddev"]), 1e-6 * abs(ref))
    assert abs(gc - gr["mean_const"]) <= 1e-8 * max(abs(gr["mean_const"]), 1e-6 * abs(ref))
    assert not np.array_equal(res[0][1], ge)  # the digit-plane model really ran in the backward pass


def test_auto_mode_guard_picks_planes_from_the_hyperparameters(lib):
    """OZ_AUTO (-1): the plane count of the int8 updates is written into the workspace by ozaki_choose_planes -- 7 for a bare
    matrix (gpb_potrf_lower), 6 only while (N variance + s) / s <= 2e6, s = obs_stddev^2 + jitter, for the fused objective --
    and the product kernels read it from there.  The host model's workspace is host memory, so the word can be inspected."""
    N, D = 700, 3
    X, y = data(N, D, N)
    ell, c = np.linspace(0.8, 1.6, D), np.array([0.0])
    assert lib.gpb_ozaki_auto_planes(50000, 1.0, 0.3, 1e-6) == 6       # the benchmark's hyper-parameters: bound 5.6e5
    assert lib.gpb_ozaki_auto_planes(100000, 1.0, 0.3, 1e-6) == 6      # config 3: 1.1e6
    assert lib.gpb_ozaki_auto_planes(8192, 1.0, 0.1, 1e-6) == 6        # 8.2e5
    assert lib.gpb_ozaki_auto_planes(8192, 1.0, 0.05, 1e-6) == 7       # 3.3e6 > 2e6
    assert lib.gpb_ozaki_auto_planes(8192, 1.0, 0.03, 1e-6) == 7       # 9.1e6
    assert lib.gpb_ozaki_auto_planes(8192, 1.0, 0.003, 1e-6) == 7      # 8.2e8
    assert lib.gpb_ozaki_auto_planes(50000, 1.0, 0.0, 0.0) == 7        # s = 0 -> inf -> 7
    nbytes = lib.gpb_mll_workspace_bytes(N, D)
    words = {}
    try:
        for tag, (var, sn, mode) in {"benign": (1.3, 0.4, -1), "ill": (1.3, 1e-3, -1), "forced5": (1.3, 0.4, 5)}.items():
            lib.gpb_set_ozaki_slices(mode)
            assert lib.gpb_get_ozaki_slices() == mode
            ws = np.zeros(nbytes // 8 + 8)
            Sig = np.full((N, N), np.nan)
            val, alpha, info = np.zeros(1), np.zeros(N), np.zeros(1, np.int32)
            assert lib.gpb_mll_forward(None, 0, N, D, p(X), D, p(y), p(ell), 0, p(np.array([var])), p(np.array([sn])), p(c), 1e-6,
                                       p(Sig), N, p(ws), nbytes, p(val), p(alpha), p(info)) == 0
            # the guard's word is the first int32 of the LAST 256-byte slot of the workspace (algorithms.cpp: off_ozp)
            words[tag] = int(ws[: nbytes // 8].view(np.int32)[-64])
            assert np.isfinite(val[0])
        lib.gpb_set_ozaki_slices(-1)
        S = o.gram("rbf", X, ell, 1.0) + 0.09 * np.eye(N)
        nb2 = lib.gpb_factor_workspace_bytes(N, D, 0)
        ws = np.zeros(nb2 // 8 + 8)
        assert lib.gpb_potrf_lower(None, N, p(S), N, 1, p(ws), nb2, N, D, 0, p(np.zeros(1, np.int32))) == 0
        words["bare"] = int(ws[: nb2 // 8].view(np.int32)[-64])
    finally:
        lib.gpb_set_ozaki_slices(0)
    assert words == {"benign": 6, "ill": 7, "forced5": 5, "bare": 7}


@pytest.mark.parametrize("raw", [False, True])
def test_sgpr_statistics_and_pass2_through_int8_digit_planes(lib, raw):
    """collapsed_elbo with the statistics SYRK (column digit planes, contraction over the block rows, K split for the int32
    headroom) and the pass-2 product dK_b = [K_b|d|1] Caug^T (M + 2 = 262 digits per plane zero-padded to 384) on the Ozaki
    model; the ragged last block (100 rows, below the threshold) stays on the GEMM model."""
    N, M, D, block = 700, 260, 3, 300
    X, y = data(N, D, N + M)
    Z = np.ascontiguousarray(np.random.default_rng(M).uniform(-2, 2, (M, D)))
    ell = np.linspace(0.8, 1.6, D)
    try:
        lib.gpb_set_ozaki_slices(0)
        v0, g0 = _sgpr_run(lib, 0, X, y, Z, ell, False, 1.3, 0.4, 0.2, 1e-6, block, 1, raw=raw)
        lib.gpb_set_ozaki_slices(7)
        v7, g7 = _sgpr_run(lib, 0, X, y, Z, ell, False, 1.3, 0.4, 0.2, 1e-6, block, 1, raw=raw)
    finally:
        lib.gpb_set_ozaki_slices(0)
    assert v0 != v7  # the int8 model really ran in the forward pass
    cond = np.linalg.cond(o.gram("rbf", Z, ell, 1.3) + 1e-6 * np.eye(M))
    amp = max(1.0, cond / 1e3) if raw else 1.0
    amp2 = max(1.0, cond / 1e6)  # two fp64-accurate evaluation orders of the statistics drift like cond * eps
    assert abs(v0 - v7) <= 1e-11 * amp * abs(v0)
    assert not np.array_equal(g0["inducing_inputs"], g7["inducing_inputs"])
    ref, gref = o.collapsed_elbo_value_and_grad_autodiff("rbf", X, y, Z, ell, 1.3, 0.4, 0.2)
    assert abs(v7 - ref) <= 1e-9 * amp * abs(ref)
    for k in gref:
        a, b = np.asarray(g7[k]).reshape(np.shape(gref[k])), np.asarray(gref[k])
        assert np.max(np.abs(a - b)) <= 1e-7 * amp * max(np.max(np.abs(b)), 1e-8 * abs(ref)), k
        a0 = np.asarray(g0[k]).reshape(np.shape(gref[k]))
        assert np.max(np.abs(a - a0)) <= 1e-10 * max(amp, amp2) * max(np.max(np.abs(b)), 1e-8 * abs(ref)), (k, cond)


def test_svgp_finish_dense_products_on_the_int8_model(lib):
    """The dense M x M x M products of the replicated SVGP finish (V = Lz^-1 W, V V^T, Phi Ttil, Phi V, Linv^T G Linv,
    -Linv^T H) as 7-plane digit products (mm_gemm in sgpr.cpp; the host model routes extents >= 256): row digit planes for
    K-contiguous operands, column digit planes for MN-layout ones, triangular K-ranges, lower mask, beta = 0.  Well-conditioned Kzz
    (cond ~ 4e1: the precondition bit 1 of the flag word states); flag 1 keeps every dense product on the FP64 GEMM model, and both
    must sit far inside the oracle tolerance."""
    N, M, D, block = 600, 260, 5, 300
    X, y = data(N, D, N + M)
    rng = np.random.default_rng(M)
    Z = np.ascontiguousarray(rng.uniform(-2, 2, (M, D)))
    mu = rng.standard_normal(M) * 0.3
    W = np.ascontiguousarray(np.tril(rng.standard_normal((M, M)) * 0.02) + 0.7 * np.eye(M))
    ell = np.full(D, 0.8)
    var_a, sn_a, c_a = np.array([1.3]), np.array([0.4]), np.array([0.2])
    ndata, jitter = 5000.0, 1e-6
    assert np.linalg.cond(o.gram("rbf", Z, ell, 1.3) + jitter * np.eye(M)) < 1e3
    ref, gref = o.svgp_elbo_value_and_grad_autodiff("rbf", X, y, Z, ell, 1.3, 0.4, 0.2, mu, np.tril(W), ndata, jitter)
    nbytes = lib.gpb_sgpr_workspace_bytes(M, D, block)
    out = {}
    for flag in (1, 3):
        ws, P = np.zeros(nbytes // 8 + 8), np.zeros(lib.gpb_sgpr_stats_count(M))
        val, info = np.zeros(1), np.zeros(2, np.int32)
        try:
            lib.gpb_set_ozaki_slices(7)
            assert lib.gpb_sgpr_stats(None, 0, N, M, D, p(X), D, p(y), p(Z), D, p(ell), 0, p(var_a), p(sn_a), p(c_a), jitter,
                                      block, p(ws), nbytes, p(P)) == 0  # whitened statistics: identical input for both flags
            assert lib.gpb_svgp_finish(None, 0, M, D, p(Z), D, p(ell), 0, p(var_a), p(sn_a), p(c_a), p(mu), p(W), M, ndata, jitter,
                                       block, p(ws), nbytes, p(P), flag, p(val), p(info)) == 0 and not info.any()
            f = np.zeros(M * D + D + 1)
            assert lib.gpb_sgpr_grad_local(None, 0, N, M, D, p(X), D, p(y), p(Z), D, p(ell), 0, p(var_a), p(sn_a), p(c_a), block,
                                           p(ws), nbytes, p(f[:M * D]), p(f[M * D:M * D + D]), p(f[M * D + D:])) == 0
            gZ, gl, gv = f[:M * D].copy(), f[M * D:M * D + D].copy(), f[M * D + D:].copy()
            gs, gc, gmu, gW = np.zeros(1), np.zeros(1), np.zeros(M), np.full((M, M), np.nan)
            assert lib.gpb_svgp_grad_finish(None, 0, M, D, p(Z), D, p(ell), 0, p(var_a), p(sn_a), jitter, block, p(ws), nbytes, None,
                                            p(W), M, p(gZ), p(gl), p(gv), p(gs), p(gc), p(gmu), p(gW), M) == 0
        finally:
            lib.gpb_set_ozaki_slices(0)
        assert abs(val[0] - ref) <= 1e-10 * abs(ref)
        got = dict(lengthscale=gl, variance=gv[0], obs_stddev=gs[0], mean_const=gc[0], inducing_inputs=gZ.reshape(M, D),
                   variational_mean=gmu, variational_root_covariance=gW)
        assert np.all(np.triu(gW, 1) == 0.0)
        for k in gref:
            a, b = np.asarray(got[k]).reshape(np.shape(gref[k])), np.asarray(gref[k])
            assert np.max(np.abs(a - b)) <= 1e-9 * max(np.max(np.abs(b)), 1e-8 * abs(ref)), (k, flag)
        out[flag] = got
    # the flag really changes the arithmetic of the dense products (digit planes vs the FP64 GEMM model) ...
    assert not np.array_equal(out[1]["variational_root_covariance"], out[3]["variational_root_covariance"])
    assert not np.array_equal(out[1]["inducing_inputs"], out[3]["inducing_inputs"])
    # ... and at this conditioning only at the 1e-12 level
    for k in gref:
        a, b = np.asarray(out[1][k]), np.asarray(out[3][k])
        assert np.max(np.abs(a - b)) <= 1e-11 * max(np.max(np.abs(np.asarray(gref[k]))), 1e-8 * abs(ref)), k


@pytest.mark.parametrize("raw", [False, True])
def test_sgpr_empty_row_shard(lib, raw):
    """A rank that holds no rows (N not divisible into non-empty shards, or a minibatch drawn elsewhere) passes N = 0 and null data
    pointers through the protocol: its statistics and local gradient are exact zeros (every word written, none left as it was), and
    the reduced result is the full-data ELBO / gradient."""
    rng = np.random.default_rng(3)
    N, M, D, block = 90, 12, 2, 32
    X, y = data(N, D, 5)
    Z = np.ascontiguousarray(rng.uniform(-2, 2, (M, D)))
    ell, var, sn, c = np.linspace(0.8, 1.4, D), np.array([1.3]), np.array([0.4]), np.array([0.2])
    nbytes = lib.gpb_sgpr_workspace_bytes(M, D, block)
    stats = lib.gpb_sgpr_stats_raw if raw else lib.gpb_sgpr_stats
    shards = [(np.ascontiguousarray(X), np.ascontiguousarray(y)), (None, None)]
    wss = [np.zeros(nbytes // 8 + 8) for _ in shards]
    Ps = []
    for (Xr, yr), ws in zip(shards, wss):
        P = np.full(lib.gpb_sgpr_stats_count(M), np.nan)
        assert stats(None, 0, 0 if Xr is None else len(Xr), M, D, p(Xr), D, p(yr), p(Z), D, p(ell), 0, p(var), p(sn), p(c), 1e-6,
                     block, p(ws), nbytes, p(P)) == 0
        Ps.append(P)
    assert not np.isnan(Ps[1]).any() and np.all(Ps[1] == 0.0)
    Pall = Ps[0] + Ps[1]
    val, info = np.zeros(1), np.zeros(2, np.int32)
    for ws in wss:
        Pc = Pall.copy()
        assert lib.gpb_sgpr_finish(None, 0, M, D, p(Z), D, p(ell), 0, p(var), p(sn), block, p(ws), nbytes, p(Pc),
                                   1 | (2 if raw else 0), p(val), p(info)) == 0
    ref, gref = o.collapsed_elbo_value_and_grad_autodiff("rbf", X, y, Z, ell, 1.3, 0.4, 0.2)
    assert abs(val[0] - ref) <= 1e-9 * abs(ref)
    tot = np.zeros(M * D + D + 1)
    for (Xr, yr), ws in zip(shards, wss):
        f = np.full(M * D + D + 1, np.nan)
        assert lib.gpb_sgpr_grad_local(None, 0, 0 if Xr is None else len(Xr), M, D, p(Xr), D, p(yr), p(Z), D, p(ell), 0, p(var),
                                       p(sn), p(c), block, p(ws), nbytes, p(f[:M * D]), p(f[M * D:M * D + D]), p(f[M * D + D:])) == 0
        if Xr is None:
            assert np.all(f == 0.0)
        tot += f
    gZ, gl, gv = tot[:M * D].copy(), tot[M * D:M * D + D].copy(), tot[M * D + D:].copy()
    gs, gc = np.zeros(1), np.zeros(1)
    assert lib.gpb_sgpr_grad_finish(None, 0, M, D, p(Z), D, p(ell), 0, p(var), p(sn), block, p(wss[1]), nbytes, None, p(gZ), p(gl),
                                    p(gv), p(gs), p(gc)) == 0
    got = dict(lengthscale=gl, variance=gv[0], obs_stddev=gs[0], mean_const=gc[0], inducing_inputs=gZ.reshape(M, D))
    for k in got:
        a, b = np.asarray(got[k]).reshape(np.shape(gref[k])), np.asarray(gref[k])
        assert np.max(np.abs(a - b)) <= 1e-7 * max(np.max(np.abs(b)), 1e-8 * abs(ref)), k
