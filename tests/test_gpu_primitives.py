"""Parity of the hand-written CUDA primitives (called through the C ABI) on a real B200."""
import numpy as np
import pytest
import torch

import oracle as o

pytestmark = pytest.mark.gpu

KINDS = [(0, "rbf"), (1, "matern32"), (2, "matern52"), (3, "matern12")]


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda")


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


# ---------------------------------------------------------------- GEMM (DMMA) -------------
@pytest.mark.parametrize("al,bl", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(128, 64, 16), (257, 131, 77), (1, 1, 1), (300, 500, 1000), (64, 2048, 5)])
def test_gemm_layouts(al, bl, M, N, K):
    from gpjax_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K + al * 2 + bl)
    A = torch.randn((M, K) if al == 0 else (K, M), dtype=torch.float64, device="cuda", generator=g)
    B = torch.randn((N, K) if bl == 0 else (K, N), dtype=torch.float64, device="cuda", generator=g)
    C0 = torch.randn((M, N), dtype=torch.float64, device="cuda", generator=g)
    Am = A if al == 0 else A.T
    Bm = B if bl == 0 else B.T
    ref = 0.5 * C0 + 1.5 * (Am @ Bm.T)
    C = C0.clone()
    ops.gemm(A, B, C, alpha=1.5, beta=0.5, a_layout=al, b_layout=bl)
    torch.cuda.synchronize()
    scale = float((Am.abs() @ Bm.abs().T).max()) + 1.0
    assert float((C - ref).abs().max()) <= 1e-13 * scale


def test_gemm_unaligned_views_and_masks():
    from gpjax_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(5)
    big = torch.randn((400, 403), dtype=torch.float64, device="cuda", generator=g)
    A = big[3:203, 1:150]       # odd leading dimension, 8-byte aligned only
    B = big[210:399, 2:151]
    Cbuf = torch.zeros((200, 191), dtype=torch.float64, device="cuda")
    C = Cbuf[:, 1:190]
    for mask, fn in [(1, torch.tril), (2, torch.triu)]:
        C.zero_()
        ops.gemm(A, B, C, mask=mask)
        torch.cuda.synchronize()
        ref = fn(A @ B.T)
        assert float((C - ref).abs().max()) <= 1e-12


# ---------------------------------------------------------------- Gram --------------------
@pytest.mark.parametrize("kind,name", KINDS)
@pytest.mark.parametrize("N,M,D,iso", [(1, 1, 1, True), (5, 3, 2, False), (200, 333, 8, False), (1000, 1000, 1, True),
                                       (513, 129, 3, False), (77, 900, 16, False), (64, 128, 20, True)])
def test_gram_vs_oracle(kind, name, N, M, D, iso):
    from gpjax_b200 import ops

    rng = np.random.default_rng(N + M + D)
    X = rng.uniform(-2, 2, (N, D))
    Z = rng.uniform(-2, 2, (M, D))
    ell = np.array(0.7) if iso else np.linspace(0.8, 1.6, D)
    var = 1.7
    K = ops.gram_forward(kind, dev(X), dev(Z), dev(ell), dev(var)).cpu().numpy()
    ref = o.cross_covariance(name, X, Z, ell, var)
    assert rel(K, ref) <= 1e-12  # north-star tolerance for Gram entries


@pytest.mark.parametrize("kind,name", KINDS)
def test_gram_adversarial_near_duplicates(kind, name):
    """l = 0.1, 1 % near-duplicate rows (x + 1e-9): checked against the 80-bit adjudicator."""
    from gpjax_b200 import ops

    rng = np.random.default_rng(11)
    N, D = 600, 8
    X = rng.uniform(-2, 2, (N, D))
    X[::100] = X[1::100] + 1e-9
    ell = np.full(D, 0.1)
    K = ops.gram_forward(kind, dev(X), dev(X), dev(ell), dev(1.0)).cpu().numpy()
    ref64 = o.gram(name, X, ell, 1.0)
    truth = o.gram_longdouble(name, X, X, ell, 1.0)
    live = np.abs(np.asarray(truth, np.float64)) > 1e-280
    assert rel(K[live], ref64[live]) <= 1e-12
    assert rel(K[live], np.asarray(truth, np.float64)[live]) <= 1e-12
    assert np.all(np.diag(K) == 1.0)  # exact on the diagonal


@pytest.mark.parametrize("kind,name", KINDS)
def test_gram_symmetric_diag_and_lower_only(kind, name):
    from gpjax_b200 import ops

    rng = np.random.default_rng(3)
    N, D = 700, 4
    X = rng.uniform(-2, 2, (N, D))
    ell = np.linspace(0.5, 1.0, D)
    sn = 0.3
    out = torch.full((N, N), float("nan"), dtype=torch.float64, device="cuda")
    ops.gram_forward(kind, dev(X), dev(X), dev(ell), dev(0.9), diag_add=1e-6, diag_add_sq=dev(sn), lower_only=True,
                     out=out)
    K = out.cpu().numpy()
    ref = o.gram(name, X, ell, 0.9) + np.eye(N) * 1e-6 + np.eye(N) * sn**2
    il = np.tril_indices(N)
    assert rel(K[il], ref[il]) <= 1e-12
    assert np.isnan(K[0, N - 1])  # far upper tiles untouched


@pytest.mark.parametrize("kind,name", KINDS)
@pytest.mark.parametrize("iso", [False, True])
def test_gram_backward_vs_autodiff(kind, name, iso):
    from gpjax_b200 import ops

    rng = np.random.default_rng(17)
    N, M, D = 150, 260, 3
    X = rng.uniform(-2, 2, (N, D))
    Z = rng.uniform(-2, 2, (M, D))
    ell = np.array(0.9) if iso else np.array([0.7, 1.1, 1.4])
    dK = rng.standard_normal((N, M))
    # torch-CPU autodiff of the literal restatement
    Xt, Zt = torch.tensor(X, requires_grad=True), torch.tensor(Z, requires_grad=True)
    et, vt = torch.tensor(ell, requires_grad=True), torch.tensor(1.3, dtype=torch.float64, requires_grad=True)
    from oracle.gp_oracle import _t_cross

    (_t_cross(torch, kind, Xt, Zt, et, vt) * torch.tensor(dK)).sum().backward()
    g_ell, g_var, g_X, g_Z = ops.gram_backward(kind, dev(X), dev(Z), dev(ell), dev(1.3), dev(dK), want_X=True,
                                               want_Z=True)
    assert rel(g_ell.cpu().numpy().reshape(et.grad.shape), et.grad.numpy()) <= 1e-10
    assert rel(g_var.cpu().numpy().reshape(()), vt.grad.numpy()) <= 1e-10
    sx = np.abs(Xt.grad.numpy()).max()
    assert np.max(np.abs(g_X.cpu().numpy() - Xt.grad.numpy())) <= 1e-11 * sx
    assert np.max(np.abs(g_Z.cpu().numpy() - Zt.grad.numpy())) <= 1e-11 * np.abs(Zt.grad.numpy()).max()


# ---------------------------------------------------------------- factorisation family ----
def _spd(n, seed, kind="rbf", D=4):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-2, 2, (n, D))
    return o.gram(kind, X, np.linspace(0.8, 1.4, D), 1.0) + 0.09 * np.eye(n)


@pytest.mark.parametrize("n", [1, 2, 5, 100, 128, 129, 256, 300, 513, 1024, 1025, 1500, 2300, 3200])  # block = 1024
def test_potrf_trsv_trsm_logdet_potri(n):
    from gpjax_b200 import ops

    S = _spd(n, n)
    A = dev(S)
    ws = ops.FactorWorkspace(n, 1, potri=True, device="cuda")
    info = ops.potrf_lower_(A, ws, zero_upper=True)
    L = A.cpu().numpy()
    assert int(info.item()) == 0
    Lref = np.linalg.cholesky(S)
    assert np.max(np.abs(L - Lref)) <= 1e-12 * np.abs(Lref).max()
    assert np.all(np.triu(L, 1) == 0.0)
    assert abs(float(ops.sum_log_diag(A)) - np.sum(np.log(np.diag(Lref)))) <= 1e-12 * max(1.0, n)
    rng = np.random.default_rng(n)
    b = rng.standard_normal(n)
    import scipy.linalg as sla

    # element-wise on every size: above n = 3072 the trailing updates run as int8 digit-plane products, and the default of 8
    # planes keeps the solves at the FP64 path's own level (profiles/r02_solve_elementwise.jsonl: 2.4e-10 vs 1.2e-10 at
    # n = 3200; the 7-plane opt-in is at 3.2e-8 and would fail here)
    close = lambda a, r: rel(a, r) <= 1e-9
    x = ops.trsv_lower_(A, dev(b), ws).cpu().numpy()
    assert close(x, sla.solve_triangular(Lref, b, lower=True))
    xt = ops.trsv_lower_(A, dev(b), ws, trans=True).cpu().numpy()
    assert close(xt, sla.solve_triangular(Lref.T, b, lower=False))
    T = min(n, 37)
    Bm = rng.standard_normal((n, T))
    Xm = ops.trsm_lower_left_(A, dev(Bm), ws).cpu().numpy()
    assert np.max(np.abs(Lref @ Xm - Bm)) <= 1e-10 * np.abs(Bm).max() * n
    Xmt = ops.trsm_lower_left_(A, dev(Bm), ws, trans=True).cpu().numpy()
    assert np.max(np.abs(Lref.T @ Xmt - Bm)) <= 1e-10 * np.abs(Bm).max() * n
    Sinv = ops.potri_lower(A, ws).cpu().numpy()
    assert np.max(np.abs(Sinv @ S - np.eye(n))) <= 1e-9


def test_potrf_not_positive_definite_nan_fills():
    from gpjax_b200 import ops

    n = 300
    S = _spd(n, 1)
    S[200, 200] = -1.0
    A = dev(S)
    ws = ops.FactorWorkspace(n, 1, device="cuda")
    info = ops.potrf_lower_(A, ws)
    assert int(info.item()) == 201
    assert torch.isnan(A[299, 299])
    assert not torch.isnan(A[100, 50])  # rows factored before the failure stay valid


@pytest.mark.parametrize("D", [33, 64])
def test_gram_and_backward_large_input_dim(D):
    """Input dimensions beyond one 8-wide chunk, up to the compiled maximum (64)."""
    from gpjax_b200 import ops
    from oracle.gp_oracle import _t_cross

    rng = np.random.default_rng(D)
    N, M = 130, 200
    X, Z = rng.uniform(-1, 1, (N, D)), rng.uniform(-1, 1, (M, D))
    ell = np.linspace(2.0, 4.0, D)
    K = ops.gram_forward(2, dev(X), dev(Z), dev(ell), dev(1.3)).cpu().numpy()
    assert rel(K, o.cross_covariance("matern52", X, Z, ell, 1.3)) <= 1e-12
    dK = rng.standard_normal((N, M))
    Xt, Zt = torch.tensor(X, requires_grad=True), torch.tensor(Z, requires_grad=True)
    et, vt = torch.tensor(ell, requires_grad=True), torch.tensor(1.3, dtype=torch.float64, requires_grad=True)
    (_t_cross(torch, 2, Xt, Zt, et, vt) * torch.tensor(dK)).sum().backward()
    g_ell, g_var, g_X, g_Z = ops.gram_backward(2, dev(X), dev(Z), dev(ell), dev(1.3), dev(dK), want_X=True, want_Z=True)
    assert np.max(np.abs(g_ell.cpu().numpy() - et.grad.numpy())) <= 1e-10 * np.abs(et.grad.numpy()).max()
    assert abs(g_var.item() - vt.grad.item()) <= 1e-10 * abs(vt.grad.item())
    assert np.max(np.abs(g_X.cpu().numpy() - Xt.grad.numpy())) <= 1e-10 * np.abs(Xt.grad.numpy()).max()
    assert np.max(np.abs(g_Z.cpu().numpy() - Zt.grad.numpy())) <= 1e-10 * np.abs(Zt.grad.numpy()).max()


def test_input_dim_above_compiled_maximum_is_rejected():
    from gpjax_b200 import ops

    X = torch.zeros((4, 65), dtype=torch.float64, device="cuda")
    with pytest.raises(RuntimeError, match="UNSUPPORTED"):
        ops.gram_forward(0, X, X, torch.ones(65, dtype=torch.float64, device="cuda"), dev(1.0))
