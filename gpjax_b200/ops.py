"""Thin torch <-> C-ABI shim: device memory, streams and autograd plumbing only.

Every function here enqueues hand-written sm_100a kernels from ``libgpjax_b200.so`` on the current
CUDA stream through the C ABI of ``include/gpjax_b200.h``; none of them computes anything with
torch ops.  Tensors must be float64, on the GPU and contiguous (asserted, never copied silently
except where a ``.contiguous()`` is explicitly documented).  The ``torch.autograd.Function``
classes are the analogue of the ``jax.custom_vjp`` registrations the north-star describes.
"""
from __future__ import annotations

import itertools
import math
from typing import Optional

import torch

from . import _abi
from ._lib import lib, require_cuda

KIND_IDS = {"RBF": 0, "Matern32": 1, "Matern52": 2, "Matern12": 3, "Matérn32": 1, "Matérn52": 2, "Matérn12": 3,
            "RationalQuadratic": 4, "PoweredExponential": 5, "Periodic": 6, "White": 7}
SHAPE_KINDS = (4, 5, 6)  # kinds whose `variance` argument is the pair [variance, shape] (include/gpjax_b200.h)


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


# Forward "generation" tokens: drawn from ONE process-wide counter, so a state object that is evicted and re-created
# (another problem size took the buffer) can never hand out a number an older forward still holds.
_GENERATION = itertools.count(1)


def next_generation() -> int:
    return next(_GENERATION)


def _no_data_grad(ctx, positions, what: str) -> None:
    """The fused objectives return no cotangent for the data (X, y): refuse instead of silently returning zero."""
    if any(ctx.needs_input_grad[i] for i in positions):
        raise NotImplementedError(
            f"{what}: gradients with respect to the data (X / y) are not produced by the fused objective; build the "
            "covariance with kernel.gram(...) / GaussianDistribution.log_prob (ops.GramFunction + GaussianLogProbFunction), "
            "which differentiate through the inputs")


def _check_mat(t: torch.Tensor, name: str) -> None:
    require_cuda(t)
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError(f"{name} must be a 2-D float64 CUDA tensor with unit column stride")


def _ell_args(ell: torch.Tensor, D: int):
    require_cuda(ell)
    if ell.numel() == 1:
        return ell.reshape(1).contiguous(), 1
    if ell.numel() != D:
        raise ValueError(f"lengthscale has {ell.numel()} entries but the inputs have {D} dimensions")
    return ell.reshape(D).contiguous(), 0


def _scalar(t: torch.Tensor, name: str) -> torch.Tensor:
    require_cuda(t)
    if t.numel() != 1:
        raise ValueError(f"{name} must hold exactly one element")
    return t.reshape(1)


def _kscalars(kind: int, t: torch.Tensor) -> torch.Tensor:
    """The kernel scalars as the C ABI wants them: [variance], or [variance, shape] for kinds 4..6."""
    require_cuda(t)
    want = 2 if kind in SHAPE_KINDS else 1
    if t.numel() != want:
        raise ValueError(f"kernel kind {kind} takes {want} scalar(s) (variance{', shape' if want == 2 else ''}); "
                         f"got {t.numel()}")
    return t.reshape(want).contiguous()


# ------------------------------------------------------------------------------------------
# K1: Gram / cross-covariance
# ------------------------------------------------------------------------------------------
def gram_forward(kind: int, X, Z, ell, variance, diag_add=0.0, diag_add_sq=None, lower_only=False, out=None):
    _check_mat(X, "X")
    _check_mat(Z, "Z")
    N, D = X.shape
    M = Z.shape[0]
    if Z.shape[1] != D:
        raise ValueError("X and Z must have the same number of columns")
    ell_v, iso = _ell_args(ell, D)
    var = _kscalars(kind, variance)
    if out is None:
        out = torch.empty((N, M), dtype=torch.float64, device=X.device)
    _check_mat(out, "out")
    rc = lib().gpb_gram(_stream(), kind, N, M, D, _p(X), X.stride(0), _p(Z), Z.stride(0), _p(ell_v), iso, _p(var),
                        float(diag_add), _p(diag_add_sq), int(lower_only), _p(out), out.stride(0))
    _abi.check(rc, "gpb_gram")
    return out


def gram_backward(kind: int, X, Z, ell, variance, dK, want_X=False, want_Z=False, scale=1.0, accum=None):
    """Returns (g_ell, g_var, g_X, g_Z); `accum` may carry existing tensors to accumulate into."""
    _check_mat(X, "X")
    _check_mat(Z, "Z")
    _check_mat(dK, "dK")
    N, D = X.shape
    M = Z.shape[0]
    ell_v, iso = _ell_args(ell, D)
    var = _kscalars(kind, variance)
    dev = X.device
    accum = accum or {}
    g_ell = accum.get("ell")
    if g_ell is None:
        g_ell = torch.zeros(1 if iso else D, dtype=torch.float64, device=dev)
    g_var = accum.get("var")
    if g_var is None:
        g_var = torch.zeros(var.numel(), dtype=torch.float64, device=dev)
    g_X = accum.get("X")
    if g_X is None and want_X:
        g_X = torch.zeros((N, D), dtype=torch.float64, device=dev)
    g_Z = accum.get("Z")
    if g_Z is None and want_Z:
        g_Z = torch.zeros((M, D), dtype=torch.float64, device=dev)
    nbytes = lib().gpb_gram_bwd_workspace_bytes(N, M, D)
    ws = torch.empty(max(nbytes // 8, 1), dtype=torch.float64, device=dev)
    rc = lib().gpb_gram_bwd(_stream(), kind, N, M, D, _p(X), X.stride(0), _p(Z), Z.stride(0), _p(ell_v), iso,
                            _p(var), _p(dK), dK.stride(0), float(scale), _p(ws), nbytes, _p(g_ell), _p(g_var),
                            _p(g_X), D, _p(g_Z), D)
    _abi.check(rc, "gpb_gram_bwd")
    return g_ell, g_var, g_X, g_Z


class GramFunction(torch.autograd.Function):
    """Differentiable K(X, Z); backward never materialises dK/dtheta."""

    @staticmethod
    def forward(ctx, kind, X, Z, ell, variance, symmetric):
        ctx.kind = kind
        ctx.symmetric = symmetric
        ctx.save_for_backward(X, Z, ell, variance)
        return gram_forward(kind, X, Z, ell, variance)

    @staticmethod
    def backward(ctx, dK):
        X, Z, ell, variance = ctx.saved_tensors
        dK = dK.contiguous()
        nx, nz = ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        g_ell, g_var, g_X, g_Z = gram_backward(ctx.kind, X, Z, ell, variance, dK, want_X=nx, want_Z=nz)
        g_ell = g_ell.reshape(ell.shape) if ctx.needs_input_grad[3] else None
        g_var = g_var.reshape(variance.shape) if ctx.needs_input_grad[4] else None
        return None, g_X, g_Z, g_ell, g_var, None


# ------------------------------------------------------------------------------------------
# K2-K5: factorisation family
# ------------------------------------------------------------------------------------------
class FactorWorkspace:
    """Device scratch for the blocked factorisation (diagonal-block inverses, panel, partials)."""

    def __init__(self, n: int, d: int = 1, potri: bool = False, device=None):
        self.n, self.d, self.potri = int(n), int(d), int(bool(potri))
        self.nbytes = lib().gpb_factor_workspace_bytes(self.n, self.d, self.potri)
        self.buf = torch.empty(max(self.nbytes // 8, 1), dtype=torch.float64, device=device)

    def args(self):
        return _p(self.buf), self.nbytes, self.n, self.d, self.potri


def potrf_lower_(A: torch.Tensor, ws: FactorWorkspace, zero_upper: bool = True, symmetrize: bool = False) -> torch.Tensor:
    """In-place lower Cholesky of the lower triangle of A.  Returns the device `info` word.  ``symmetrize`` first replaces
    the lower triangle by (A + A^T) / 2, the ``symmetrize_input=True`` default of ``jnp.linalg.cholesky``
    (gpjax/linalg/operations.py:54-55 calls it on whatever dense array it is given)."""
    _check_mat(A, "A")
    n = A.shape[0]
    info = torch.zeros(1, dtype=torch.int32, device=A.device)
    rc = lib().gpb_potrf_lower(_stream(), n, _p(A), A.stride(0), int(bool(zero_upper)) | (2 if symmetrize else 0), *ws.args(),
                               _p(info))
    _abi.check(rc, "gpb_potrf_lower")
    return info


def diag_inverses(L: torch.Tensor, ws: FactorWorkspace) -> None:
    _check_mat(L, "L")
    rc = lib().gpb_diag_inverses(_stream(), L.shape[0], _p(L), L.stride(0), *ws.args())
    _abi.check(rc, "gpb_diag_inverses")


def trsv_lower_(L: torch.Tensor, x: torch.Tensor, ws: FactorWorkspace, trans: bool = False) -> torch.Tensor:
    _check_mat(L, "L")
    require_cuda(x)
    assert x.dim() == 1 and x.is_contiguous()
    rc = lib().gpb_trsv_lower(_stream(), L.shape[0], _p(L), L.stride(0), int(trans), _p(x), *ws.args())
    _abi.check(rc, "gpb_trsv_lower")
    return x


def trsm_lower_left_(L: torch.Tensor, B: torch.Tensor, ws: FactorWorkspace, trans: bool = False) -> torch.Tensor:
    _check_mat(L, "L")
    _check_mat(B, "B")
    rc = lib().gpb_trsm_lower_left(_stream(), L.shape[0], B.shape[1], _p(L), L.stride(0), int(trans), _p(B),
                                   B.stride(0), *ws.args())
    _abi.check(rc, "gpb_trsm_lower_left")
    return B


def sum_log_diag(L: torch.Tensor) -> torch.Tensor:
    _check_mat(L, "L")
    out = torch.empty(1, dtype=torch.float64, device=L.device)
    rc = lib().gpb_sum_log_diag(_stream(), L.shape[0], _p(L), L.stride(0), _p(out))
    _abi.check(rc, "gpb_sum_log_diag")
    return out.reshape(())


def potri_lower(Lbuf: torch.Tensor, ws: FactorWorkspace) -> torch.Tensor:
    """Sigma^-1 (full symmetric) from a factor produced by potrf_lower_ with the same workspace.
    The strict upper blocks of `Lbuf` are used as scratch."""
    _check_mat(Lbuf, "L")
    n = Lbuf.shape[0]
    out = torch.empty((n, n), dtype=torch.float64, device=Lbuf.device)
    rc = lib().gpb_potri_lower(_stream(), n, _p(Lbuf), Lbuf.stride(0), _p(out), out.stride(0), *ws.args())
    _abi.check(rc, "gpb_potri_lower")
    return out


def gemm(A, B, C=None, alpha=1.0, beta=0.0, a_layout=0, b_layout=0, mask=0):
    """C = beta*C + alpha * op(A) op(B)^T on the DMMA pipe (layout 0: [rows,K]; 1: [K,rows])."""
    _check_mat(A, "A")
    _check_mat(B, "B")
    M, K = (A.shape if a_layout == 0 else (A.shape[1], A.shape[0]))
    N, K2 = (B.shape if b_layout == 0 else (B.shape[1], B.shape[0]))
    if K != K2:
        raise ValueError("inner dimensions differ")
    if C is None:
        C = torch.empty((M, N), dtype=torch.float64, device=A.device)
        beta = 0.0
    _check_mat(C, "C")
    rc = lib().gpb_gemm(_stream(), M, N, K, float(alpha), _p(A), A.stride(0), a_layout, _p(B), B.stride(0), b_layout,
                        float(beta), _p(C), C.stride(0), mask)
    _abi.check(rc, "gpb_gemm")
    return C


# ------------------------------------------------------------------------------------------
# conjugate_mll: fused forward + analytic backward
# ------------------------------------------------------------------------------------------
# ---- FP64 products on the INT8 tensor pipe (Ozaki scheme; csrc/ozaki_i8.cu) ------------------------------------------
def ozaki_available() -> bool:
    return bool(lib().gpb_ozaki_available())


OZAKI_AUTO = -1


def set_ozaki_slices(nslices: int) -> None:
    """Arithmetic of the large rank-NB trailing updates of ``lower_cholesky`` / the inverse (process-wide switch):
    ``OZAKI_AUTO`` (-1, library default): exact int8 digit-plane products (``tcgen05.mma kind::i8``) whose plane count is chosen
    on the device per call -- 7 radix-256 planes (56 bits, fp64-rounding-level) unless the hyper-parameters of a fused
    objective bound cond(Sigma) by 2e6, then 6; 4..7: that many planes everywhere; 0: FP64 DMMA everywhere.
    ``GPB_OZAKI`` in the environment sets the initial value."""
    lib().gpb_set_ozaki_slices(int(nslices))


def ozaki_auto_planes(n: int, variance: float, obs_stddev: float, jitter: float) -> int:
    """The plane count the device-side guard picks for a fused objective with these hyper-parameters (reporting only)."""
    return int(lib().gpb_ozaki_auto_planes(int(n), float(variance), float(obs_stddev), float(jitter)))


def get_ozaki_slices() -> int:
    return int(lib().gpb_get_ozaki_slices())


def ozaki_slice(X: torch.Tensor, nslices: int):
    """(Q int8 [rows, nslices*K], scale [rows]) digit planes of a row-major fp64 panel."""
    _check_mat(X, "X")
    rows, k = X.shape
    Q = torch.empty((rows, nslices * k), dtype=torch.int8, device=X.device)
    scale = torch.empty(rows, dtype=torch.float64, device=X.device)
    rc = lib().gpb_ozaki_slice(_stream(), rows, k, _p(X), X.stride(0), int(nslices), _p(Q), Q.stride(0), _p(scale))
    _abi.check(rc, "gpb_ozaki_slice")
    return Q, scale


def ozaki_gemm_(C: torch.Tensor, Qa, sa, Qb, sb, K: int, nslices: int, alpha: float = 1.0, mask_lower: bool = False):
    """C += alpha * A B^T from digit planes (in place)."""
    _check_mat(C, "C")
    rc = lib().gpb_ozaki_gemm(_stream(), C.shape[0], C.shape[1], int(K), int(nslices), _p(Qa), Qa.stride(0), _p(sa),
                              _p(Qb), Qb.stride(0), _p(sb), float(alpha), _p(C), C.stride(0), int(mask_lower))
    _abi.check(rc, "gpb_ozaki_gemm")
    return C


def igemm_i8(A: torch.Tensor, B: torch.Tensor) -> torch.Tensor:
    """int32 [m, n] = A[m, k] B[n, k]^T for int8 CUDA operands (the raw tcgen05 kind::i8 product)."""
    if A.dtype != torch.int8 or B.dtype != torch.int8 or not A.is_cuda or not B.is_cuda:
        raise TypeError("igemm_i8 takes int8 CUDA tensors")
    if A.stride(1) != 1 or B.stride(1) != 1 or A.shape[1] != B.shape[1]:
        raise ValueError("operands must be K-contiguous with equal K")
    Cc = torch.empty((A.shape[0], B.shape[0]), dtype=torch.int32, device=A.device)
    rc = lib().gpb_igemm_i8(_stream(), A.shape[0], B.shape[0], A.shape[1], _p(A), A.stride(0), _p(B), B.stride(0), _p(Cc),
                            Cc.stride(0))
    _abi.check(rc, "gpb_igemm_i8")
    return Cc


class _MllState:
    """Per-(device, N, D) buffers reused across iterations: the N x N Sigma/L/Sigma^-1 buffer and
    the factorisation workspace.  `generation` guards against backward being called on a stale
    forward (then the forward is simply replayed)."""

    def __init__(self, n, d, device):
        self.sigma = torch.empty((n, n), dtype=torch.float64, device=device)
        self.nbytes = lib().gpb_mll_workspace_bytes(n, d)
        self.ws = torch.empty(max(self.nbytes // 8, 1), dtype=torch.float64, device=device)
        self.generation = 0


_MLL_CACHE: dict = {}


def _mll_state(n, d, device) -> _MllState:
    key = (device.index if device.index is not None else torch.cuda.current_device(), n, d)
    st = _MLL_CACHE.get(key)
    if st is None:
        _MLL_CACHE.clear()  # one live problem size at a time: the buffer is O(N^2)
        st = _MllState(n, d, device)
        _MLL_CACHE[key] = st
    return st


def release_buffers() -> None:
    _MLL_CACHE.clear()


def _mll_forward_raw(st, kind, X, y, ell_v, iso, var, sn, mean, jitter):
    N, D = X.shape
    val = torch.empty(1, dtype=torch.float64, device=X.device)
    alpha = torch.empty(N, dtype=torch.float64, device=X.device)
    info = torch.zeros(1, dtype=torch.int32, device=X.device)
    rc = lib().gpb_mll_forward(_stream(), kind, N, D, _p(X), X.stride(0), _p(y), _p(ell_v), iso, _p(var), _p(sn),
                               _p(mean), float(jitter), _p(st.sigma), st.sigma.stride(0), _p(st.ws), st.nbytes,
                               _p(val), _p(alpha), _p(info))
    _abi.check(rc, "gpb_mll_forward")
    st.generation = next_generation()
    return val, alpha, info


class ConjugateMllFunction(torch.autograd.Function):
    """log N(y | m, K + (jitter + s^2) I) with the analytic gradient
    W = 1/2 (alpha alpha^T - Sigma^-1) : dK/dtheta  (custom_vjp analogue)."""

    @staticmethod
    def forward(ctx, kind, X, y, ell, variance, obs_stddev, mean_const, jitter):
        _check_mat(X, "X")
        require_cuda(y)
        N, D = X.shape
        y = y.reshape(-1).contiguous()
        if y.numel() != N:
            raise ValueError("conjugate_mll supports a single output column (y of shape [N, 1])")
        ell_v, iso = _ell_args(ell, D)
        var = _kscalars(kind, variance)
        sn = _scalar(obs_stddev, "obs_stddev")
        mean = None if mean_const is None else _scalar(mean_const, "mean constant")
        st = _mll_state(N, D, X.device)
        val, alpha, info = _mll_forward_raw(st, kind, X, y, ell_v, iso, var, sn, mean, jitter)
        ctx.kind, ctx.jitter, ctx.iso = kind, jitter, iso
        ctx.gen = st.generation
        ctx.has_mean = mean is not None
        ctx.shapes = (ell.shape, variance.shape, obs_stddev.shape, None if mean_const is None else mean_const.shape)
        ctx.save_for_backward(X, y, ell_v, var, sn, mean if mean is not None else var, alpha)
        return val.reshape(())

    @staticmethod
    def backward(ctx, gout):
        _no_data_grad(ctx, (1, 2), "conjugate_mll")
        X, y, ell_v, var, sn, mean, alpha = ctx.saved_tensors
        N, D = X.shape
        st = _mll_state(N, D, X.device)
        if st.generation != ctx.gen:  # buffer was reused by another forward: replay this one
            _, alpha, _ = _mll_forward_raw(st, ctx.kind, X, y, ell_v, ctx.iso, var, sn,
                                           mean if ctx.has_mean else None, ctx.jitter)
        dev = X.device
        g_ell = torch.empty(1 if ctx.iso else D, dtype=torch.float64, device=dev)
        g_var = torch.empty(var.numel(), dtype=torch.float64, device=dev)
        g_sn = torch.empty(1, dtype=torch.float64, device=dev)
        g_mean = torch.empty(1, dtype=torch.float64, device=dev) if ctx.has_mean else None
        g = gout.reshape(1).contiguous()
        rc = lib().gpb_mll_backward(_stream(), ctx.kind, N, D, _p(X), X.stride(0), _p(ell_v), ctx.iso, _p(var),
                                    _p(sn), _p(st.sigma), st.sigma.stride(0), _p(st.ws), st.nbytes, _p(alpha), _p(g),
                                    _p(g_ell), _p(g_var), _p(g_sn), _p(g_mean))
        _abi.check(rc, "gpb_mll_backward")
        s_ell, s_var, s_sn, s_mean = ctx.shapes
        return (None, None, None, g_ell.reshape(s_ell), g_var.reshape(s_var), g_sn.reshape(s_sn),
                g_mean.reshape(s_mean) if ctx.has_mean else None, None)


def conjugate_mll_fused(kind, X, y, ell, variance, obs_stddev, mean_const=None, jitter=1e-6):
    return ConjugateMllFunction.apply(kind, X, y, ell, variance, obs_stddev, mean_const, jitter)


LOG_2PI = math.log(2.0 * math.pi)


# ------------------------------------------------------------------------------------------
# Dense-covariance objectives: the composable route for kernels without a fused objective
# (sum / product kernels).  Sigma arrives as a tensor built by differentiable Gram launches; the
# factorisation, solves and the N^3 part of the backward run on the same CUDA path as the fused one.
# ------------------------------------------------------------------------------------------
def _factor_for_backward(Sigma: torch.Tensor):
    n = Sigma.shape[0]
    L = Sigma.detach().clone().contiguous()
    ws = FactorWorkspace(n, 1, potri=True, device=Sigma.device)
    info = potrf_lower_(L, ws, zero_upper=False)
    return L, ws, info


class GaussianLogProbFunction(torch.autograd.Function):
    """log N(y; mu, Sigma) as a function of (Sigma, diff = y - mu)  -- gpjax/distributions.py:115-134.
    Backward: dSigma = 1/2 (alpha alpha^T - Sigma^-1), ddiff = -alpha, with Sigma^-1 from TRTRI + LAUUM."""

    @staticmethod
    def forward(ctx, Sigma, diff):
        _check_mat(Sigma, "Sigma")
        n = Sigma.shape[0]
        d = diff.detach().reshape(-1).contiguous()
        if d.numel() != n or Sigma.shape[1] != n:
            raise ValueError("Sigma must be [n, n] and diff [n]")
        L, ws, info = _factor_for_backward(Sigma)
        w = trsv_lower_(L, d.clone(), ws, trans=False)
        quad = gemm(w.reshape(1, -1), w.reshape(1, -1)).reshape(())
        alpha = trsv_lower_(L, w, ws, trans=True)  # w is overwritten: alpha = Sigma^-1 diff
        val = -0.5 * (n * LOG_2PI + 2.0 * sum_log_diag(L) + quad)
        ctx.state = (L, ws)
        ctx.diff_shape = diff.shape
        ctx.save_for_backward(alpha)
        return val

    @staticmethod
    def backward(ctx, gout):
        (alpha,) = ctx.saved_tensors
        L, ws = ctx.state
        g_sigma = g_diff = None
        if ctx.needs_input_grad[0]:
            P = potri_lower(L, ws)
            a = alpha.reshape(-1, 1)
            gemm(a, a, C=P, alpha=0.5, beta=-0.5)  # P <- 1/2 (alpha alpha^T - Sigma^-1)
            g_sigma = P.mul_(gout)
        if ctx.needs_input_grad[1]:
            g_diff = (-gout * alpha).reshape(ctx.diff_shape)
        return g_sigma, g_diff


class LoocvFunction(torch.autograd.Function):
    """Leave-one-out log predictive probability (gpjax/objectives.py:161-178) as a function of (Sigma, diff):
    with P = Sigma^-1, alpha = P diff:  sum_i [ -1/2 log 2pi + 1/2 log P_ii - 1/2 alpha_i^2 / P_ii ].
    Backward: dSigma = -P R P + sym((P s) alpha^T), R = diag(1/(2 P_ii) + alpha_i^2/(2 P_ii^2)), s = alpha / diag(P);
    ddiff = -P s.  The N^3 term P R P = (P R^1/2)(P R^1/2)^T is one DMMA product."""

    @staticmethod
    def forward(ctx, Sigma, diff):
        _check_mat(Sigma, "Sigma")
        n = Sigma.shape[0]
        d = diff.detach().reshape(-1).contiguous()
        if d.numel() != n or Sigma.shape[1] != n:
            raise ValueError("Sigma must be [n, n] and diff [n]")
        L, ws, info = _factor_for_backward(Sigma)
        alpha = trsv_lower_(L, trsv_lower_(L, d.clone(), ws, trans=False), ws, trans=True)
        P = potri_lower(L, ws)
        pd = torch.diagonal(P)
        val = torch.sum(-0.5 * LOG_2PI + 0.5 * torch.log(pd) - 0.5 * alpha * alpha / pd)
        ctx.diff_shape = diff.shape
        ctx.save_for_backward(P, alpha)
        return val

    @staticmethod
    def backward(ctx, gout):
        P, alpha = ctx.saved_tensors
        pd = torch.diagonal(P)
        s = (alpha / pd).contiguous()
        Ps = gemm(P, s.reshape(1, -1)).reshape(-1)  # P s (P symmetric)
        g_sigma = g_diff = None
        if ctx.needs_input_grad[0]:
            r = 0.5 / pd + 0.5 * alpha * alpha / (pd * pd)
            A = P * torch.sqrt(r)[None, :]
            G = gemm(A, A, alpha=-1.0)  # -P R P
            u, a = Ps.reshape(-1, 1), alpha.reshape(-1, 1)
            gemm(u, a, C=G, alpha=0.5, beta=1.0)
            gemm(a, u, C=G, alpha=0.5, beta=1.0)
            g_sigma = G.mul_(gout)
        if ctx.needs_input_grad[1]:
            g_diff = (-gout * Ps).reshape(ctx.diff_shape)
        return g_sigma, g_diff


# ------------------------------------------------------------------------------------------
# Reverse mode for the free linalg functions (gpjax/linalg/operations.py) -- what jax.grad derives through
# jnp.linalg.cholesky / jsp.linalg.solve_triangular.  All N^3 work stays on the DMMA GEMM + blocked solves.
# ------------------------------------------------------------------------------------------
class CholeskyFunction(torch.autograd.Function):
    """L = chol(A) (lower, strict upper zero).  Backward (symmetrised, as JAX's cholesky JVP/VJP):
    Abar = sym( L^-T Phi(L^T Lbar) L^-1 ),  Phi = lower triangle with the diagonal halved."""

    @staticmethod
    def forward(ctx, A):
        _check_mat(A, "A")
        n = A.shape[0]
        L = A.detach().clone().contiguous()
        ws = FactorWorkspace(n, 1, potri=False, device=A.device)
        potrf_lower_(L, ws, zero_upper=True, symmetrize=True)
        ctx.ws = ws
        ctx.save_for_backward(L)
        return L

    @staticmethod
    def backward(ctx, Lbar):
        (L,) = ctx.saved_tensors
        ws = ctx.ws
        P = gemm(L, torch.tril(Lbar).contiguous(), a_layout=1, b_layout=1)  # L^T Lbar
        P = torch.tril(P)
        torch.diagonal(P).mul_(0.5)
        Y = trsm_lower_left_(L, P.contiguous(), ws, trans=True)             # L^-T Phi
        St = trsm_lower_left_(L, Y.T.contiguous(), ws, trans=True)          # (Y L^-1)^T
        return 0.5 * (St + St.T)


class GemmFunction(torch.autograd.Function):
    """C = op(A) op(B)^T on the DMMA pipe, differentiable (layout 0: the tensor is [rows, K]; 1: it is [K, rows]) -- the
    jnp.matmul of the composable objectives.  Backward: d op(A) = G op(B), d op(B) = G^T op(A), two more launches."""

    @staticmethod
    def forward(ctx, A, B, a_layout, b_layout):
        ctx.layouts = (int(a_layout), int(b_layout))
        A, B = A.detach().contiguous(), B.detach().contiguous()
        ctx.save_for_backward(A, B)
        return gemm(A, B, a_layout=a_layout, b_layout=b_layout)

    @staticmethod
    def backward(ctx, G):
        A, B = ctx.saved_tensors
        al, bl = ctx.layouts
        G = G.contiguous()
        gA = gB = None
        if ctx.needs_input_grad[0]:
            gA = gemm(G, B, a_layout=0, b_layout=1 - bl) if al == 0 else gemm(B, G, a_layout=1 - bl, b_layout=0)
        if ctx.needs_input_grad[1]:
            gB = gemm(G, A, a_layout=1, b_layout=1 - al) if bl == 0 else gemm(A, G, a_layout=1 - al, b_layout=1)
        return gA, gB, None, None


def matmul_nt(A: torch.Tensor, B: torch.Tensor, a_layout: int = 0, b_layout: int = 0) -> torch.Tensor:
    """op(A) op(B)^T through GemmFunction when either operand requires grad, else one plain launch."""
    if torch.is_grad_enabled() and (A.requires_grad or B.requires_grad):
        return GemmFunction.apply(A, B, a_layout, b_layout)
    return gemm(A.contiguous(), B.contiguous(), a_layout=a_layout, b_layout=b_layout)


class TriangularSolveFunction(torch.autograd.Function):
    """X = op(L)^-1 B for a lower-triangular L (op = transpose when `trans`).  Backward:
    Bbar = op(L)^-T Xbar;  Lbar = -tril(Bbar X^T)  (or -tril(X Bbar^T) when `trans`)."""

    @staticmethod
    def forward(ctx, L, B, trans, ws):
        _check_mat(L, "L")
        vec = B.dim() == 1
        X = B.detach().reshape(L.shape[0], -1).clone().contiguous()
        if vec:
            X = trsv_lower_(L.detach(), X.reshape(-1), ws, trans=trans).reshape(-1, 1)
        else:
            X = trsm_lower_left_(L.detach(), X, ws, trans=trans)
        ctx.trans, ctx.ws, ctx.vec = bool(trans), ws, vec
        ctx.save_for_backward(L.detach(), X)
        return X.reshape(-1) if vec else X

    @staticmethod
    def backward(ctx, Xbar):
        L, X = ctx.saved_tensors
        G = Xbar.detach().reshape(L.shape[0], -1).clone().contiguous()
        if ctx.vec:
            G = trsv_lower_(L, G.reshape(-1), ctx.ws, trans=not ctx.trans).reshape(-1, 1)
        else:
            G = trsm_lower_left_(L, G, ctx.ws, trans=not ctx.trans)
        g_L = None
        if ctx.needs_input_grad[0]:
            outer = gemm(X, G) if ctx.trans else gemm(G, X)   # X Bbar^T  or  Bbar X^T
            g_L = -torch.tril(outer)
        g_B = (G.reshape(-1) if ctx.vec else G) if ctx.needs_input_grad[1] else None
        return g_L, g_B, None, None
