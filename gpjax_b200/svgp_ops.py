"""Uncollapsed minibatch ELBO (SVGP) on the C ABI.  The batch enters only through the same row-additive
statistics as the collapsed bound, so the streamed passes and both all-reduces are the SGPR entry points;
``gpb_svgp_finish`` / ``gpb_svgp_grad_finish`` add the replicated M x M part (KL, variational-parameter
gradients).  Data-parallel: every rank passes ITS minibatch; the effective batch is the sum over ranks."""
from __future__ import annotations

import torch

from . import _abi
from ._lib import lib
from .ops import _check_mat, _ell_args, _kscalars, _no_data_grad, _p, _scalar, _stream, next_generation, require_cuda
from .sgpr_ops import DEFAULT_BLOCK_ROWS, _all_reduce, _state, _stats, _use_raw_statistics, finish_flags


def _forward_raw(st, kind, X, y, Z, ell_v, iso, var, sn, mean, mu, W, ndata, jitter, block_rows, group, need_grad,
                 raw=False):
    n_loc, D = X.shape
    M = Z.shape[0]
    L = lib()
    P = torch.empty(L.gpb_sgpr_stats_count(M), dtype=torch.float64, device=Z.device)
    fn, name = _stats(L, raw)
    rc = fn(_stream(), kind, n_loc, M, D, _p(X), X.stride(0) if n_loc else D, _p(y), _p(Z), Z.stride(0),
            _p(ell_v), iso, _p(var), _p(sn), _p(mean), float(jitter), block_rows, _p(st.ws), st.nbytes, _p(P))
    _abi.check(rc, name)
    _all_reduce(P, group)
    val = torch.empty(1, dtype=torch.float64, device=Z.device)
    info = torch.zeros(2, dtype=torch.int32, device=Z.device)
    rc = L.gpb_svgp_finish(_stream(), kind, M, D, _p(Z), Z.stride(0), _p(ell_v), iso, _p(var), _p(sn), _p(mean), _p(mu),
                           _p(W), W.stride(0), float(ndata), float(jitter), block_rows, _p(st.ws), st.nbytes, _p(P),
                           finish_flags(need_grad, raw), _p(val), _p(info))
    _abi.check(rc, "gpb_svgp_finish")
    st.generation = next_generation()
    return val


class SvgpElboFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, kind, X, y, Z, ell, variance, obs_stddev, mean_const, var_mean, var_sqrt, num_datapoints, jitter,
                block_rows, group, statistics="auto"):
        _check_mat(X, "X")
        _check_mat(Z, "Z")
        _check_mat(var_sqrt, "variational_root_covariance")
        require_cuda(y, var_mean)
        n_loc, D = X.shape
        M = Z.shape[0]
        y = y.reshape(-1).contiguous()
        if y.numel() != n_loc:
            raise ValueError("elbo supports a single output column (y of shape [B, 1])")
        Z = Z.contiguous()
        mu = var_mean.reshape(-1).contiguous()
        if mu.numel() != M or tuple(var_sqrt.shape) != (M, M):
            raise ValueError("variational parameters do not match the number of inducing points")
        W = var_sqrt.contiguous()
        ell_v, iso = _ell_args(ell, D)
        var = _kscalars(kind, variance)
        sn = _scalar(obs_stddev, "obs_stddev")
        mean = None if mean_const is None else _scalar(mean_const, "mean constant")
        block_rows = int(min(block_rows, max(n_loc, 1)))
        st = _state(M, D, block_rows, Z.device)
        need_grad = any(ctx.needs_input_grad)
        raw = _use_raw_statistics(statistics, kind, Z, ell_v, var, jitter)
        val = _forward_raw(st, kind, X, y, Z, ell_v, iso, var, sn, mean, mu, W, num_datapoints, jitter, block_rows, group,
                           need_grad, raw)
        ctx.cfg = (kind, iso, jitter, block_rows, group, mean is not None, float(num_datapoints), raw)
        ctx.gen = st.generation
        ctx.shapes = (ell.shape, variance.shape, obs_stddev.shape, None if mean_const is None else mean_const.shape,
                      var_mean.shape)
        ctx.save_for_backward(X, y, Z, ell_v, var, sn, mean if mean is not None else var, mu, W)
        return val.reshape(())

    @staticmethod
    def backward(ctx, gout):
        _no_data_grad(ctx, (1, 2), "elbo")
        X, y, Z, ell_v, var, sn, mean, mu, W = ctx.saved_tensors
        kind, iso, jitter, block_rows, group, has_mean, ndata, raw = ctx.cfg
        n_loc, D = X.shape
        M = Z.shape[0]
        st = _state(M, D, block_rows, Z.device)
        if st.generation != ctx.gen:
            _forward_raw(st, kind, X, y, Z, ell_v, iso, var, sn, mean if has_mean else None, mu, W, ndata, jitter,
                         block_rows, group, True, raw)
        L = lib()
        nl = 1 if iso else D
        flat = torch.empty(M * D + nl + var.numel(), dtype=torch.float64, device=Z.device)
        g_Z, g_ell, g_var = flat[: M * D], flat[M * D: M * D + nl], flat[M * D + nl:]
        rc = L.gpb_sgpr_grad_local(_stream(), kind, n_loc, M, D, _p(X), X.stride(0) if n_loc else D, _p(y), _p(Z),
                                   Z.stride(0), _p(ell_v), iso, _p(var), _p(sn), _p(mean if has_mean else None),
                                   block_rows, _p(st.ws), st.nbytes, _p(g_Z), _p(g_ell), _p(g_var))
        _abi.check(rc, "gpb_sgpr_grad_local")
        _all_reduce(flat, group)
        dev = Z.device
        g_sn = torch.empty(1, dtype=torch.float64, device=dev)
        g_mean = torch.empty(1, dtype=torch.float64, device=dev)
        g_mu = torch.empty(M, dtype=torch.float64, device=dev)
        g_W = torch.empty((M, M), dtype=torch.float64, device=dev)
        g = gout.reshape(1).contiguous()
        rc = L.gpb_svgp_grad_finish(_stream(), kind, M, D, _p(Z), Z.stride(0), _p(ell_v), iso, _p(var), _p(sn),
                                    float(jitter), block_rows, _p(st.ws), st.nbytes, _p(g), _p(W), W.stride(0), _p(g_Z),
                                    _p(g_ell), _p(g_var), _p(g_sn), _p(g_mean), _p(g_mu), _p(g_W), M)
        _abi.check(rc, "gpb_svgp_grad_finish")
        s_ell, s_var, s_sn, s_mean, s_mu = ctx.shapes
        return (None, None, None, g_Z.reshape(M, D), g_ell.reshape(s_ell), g_var.reshape(s_var), g_sn.reshape(s_sn),
                g_mean.reshape(s_mean) if has_mean else None, g_mu.reshape(s_mu), g_W, None, None, None, None, None)


def svgp_elbo_fused(kind, X, y, Z, ell, variance, obs_stddev, mean_const, var_mean, var_sqrt, num_datapoints,
                    jitter=1e-6, block_rows: int = DEFAULT_BLOCK_ROWS, group=None, statistics: str = "auto"):
    """`statistics`: see sgpr_ops.collapsed_elbo_fused."""
    return SvgpElboFunction.apply(kind, X, y, Z, ell, variance, obs_stddev, mean_const, var_mean, var_sqrt,
                                  num_datapoints, jitter, block_rows, group, statistics)
