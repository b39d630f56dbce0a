"""collapsed_elbo (SGPR) on the C ABI: two streamed passes over the local rows, one all-reduce each.

Row sharding: every rank holds a contiguous shard of (X, y); Z and the hyper-parameters
are replicated.  The only exchange steps are
  * forward : all-reduce(sum) of the (M+2)^2 augmented statistics  (33.6 MB at M = 2048),
  * backward: all-reduce(sum) of [g_Z (M x D), g_lengthscale, g_variance].
``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the CPU tests of the host logic)
is only the plumbing for those two calls; everything else is the hand-written CUDA path.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _abi
from ._lib import lib
from .ops import _check_mat, _ell_args, _kscalars, _no_data_grad, _p, _scalar, _stream, next_generation, require_cuda

DEFAULT_BLOCK_ROWS = 32768
# kernel kinds the streamed sparse path evaluates (csrc/sgpr.cpp check_args: every fused stationary kernel except PoweredExponential,
# whose clamped diagonal k(x, x) != variance)
FUSED_SPARSE_KINDS = frozenset({0, 1, 2, 3, 4, 6, 7})


def _world(group) -> int:
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group)
    return 1


# ---- exchange step: NCCL through the C ABI (gpb_allreduce_f64) when a native communicator exists, else torch.distributed ----
_NATIVE_COMMS: dict = {}  # id(group) or None -> ncclComm_t (int)


def init_native_collective(group=None) -> bool:
    """Create an NCCL communicator for `group` through the C ABI (``gpb_nccl_unique_id`` / ``gpb_nccl_comm_init_rank``) so the two
    all-reduces of the sharded path are issued by ``gpb_allreduce_f64`` on the launching stream -- the same call a non-torch
    binding makes.  ``torch.distributed`` only ships the 128-byte unique id (rendezvous).  Returns False (and changes
    nothing) when the process has no NCCL or a single rank."""
    import ctypes

    import torch.distributed as dist

    L = lib()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) < 2 or L.gpb_nccl_version() == 0:
        return False
    key = None if group is None else id(group)
    if key in _NATIVE_COMMS:
        return True
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    idbuf = (ctypes.c_char * 128)()
    if rank == 0:
        _abi.check(L.gpb_nccl_unique_id(idbuf), "gpb_nccl_unique_id")
    box = [bytes(idbuf.raw) if rank == 0 else None]
    src = 0 if group is None else dist.get_global_rank(group, 0)
    dist.broadcast_object_list(box, src=src, group=group)
    comm = ctypes.c_void_p()
    _abi.check(L.gpb_nccl_comm_init_rank(ctypes.byref(comm), world, box[0], rank), "gpb_nccl_comm_init_rank")
    _NATIVE_COMMS[key] = comm.value
    return True


def destroy_native_collectives() -> None:
    for comm in _NATIVE_COMMS.values():
        lib().gpb_nccl_comm_destroy(comm)
    _NATIVE_COMMS.clear()


def _all_reduce(t: torch.Tensor, group) -> None:
    import torch.distributed as dist

    if _world(group) <= 1:
        return
    comm = _NATIVE_COMMS.get(None if group is None else id(group))
    if comm is not None and t.is_cuda and t.dtype == torch.float64 and t.is_contiguous():
        _abi.check(lib().gpb_allreduce_f64(comm, _stream(), t.data_ptr(), t.numel()), "gpb_allreduce_f64")
    else:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)


class _SgprState:
    def __init__(self, m, d, block_rows, device):
        self.nbytes = lib().gpb_sgpr_workspace_bytes(m, d, block_rows)
        self.ws = torch.empty(max(self.nbytes // 8, 1), dtype=torch.float64, device=device)
        self.generation = 0


_CACHE: dict = {}


def _state(m, d, block_rows, device) -> _SgprState:
    key = (device.index if device.index is not None else torch.cuda.current_device(), m, d, block_rows)
    st = _CACHE.get(key)
    if st is None:
        _CACHE.clear()
        st = _SgprState(m, d, block_rows, device)
        _CACHE[key] = st
    return st


def release_buffers() -> None:
    _CACHE.clear()
    _ROUTE_CACHE.clear()


# ---- which route to the statistics (csrc/sgpr.cpp: whiten-first vs raw products + one whitening at the end) ----------
RAW_STATISTICS_COND_LIMIT = 1e3  # "auto": raw route only while cond(Kzz + jitter I) is estimated below this
RAW_STATISTICS_RECHECK = 10      # "auto": a new estimate is launched every this many evaluations
_ROUTE_CACHE: dict = {}          # (device, kind, M, D, jitter) -> _RouteSlot


class _RouteSlot:
    """Decision state of statistics="auto" for one problem shape.  The estimate is computed on the DEVICE and travels to a pinned
    host word with a non-blocking copy; a forward only ever *polls* the copy's event, so the training loop never synchronises
    with the host.  Until the first estimate has landed the reference's order ("whitened") is used."""

    def __init__(self):
        self.decision = False   # raw route allowed?
        self.countdown = 0      # evaluations until the next estimate is launched
        self.pending = None     # (event, pinned host tensor) of an estimate in flight
        self.estimates = 0      # completed estimates (tests / reporting)
        self.last = None        # last completed estimate (host float)

    def harvest(self) -> None:
        if self.pending is not None and self.pending[0].query():
            est = float(self.pending[1][0])
            self.last = est if est == est else float("inf")  # NaN (Kzz not positive definite) -> never the raw route
            self.decision = self.last <= RAW_STATISTICS_COND_LIMIT
            self.estimates += 1
            self.pending = None


def kzz_condition_estimate_device(kind, Z, ell, var, jitter, iters: int = 8) -> torch.Tensor:
    """Estimate of cond_2(Kzz + jitter I) as a DEVICE scalar: lambda_max by power iteration on the matrix, lambda_min by
    inverse iteration through its Cholesky factor (a few GEMV / TRSV launches on the M x M matrix; no host read)."""
    from . import ops

    M = Z.shape[0]
    K = ops.gram_forward(kind, Z, Z, ell, var, diag_add=float(jitter))
    Lf = K.clone()
    ws = ops.FactorWorkspace(M, 1, device=Z.device)
    ops.potrf_lower_(Lf, ws, zero_upper=False)
    g = torch.Generator(device=Z.device)
    g.manual_seed(0)
    v = torch.randn(M, dtype=torch.float64, device=Z.device, generator=g)
    u = v.clone()
    for _ in range(iters):
        v = ops.gemm(K, (v / v.norm()).reshape(1, -1).contiguous()).reshape(-1)                    # K v
        u = u / u.norm()
        u = ops.trsv_lower_(Lf, ops.trsv_lower_(Lf, u.contiguous(), ws, trans=False), ws, trans=True)  # K^-1 u
    return (v.norm() * u.norm()).reshape(1)


def kzz_condition_estimate(kind, Z, ell, var, jitter, iters: int = 8) -> float:
    """Host value of :func:`kzz_condition_estimate_device` (one host read: tests and bench reporting only)."""
    est = kzz_condition_estimate_device(kind, Z, ell, var, jitter, iters).item()
    return est if est == est else float("inf")


def _use_raw_statistics(mode: str, kind, Z, ell_v, var, jitter) -> bool:
    if mode == "whitened":
        return False
    if mode == "raw":
        return True
    if mode != "auto":
        raise ValueError("statistics must be 'auto', 'whitened' or 'raw'")
    if torch.cuda.is_current_stream_capturing():
        # fit(cuda_graph=True): a captured step cannot poll an event or re-estimate cond(Kzz) as training moves Z, and a decision
        # frozen at capture time would outlive its guard -- a replayed step takes the reference's order unless told otherwise
        return False
    key = (Z.device.index, kind, Z.shape[0], Z.shape[1], float(jitter))
    slot = _ROUTE_CACHE.get(key)
    if slot is None:
        slot = _ROUTE_CACHE[key] = _RouteSlot()
    slot.harvest()
    if slot.pending is None and slot.countdown <= 0:
        # hyper-parameters move slowly between optimiser steps and the limit leaves two orders of magnitude of margin to the
        # stated tolerance, so a decision that lags the parameters by a few evaluations is safe
        with torch.no_grad():
            est = kzz_condition_estimate_device(kind, Z.detach(), ell_v.detach(), var.detach(), jitter)
            host = torch.empty(1, dtype=torch.float64, pin_memory=True)
            host.copy_(est, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
        slot.pending = (ev, host)
        slot.countdown = RAW_STATISTICS_RECHECK
    slot.countdown -= 1
    return slot.decision


def route_state(kind, Z, jitter):
    """The `_RouteSlot` of a problem shape (tests / bench reporting), or None."""
    return _ROUTE_CACHE.get((Z.device.index, kind, Z.shape[0], Z.shape[1], float(jitter)))


def _stats(L, raw: bool):
    return (L.gpb_sgpr_stats_raw, "gpb_sgpr_stats_raw") if raw else (L.gpb_sgpr_stats, "gpb_sgpr_stats")


FINISH_DENSE_INT8 = 2  # include/gpjax_b200.h: GPB_FINISH_DENSE_INT8


def finish_flags(need_grad, raw) -> int:
    """Flag word of gpb_sgpr_finish / gpb_svgp_finish: bit 0 = prepare the gradient pass, bit 1 = Kzz is well conditioned (the
    raw-statistics route was chosen: estimated cond <= RAW_STATISTICS_COND_LIMIT, or the caller asked for it), so the dense M^3
    products of the replicated finish may run as int8 digit-plane products."""
    return (1 if need_grad else 0) | (FINISH_DENSE_INT8 if raw else 0)


def _forward_raw(st, kind, X, y, Z, ell_v, iso, var, sn, mean, jitter, block_rows, group, need_grad, raw=False):
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
    rc = L.gpb_sgpr_finish(_stream(), kind, M, D, _p(Z), Z.stride(0), _p(ell_v), iso, _p(var), _p(sn), block_rows,
                           _p(st.ws), st.nbytes, _p(P), finish_flags(need_grad, raw), _p(val), _p(info))
    _abi.check(rc, "gpb_sgpr_finish")
    st.generation = next_generation()
    return val, info


class CollapsedElboFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, kind, X, y, Z, ell, variance, obs_stddev, mean_const, jitter, block_rows, group, statistics="auto"):
        _check_mat(X, "X")
        _check_mat(Z, "Z")
        require_cuda(y)
        n_loc, D = X.shape
        M = Z.shape[0]
        y = y.reshape(-1).contiguous()
        if y.numel() != n_loc:
            raise ValueError("collapsed_elbo supports a single output column (y of shape [N, 1])")
        Z = Z.contiguous()
        ell_v, iso = _ell_args(ell, D)
        var = _kscalars(kind, variance)
        sn = _scalar(obs_stddev, "obs_stddev")
        mean = None if mean_const is None else _scalar(mean_const, "mean constant")
        block_rows = int(min(block_rows, max(n_loc, 1)))
        st = _state(M, D, block_rows, Z.device)
        need_grad = any(ctx.needs_input_grad)
        raw = _use_raw_statistics(statistics, kind, Z, ell_v, var, jitter)
        val, info = _forward_raw(st, kind, X, y, Z, ell_v, iso, var, sn, mean, jitter, block_rows, group, need_grad, raw)
        ctx.cfg = (kind, iso, jitter, block_rows, group, mean is not None, raw)
        ctx.gen = st.generation
        ctx.shapes = (ell.shape, variance.shape, obs_stddev.shape, None if mean_const is None else mean_const.shape)
        ctx.save_for_backward(X, y, Z, ell_v, var, sn, mean if mean is not None else var)
        return val.reshape(())

    @staticmethod
    def backward(ctx, gout):
        _no_data_grad(ctx, (1, 2), "collapsed_elbo")
        X, y, Z, ell_v, var, sn, mean = ctx.saved_tensors
        kind, iso, jitter, block_rows, group, has_mean, raw = ctx.cfg
        n_loc, D = X.shape
        M = Z.shape[0]
        st = _state(M, D, block_rows, Z.device)
        if st.generation != ctx.gen:
            _forward_raw(st, kind, X, y, Z, ell_v, iso, var, sn, mean if has_mean else None, jitter, block_rows, group,
                         True, raw)
        L = lib()
        nl = 1 if iso else D
        flat = torch.empty(M * D + nl + var.numel(), dtype=torch.float64, device=Z.device)
        g_Z, g_ell, g_var = flat[: M * D], flat[M * D: M * D + nl], flat[M * D + nl:]
        rc = L.gpb_sgpr_grad_local(_stream(), kind, n_loc, M, D, _p(X), X.stride(0) if n_loc else D, _p(y), _p(Z),
                                   Z.stride(0), _p(ell_v), iso, _p(var), _p(sn), _p(mean if has_mean else None),
                                   block_rows, _p(st.ws), st.nbytes, _p(g_Z), _p(g_ell), _p(g_var))
        _abi.check(rc, "gpb_sgpr_grad_local")
        _all_reduce(flat, group)
        g_sn = torch.empty(1, dtype=torch.float64, device=Z.device)
        g_mean = torch.empty(1, dtype=torch.float64, device=Z.device)
        g = gout.reshape(1).contiguous()
        rc = L.gpb_sgpr_grad_finish(_stream(), kind, M, D, _p(Z), Z.stride(0), _p(ell_v), iso, _p(var), _p(sn),
                                    block_rows, _p(st.ws), st.nbytes, _p(g), _p(g_Z), _p(g_ell), _p(g_var), _p(g_sn),
                                    _p(g_mean))
        _abi.check(rc, "gpb_sgpr_grad_finish")
        s_ell, s_var, s_sn, s_mean = ctx.shapes
        return (None, None, None, g_Z.reshape(M, D), g_ell.reshape(s_ell), g_var.reshape(s_var), g_sn.reshape(s_sn),
                g_mean.reshape(s_mean) if has_mean else None, None, None, None, None)


def profile_phases(kind, X, y, Z, ell, variance, obs_stddev, mean_const=None, jitter=1e-6,
                   block_rows: int = DEFAULT_BLOCK_ROWS, group=None, raw: bool = True) -> dict:
    """One value+gradient evaluation with CUDA events between the six protocol steps (measurement only: bench.py reports where
    the time of a sharded evaluation goes).  Returns milliseconds per phase on this rank and the ELBO."""
    L = lib()
    n_loc, D = X.shape
    M = Z.shape[0]
    y = y.reshape(-1).contiguous()
    Z = Z.detach().contiguous()
    ell_v, iso = _ell_args(ell.detach(), D)
    var = _kscalars(kind, variance.detach())
    sn = _scalar(obs_stddev.detach(), "obs_stddev")
    mean = None if mean_const is None else _scalar(mean_const.detach(), "mean constant")
    block_rows = int(min(block_rows, max(n_loc, 1)))
    st = _state(M, D, block_rows, Z.device)
    st.generation = next_generation()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(7)]
    P = torch.empty(L.gpb_sgpr_stats_count(M), dtype=torch.float64, device=Z.device)
    val = torch.empty(1, dtype=torch.float64, device=Z.device)
    info = torch.zeros(2, dtype=torch.int32, device=Z.device)
    nl = 1 if iso else D
    flat = torch.empty(M * D + nl + var.numel(), dtype=torch.float64, device=Z.device)
    g_Z, g_ell, g_var = flat[: M * D], flat[M * D: M * D + nl], flat[M * D + nl:]
    g_sn = torch.empty(1, dtype=torch.float64, device=Z.device)
    g_mean = torch.empty(1, dtype=torch.float64, device=Z.device)
    fn, name = _stats(L, raw)
    ev[0].record()
    _abi.check(fn(_stream(), kind, n_loc, M, D, _p(X), X.stride(0) if n_loc else D, _p(y), _p(Z), Z.stride(0), _p(ell_v), iso,
                  _p(var), _p(sn), _p(mean), float(jitter), block_rows, _p(st.ws), st.nbytes, _p(P)), name)
    ev[1].record()
    _all_reduce(P, group)
    ev[2].record()
    _abi.check(L.gpb_sgpr_finish(_stream(), kind, M, D, _p(Z), Z.stride(0), _p(ell_v), iso, _p(var), _p(sn), block_rows,
                                 _p(st.ws), st.nbytes, _p(P), finish_flags(True, raw), _p(val), _p(info)), "gpb_sgpr_finish")
    ev[3].record()
    _abi.check(L.gpb_sgpr_grad_local(_stream(), kind, n_loc, M, D, _p(X), X.stride(0) if n_loc else D, _p(y), _p(Z),
                                     Z.stride(0), _p(ell_v), iso, _p(var), _p(sn), _p(mean), block_rows, _p(st.ws), st.nbytes,
                                     _p(g_Z), _p(g_ell), _p(g_var)), "gpb_sgpr_grad_local")
    ev[4].record()
    _all_reduce(flat, group)
    ev[5].record()
    _abi.check(L.gpb_sgpr_grad_finish(_stream(), kind, M, D, _p(Z), Z.stride(0), _p(ell_v), iso, _p(var), _p(sn), block_rows,
                                      _p(st.ws), st.nbytes, None, _p(g_Z), _p(g_ell), _p(g_var), _p(g_sn), _p(g_mean)),
               "gpb_sgpr_grad_finish")
    ev[6].record()
    torch.cuda.synchronize()
    names = ["pass1_statistics", "allreduce_statistics", "finish_replicated", "pass2_gradient", "allreduce_gradient",
             "grad_finish_replicated"]
    out = {nm: ev[i].elapsed_time(ev[i + 1]) for i, nm in enumerate(names)}
    out["elbo"] = float(val.item())
    return out


def collapsed_elbo_fused(kind, X, y, Z, ell, variance, obs_stddev, mean_const=None, jitter=1e-6,
                         block_rows: int = DEFAULT_BLOCK_ROWS, group=None, statistics: str = "auto"):
    """ELBO of the collapsed (Titsias) bound for the rows held by this rank, all-reduced over `group`.

    statistics: "whitened" = the reference's order (A = Lz^-1 Kzx per block, objectives.py:387-390);
    "raw" = accumulate Kzx Kxz and whiten the M x M sums once (25 % fewer flop per value+gradient, rounding amplified by
    cond(Kzz)); "auto" (default) = raw only while `kzz_condition_estimate` <= RAW_STATISTICS_COND_LIMIT."""
    return CollapsedElboFunction.apply(kind, X, y, Z, ell, variance, obs_stddev, mean_const, jitter, block_rows, group,
                                       statistics)
