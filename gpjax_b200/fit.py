"""Optimisation drivers -- gpjax/fit.py:44-256,364-453.  Host-side glue around the fused objectives:
the gradient consumer (`value_and_grad` of `objective(model, batch)` w.r.t. the unconstrained values
of every trainable Parameter)."""
from __future__ import annotations

import copy
import os
import typing as tp

import torch

from .dataset import Dataset
from .optim import GradientTransformation, apply_updates
from .parameters import DEFAULT_BIJECTION, IdentityTransform, Module, Parameter


def _select(model: Module, trainable) -> tp.Dict[str, Parameter]:
    """nnx.split(model, trainable, ...) analogue: `trainable` is a Parameter subclass, a tuple of them,
    or a predicate (path, parameter) -> bool."""
    out = {}
    for path, p in model.named_parameters():
        if isinstance(trainable, type) or isinstance(trainable, tuple):
            ok = isinstance(p, trainable)
        elif callable(trainable):
            ok = bool(trainable(path, p))
        else:
            raise TypeError(f"unsupported trainable filter {trainable!r}")
        if ok:
            out[path] = p
    return out


class _Loss:
    def __init__(self, model, objective, params, bijection):
        self.model, self.objective, self.params = model, objective, params
        self.bij = {k: (bijection or {}).get(p.tag, IdentityTransform()) for k, p in params.items()}

    def unconstrained(self) -> dict:
        return {k: (self.bij[k].inv(p.value.detach()) if self.bij else p.value.detach()).clone()
                for k, p in self.params.items()}

    def value_and_grad(self, u: dict, batch: Dataset):
        leaves = {k: v.detach().requires_grad_(True) for k, v in u.items()}
        for k, p in self.params.items():
            p.value = self.bij[k](leaves[k])
        loss = self.objective(self.model, batch)
        grads = torch.autograd.grad(loss, list(leaves.values()), allow_unused=True)
        g = {k: (gi if gi is not None else torch.zeros_like(leaves[k])) for k, gi in zip(leaves, grads)}
        return loss.detach(), g

    def commit(self, u: dict) -> None:
        for k, p in self.params.items():
            p.value = self.bij[k](u[k].detach()).detach()


def get_batch(train_data: Dataset, batch_size: int, key) -> Dataset:
    """fit.py:364-381: indices drawn uniformly WITH replacement."""
    gen = key if isinstance(key, torch.Generator) else torch.Generator(device=train_data.X.device).manual_seed(int(key))
    idx = torch.randint(0, train_data.n, (batch_size,), generator=gen, device=train_data.X.device)
    return Dataset(X=train_data.X[idx], y=train_data.y[idx])


def fit(*, model: Module, objective, train_data: Dataset, optim: GradientTransformation,
        params_bijection: tp.Optional[dict] = DEFAULT_BIJECTION, trainable=Parameter, key=42, num_iters: int = 100,
        batch_size: int = -1, log_rate: int = 10, verbose: bool = True, unroll: int = 1, safe: bool = True,
        cuda_graph: tp.Optional[bool] = None):
    """Minimises `objective(model, batch)`; returns (optimised copy of the model, history[num_iters]).

    `cuda_graph` (not a reference argument; default: the environment switch GPB_FIT_CUDA_GRAPH, off when unset) plays the
    part `jax.jit` + `lax.scan` play in the reference's loop (fit.py:160-170): after `_GRAPH_WARMUP` ordinary iterations one
    optimisation step -- objective, its backward, the optimiser update, the history entry -- is captured once as a CUDA graph and
    replayed for the remaining iterations, so a step costs one graph launch instead of a few hundred kernel launches from Python
    (the launch-bound regime: N of a few thousand, BASELINE config 1).  Requirements, all checked or failing loudly at capture:
    full batches (`batch_size == -1`), everything on one CUDA device, an objective whose evaluation only enqueues device work
    (conjugate_mll, conjugate_loocv, and the sparse objectives: their statistics="auto" route decision needs a host poll, so a
    captured step takes the reference's whitened order unless a route is named), and an optimiser whose state is made of
    tensors (the optimisers in `gpjax_b200.optim` are)."""
    if safe:
        _check_model(model)
        _check_train_data(train_data)
        _check_optim(optim)
        _check_num_iters(num_iters)
        _check_batch_size(batch_size)
        _check_log_rate(log_rate)
        _check_verbose(verbose)
    model = copy.deepcopy(model)
    loss = _Loss(model, objective, _select(model, trainable), params_bijection)
    u = loss.unconstrained()
    opt_state = optim.init(u)
    history = torch.empty(num_iters, dtype=torch.float64, device=train_data.X.device)
    gen = None
    if batch_size != -1:
        gen = torch.Generator(device=train_data.X.device).manual_seed(int(key))
    bar = None
    if verbose:
        try:
            from tqdm import trange

            bar = trange(num_iters)
        except Exception:  # pragma: no cover
            bar = None
    if cuda_graph is None:  # the environment default only applies where a step can be captured; an explicit True is strict
        cuda_graph = (os.environ.get("GPB_FIT_CUDA_GRAPH", "0") not in ("", "0") and batch_size == -1
                      and train_data.X.device.type == "cuda")
    if cuda_graph:
        if batch_size != -1:
            raise NotImplementedError("fit(cuda_graph=True) replays one captured step: it needs full batches (batch_size=-1)")
        if train_data.X.device.type != "cuda":
            raise RuntimeError("fit(cuda_graph=True) needs the training data on a CUDA device")
    n_eager = min(num_iters, _GRAPH_WARMUP) if cuda_graph else num_iters
    it = 0
    for it in range(n_eager):
        batch = get_batch(train_data, batch_size, gen) if batch_size != -1 else train_data
        val, grads = loss.value_and_grad(u, batch)
        updates, opt_state = optim.update(grads, opt_state, u)
        u = apply_updates(u, updates)
        history[it] = val
        if bar is not None:
            bar.update(1)
            if it % log_rate == 0:
                bar.set_postfix(Value=f"{val.item():.2f}")
    if n_eager < num_iters:
        u = _replay_captured_steps(loss, optim, u, opt_state, train_data, history, n_eager, num_iters, bar, log_rate)
    if bar is not None:
        bar.close()
    loss.commit(u)
    return model, history


_GRAPH_WARMUP = 3  # ordinary iterations before the capture: they size the cached workspaces and create the helper streams


def _tree_tensors(old, new, where="optimiser state"):
    """Pairs (old tensor, new tensor) of two optimiser states of equal structure; a host-side leaf that changes from step to step
    (a Python step counter) cannot be replayed and is refused."""
    if isinstance(old, torch.Tensor):
        if not isinstance(new, torch.Tensor) or new.shape != old.shape or new.dtype != old.dtype:
            raise TypeError(f"fit(cuda_graph=True): {where} changes type or shape between steps")
        return [(old, new)]
    if isinstance(old, dict):
        if not isinstance(new, dict) or old.keys() != new.keys():
            raise TypeError(f"fit(cuda_graph=True): {where} changes structure between steps")
        return [pair for k in old for pair in _tree_tensors(old[k], new[k], f"{where}[{k!r}]")]
    if isinstance(old, (tuple, list)):
        if not isinstance(new, (tuple, list)) or len(old) != len(new):
            raise TypeError(f"fit(cuda_graph=True): {where} changes structure between steps")
        return [pair for i, (a, b) in enumerate(zip(old, new)) for pair in _tree_tensors(a, b, f"{where}[{i}]")]
    if old != new:
        raise TypeError(f"fit(cuda_graph=True): {where} holds a host-side value that changes every step ({old!r} -> {new!r}); "
                        "keep optimiser state in device tensors")
    return []


def _tree_clone(t):
    if isinstance(t, torch.Tensor):
        return t.clone()
    if isinstance(t, dict):
        return {k: _tree_clone(v) for k, v in t.items()}
    if isinstance(t, (tuple, list)):
        return type(t)(_tree_clone(v) for v in t) if not hasattr(t, "_fields") else type(t)(*(_tree_clone(v) for v in t))
    return t


def _replay_captured_steps(loss, optim, u, opt_state, train_data, history, first, num_iters, bar, log_rate):
    """Iterations first .. num_iters-1 as replays of ONE captured step.  The step reads and writes fixed buffers: the unconstrained
    parameters `u`, the optimiser state, a device-side iteration index and `history`; the objective's own buffers (the cached
    N x N workspace of ops._mll_state, the streamed SGPR state) were sized by the ordinary iterations and stay alive in their
    caches while the graph exists."""
    dev = train_data.X.device
    with torch.cuda.device(dev):
        u = {k: v.detach().clone() for k, v in u.items()}
        opt_state = _tree_clone(opt_state)
        slot = torch.full((1,), first, dtype=torch.int64, device=dev)
        graph = torch.cuda.CUDAGraph()
        torch.cuda.synchronize(dev)
        with torch.cuda.graph(graph):
            val, grads = loss.value_and_grad(u, train_data)
            updates, new_state = optim.update(grads, opt_state, u)
            new_u = apply_updates(u, updates)
            history.index_copy_(0, slot, val.reshape(1).to(history.dtype))
            slot.add_(1)
            for k in u:
                u[k].copy_(new_u[k])
            for old, new in _tree_tensors(opt_state, new_state):
                old.copy_(new)
        for it in range(first, num_iters):
            graph.replay()
            if bar is not None:
                bar.update(1)
                if it % log_rate == 0:
                    bar.set_postfix(Value=f"{history[it].item():.2f}")
        torch.cuda.synchronize(dev)  # the graph (and its private pool) is released on return
    return u


def fit_scipy(*, model: Module, objective, train_data: Dataset, trainable=Parameter, max_iters: int = 500,
              verbose: bool = True, safe: bool = True):
    """fit.py:183-256: scipy.optimize.minimize on the raveled unconstrained parameters (value and
    gradient copied to the host every iteration, exactly as the reference does)."""
    import numpy as np
    from scipy.optimize import minimize

    if safe:
        _check_model(model)
        _check_train_data(train_data)
        _check_num_iters(max_iters)
        _check_verbose(verbose)
    model = copy.deepcopy(model)
    loss = _Loss(model, objective, _select(model, trainable), DEFAULT_BIJECTION)
    u0 = loss.unconstrained()
    keys = list(u0)
    shapes = [u0[k].shape for k in keys]
    sizes = [int(u0[k].numel()) for k in keys]
    dev = train_data.X.device

    def unravel(x):
        out, o = {}, 0
        for k, sh, sz in zip(keys, shapes, sizes):
            out[k] = torch.as_tensor(x[o:o + sz], dtype=torch.float64, device=dev).reshape(sh)
            o += sz
        return out

    def wrapper(x):
        val, g = loss.value_and_grad(unravel(x), train_data)
        return float(val), np.concatenate([g[k].reshape(-1).cpu().numpy() for k in keys])

    x0 = np.concatenate([u0[k].reshape(-1).cpu().numpy() for k in keys])
    history = [wrapper(x0)[0]]
    result = minimize(fun=wrapper, x0=x0, jac=True, callback=lambda xk: history.append(wrapper(xk)[0]),
                      options={"maxiter": max_iters, "disp": verbose})
    loss.commit(unravel(result.x))
    return model, torch.as_tensor(history, dtype=torch.float64)


def fit_lbfgs(*, model: Module, objective, train_data: Dataset, params_bijection: tp.Optional[dict] = DEFAULT_BIJECTION,
              trainable=Parameter, max_iters: int = 100, safe: bool = True, max_linesearch_steps: int = 32,
              gtol: float = 1e-5):
    """fit.py:259-361.  The reference drives optax's L-BFGS (memory 10, zoom line search from step 1) inside a lax.while_loop
    that runs while  n == 0 or (n < max_iters and |grad|_2 >= gtol);  the same optimiser is restated as host glue in
    optim.lbfgs_minimize on the raveled unconstrained parameters.  Returns (optimised model, final loss)."""
    import numpy as np

    from .optim import lbfgs_minimize

    if safe:
        _check_model(model)
        _check_train_data(train_data)
        _check_num_iters(max_iters)
    model = copy.deepcopy(model)
    loss = _Loss(model, objective, _select(model, trainable), params_bijection)
    u0 = loss.unconstrained()
    keys = list(u0)
    shapes = [u0[k].shape for k in keys]
    sizes = [int(u0[k].numel()) for k in keys]
    dev = train_data.X.device

    def unravel(x):
        out, o = {}, 0
        for k, sh, sz in zip(keys, shapes, sizes):
            out[k] = torch.as_tensor(x[o:o + sz], dtype=torch.float64, device=dev).reshape(sh)
            o += sz
        return out

    def wrapper(x):
        val, g = loss.value_and_grad(unravel(x), train_data)
        return float(val), np.concatenate([g[k].reshape(-1).cpu().numpy() for k in keys])

    x0 = np.concatenate([u0[k].reshape(-1).cpu().numpy() for k in keys])
    x, fval, _, _ = lbfgs_minimize(wrapper, x0, max_iters=max_iters, max_linesearch_steps=max_linesearch_steps, gtol=gtol)
    loss.commit(unravel(x))
    return model, torch.as_tensor(fval, dtype=torch.float64)


def _check_model(model) -> None:
    if not isinstance(model, Module):
        raise TypeError(f"Expected model to be a subclass of nnx.Module. Got {model} of type {type(model)}.")


def _check_train_data(train_data) -> None:
    if not isinstance(train_data, Dataset):
        raise TypeError(f"Expected train_data to be of type gpjax.Dataset. Got {train_data} of type {type(train_data)}.")


def _check_optim(optim) -> None:
    if not isinstance(optim, GradientTransformation):
        raise TypeError(f"Expected optim to be of type optax.GradientTransformation. Got {optim} of type {type(optim)}.")


def _check_num_iters(num_iters) -> None:
    if not isinstance(num_iters, int):
        raise TypeError(f"Expected num_iters to be of type int. Got {num_iters} of type {type(num_iters)}.")
    if num_iters <= 0:
        raise ValueError(f"Expected num_iters to be positive. Got {num_iters}.")


def _check_log_rate(log_rate) -> None:
    if not isinstance(log_rate, int):
        raise TypeError(f"Expected log_rate to be of type int. Got {log_rate} of type {type(log_rate)}.")
    if not log_rate > 0:
        raise ValueError(f"Expected log_rate to be positive. Got {log_rate}.")


def _check_verbose(verbose) -> None:
    if not isinstance(verbose, bool):
        raise TypeError(f"Expected verbose to be of type bool. Got {verbose} of type {type(verbose)}.")


def _check_batch_size(batch_size) -> None:
    if not isinstance(batch_size, int):
        raise TypeError(f"Expected batch_size to be of type int. Got {batch_size} of type {type(batch_size)}.")
    if not batch_size == -1 and not batch_size > 0:
        raise ValueError(f"Expected batch_size to be positive or -1. Got {batch_size}.")


__all__ = ["fit", "fit_scipy", "fit_lbfgs", "get_batch"]
