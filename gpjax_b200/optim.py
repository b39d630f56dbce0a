"""Tiny optax-shaped optimisers (init/update pairs) so `fit` keeps the reference's calling convention
(`optim=ox.adam(1e-2)`, gpjax/fit.py:49,146,161).  State and updates are device tensors; nothing here
is on the hot path."""
from __future__ import annotations

import typing as tp

import torch


class GradientTransformation(tp.NamedTuple):
    init: tp.Callable
    update: tp.Callable


def apply_updates(params: dict, updates: dict) -> dict:
    return {k: params[k] + updates[k] for k in params}


def sgd(learning_rate: float) -> GradientTransformation:
    def init(params):
        return ()

    def update(grads, state, params=None):
        return {k: -learning_rate * g for k, g in grads.items()}, state

    return GradientTransformation(init, update)


def adam(learning_rate: float, b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8,
         weight_decay: float = 0.0) -> GradientTransformation:
    """optax.adam (weight_decay = 0) / optax.adamw semantics: bias-corrected moments, eps outside the sqrt."""

    def init(params):
        return {"count": 0, "mu": {k: torch.zeros_like(v) for k, v in params.items()},
                "nu": {k: torch.zeros_like(v) for k, v in params.items()}}

    def update(grads, state, params=None):
        t = state["count"] + 1
        mu = {k: b1 * state["mu"][k] + (1 - b1) * g for k, g in grads.items()}
        nu = {k: b2 * state["nu"][k] + (1 - b2) * g * g for k, g in grads.items()}
        upd = {}
        for k in grads:
            u = (mu[k] / (1 - b1**t)) / (torch.sqrt(nu[k] / (1 - b2**t)) + eps)
            if weight_decay and params is not None:
                u = u + weight_decay * params[k]
            upd[k] = -learning_rate * u
        return upd, {"count": t, "mu": mu, "nu": nu}

    return GradientTransformation(init, update)


def adamw(learning_rate: float, b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8,
          weight_decay: float = 1e-4) -> GradientTransformation:
    return adam(learning_rate, b1, b2, eps, weight_decay)
