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
    """optax.adam (weight_decay = 0) / optax.adamw semantics: bias-corrected moments, eps outside the sqrt.  The step count
    lives on the parameters' device (a 0-d tensor, like optax's `count` array), so one update is a fixed sequence of device
    operations that `fit(..., cuda_graph=True)` can capture and replay."""

    def init(params):
        ref = next(iter(params.values()), None)
        count = torch.zeros((), dtype=torch.float64, device=None if ref is None else ref.device)
        return {"count": count, "mu": {k: torch.zeros_like(v) for k, v in params.items()},
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


# ---- L-BFGS with a zoom line search: the optimiser gpjax/fit.py:259-361 builds from optax ----------------------------------
# ox.lbfgs(linesearch=ox.scale_by_zoom_linesearch(max_linesearch_steps, initial_guess_strategy="one")): memory 10, initial inverse
# Hessian scaled by s.y / y.y of the latest pair, strong-Wolfe zoom search (sufficient decrease 1e-4, curvature 0.9) started at
# step 1, and the loop of fit_lbfgs: continue while  n == 0 or (n < max_iters and |grad|_2 >= gtol).  Host glue on the raveled
# unconstrained parameters (a few dozen numbers): numpy float64, one objective evaluation per trial step.
def _two_loop(g, S, Y):
    import numpy as np

    q = g.copy()
    alphas = []
    for s, y in zip(reversed(S), reversed(Y)):
        a = float(s @ q) / float(s @ y)
        alphas.append(a)
        q -= a * y
    if S:
        q *= float(S[-1] @ Y[-1]) / float(Y[-1] @ Y[-1])
    for (s, y), a in zip(zip(S, Y), reversed(alphas)):
        b = float(y @ q) / float(s @ y)
        q += (a - b) * s
    return q


def _zoom_linesearch(fun, x, f0, g0, d, max_steps, c1=1e-4, c2=0.9, increase=2.0):
    """Strong-Wolfe step along d from (x, f0, g0).  Returns (step, f, g, evaluations).  Bracketing by doubling, then zoom by safeguarded
    cubic interpolation; when the budget runs out the best point with sufficient decrease seen so far is taken (step 0 if none)."""
    import numpy as np

    slope0 = float(g0 @ d)
    if not (slope0 < 0.0):  # not a descent direction (curvature pair went bad): steepest descent
        d = -g0
        slope0 = float(g0 @ d)
    evals = 0
    best = (0.0, f0, g0)

    def phi(t):
        nonlocal evals, best
        f, g = fun(x + t * d)
        evals += 1
        if np.isfinite(f) and f <= f0 + c1 * t * slope0 and f < best[1]:
            best = (t, f, g)
        return f, g, float(g @ d)

    def cubic(a, fa, da, b, fb, db):
        z = 3.0 * (fa - fb) / (b - a) + da + db
        w2 = z * z - da * db
        if w2 < 0.0:
            return 0.5 * (a + b)
        w = np.sqrt(w2) * (1.0 if b > a else -1.0)
        t = b - (b - a) * (db + w - z) / (db - da + 2.0 * w)
        lo, hi = min(a, b), max(a, b)
        return t if lo + 0.1 * (hi - lo) <= t <= hi - 0.1 * (hi - lo) else 0.5 * (a + b)

    t_prev, f_prev, s_prev = 0.0, f0, slope0
    t = 1.0
    lo = hi = None
    while evals < max_steps:
        f, g, sl = phi(t)
        if not np.isfinite(f) or f > f0 + c1 * t * slope0 or (evals > 1 and f >= f_prev):
            lo, hi = (t_prev, f_prev, s_prev), (t, f, sl)
            break
        if abs(sl) <= -c2 * slope0:
            return t, f, g, evals, d
        if sl >= 0.0:
            lo, hi = (t, f, sl), (t_prev, f_prev, s_prev)
            break
        t_prev, f_prev, s_prev = t, f, sl
        t *= increase
    while lo is not None and evals < max_steps:
        t = cubic(lo[0], lo[1], lo[2], hi[0], hi[1], hi[2]) if np.isfinite(hi[1]) else 0.5 * (lo[0] + hi[0])
        f, g, sl = phi(t)
        if not np.isfinite(f) or f > f0 + c1 * t * slope0 or f >= lo[1]:
            hi = (t, f, sl)
        else:
            if abs(sl) <= -c2 * slope0:
                return t, f, g, evals, d
            if sl * (hi[0] - lo[0]) >= 0.0:
                hi = lo
            lo = (t, f, sl)
        if abs(hi[0] - lo[0]) <= 1e-12 * max(1.0, abs(lo[0])):
            break
    return best[0], best[1], best[2], evals, d


def lbfgs_minimize(fun, x0, max_iters: int = 100, max_linesearch_steps: int = 32, gtol: float = 1e-5, memory: int = 10):
    """Minimise fun(x) -> (value, gradient) from x0 (1-D float64 array).  Returns (x, value, gradient, iterations)."""
    import numpy as np

    x = np.asarray(x0, np.float64).copy()
    f, g = fun(x)
    S, Y = [], []
    n = 0
    while n == 0 or (n < max_iters and float(np.linalg.norm(g)) >= gtol):
        d = -_two_loop(g, S, Y)
        t, f_new, g_new, _, d = _zoom_linesearch(fun, x, f, g, d, max_linesearch_steps)
        n += 1
        if t == 0.0:  # no acceptable step: the search direction is exhausted at this precision
            S, Y = [], []
            if n > 1 and float(np.linalg.norm(g)) < 1e3 * gtol:
                break
            continue
        s = t * d
        y = g_new - g
        x = x + s
        f, g = f_new, g_new
        if float(s @ y) > 1e-12 * float(np.linalg.norm(s)) * float(np.linalg.norm(y)):
            S.append(s); Y.append(y)
            if len(S) > memory:
                S.pop(0); Y.pop(0)
    return x, f, g, n

