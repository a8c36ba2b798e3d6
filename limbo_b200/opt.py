"""Inner optimisers on the hot path's caller side, mirroring src/limbo/opt/:
eval_t helpers (optimizer.hpp:60-96), Rprop (rprop.hpp:82-145) and
ParallelRepeater (parallel_repeater.hpp:76-107).  Host control flow only."""
from __future__ import annotations

import math

import numpy as np

from .params import get


def no_grad(x: float):  # optimizer.hpp:66-69
    return (x, None)


def fun(fg) -> float:  # optimizer.hpp:77-81
    return fg[0]


def grad(fg) -> np.ndarray:  # optimizer.hpp:71-75
    assert fg[1] is not None
    return fg[1]


def eval(f, x) -> float:  # optimizer.hpp:84-88
    return f(x, False)[0]


def eval_grad(f, x):  # optimizer.hpp:91-95
    return f(x, True)


def _signum(x: float) -> int:  # tools/math.hpp:71-91
    return int(0 < x) - int(x < 0)


class Rprop:
    def __init__(self, params=None):
        self._params = params

    def __call__(self, f, init, bounded: bool) -> np.ndarray:
        iterations = int(get(self._params, "opt_rprop", "iterations"))
        eps_stop = float(get(self._params, "opt_rprop", "eps_stop"))
        assert eps_stop >= 0.0
        init = np.asarray(init, dtype=np.float64)
        n = init.size
        delta0, deltamin, deltamax, etaminus, etaplus = 0.1, 1e-6, 50.0, 0.5, 1.2
        delta = np.ones(n) * delta0
        grad_old = np.zeros(n)
        params = init.copy()
        if bounded:
            params = np.clip(params, 0.0, 1.0)
        best_params = params.copy()
        best = -math.inf  # log(0)
        for _ in range(iterations):
            perf = eval_grad(f, params)
            lik = fun(perf)
            if lik > best:
                best = lik
                best_params = params.copy()
            g = -np.asarray(grad(perf), dtype=np.float64)
            grad_old = grad_old * g
            for j in range(n):
                if grad_old[j] > 0:
                    delta[j] = min(delta[j] * etaplus, deltamax)
                elif grad_old[j] < 0:
                    delta[j] = max(delta[j] * etaminus, deltamin)
                    g[j] = 0
                params[j] += -_signum(g[j]) * delta[j]
                if bounded and params[j] < 0:
                    params[j] = 0
                if bounded and params[j] > 1:
                    params[j] = 1
            grad_old = g
            if np.linalg.norm(grad_old) < eps_stop:
                break
        return best_params


class ParallelRepeater:
    """Restarts run one after another on the device (each evaluation already fills
    the GPU); multi-GPU runs shard restarts across ranks (limbo_b200.dist)."""

    def __init__(self, params=None, optimizer=None, rng: np.random.Generator | None = None):
        self._params = params
        self._optimizer = optimizer if optimizer is not None else Rprop(params)
        self._rng = rng if rng is not None else np.random.default_rng()

    def __call__(self, f, init, bounded: bool) -> np.ndarray:
        repeats = int(get(self._params, "opt_parallelrepeater", "repeats"))
        epsilon = float(get(self._params, "opt_parallelrepeater", "epsilon"))
        assert repeats > 0
        assert epsilon > 0.0
        init = np.asarray(init, dtype=np.float64)
        best_v, best_val = init, -float(np.finfo(np.float32).max)  # parallel_repeater.hpp:102
        for _ in range(repeats):
            r_deviation = self._rng.random(init.size) * 2.0 * epsilon - epsilon
            v = self._optimizer(f, init + r_deviation, bounded)
            val = eval(f, v)
            if val > best_val:
                best_v, best_val = v, val
        return best_v
