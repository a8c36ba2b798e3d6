"""The outer loop that drives the hot path, mirroring bayes_opt::BOptimizer::optimize
(src/limbo/bayes_opt/boptimizer.hpp:139-170, bo_base.hpp:220-283) with a device-aware inner acquisition
optimiser: the reference's optimisers evaluate one point at a time (opt/optimizer.hpp:84-96), which would leave the GPU
idle; `BatchedRandomSearch` scores a whole candidate set per call through lb_acq_argmax."""
from __future__ import annotations

import numpy as np

from . import acqui as _acqui
from .params import get


class defaults_bo:
    class init_randomsampling:  # init/random_sampling.hpp:57-60
        samples = 10

    class stop_maxiterations:  # stop/max_iterations.hpp:56-59
        iterations = 190

    class opt_batchedrandom:
        candidates = 20000
        refinements = 2
        shrink = 0.1


def _get(params, section, name):
    sec = getattr(params, section, None) if params is not None else None
    if sec is not None and hasattr(sec, name):
        return getattr(sec, name)
    return getattr(getattr(defaults_bo, section), name)


class EvaluationError(Exception):  # bo_base.hpp:106-107
    pass


class BatchedRandomSearch:
    """acquiopt<> policy: uniform candidates in [0,1]^D scored in one device pass, then `refinements` rounds of
    candidates drawn in a box shrinking around the incumbent."""

    def __init__(self, params=None, rng: np.random.Generator | None = None):
        self._params = params
        self._rng = rng if rng is not None else np.random.default_rng()

    def __call__(self, acqui, dim: int, bounded: bool = True) -> np.ndarray:
        m = int(_get(self._params, "opt_batchedrandom", "candidates"))
        cand = self._rng.random((m, dim))
        best, idx = acqui.argmax_batch(cand)
        x = cand[idx].copy()
        radius = 1.0
        for _ in range(int(_get(self._params, "opt_batchedrandom", "refinements"))):
            radius *= float(_get(self._params, "opt_batchedrandom", "shrink"))
            cand = x + (self._rng.random((m, dim)) * 2.0 - 1.0) * radius
            if bounded:
                cand = np.clip(cand, 0.0, 1.0)
            cand[0] = x  # keep the incumbent in the set
            b2, i2 = acqui.argmax_batch(cand)
            if b2 >= best:
                best, x = b2, cand[i2].copy()
        return x


class BOptimizer:
    def __init__(self, model, params=None, acqui=_acqui.UCB, acqui_opt=None, rng: np.random.Generator | None = None):
        self._model, self._params, self._acqui_cls = model, params, acqui
        self._rng = rng if rng is not None else np.random.default_rng()
        self._acqui_opt = acqui_opt if acqui_opt is not None else BatchedRandomSearch(params, self._rng)
        self._samples: list[np.ndarray] = []
        self._observations: list[np.ndarray] = []
        self._current_iteration = 0
        self._total_iterations = 0

    # bo_base.hpp:220-245
    def add_new_sample(self, s, v) -> None:
        self._samples.append(np.asarray(s, dtype=np.float64))
        self._observations.append(np.atleast_1d(np.asarray(v, dtype=np.float64)))

    def eval_and_add(self, sfun, sample) -> None:
        v = np.atleast_1d(np.asarray(sfun(sample), dtype=np.float64))
        if np.any(~np.isfinite(v)):
            raise EvaluationError("the evaluation function returned NaN or inf")
        self.add_new_sample(sample, v)

    def optimize(self, sfun, dim_in: int, afun=_acqui.first_elem, reset: bool = True) -> None:
        # boptimizer.hpp:139-170
        if reset:
            self._samples, self._observations = [], []
            self._current_iteration = 0
        if self._total_iterations == 0 or reset:
            for _ in range(int(_get(self._params, "init_randomsampling", "samples"))):  # init/random_sampling.hpp:73-79
                self.eval_and_add(sfun, self._rng.random(dim_in))
        if self._observations:
            self._model.compute(np.stack(self._samples), np.stack(self._observations))
        hp_period = int(get(self._params, "bayes_opt_boptimizer", "hp_period"))
        max_it = int(_get(self._params, "stop_maxiterations", "iterations"))
        while self._current_iteration < max_it:
            acqui = self._acqui_cls(self._model, self._current_iteration, params=self._params)
            x = self._acqui_opt(acqui, dim_in, True)
            self.eval_and_add(sfun, x)
            self._model.add_sample(self._samples[-1], self._observations[-1])
            if hp_period > 0 and (self._current_iteration + 1) % hp_period == 0:
                self._model.optimize_hyperparams()
            self._current_iteration += 1
            self._total_iterations += 1

    def best_observation(self, afun=_acqui.first_elem) -> np.ndarray:  # boptimizer.hpp:173-180
        return max(self._observations, key=afun)

    def best_sample(self, afun=_acqui.first_elem) -> np.ndarray:  # boptimizer.hpp:183-190
        i = int(np.argmax([afun(o) for o in self._observations]))
        return self._samples[i]

    def model(self):
        return self._model

    def samples(self):
        return self._samples

    def observations(self):
        return self._observations
