"""Acquisition functions mirroring src/limbo/acqui/{ucb,gp_ucb,ei}.hpp.  The scalar
``__call__(v, afun, gradient)`` keeps the reference contract (one point, any host
aggregator); ``argmax_batch`` is the batched device path (FirstElem aggregator)."""
from __future__ import annotations

import math

import numpy as np

from . import _lib, opt
from .params import get


def first_elem(x):  # bayes_opt/bo_base.hpp:99-105
    return float(np.asarray(x)[0])


class UCB:
    def __init__(self, model, iteration: int = 0, params=None):
        self._model, self._params = model, params

    def dim_in(self):
        return self._model.dim_in()

    def dim_out(self):
        return self._model.dim_out()

    def _alpha(self) -> float:
        return float(get(self._params, "acqui_ucb", "alpha"))

    def __call__(self, v, afun=first_elem, gradient: bool = False):  # ucb.hpp:83-90
        assert not gradient
        mu, sigma = self._model.query(v)
        return opt.no_grad(afun(mu) + self._alpha() * math.sqrt(sigma))

    def argmax_batch(self, Xq, return_values: bool = False):
        return self._model.acq_argmax_batch(_lib.ACQ_UCB, [self._alpha(), 0.0], Xq, return_values)


class GP_UCB(UCB):
    def __init__(self, model, iteration: int, params=None):  # gp_ucb.hpp:83-88
        super().__init__(model, iteration, params)
        nt = math.pow(iteration, model.dim_in() / 2.0 + 2.0)
        delta3 = float(get(params, "acqui_gpucb", "delta")) * 3
        self._beta = math.sqrt(2.0 * math.log(nt * math.pi * math.pi / delta3))

    def _alpha(self) -> float:
        return self._beta


class EI:
    def __init__(self, model, iteration: int = 0, params=None):
        self._model, self._params = model, params
        self._nb_samples = -1
        self._f_max = 0.0

    def dim_in(self):
        return self._model.dim_in()

    def dim_out(self):
        return self._model.dim_out()

    def _update_f_max(self, afun) -> None:  # ei.hpp:100-108, batched: N mu() calls in one pass
        if self._nb_samples != self._model.nb_samples():
            mu, _ = self._model.query_batch(np.stack(self._model.samples(), axis=0))
            self._f_max = max(afun(m) for m in mu)
            self._nb_samples = self._model.nb_samples()

    def __call__(self, v, afun=first_elem, gradient: bool = False):  # ei.hpp:85-116
        assert not gradient
        mu, sigma_sq = self._model.query(v)
        sigma = math.sqrt(sigma_sq)
        if sigma < 1e-10 or len(self._model.samples()) < 1:
            return opt.no_grad(0.0)
        self._update_f_max(afun)
        X = afun(mu) - self._f_max - float(get(self._params, "acqui_ei", "jitter"))
        Z = X / sigma
        phi = math.exp(-0.5 * math.pow(Z, 2.0)) / math.sqrt(2.0 * math.pi)
        Phi = 0.5 * math.erfc(-Z / math.sqrt(2))
        return opt.no_grad(X * Phi + sigma * phi)

    def argmax_batch(self, Xq, return_values: bool = False):
        if len(self._model.samples()) < 1:
            vals = np.zeros(len(Xq))
            return (0.0, 0, vals) if return_values else (0.0, 0)
        self._update_f_max(first_elem)
        return self._model.acq_argmax_batch(_lib.ACQ_EI, [self._f_max, float(get(self._params, "acqui_ei", "jitter"))], Xq,
                                            return_values)
