"""model::MultiGP (src/limbo/model/multi_gp.hpp:60-300): P independent single-output GPs sharing the samples, one
per output dimension (the reference loops over them with tools::par; here each GP is its own device handle)."""
from __future__ import annotations

import numpy as np

from .. import kernel as _kernel
from .. import mean as _mean
from .gp import GP


class MultiGP:
    def __init__(self, dim_in: int = -1, dim_out: int = -1, params=None, kernel=_kernel.MaternFiveHalves, mean=_mean.Data,
                 hp_opt_factory=None, device: int = 0):
        self._dim_in, self._dim_out = dim_in, dim_out
        self._mk = lambda: GP(dim_in if dim_in > 0 else -1, 1, params=params, kernel=kernel, mean=mean,
                              hp_opt=hp_opt_factory() if hp_opt_factory else None, device=device)
        self._gp_models = [self._mk() for _ in range(max(dim_out, 0))]
        self._observations = np.zeros((0, max(dim_out, 1)))

    def compute(self, samples, observations, compute_kernel: bool = True) -> None:  # multi_gp.hpp:87-127
        assert len(samples) != 0 and len(samples) == len(observations)
        Y = np.ascontiguousarray(observations, dtype=np.float64)
        if Y.ndim == 1:
            Y = Y[:, None]
        if self._dim_out != Y.shape[1]:
            self._dim_out = Y.shape[1]
            self._gp_models = [self._mk() for _ in range(self._dim_out)]
        self._observations = Y
        X = np.ascontiguousarray(samples, dtype=np.float64)
        self._dim_in = X.shape[1]
        for p, gp in enumerate(self._gp_models):
            gp.compute(X, Y[:, p:p + 1], compute_kernel)

    def add_sample(self, sample, observation) -> None:  # multi_gp.hpp:149-176
        observation = np.atleast_1d(np.asarray(observation, dtype=np.float64))
        if not self._gp_models:
            self._dim_out = observation.size
            self._gp_models = [self._mk() for _ in range(self._dim_out)]
        for p, gp in enumerate(self._gp_models):
            gp.add_sample(sample, observation[p:p + 1])
        self._observations = np.vstack([self._observations.reshape(-1, self._dim_out), observation[None, :]])

    def query(self, v):  # multi_gp.hpp:183-203: per-output mu, sigma^2 of the FIRST output's... (max is not taken: each GP's own)
        mus, sig = [], []
        for gp in self._gp_models:
            m, s = gp.query(v)
            mus.append(m[0])
            sig.append(s)
        return np.array(mus), np.array(sig)

    def query_batch(self, Xq):
        res = [gp.query_batch(Xq) for gp in self._gp_models]
        return np.concatenate([r[0] for r in res], axis=1), np.stack([r[1] for r in res], axis=1)

    def mu(self, v):
        return self.query(v)[0]

    def sigma(self, v):
        return self.query(v)[1]

    def dim_in(self):
        return self._dim_in

    def dim_out(self):
        return self._dim_out

    def nb_samples(self):
        return self._gp_models[0].nb_samples() if self._gp_models else 0

    def samples(self):
        return self._gp_models[0].samples() if self._gp_models else []

    def gp_models(self):
        return self._gp_models

    def optimize_hyperparams(self) -> None:  # multi_gp.hpp:256-266 (ParallelLFOpt: every GP optimises its own kernel)
        for gp in self._gp_models:
            gp.optimize_hyperparams()

    def recompute(self, update_obs_mean: bool = True, update_full_kernel: bool = True) -> None:
        for gp in self._gp_models:
            gp.recompute(update_obs_mean, update_full_kernel)
