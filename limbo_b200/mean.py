"""Mean-function policies mirroring limbo::mean::* (src/limbo/mean/).  Means are
arbitrary host functors in the reference (mean/mean.hpp:60-77); they stay on
the host here: the GP wrapper subtracts mean(x_i) before handing obs_mean to
the device and adds mean(v) to the returned mu.  ``batch`` evaluates M points."""
from __future__ import annotations

import numpy as np

from .params import get


class BaseMean:
    def __init__(self, params=None, dim_out: int = 1):
        self._dim_out = dim_out

    def h_params_size(self) -> int:
        return 0

    def h_params(self) -> np.ndarray:
        return np.zeros(0)

    def set_h_params(self, p) -> None:
        pass

    def grad(self, x, gp) -> np.ndarray:  # mean/mean.hpp:72-76
        return np.zeros((self._dim_out, 0))

    def batch(self, xs: np.ndarray, gp) -> np.ndarray:
        return np.stack([np.asarray(self(x, gp), dtype=np.float64) for x in xs], axis=0)

    def is_constant(self) -> bool:
        return False


class NullFunction(BaseMean):
    """mean/null_function.hpp:57-63"""

    def __call__(self, v, gp) -> np.ndarray:
        return np.zeros(self._dim_out)

    def batch(self, xs, gp):
        return np.zeros((len(xs), self._dim_out))

    def is_constant(self) -> bool:
        return True


class Constant(BaseMean):
    """mean/constant.hpp:66-94"""

    def __init__(self, params=None, dim_out: int = 1):
        super().__init__(params, dim_out)
        self._constant = float(get(params, "mean_constant", "constant"))

    def __call__(self, v, gp) -> np.ndarray:
        return np.full(self._dim_out, self._constant)

    def batch(self, xs, gp):
        return np.full((len(xs), self._dim_out), self._constant)

    def grad(self, x, gp):
        return np.ones((self._dim_out, 1))

    def h_params_size(self) -> int:
        return 1

    def h_params(self):
        return np.array([self._constant])

    def set_h_params(self, p):
        self._constant = float(np.asarray(p)[0])

    def is_constant(self) -> bool:
        return True


class Data(BaseMean):
    """mean/data.hpp:55-64: the mean of the observations."""

    def __call__(self, v, gp) -> np.ndarray:
        return np.asarray(gp.mean_observation(), dtype=np.float64)

    def batch(self, xs, gp):
        return np.tile(np.asarray(gp.mean_observation(), dtype=np.float64), (len(xs), 1))

    def is_constant(self) -> bool:
        return True
