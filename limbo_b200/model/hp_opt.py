"""Hyper-parameter optimisation functors mirroring src/limbo/model/gp/*.hpp."""
from __future__ import annotations

import sys

import numpy as np

from .. import opt as _opt


class HPOpt:  # model/gp/hp_opt.hpp:57-76
    def __init__(self, params=None, optimizer=None):
        self._params = params
        self._optimizer = optimizer if optimizer is not None else _opt.Rprop(params)
        self._called = False


class NoLFOpt(HPOpt):  # model/gp/no_lf_opt.hpp:55-65
    def __call__(self, gp) -> None:
        print("'NoLFOpt' was called: nothing to optimize", file=sys.stderr)


class KernelLFOpt(HPOpt):
    """model/gp/kernel_lf_opt.hpp:57-97: maximise the log marginal likelihood over the
    kernel h-params.  Each evaluation is one device pipeline
    K -> L -> alpha -> log-lik [-> K^-1 -> gradient] on a private copy of the GP."""

    def __call__(self, gp) -> None:
        self._called = True
        optimization = _KernelLFOptimization(gp)
        params = self._optimizer(optimization, gp.kernel_function().h_params(), False)
        gp.kernel_function().set_h_params(params)
        gp.recompute(False)
        gp.compute_log_lik()


class _KernelLFOptimization:
    def __init__(self, gp):
        self._original_gp = gp
        self._work = None

    def __call__(self, params, compute_grad: bool):
        # kernel_lf_opt.hpp:79: GP gp(this->_original_gp).  The copy shares nothing
        # mutable with the original; we keep ONE workspace copy per functor (a
        # fresh device clone per evaluation would only re-copy identical X/Y).
        if self._work is None:
            self._work = self._original_gp.copy()
        gp = self._work
        gp.kernel_function().set_h_params(np.asarray(params, dtype=np.float64))
        gp.recompute(False)
        lik = gp.compute_log_lik()
        if not compute_grad:
            return _opt.no_grad(lik)
        return lik, gp.compute_kernel_grad_log_lik()
