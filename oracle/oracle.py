"""ctypes front-end of oracle/_build/liblimbo_oracle.so (see oracle/limbo_oracle.hpp).

TEST INFRASTRUCTURE ONLY: the product package limbo_b200/ never imports this."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "liblimbo_oracle.so")
_lib = None

PREC_DOUBLE, PREC_LONG_DOUBLE, PREC_QUAD = 0, 1, 2
K_SE_ARD, K_MATERN52, K_MATERN32, K_EXP = 0, 1, 2, 3


def build(force: bool = False) -> str:
    src = [os.path.join(HERE, f) for f in ("oracle_capi.cpp", "limbo_oracle.hpp", "Makefile")]
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in src):
        subprocess.run(["make", "-C", HERE, "-B", "CXX=g++"], check=True, capture_output=True)
    return LIB_PATH


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        lib = C.CDLL(LIB_PATH)
        vp, lg, i, d = C.c_void_p, C.c_long, C.c_int, C.c_double
        sig = {
            "lbo_create": ([i], vp), "lbo_destroy": ([vp], None), "lbo_clone": ([vp], vp),
            "lbo_set_data": ([vp, lg, i, i, vp, vp], None), "lbo_set_kernel": ([vp, i, vp, i, d], None),
            "lbo_fit": ([vp], lg), "lbo_refit_alpha": ([vp, vp], None), "lbo_append": ([vp, vp, vp], None),
            "lbo_query": ([vp, lg, vp, vp, vp, i], None), "lbo_log_lik": ([vp], d), "lbo_grad": ([vp, vp, i], None),
            "lbo_loo_cv": ([vp], d), "lbo_loo_grad": ([vp, vp, i], None), "lbo_kinv_obs": ([vp, vp], None),
            "lbo_get": ([vp, i, vp], None), "lbo_n": ([vp], lg),
            "lbo_kernel_eval": ([i, i, vp, d, vp, vp, i], d), "lbo_kernel_grad": ([i, i, vp, vp, vp, vp], None),
            "lbo_ucb": ([lg, vp, vp, d, vp], None), "lbo_ei": ([lg, vp, vp, d, d, vp], None),
            "lbo_gp_ucb_beta": ([i, i, d], d),
            "lbo_lml_eval": ([vp, vp, i, d, i, vp, i], d),
            "lbo_rprop_lml": ([vp, i, vp, i, d, i, d, i, vp, vp], None),
            "lbo_cholesky": ([lg, vp], lg), "lbo_has_quad": ([], i),
        }
        for n, (a, r) in sig.items():
            f = getattr(lib, n)
            f.argtypes, f.restype = a, r
        _lib = lib
    return _lib


def _p(a):
    return a.ctypes.data if a is not None else None


class OracleGP:
    """Restated limbo::model::GP (mean handled by the caller: pass obs_mean, add mean(v) to mu)."""

    def __init__(self, prec: int = PREC_DOUBLE):
        self.lib = load()
        self.h = self.lib.lbo_create(prec)
        self.N = self.D = self.P = 0
        self.kernel_id = K_SE_ARD
        self.noise = 0.01

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.lbo_destroy(self.h)
            self.h = None

    def set_data(self, X, obs_mean):
        X = np.ascontiguousarray(X, dtype=np.float64)
        Y = np.asfortranarray(np.atleast_2d(obs_mean.T).T if obs_mean.ndim == 1 else obs_mean, dtype=np.float64)
        if Y.ndim == 1:
            Y = Y[:, None]
        self.N, self.D = X.shape
        self.P = Y.shape[1]
        self.lib.lbo_set_data(self.h, self.N, self.D, self.P, _p(X), _p(Y))

    def set_kernel(self, kernel_id, hparams, noise):
        hp = np.ascontiguousarray(hparams, dtype=np.float64)
        self.kernel_id, self.noise, self.hp = kernel_id, float(noise), hp
        self.lib.lbo_set_kernel(self.h, kernel_id, _p(hp), hp.size, float(noise))

    def fit(self) -> int:
        return int(self.lib.lbo_fit(self.h))

    def refit_alpha(self, obs_mean):
        Y = np.asfortranarray(obs_mean, dtype=np.float64)
        self.lib.lbo_refit_alpha(self.h, _p(Y))

    def append(self, x, obs_mean_all):
        x = np.ascontiguousarray(x, dtype=np.float64)
        Y = np.asfortranarray(obs_mean_all, dtype=np.float64)
        self.lib.lbo_append(self.h, _p(x), _p(Y))
        self.N += 1

    def query(self, Xq, nthreads: int = 1):
        Xq = np.ascontiguousarray(np.atleast_2d(Xq), dtype=np.float64)
        M = Xq.shape[0]
        mu = np.empty((M, max(self.P, 1)))
        s2 = np.empty(M)
        self.lib.lbo_query(self.h, M, _p(Xq), _p(mu), _p(s2), nthreads)
        return mu, s2

    def log_lik(self) -> float:
        return float(self.lib.lbo_log_lik(self.h))

    def grad(self, optimize_noise: bool = False):
        nh = self.hp.size + (1 if optimize_noise else 0)
        g = np.empty(nh)
        self.lib.lbo_grad(self.h, _p(g), int(optimize_noise))
        return g

    def loo_cv(self) -> float:
        return float(self.lib.lbo_loo_cv(self.h))

    def loo_grad(self, optimize_noise: bool = False):
        nh = self.hp.size + (1 if optimize_noise else 0)
        g = np.empty(nh)
        self.lib.lbo_loo_grad(self.h, _p(g), int(optimize_noise))
        return g

    def kinv_obs(self):
        n = int(self.lib.lbo_n(self.h))
        out = np.empty((n, self.P), order="F")
        self.lib.lbo_kinv_obs(self.h, _p(out))
        return out

    def get(self, what: int):
        n = int(self.lib.lbo_n(self.h))
        shape = (n, self.P) if what == 2 else (n, n)
        out = np.empty(shape, order="F")
        self.lib.lbo_get(self.h, what, _p(out))
        return out

    def lml_eval(self, hparams, compute_grad: bool = True, optimize_noise: bool = False):
        hp = np.ascontiguousarray(hparams, dtype=np.float64)
        g = np.empty(hp.size) if compute_grad else None
        v = self.lib.lbo_lml_eval(self.h, _p(hp), hp.size, self.noise, self.kernel_id, _p(g), int(optimize_noise))
        return float(v), g

    def rprop_lml(self, init, iterations: int, eps_stop: float = 0.0, optimize_noise: bool = False):
        hp = np.ascontiguousarray(init, dtype=np.float64)
        out = np.empty(hp.size)
        ne = C.c_long(0)
        self.lib.lbo_rprop_lml(self.h, self.kernel_id, _p(hp), hp.size, self.noise, iterations, eps_stop, int(optimize_noise),
                               _p(out), C.addressof(ne))
        return out, ne.value


def kernel_eval(kernel_id, hparams, noise, x1, x2, same_index=False) -> float:
    lib = load()
    hp = np.ascontiguousarray(hparams, dtype=np.float64)
    x1 = np.ascontiguousarray(x1, dtype=np.float64)
    x2 = np.ascontiguousarray(x2, dtype=np.float64)
    return float(lib.lbo_kernel_eval(kernel_id, x1.size, _p(hp), float(noise), _p(x1), _p(x2), int(same_index)))


def kernel_grad(kernel_id, hparams, x1, x2):
    lib = load()
    hp = np.ascontiguousarray(hparams, dtype=np.float64)
    x1 = np.ascontiguousarray(x1, dtype=np.float64)
    x2 = np.ascontiguousarray(x2, dtype=np.float64)
    g = np.empty(hp.size)
    lib.lbo_kernel_grad(kernel_id, x1.size, _p(hp), _p(x1), _p(x2), _p(g))
    return g


def ucb(mu0, s2, alpha):
    lib = load()
    mu0 = np.ascontiguousarray(mu0, dtype=np.float64)
    s2 = np.ascontiguousarray(s2, dtype=np.float64)
    out = np.empty_like(mu0)
    lib.lbo_ucb(mu0.size, _p(mu0), _p(s2), float(alpha), _p(out))
    return out


def ei(mu0, s2, f_max, jitter=0.0):
    lib = load()
    mu0 = np.ascontiguousarray(mu0, dtype=np.float64)
    s2 = np.ascontiguousarray(s2, dtype=np.float64)
    out = np.empty_like(mu0)
    lib.lbo_ei(mu0.size, _p(mu0), _p(s2), float(f_max), float(jitter), _p(out))
    return out


def cholesky(A):
    lib = load()
    L = np.asfortranarray(A, dtype=np.float64).copy(order="F")
    info = lib.lbo_cholesky(L.shape[0], _p(L))
    return np.tril(L), int(info)
