"""Small end-to-end case for compute-sanitizer (memcheck / racecheck / synccheck): fit, append, query (fused and
multi-launch), acquisition, log-lik, gradient, TF32 query."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from limbo_b200 import _lib, acqui, kernel, mean, model, synth  # noqa: E402

N, D, M = 300, 6, 700
X = synth.points(1, N + 2, D)
y = synth.targets(X)
Xq = synth.points(2, M, D)
for prec in ("fp64", "tf32"):
    gp = model.GP(D, 1, kernel=kernel.SquaredExpARD, mean=mean.Data, precision=prec)
    gp.compute(X[:N], y[:N, None])
    gp.add_sample(X[N], y[N:N + 1])
    gp.add_sample(X[N + 1], y[N + 1:N + 2])
    mu, s2 = gp.query_batch(Xq)
    print(prec, "query ok", float(mu.sum()), float(s2.sum()))
    print(acqui.EI(gp).argmax_batch(Xq))
    if prec == "fp64":
        lib = _lib.load()
        lib.lb_debug_force_unfused_query.argtypes = [C.c_void_p, C.c_int]
        lib.lb_debug_force_unfused_query(gp._h, 1)
        mu2, s22 = gp.query_batch(Xq)
        assert np.abs(mu - mu2).max() < 1e-10 and np.abs(s2 - s22).max() < 1e-10
        print(gp.compute_log_lik(), gp.compute_kernel_grad_log_lik())
print("SANITIZE CASE DONE")
