"""Small end-to-end case for compute-sanitizer (memcheck / racecheck / synccheck): fit (quad-panel Cholesky), append, query
(panel path, fused slab kernel, multi-launch, one-point kernel), acquisition, log-lik, gradient, LOO value / gradient,
K^-1 obs_mean, SE-ARD with Lambda columns, Matern K build / gradient, tf32 / fp16 / fp16x3 queries, copy-on-write clones, and
the multi-GPU Cholesky blocks, the distributed fit and the inversion by column tiles at world = 1."""
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
for prec in ("fp64", "tf32", "fp16", "fp16x3"):
    gp = model.GP(D, 1, kernel=kernel.SquaredExpARD, mean=mean.Data, precision=prec)
    gp.compute(X[:N], y[:N, None])
    gp.add_sample(X[N], y[N:N + 1])
    gp.add_sample(X[N + 1], y[N + 1:N + 2])
    mu, s2 = gp.query_batch(Xq)
    print(prec, "query ok", float(mu.sum()), float(s2.sum()))
    print(acqui.EI(gp).argmax_batch(Xq))
    if prec != "fp64":  # inversion by column tiles (one rank) + cast + adoption, then the same scoring
        from limbo_b200 import dist_inv
        di = dist_inv.DistInverse(gp, 0, 1, "cuda:0")
        gp.recompute(False)
        di.prepare(gp)
        print(prec, "column inverse", float(np.abs(gp.query_batch(Xq)[1] - s2).max()))
        di.close()
    if prec == "fp64":
        lib = _lib.load()
        m1, v1 = gp.query(Xq[0])  # one-point kernel
        lib.lb_debug_set_query_panel_min(1 << 40)  # the same batch on the fused slab kernel
        mu_s, s2_s = gp.query_batch(Xq)
        lib.lb_debug_set_query_panel_min(0)
        assert np.abs(mu - mu_s).max() < 1e-10 and np.abs(s2 - s2_s).max() < 1e-10 and abs(v1 - s2_s[0]) < 1e-12
        c = gp.copy()  # copy-on-write clone: refit with other hyper-parameters, the source stays intact
        c.kernel_function().set_h_params(c.kernel_function().h_params() - 0.2)
        c.recompute(False)
        assert np.array_equal(gp.query_batch(Xq[:100])[1], s2_s[:100])
        del c
        lib.lb_debug_force_unfused_query.argtypes = [C.c_void_p, C.c_int]
        lib.lb_debug_force_unfused_query(gp._h, 1)
        mu2, s22 = gp.query_batch(Xq)
        assert np.abs(mu - mu2).max() < 1e-10 and np.abs(s2 - s22).max() < 1e-10
        print(gp.compute_log_lik(), gp.compute_kernel_grad_log_lik())
        print(gp.compute_log_loo_cv(), gp.compute_kernel_grad_log_loo_cv())
        w = np.empty((N + 2, 1), order="F")
        _lib.check(lib.lb_kinv_obs_mean(gp._h, w.ctypes.data), "kinv_obs")


class PL:
    class kernel_squared_exp_ard:
        k = 2
        sigma_sq = 1.0


gl = model.GP(D, 1, params=PL, kernel=kernel.SquaredExpARD, mean=mean.Data)
hp = gl.kernel_function().h_params()
hp[D:3 * D] = np.linspace(-0.5, 0.5, 2 * D)
gl.kernel_function().set_h_params(hp)
gl.compute(X[:N], y[:N, None])
print("lambda", gl.compute_log_lik(), gl.compute_kernel_grad_log_lik()[:3], gl.query_batch(Xq[:50])[1].sum())

gm = model.GP(D, 1, kernel=kernel.MaternFiveHalves, mean=mean.Data)
gm.compute(X[:N], y[:N, None])
print("matern", gm.compute_log_lik(), gm.compute_kernel_grad_log_lik(), gm.query_batch(Xq)[1].sum())

from limbo_b200 import dist_chol, dist_fit  # noqa: E402
gd = model.GP(D, 1, kernel=kernel.SquaredExpARD, mean=mean.Data)
gd.compute(X[:256], y[:256, None], compute_kernel=False)
fitter = dist_fit.DistFit(gd, 0, 1, "cuda:0")
print("dist_fit", fitter.fit(gd), gd.compute_log_lik())
fitter.close()
dc = dist_chol.DistCholesky(X[:N], gp.kernel_function(), 0, 1, "cuda:0")
dc.build()
print("dchol", dc.factor())
dc.close()
print("SANITIZE CASE DONE")
