"""Times the leave-one-out objective and its gradient (KernelLooOpt's evaluation, gp.hpp:339-399) and the mean-gradient factor
at N x D on one GPU.  usage: python tools/loo_timing.py [N] [D]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from limbo_b200 import _lib, kernel, mean, model, synth  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
D = int(sys.argv[2]) if len(sys.argv) > 2 else 6
X = synth.points(1234, N, D)
y = synth.targets(X)
gp = model.GP(D, 1, kernel=kernel.SquaredExpARD, mean=mean.Data)
gp.kernel_function().set_h_params(np.concatenate([np.full(D, np.log(0.3)), [0.0]]))
gp.compute(X, y[:, None])
lib = _lib.load()


def timed(fn, reps=2):
    fn()
    lib.lb_sync(gp._h)
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    lib.lb_sync(gp._h)
    return (time.perf_counter() - t0) / reps, out


gp.compute_inv_kernel()
t_val, v = timed(gp.compute_log_loo_cv, 5)
t_grad, g = timed(gp.compute_kernel_grad_log_loo_cv, 2)
w = np.empty((N, 1), order="F")
t_kinv, _ = timed(lambda: _lib.check(lib.lb_kinv_obs_mean(gp._h, w.ctypes.data), "kinv_obs"), 5)
nh = D + 1
print(json.dumps({"N": N, "D": D, "loo_value_ms": t_val * 1e3, "loo_grad_ms": t_grad * 1e3, "n_hparams": nh,
                  "loo_grad_tflops": nh * 2.0 * N ** 3 / t_grad / 1e12, "reference_formulation_flops": nh * 6.0 * N ** 3,
                  "kinv_obs_mean_ms": t_kinv * 1e3, "kinv_read_gbs": 8.0 * N * N / t_kinv / 1e9, "loo": v, "grad": g.tolist()}))
