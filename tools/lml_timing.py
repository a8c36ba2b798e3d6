"""Times one KernelLFOpt objective evaluation (config 3 unit: recompute -> log-lik -> K^-1 -> gradient) per stage.
usage: python tools/lml_timing.py [N] [reps]"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from limbo_b200 import _lib, kernel, mean, model, synth  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
D = 6
PC = ["kbuild", "potf2", "trsm_panel", "syrk", "syrk_col", "trsv", "kstar", "qstep", "qreduce", "acq", "trtri", "lauum", "grad", "other"]
X = synth.points(1234, N, D)
y = synth.targets(X)
gp = model.GP(D, 1, kernel=kernel.SquaredExpARD, mean=mean.Data)
gp.compute(list(X), list(y[:, None]))
lib = _lib.load()
lib.lb_profile_enable.argtypes = [C.c_void_p, C.c_int]
lib.lb_profile_read.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
lib.lb_sync.argtypes = [C.c_void_p]
hp = np.zeros(D + 1)
gp.kernel_function().set_h_params(hp + 0.05)
gp.recompute(False); gp.compute_log_lik(); gp.compute_kernel_grad_log_lik()  # warm-up (allocations)
lib.lb_profile_enable(gp._h, 1)
ms = (C.c_double * len(PC))(); cnt = (C.c_longlong * len(PC))()
lib.lb_profile_read(gp._h, ms, cnt, 1)
t0 = time.perf_counter()
for r in range(reps):
    gp.kernel_function().set_h_params(hp + 0.01 * r)
    gp.recompute(False)
    ll = gp.compute_log_lik()
    g = gp.compute_kernel_grad_log_lik()
lib.lb_sync(gp._h)
wall = (time.perf_counter() - t0) / reps
lib.lb_profile_read(gp._h, ms, cnt, 1)
print({"N": N, "wall_ms_per_eval": wall * 1e3, "evals_per_s": 1.0 / wall, "loglik": ll, "grad": g.tolist(),
       "stage_ms": {PC[i]: ms[i] / reps for i in range(len(PC)) if cnt[i]}})
fl = N ** 3
print("flops/eval N^3 = %.3e -> %.2f TFLOP/s overall" % (fl, fl / wall / 1e12))
