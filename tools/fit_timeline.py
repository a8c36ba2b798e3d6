"""Per-launch timeline of one fit (lb_profile_* with LB_PROF_TIMELINE=1): where the main stream waits for the panel chain.
usage: LB_PROF_TIMELINE=1 python tools/fit_timeline.py [N] 2> timeline.txt"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from limbo_b200 import _lib, kernel, mean, model, synth  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
X = synth.points(1234, N, 6)
y = synth.targets(X)
gp = model.GP(6, 1, kernel=kernel.SquaredExpARD, mean=mean.Data)
gp.compute(X, y[:, None])
lib = _lib.load()
lib.lb_profile_enable.argtypes = [C.c_void_p, C.c_int]
lib.lb_profile_read.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
for _ in range(2):
    gp.recompute(False)
lib.lb_profile_enable(gp._h, 1)
gp.recompute(False)
ms = (C.c_double * 14)(); cnt = (C.c_longlong * 14)()
lib.lb_profile_read(gp._h, ms, cnt, 1)
