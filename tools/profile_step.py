"""One fit (+ optional query / LML gradient) at a chosen size, for ncu captures.
usage: python tools/profile_step.py [N] [M] [--grad]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from limbo_b200 import acqui, kernel, mean, model, synth  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
M = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
D = 6
X = synth.points(1234, N, D)
y = synth.targets(X)
gp = model.GP(D, 1, kernel=kernel.SquaredExpARD, mean=mean.Data)
gp.compute(list(X), list(y[:, None]))
if M > 0:
    Xq = synth.points(1235, M, D)
    print(acqui.UCB(gp).argmax_batch(Xq))
if "--grad" in sys.argv:
    print(gp.compute_log_lik(), gp.compute_kernel_grad_log_lik())
print("launches", gp.launch_count())
