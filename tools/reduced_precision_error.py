"""Max / mean error of the reduced-precision (tf32, fp16) variance and mean against the fp64 path on the same model.
usage: python tools/reduced_precision_error.py"""
import numpy as np, sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from limbo_b200 import kernel, mean, model, synth
for kname, N, D in [("Exp", 513, 5), ("SquaredExpARD", 1000, 12), ("MaternFiveHalves", 700, 6), ("SquaredExpARD", 4096, 6)]:
    X = synth.points(77, N, D); y = np.cos(3 * X.sum(1)); Xq = synth.points(78, 2000, D)
    kw = dict(kernel=getattr(kernel, kname), mean=mean.Data)
    g64 = model.GP(D, 1, **kw); g64.compute(X, y[:, None]); m64, s64 = g64.query_batch(Xq)
    for prec in ("tf32", "fp16"):
        g = model.GP(D, 1, precision=prec, **kw); g.compute(X, y[:, None]); m, s = g.query_batch(Xq)
        print(kname, N, prec, "max|ds2|", np.abs(s - s64).max(), "mean ds2", (s - s64).mean(), "max|dmu|", np.abs(m - m64).max())
