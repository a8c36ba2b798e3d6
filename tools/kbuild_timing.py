"""Times the K-build kernel alone (lb_stage_kbuild) at N x D with CUDA events; prints GB/s vs the HBM peak."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from limbo_b200 import _lib, kernel, mean, model, synth  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
D = int(sys.argv[2]) if len(sys.argv) > 2 else 6
kname = sys.argv[3] if len(sys.argv) > 3 else "SquaredExpARD"
X = synth.points(1234, N, D)
y = synth.targets(X)
gp = model.GP(D, 1, kernel=getattr(kernel, kname), mean=mean.Data)
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
gp.set_stream(st.cuda_stream)
gp.compute(list(X), list(y[:, None]))
lib = _lib.load()
lib.lb_stage_kbuild.argtypes = [C.c_void_p]
ts = []
for r in range(25):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    lib.lb_stage_kbuild(gp._h)
    e1.record(st)
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts = np.array(ts[5:])
byts = 8.0 * N * N + 8.0 * N * D
peak = 6577.7
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
print(f"kbuild {kname} N={N} D={D}: median {np.median(ts):.4f} ms  min {ts.min():.4f} ms  -> {byts / np.median(ts) / 1e6:.0f} GB/s "
      f"= {byts / np.median(ts) / 1e6 / peak:.3f} of {peak} GB/s (includes the tiny scale_x launch)")
