import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from limbo_b200 import kernel, mean, model, synth, _lib
N, D = 512, 6
X = synth.points(1234, N, D); y = synth.targets(X)
gp = model.GP(D, 1, kernel=kernel.SquaredExpARD, mean=mean.Data)
gp.compute(list(X), list(y[:, None]))
lib = _lib.load()
lib.lb_stage_kbuild.argtypes = [C.c_void_p]
lib.lb_debug_potf2.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
out = np.zeros(16, dtype=np.int64)
for rep in range(3):
    lib.lb_stage_kbuild(gp._h)   # fresh K in the buffer; block 0 is a valid SPD block
    lib.lb_debug_potf2(gp._h, 0, out.ctypes.data, 16)
    d = np.diff(out[:8])
    print("cycles: load %d | jb0 factor %d | jb0 inverse %d | jb0 panel %d | rest of jb loop %d | inv assembly %d | writeback %d | total %d"
          % (d[0], d[1], d[2], d[3], d[4], d[5], d[6], out[7] - out[0]))
