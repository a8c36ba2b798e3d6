"""world_size-2 gloo test of the sharded-argmax host logic (no GPU): each rank owns a contiguous
candidate range, evaluates a stand-in acquisition, and the single all_gather collective returns the
same global (value, index) on both ranks — equal to the sequential scan, lowest index on ties."""
import os
import socket
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys, json
import numpy as np
import torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from limbo_b200 import dist as lbd
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{sys.argv[2]}", rank=int(sys.argv[3]), world_size=int(sys.argv[4]))
rank, world = dist.get_rank(), dist.get_world_size()
rng = np.random.default_rng(7)
Xq = rng.random((1001, 3))
vals = np.round(np.sin(Xq.sum(1) * 3.0), 2)            # rounded -> ties exist
class FakeAcq:
    def argmax_batch(self, X):
        v = np.round(np.sin(X.sum(1) * 3.0), 2)
        i = int(np.argmax(v))                         # first maximum = lowest index
        return float(v[i]), i
best, idx = lbd.sharded_acq_argmax(FakeAcq(), Xq, rank, world)
ref_i = int(np.argmax(vals))
ok = (idx == ref_i) and (best == float(vals[ref_i]))
lo, hi = lbd.shard_range(1001, rank, world)
print(json.dumps({"rank": rank, "ok": bool(ok), "idx": idx, "ref": ref_i, "lo": lo, "hi": hi}))
dist.destroy_process_group()
"""


WORKER_RESTARTS = r"""
import sys, json
import numpy as np
import torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from limbo_b200 import dist as lbd, opt
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{sys.argv[2]}", rank=int(sys.argv[3]), world_size=int(sys.argv[4]))
class P:
    class opt_rprop:
        iterations = 20
    class opt_parallelrepeater:
        repeats = 5
        epsilon = 0.5
def f(x, g):
    v = -float(((x - np.array([0.3, -0.7])) ** 2).sum()) + 0.1 * float(np.cos(5 * x).sum())
    gr = -2 * (x - np.array([0.3, -0.7])) - 0.5 * np.sin(5 * x)
    return (v, gr) if g else (v, None)
best = lbd.ShardedRepeater(P, opt.Rprop(P), seed=11)(f, np.zeros(2), False)
print(json.dumps({"rank": dist.get_rank(), "best": best.tolist()}))
dist.destroy_process_group()
"""


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_range_partitions():
    from limbo_b200 import dist as lbd
    for n in (0, 1, 7, 1000, 1001):
        for w in (1, 2, 3, 8):
            spans = [lbd.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_reduce_records_ties_and_nan():
    from limbo_b200 import dist as lbd
    v, i = lbd.reduce_records(np.array([1.0, 3.0, 3.0, np.nan]), np.array([10, 30, 20, 5]))
    assert (v, i) == (3.0, 20)
    v, i = lbd.reduce_records(np.array([-np.inf, 0.5]), np.array([-1, 4]))
    assert (v, i) == (0.5, 4)


def test_sharded_argmax_world2_gloo(tmp_path):
    import json
    port = _free_port()
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, OMP_NUM_THREADS="1")
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(port), str(r), "2"], stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True, env=env) for r in range(2)]
    outs = [p.communicate(timeout=120) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-2000:]
    recs = [json.loads(o.strip().splitlines()[-1]) for o, _ in outs]
    assert all(r["ok"] for r in recs), recs
    assert recs[0]["idx"] == recs[1]["idx"]
    assert recs[0]["hi"] == recs[1]["lo"]


def test_sharded_restarts_world2_equals_world1(tmp_path):
    """ParallelRepeater restarts spread over 2 ranks give the same winner as a single rank."""
    import json
    sys.path.insert(0, ROOT)
    from limbo_b200 import dist as lbd, opt

    class P:
        class opt_rprop:
            iterations = 20

        class opt_parallelrepeater:
            repeats = 5
            epsilon = 0.5

    def f(x, g):
        v = -float(((x - np.array([0.3, -0.7])) ** 2).sum()) + 0.1 * float(np.cos(5 * x).sum())
        gr = -2 * (x - np.array([0.3, -0.7])) - 0.5 * np.sin(5 * x)
        return (v, gr) if g else (v, None)
    single = lbd.ShardedRepeater(P, opt.Rprop(P), seed=11)(f, np.zeros(2), False)
    port = _free_port()
    script = tmp_path / "worker_r.py"
    script.write_text(WORKER_RESTARTS)
    env = dict(os.environ, OMP_NUM_THREADS="1")
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(port), str(r), "2"], stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True, env=env) for r in range(2)]
    outs = [p.communicate(timeout=120) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-2000:]
    recs = [json.loads(o.strip().splitlines()[-1]) for o, _ in outs]
    assert np.allclose(recs[0]["best"], recs[1]["best"], rtol=0, atol=0)
    assert np.allclose(recs[0]["best"], single, rtol=0, atol=1e-15)
