"""world_size-2 gloo test of the sharded-argmax host logic (no GPU): each rank owns a contiguous
candidate range, evaluates a stand-in acquisition, and the single all_gather collective returns the
same global (value, index) on both ranks — equal to the sequential scan, lowest index on ties."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

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


# ---- multi-GPU Cholesky schedule (limbo_b200/dist_chol.py), executed with NumPy over gloo ---------------------------
WORKER_DCHOL = r"""
import sys, json
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from limbo_b200 import dist_chol as dc
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{sys.argv[2]}", rank=int(sys.argv[3]), world_size=int(sys.argv[4]))
rank, world = dist.get_rank(), dist.get_world_size()
N = int(sys.argv[5])
Nd = dc.padded_order(N); T = Nd // dc.TILE; npairs = T // 2
rng = np.random.default_rng(3)
B = rng.standard_normal((N, N + 5))
K = np.eye(Nd); K[:N, :N] = B @ B.T / N + 0.5 * np.eye(N)          # SPD, identity in the padding (as the device build)
pairs = dc.local_pairs(npairs, rank, world)
cols = np.concatenate([np.arange(dc.global_block(l, rank, world) * dc.TILE, (dc.global_block(l, rank, world) + 1) * dc.TILE)
                       for l in range(2 * len(pairs))]) if pairs else np.zeros(0, dtype=int)
L = K[:, cols].copy()                                               # this rank's columns, full height
panels = [None, None]
log = []
for act in dc.schedule(npairs, rank, world):
    kind, p = act[0], act[1]
    r0 = (2 * p + 2) * dc.TILE
    if kind == "panel":
        lp = p // world
        c = slice(lp * dc.PAIR, (lp + 1) * dc.PAIR)
        d0 = 2 * p * dc.TILE
        Ld = np.linalg.cholesky(L[d0:d0 + dc.PAIR, c])
        L[d0:d0 + dc.PAIR, c] = Ld
        L[r0:, c] = np.linalg.solve(Ld, L[r0:, c].T).T
        panels[p % 2] = L[r0:, c].copy()
    elif kind == "bcast":
        t = torch.from_numpy(panels[p % 2] if act[2] == rank else np.empty((Nd - r0, dc.PAIR)))
        dist.broadcast(t, src=act[2])
        panels[p % 2] = t.numpy()
    else:
        _, _, l0, l1, tag = act
        P = panels[p % 2]
        for l in range(l0, l1):
            j = dc.global_block(l, rank, world)
            assert j > 2 * p + 1
            rows = slice(j * dc.TILE, Nd)
            L[rows, l * dc.TILE:(l + 1) * dc.TILE] -= P[j * dc.TILE - r0:, :] @ P[j * dc.TILE - r0:(j + 1) * dc.TILE - r0, :].T
    log.append(act[0])
ref = np.linalg.cholesky(K)
err = 0.0
for ci, j in enumerate(cols):
    err = max(err, float(np.abs(L[j:, ci] - ref[j:, j]).max()))
print(json.dumps({"rank": rank, "err": err, "ncols": int(cols.size), "n_panel": log.count("panel"), "n_bcast": log.count("bcast")}))
dist.destroy_process_group()
"""


def test_dist_cholesky_schedule_structure():
    """Every pair is factored exactly once by its owner, every panel but the last is broadcast once per rank, and every
    local block column right of a panel is updated exactly once by it (a and b parts are disjoint)."""
    sys.path.insert(0, ROOT)
    from limbo_b200 import dist_chol as dc
    for world in (1, 2, 3, 8):
        for npairs in (1, 2, 5, 16):
            panels = []
            for rank in range(world):
                acts = list(dc.schedule(npairs, rank, world))
                panels += [a[1] for a in acts if a[0] == "panel"]
                assert [a[1] for a in acts if a[0] == "bcast"] == list(range(npairs - 1))
                nlb = 2 * len(dc.local_pairs(npairs, rank, world))
                for p in range(npairs - 1):
                    upd = [a for a in acts if a[0] == "update" and a[1] == p]
                    touched = [l for a in upd for l in range(a[2], a[3])]
                    want = [l for l in range(nlb) if dc.global_block(l, rank, world) > 2 * p + 1]
                    assert sorted(touched) == want and len(set(touched)) == len(touched)
                    # look-ahead: the a part (columns of pair p + 1) comes first on its owner
                    if (p + 1) % world == rank:
                        assert upd[0][4] == "a" and [dc.global_block(l, rank, world) for l in range(upd[0][2], upd[0][3])] == [2 * p + 2, 2 * p + 3]
                # a panel is issued only after the a-part update of the previous step
                order = [(a[0], a[1]) for a in acts]
                for p in [a[1] for a in acts if a[0] == "panel" and a[1] > 0]:
                    assert order.index(("update", p - 1)) < order.index(("panel", p))
            assert sorted(panels) == list(range(npairs))


def test_dist_cholesky_world2_gloo(tmp_path):
    """The schedule executed with NumPy on 2 gloo ranks (ragged N: 700 -> 3 pairs) reproduces numpy.linalg.cholesky."""
    import json
    port = _free_port()
    script = tmp_path / "worker_dchol.py"
    script.write_text(WORKER_DCHOL)
    env = dict(os.environ, OMP_NUM_THREADS="2")
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(port), str(r), "2", "700"], stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True, env=env) for r in range(2)]
    outs = [p.communicate(timeout=180) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-2000:]
    recs = [json.loads(o.strip().splitlines()[-1]) for o, _ in outs]
    assert all(r["err"] < 1e-11 for r in recs), recs
    assert sum(r["ncols"] for r in recs) == 768 and sum(r["n_panel"] for r in recs) == 3
    assert all(r["n_bcast"] == 2 for r in recs)


# ---- inversion of the factor by column tiles (limbo_b200/dist_inv.py): chunk layout + all_gather + un-permutation, executed with
# NumPy over gloo (the CUDA side is tests/test_gpu_dist_inv.py and tests/test_gpu_multirank.py) ------------------------------------
WORKER_DINV = r"""
import sys, json
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from limbo_b200 import dist_inv
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{sys.argv[2]}", rank=int(sys.argv[3]), world_size=int(sys.argv[4]))
rank, world = dist.get_rank(), dist.get_world_size()
precision = sys.argv[5]
TILE = 128
N = 5 * TILE
rng = np.random.default_rng(3)
A = rng.standard_normal((N, N)); K = A @ A.T / N + np.eye(N)
L = np.linalg.cholesky(K)
lay = dist_inv.chunk_layout(N, world, precision)
W, T = lay["width"], N // TILE
mine = dist_inv.owned_tiles(T, rank, world)
# this rank's columns of L^-1: forward solves of identity columns (the columns of a triangular inverse are independent)
from scipy.linalg import solve_triangular
V = np.zeros((N, W))
for t, c in enumerate(mine):
    E = np.zeros((N, TILE)); E[c * TILE:(c + 1) * TILE] = np.eye(TILE)
    V[:, t * TILE:(t + 1) * TILE] = solve_triangular(L, E, lower=True)
amax = torch.tensor([np.abs(V).max()], dtype=torch.float64)
dist.all_reduce(amax, op=dist.ReduceOp.MAX)                      # the fp16 scale needs the maximum over all ranks
scale = 1.0 if precision == "tf32" else 2.0 ** (14 - int(np.ceil(np.log2(amax.item()))))
chunk = np.zeros(lay["chunk_bytes"], dtype=np.uint8)
et = np.float32 if precision == "tf32" else np.float16
hi = (V * scale).astype(et)
chunk[:lay["plane_bytes"]] = hi.view(np.uint8).reshape(-1)
if lay["planes"] == 2:
    lo = ((V * scale - hi.astype(np.float64)) * 2048.0).astype(np.float16)
    chunk[lay["plane_bytes"]:2 * lay["plane_bytes"]] = lo.view(np.uint8).reshape(-1)
chunk[lay["weights_offset"]:] = (V ** 2).sum(0).view(np.uint8)
everything = torch.zeros(world * lay["chunk_bytes"], dtype=torch.uint8)
dist.all_gather_into_tensor(everything, torch.from_numpy(chunk))
allb = everything.numpy()
# un-permute (lb_dinv_adopt): global tile c lives in the chunk of rank c mod world at local tile c // world
out = np.zeros((N, N), dtype=et); w = np.zeros(N)
for c in range(T):
    base = (c % world) * lay["chunk_bytes"]
    plane = allb[base:base + lay["plane_bytes"]].view(et).reshape(N, W)
    out[:, c * TILE:(c + 1) * TILE] = plane[:, (c // world) * TILE:(c // world + 1) * TILE]
    ww = allb[base + lay["weights_offset"]:base + lay["chunk_bytes"]].view(np.float64)
    w[c * TILE:(c + 1) * TILE] = ww[(c // world) * TILE:(c // world + 1) * TILE]
Linv = np.linalg.inv(L)
ref = (np.tril(Linv) * scale).astype(et)
err = float(np.abs(out.astype(np.float64) - ref.astype(np.float64)).max()) / float(np.abs(ref).max())
werr = float(np.abs(w - (np.tril(Linv) ** 2).sum(0)).max())
print(json.dumps({"rank": rank, "rel_err": err, "w_err": werr, "tiles": mine, "scale": scale, "upper_zero": bool((np.triu(out.astype(np.float64), 1) == 0).all())}))
dist.destroy_process_group()
"""


def test_chunk_layout_and_tile_ownership():
    from limbo_b200 import dist_inv
    for T in (1, 5, 8, 128):
        for w in (1, 2, 3, 8):
            tiles = [dist_inv.owned_tiles(T, r, w) for r in range(w)]
            assert sorted(c for ts in tiles for c in ts) == list(range(T))          # a partition of the column tiles
            assert all(c % w == r for r, ts in enumerate(tiles) for c in ts)
            lay = dist_inv.chunk_layout(T * 128, w, "fp16")
            assert lay["width"] == 128 * max(len(ts) for ts in tiles)
    lay = dist_inv.chunk_layout(16384, 4, "fp16x3")
    assert lay == {"width": 4096, "elem_bytes": 2, "planes": 2, "plane_bytes": 16384 * 4096 * 2, "weights_offset": 2 * 16384 * 4096 * 2,
                   "chunk_bytes": 2 * 16384 * 4096 * 2 + 8 * 4096}
    assert dist_inv.chunk_layout(16384, 2, "tf32")["chunk_bytes"] == 16384 * 8192 * 4 + 8 * 8192



@pytest.mark.parametrize("precision", ["fp16", "fp16x3", "tf32"])
def test_column_tile_inverse_world2_gloo(tmp_path, precision):
    """every rank inverts its column tiles, one all_gather, un-permutation: the assembled reduced-precision L^-1 equals the cast of
    the whole inverse (up to the last bit of the storage type: the two fp64 results differ by rounding before the cast)"""
    import json
    port = _free_port()
    script = tmp_path / "worker_dinv.py"
    script.write_text(WORKER_DINV)
    env = dict(os.environ, OMP_NUM_THREADS="1")
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(port), str(r), "2", precision], stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True, env=env) for r in range(2)]
    outs = [p.communicate(timeout=180) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-2000:]
    recs = [json.loads(o.strip().splitlines()[-1]) for o, _ in outs]
    assert recs[0]["tiles"] == [0, 2, 4] and recs[1]["tiles"] == [1, 3]
    for r in recs:
        assert r["upper_zero"] and r["w_err"] <= 1e-9
        assert r["rel_err"] <= (2.0 ** -10 if precision != "tf32" else 2.0 ** -22), r  # at most one unit in the last place of the storage type
    assert recs[0]["scale"] == recs[1]["scale"]
