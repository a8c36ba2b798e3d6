"""Inversion of the factor spread over the ranks (limbo_b200/dist_inv.py) under torchrun: the reduced-precision variances scored
against the assembled copy of L^-1 must agree with the ones of the replicated inversion (lb_tf32_prepare) on every rank, and the
sharded EI argmax must be the unsharded one.  Also times the distributed preparation against the replicated one.  Prints one JSON
line (rank 0).
usage: torchrun --nproc-per-node G tools/dist_inv_check.py [--size 16384] [--cands 20000] [--precision fp16] [--dim 12]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=16384)
    ap.add_argument("--cands", type=int, default=20000)
    ap.add_argument("--dim", type=int, default=12)
    ap.add_argument("--precision", default="fp16")
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from limbo_b200 import _lib, acqui, dist_inv, kernel, mean, model, synth
    from limbo_b200 import dist as lbd
    lib = _lib.load()
    X = synth.points(1234, a.size, a.dim)
    y = synth.targets(X)
    Xq = synth.points(1235, a.cands, a.dim)
    gp = model.GP(a.dim, 1, kernel=kernel.SquaredExpARD, mean=mean.Data, device=lr, precision=a.precision)
    gp.compute(X, y[:, None])
    dinv = dist_inv.DistInverse(gp, rank, world, dev)
    res = {"n_gpus": world, "n": a.size, "m": a.cands, "precision": a.precision, "supported": bool(dinv.supported(gp))}
    ts = []
    for _ in range(a.reps + 1):  # the first pass is the warm-up (communicator, attributes, allocations)
        gp.recompute(False)      # a fit invalidates the reduced-precision copy
        _lib.check(lib.lb_sync(gp._h), "lb_sync")
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        dinv.prepare(gp)
        _lib.check(lib.lb_sync(gp._h), "lb_sync")
        ts.append((time.perf_counter() - t0) * 1e3)
    t = torch.tensor([min(ts[1:])], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res["dist_prepare_ms"] = float(t.item())
    best, idx = lbd.sharded_acq_argmax(acqui.EI(gp), Xq, rank, world, device=dev)
    mu_d, s2_d = gp.query_batch(Xq[:4096])
    # reference: the replicated inversion of the same model
    ref = model.GP(a.dim, 1, kernel=kernel.SquaredExpARD, mean=mean.Data, device=lr, precision=a.precision)
    ref.compute(X, y[:, None])
    ts = []
    for _ in range(a.reps + 1):
        ref.recompute(False)
        _lib.check(lib.lb_sync(ref._h), "lb_sync")
        t0 = time.perf_counter()
        ref.query_batch(Xq[:256])  # inversion + cast + one small chunk
        ts.append((time.perf_counter() - t0) * 1e3)
    res["replicated_prepare_plus_256_queries_ms"] = min(ts[1:])
    mu_r, s2_r = ref.query_batch(Xq[:4096])
    b1, i1 = acqui.EI(ref).argmax_batch(Xq)
    d = float(np.abs(s2_d - s2_r).max())
    ok = bool(np.array_equal(mu_d, mu_r) and idx == i1 and d <= 1e-6)
    flag = torch.tensor([1.0 if ok else 0.0, d], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(flag[:1], op=dist.ReduceOp.MIN)
        dist.all_reduce(flag[1:], op=dist.ReduceOp.MAX)
    res["ok_on_every_rank"] = bool(flag[0].item() == 1.0)
    res["max_abs_sigma2_diff_vs_replicated"] = float(flag[1].item())
    res["sigma2_bit_identical"] = bool(np.array_equal(s2_d, s2_r))
    res["argmax"] = {"sharded": [best, idx], "single": [b1, i1]}
    if rank == 0:
        print(json.dumps(res), flush=True)
    dinv.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
