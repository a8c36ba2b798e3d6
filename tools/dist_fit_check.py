"""Distributed fit of one GP (limbo_b200/dist_fit.py) under torchrun: every rank must end with the factor lb_fit produces
(bit-identical L, alpha, predictions), and the sharded acquisition on top of it must pick the unsharded argmax.  Also times the
distributed fit against the replicated one.  Prints one JSON line (rank 0).
usage: torchrun --nproc-per-node G tools/dist_fit_check.py [--size 16384] [--cands 10000] [--kernel SquaredExpARD]"""
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
    ap.add_argument("--cands", type=int, default=10000)
    ap.add_argument("--dim", type=int, default=6)
    ap.add_argument("--kernel", default="SquaredExpARD")
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from limbo_b200 import acqui, dist_fit, kernel, mean, model, synth
    from limbo_b200 import dist as lbd
    X = synth.points(1234, a.size, a.dim)
    y = synth.targets(X)
    Xq = synth.points(1235, a.cands, a.dim)
    kcls = getattr(kernel, a.kernel)
    st = torch.cuda.Stream(dev)
    torch.cuda.set_stream(st)
    gp = model.GP(a.dim, 1, kernel=kcls, mean=mean.Data, device=lr)
    gp.set_stream(st.cuda_stream)
    gp.compute(X, y[:, None], compute_kernel=False)
    fitter = dist_fit.DistFit(gp, rank, world, dev)
    res = {"n_gpus": world, "n": a.size, "m": a.cands, "kernel": a.kernel, "supported": bool(fitter.supported(gp))}
    info = fitter.fit(gp)  # warm-up (communicator, attributes, allocations)
    res["info"] = info
    ts = []
    for _ in range(a.reps):
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        fitter.fit(gp)
        torch.cuda.synchronize(dev)
        ts.append((time.perf_counter() - t0) * 1e3)
    t = torch.tensor([min(ts)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res["dist_fit_ms"] = float(t.item())
    best, idx = lbd.sharded_acq_argmax(acqui.UCB(gp), Xq, rank, world, device=dev)
    mu_d, s2_d = gp.query_batch(Xq[:2000])
    L_d, A_d = (gp.matrixL(), gp.alpha()) if a.size <= 8192 else (None, gp.alpha())
    # reference: the replicated single-GPU fit of the same model
    ref = model.GP(a.dim, 1, kernel=kcls, mean=mean.Data, device=lr)
    ref.set_stream(st.cuda_stream)
    ref.compute(X, y[:, None])
    ts = []
    for _ in range(a.reps):
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        ref.recompute(False)
        torch.cuda.synchronize(dev)
        ts.append((time.perf_counter() - t0) * 1e3)
    res["replicated_fit_ms"] = min(ts)
    mu_r, s2_r = ref.query_batch(Xq[:2000])
    b1, i1 = acqui.UCB(ref).argmax_batch(Xq)
    ok = bool(np.array_equal(A_d, ref.alpha()) and np.array_equal(mu_d, mu_r) and np.array_equal(s2_d, s2_r) and best == b1 and idx == i1)
    if L_d is not None:
        ok = ok and bool(np.array_equal(L_d, ref.matrixL()))
    flag = torch.tensor([1.0 if ok else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    res["bit_identical_on_every_rank"] = bool(flag.item() == 1.0)
    res["argmax"] = {"sharded": [best, idx], "single": [b1, i1]}
    res["loglik_rel_diff"] = abs(gp.compute_log_lik() - ref.compute_log_lik()) / abs(ref.compute_log_lik())
    if rank == 0:
        print(json.dumps(res))
    fitter.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
