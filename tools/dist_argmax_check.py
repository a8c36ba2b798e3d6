"""Sharded acquisition argmax under torchrun (limbo_b200.dist.sharded_acq_argmax: contiguous candidate ranges, one
all_gather of (value, global index) records) against the unsharded device argmax on rank 0.  Prints one JSON line (rank 0).
usage: torchrun --nproc-per-node G tools/dist_argmax_check.py [--size 1500] [--cands 20001]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=1500)
    ap.add_argument("--cands", type=int, default=20001)
    ap.add_argument("--dim", type=int, default=6)
    a = ap.parse_args()
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from limbo_b200 import acqui, kernel, mean, model, synth
    from limbo_b200 import dist as lbd
    X = synth.points(1234, a.size, a.dim)
    y = synth.targets(X)
    Xq = synth.points(99, a.cands, a.dim)
    Xq[a.cands // 3] = Xq[a.cands - 5]  # a duplicated candidate in two different shards: ties resolve to the lowest global index
    res = {}
    for name, mk in (("UCB", lambda g: acqui.UCB(g)), ("EI", lambda g: acqui.EI(g))):
        gp = model.GP(a.dim, 1, kernel=kernel.SquaredExpARD, mean=mean.Data, device=lr)
        gp.compute(X, y[:, None])
        v, i = lbd.sharded_acq_argmax(mk(gp), Xq, rank, world, device=dev)
        if rank == 0:
            v1, i1, vals = mk(gp).argmax_batch(Xq, return_values=True)
            res[name] = {"sharded": [v, i], "single": [v1, i1], "same": bool(v == v1 and i == i1 and i1 == int(np.argmax(vals)))}
    if rank == 0:
        print(json.dumps({"n_gpus": world, "n": a.size, "m": a.cands, "results": res, "ok": all(r["same"] for r in res.values())}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
