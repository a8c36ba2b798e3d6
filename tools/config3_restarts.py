"""BASELINE.json config 3: N=16384, D=6, SquaredExpARD, fp64 - model::gp::KernelLFOpt with 20 random restarts
(opt::ParallelRepeater<Rprop>, model/gp/kernel_lf_opt.hpp:59-92, opt/parallel_repeater.hpp:76-107), a few Rprop iterations
per restart.  Alone: restarts run one after another on one GPU; under torchrun: limbo_b200.dist.ShardedRepeater spreads
them over the ranks (restart r on rank r % G) and one all_gather picks the winner.  Every objective evaluation copies the GP
(kernel_lf_opt.hpp:79) - lb_clone shares buffers copy-on-write and the pool serves the N x N buffers, so the run reports the
cudaMalloc count inside the timed region.  Prints one JSON line (rank 0).
usage: [torchrun --nproc-per-node G] python tools/config3_restarts.py [--size 16384] [--restarts 20] [--iterations 5]"""
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
    ap.add_argument("--dim", type=int, default=6)
    ap.add_argument("--restarts", type=int, default=20)
    ap.add_argument("--iterations", type=int, default=5)
    a = ap.parse_args()
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from limbo_b200 import _lib, kernel, mean, model, opt, synth
    from limbo_b200 import dist as lbd
    from limbo_b200.model.hp_opt import KernelLFOpt, _KernelLFOptimization

    class P:
        class opt_rprop:
            iterations = a.iterations
            eps_stop = 0.0

        class opt_parallelrepeater:
            repeats = a.restarts
            epsilon = 1e-2
    lib = _lib.load()
    X = synth.points(1234, a.size, a.dim)
    y = synth.targets(X)
    evals = {"n": 0}

    class Counting(_KernelLFOptimization):
        def __call__(self, params, compute_grad):
            evals["n"] += 1
            return super().__call__(params, compute_grad)
    import limbo_b200.model.hp_opt as hp
    hp._KernelLFOptimization = Counting
    inner = opt.Rprop(P)
    rep = lbd.ShardedRepeater(P, inner, seed=2000, device=dev) if world > 1 else opt.ParallelRepeater(P, inner, rng=np.random.default_rng(2000))
    gp = model.GP(a.dim, 1, params=P, kernel=kernel.SquaredExpARD, mean=mean.Data, hp_opt=KernelLFOpt(P, rep), device=lr)
    gp.compute(X, y[:, None])
    # warm-up: one evaluation with gradient (allocations, kernel attributes)
    Counting(gp)(gp.kernel_function().h_params(), True)
    evals["n"] = 0
    torch.cuda.synchronize(dev)
    if world > 1:  # communicator start-up (lazy in NCCL) stays outside the timed region
        lbd.allgather_argmax(0.0, rank, device=dev)
        wb = torch.zeros(8, dtype=torch.float64, device=dev)
        dist.broadcast(wb, src=0)
        dist.barrier()
    m0 = lib.lb_debug_pool_mallocs()
    t0 = time.perf_counter()
    gp.optimize_hyperparams()
    torch.cuda.synchronize(dev)
    wall = time.perf_counter() - t0
    mallocs = lib.lb_debug_pool_mallocs() - m0
    tt = torch.tensor([wall, float(evals["n"]), float(mallocs)], dtype=torch.float64, device=dev)
    if world > 1:
        red = [torch.zeros_like(tt) for _ in range(world)]
        dist.all_gather(red, tt)
        allv = torch.stack(red).cpu().numpy()
    else:
        allv = tt.cpu().numpy()[None]
    if rank == 0:
        wall_max = float(allv[:, 0].max())
        ev_total = int(allv[:, 1].sum())
        print(json.dumps({
            "config": f"config 3: N={a.size}, D={a.dim}, SquaredExpARD fp64, KernelLFOpt, {a.restarts} restarts x {a.iterations} Rprop iterations, {world} GPU(s)",
            "n_gpus": world, "wall_s": wall_max, "objective_evaluations_total": ev_total, "evaluations_per_s": ev_total / wall_max,
            "evaluations_per_rank": allv[:, 1].astype(int).tolist(), "cuda_mallocs_in_timed_region_per_rank": allv[:, 2].astype(int).tolist(),
            "ms_per_evaluation_rank0": 1e3 * float(allv[0, 0]) / max(1.0, float(allv[0, 1])),
            "flops_per_evaluation": float(a.size) ** 3, "tflops_per_gpu": float(allv[0, 1]) * float(a.size) ** 3 / float(allv[0, 0]) / 1e12,
            "h_params": gp.kernel_function().h_params().tolist(), "log_lik": gp.get_log_lik(),
            "parallelism": ("restarts sharded over ranks (ShardedRepeater: restart r on rank r % G), one all_gather of (value, restart) + one broadcast of the winner"
                            if world > 1 else "restarts sequential on one GPU")}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
