"""BASELINE.json config 4: N=16384, D=12, TF32 candidate scoring, EI over M candidates sharded across the ranks,
one all_gather for the argmax.  Launch alone (1 GPU) or under torchrun.  Prints one JSON line (rank 0).
usage: python tools/bench_config4.py [--m-total 1000000] [--n 16384] [--steps 2]"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m-total", type=int, default=1_000_000)
    ap.add_argument("--n", type=int, default=16384)
    ap.add_argument("--d", type=int, default=12)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--precision", default="tf32")
    a = ap.parse_args()
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from limbo_b200 import _lib, acqui, kernel, mean, model, synth
    from limbo_b200 import dist as lbd
    N, D = a.n, a.d
    X = synth.points(1234, N, D)
    y = synth.targets(X)
    lo, hi = lbd.shard_range(a.m_total, rank, world)
    Xq = synth.points(1235, a.m_total, D)[lo:hi] if a.m_total <= 2_000_000 else synth.points(1235 + rank, hi - lo, D)
    gp = model.GP(D, 1, kernel=kernel.SquaredExpARD, mean=mean.Data, device=lr, precision=a.precision)
    st = torch.cuda.Stream(dev)
    torch.cuda.set_stream(st)
    gp.set_stream(st.cuda_stream)
    t0 = time.perf_counter()
    gp.compute(X, y[:, None])
    ei = acqui.EI(gp)
    ei._update_f_max(acqui.first_elem)
    t_fit = time.perf_counter() - t0
    lib = _lib.load()
    dXq = torch.from_numpy(np.ascontiguousarray(Xq)).to(dev)
    dBest = torch.zeros(1, dtype=torch.float64, device=dev)
    dIdx = torch.zeros(1, dtype=torch.int64, device=dev)
    ap_ = np.array([ei._f_max, 0.0])
    mean_const = float(y.mean())
    M = hi - lo

    def step():
        _lib.check(lib.lb_acq_argmax_dev(gp._h, 1, ap_.ctypes.data, M, dXq.data_ptr(), None, mean_const, None, dBest.data_ptr(),
                                         dIdx.data_ptr()), "acq")
        torch.cuda.synchronize(dev)
        return lbd.allgather_argmax(float(dBest.item()), int(dIdx.item()) + lo, device=dev)

    step()  # warm-up: inverts the factor, allocates
    torch.cuda.synchronize(dev)
    PC = ["kbuild", "potf2", "trsm_panel", "syrk", "syrk_col", "trsv", "kstar", "qstep", "qreduce", "acq", "trtri", "lauum", "grad", "other"]
    lib.lb_profile_enable.argtypes = [C.c_void_p, C.c_int]
    lib.lb_profile_read.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    lib.lb_profile_enable(gp._h, 1)
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(a.steps):
        best = step()
    e1.record(st)
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / a.steps
    pms = (C.c_double * len(PC))(); pcnt = (C.c_longlong * len(PC))()
    lib.lb_profile_read(gp._h, pms, pcnt, 1)
    stage = {PC[i]: pms[i] / a.steps for i in range(len(PC)) if pcnt[i]}
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    if rank == 0:
        flops = float(a.m_total) * N * N  # M N^2 (triangular GEMM, 2 flops per MAC on N^2/2)
        print(json.dumps({"config": f"N={N}, D={D}, {a.precision}, EI over {a.m_total} candidates on {world} GPU(s)",
                          "candidates_per_s": a.m_total / (ms * 1e-3), "ms_per_batch": ms, "fit_s": t_fit, "tflops_total": flops / (ms * 1e-3) / 1e12,
                          "tflops_per_gpu": flops / (ms * 1e-3) / 1e12 / world, "best": best, "n_gpus": world, "stage_ms_rank0": stage}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
