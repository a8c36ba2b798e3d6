"""Latency sweep in the regime Limbo is normally run in and benchmarked in by the reference itself (N in {50, ..., 600},
learning time and 10^4 SEQUENTIAL query() calls: waf_tools/benchmark_template.cpp:95-120, regression_benchmarks.json):
device-timed fit and 10^4-candidate UCB argmax, host-visible compute() / add_sample() / one-point query() latency (the
one-launch path: query_point_kernel) per N, with the CPU restatement (oracle: 1-thread fit, 1-thread one-point queries, the
reference's protocol) beside it.
usage: python tools/small_n_sweep.py [--sizes 50,100,...] [--m 10000]"""
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
    ap.add_argument("--sizes", default="50,100,200,300,400,500,600,1024,2048,4096")
    ap.add_argument("--m", type=int, default=10000)
    ap.add_argument("--dim", type=int, default=6)
    ap.add_argument("--cpu", type=int, default=1)
    a = ap.parse_args()
    from limbo_b200 import _lib, acqui, kernel, mean, model, synth
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    st = torch.cuda.Stream(dev)
    torch.cuda.set_stream(st)
    D, M = a.dim, a.m
    Xq = synth.points(1235, M, D)
    dXq = torch.from_numpy(Xq).to(dev)
    dBest = torch.zeros(1, dtype=torch.float64, device=dev)
    dIdx = torch.zeros(1, dtype=torch.int64, device=dev)
    ap_ = np.array([0.5, 0.0])
    rows = []
    for N in [int(s) for s in a.sizes.split(",")]:
        X = synth.points(1234, N + 1, D)
        y = synth.targets(X)
        gp = model.GP(D, 1, kernel=kernel.SquaredExpARD, mean=mean.Data)
        gp.set_stream(st.cuda_stream)
        gp.compute(X[:N], y[:N, None])

        def timed(fn, reps):
            fn()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(reps):
                fn()
            e1.record(st)
            torch.cuda.synchronize(dev)
            return e0.elapsed_time(e1) / reps
        fit_ms = timed(lambda: _lib.check(lib.lb_fit_async(gp._h), "fit"), 20)
        q_ms = timed(lambda: _lib.check(lib.lb_acq_argmax_dev(gp._h, 0, ap_.ctypes.data, M, dXq.data_ptr(), None, float(y[:N].mean()), None,
                                                                dBest.data_ptr(), dIdx.data_ptr()), "acq"), 10)
        # host-visible latencies through the public API (what a BO iteration pays)
        t0 = time.perf_counter(); gp.compute(X[:N], y[:N, None]); t_fit_api = (time.perf_counter() - t0) * 1e3
        t0 = time.perf_counter(); gp.add_sample(X[N], y[N:N + 1]); t_add_api = (time.perf_counter() - t0) * 1e3
        t0 = time.perf_counter(); acqui.UCB(gp).argmax_batch(Xq); t_acq_api = (time.perf_counter() - t0) * 1e3
        gp.query(Xq[0])
        nq = 300
        t0 = time.perf_counter()
        for i in range(nq):
            gp.query(Xq[i])  # sequential one-point calls through the public API, like the reference benchmark
        t_q1 = (time.perf_counter() - t0) / nq * 1e6
        t0 = time.perf_counter(); gp.query_batch(Xq); t_qb = (time.perf_counter() - t0) * 1e3
        row = {"N": N, "fit_ms_dev": fit_ms, "acq10k_ms_dev": q_ms, "compute_ms_api": t_fit_api, "add_sample_ms_api": t_add_api,
               "acq10k_ms_api": t_acq_api, "query1_us_api_sequential": t_q1, "query10k_sequential_ms_extrapolated": t_q1 * M / 1e3,
               "query10k_batched_ms_api": t_qb}
        if a.cpu:
            from oracle import oracle as O
            og = O.OracleGP()
            og.set_data(X[:N], (y[:N] - y[:N].mean())[:, None])
            og.set_kernel(0, np.zeros(D + 1), 0.01)
            t0 = time.perf_counter(); og.fit(); row["cpu_fit_ms_1thread"] = (time.perf_counter() - t0) * 1e3
            mq = min(M, 2000)
            t0 = time.perf_counter(); og.query(Xq[:mq], nthreads=os.cpu_count()); row["cpu_query10k_ms_allcores"] = (time.perf_counter() - t0) * 1e3 * M / mq
            m1 = min(M, 300)
            t0 = time.perf_counter(); og.query(Xq[:m1], nthreads=1); row["cpu_query1_us_1thread"] = (time.perf_counter() - t0) * 1e6 / m1
        rows.append(row)
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
