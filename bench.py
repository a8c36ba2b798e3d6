#!/usr/bin/env python
"""bench.py — GP fit + batched query throughput on B200 (BASELINE.json's metric).

One "step" = one pass of the hot path over one batch of synthetic input:
    fit  : K-build (N x N) -> blocked Cholesky -> alpha            (model/gp.hpp:550-571)
    query: M UCB candidates: K*, L^-1 K*, mu, sigma^2, UCB, argmax (gp.hpp:159-167, acqui/ucb.hpp:83-90)
at N = 16384, D = 6, SquaredExpARD, fp64, M = 10000 per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

`value`  : steps/s with inputs already resident in HBM (device-pointer ABI), CUDA events.
`e2e`    : the same step through the public host API (limbo_b200.model.GP.compute +
           acqui.UCB.argmax_batch): host buffers in, host scalars out, copies inside the timed region.
`roofline`: dominant kernel (Cholesky trailing update, fp64 DMMA) timed live with CUDA events around
           every launch; `roofline_kbuild` is the HBM-bound K-build kernel.
`cpu_baseline` / `--impl reference`: the CPU restatement of the reference path (oracle/, "port":
           the reference itself needs Eigen/Boost/TBB which this image lacks) on the host cores, on a
           bounded sample extrapolated to the full size (stated in `sample`).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_TRAIN, DIM, M_CAND = 16384, 6, 10000
KERNEL_NAME = "SquaredExpARD"
WORKLOADS = {  # --workload: BASELINE.json configs that fit one GPU; the default is the configuration the metric is quoted on
    "n16384_se_ard": (16384, 6, 10000, "SquaredExpARD"),
    "config2_n4096_matern": (4096, 6, 10000, "MaternFiveHalves"),
}
NOISE = 0.01
UCB_ALPHA = 0.5
METRIC = "GP fit+query/s at N=16384,D=6 (fit = K-build+Cholesky+alpha, query = 10k UCB candidates+argmax, fp64)"
UNIT = "fit+query/s"


def workload_config(n_gpus: int) -> dict:
    return {
        "workload": f"N={N_TRAIN}, D={DIM}, {KERNEL_NAME}, fp64, fit + {M_CAND} UCB queries + argmax per GPU",
        "n_train": N_TRAIN, "dim": DIM, "kernel": KERNEL_NAME, "noise": NOISE,
        "candidates_per_gpu": M_CAND, "global_candidates": M_CAND * n_gpus,
        "parallelism": f"fit replicated per GPU, candidates sharded x{n_gpus}, one all_gather for the argmax" if n_gpus > 1 else "single GPU",
        "l2": "inputs larger than L2 (factor 2.1 GB, K* 1.3 GB vs 126 MB L2); no explicit flush",
        "hyperparams": "reference defaults: log ell_d = 0, log sigma_f = 0, noise 0.01, UCB alpha 0.5, mean::Data",
        "seeds": {"data": 1234, "candidates": 1235},
    }


# --------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi during the timed region)
# --------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines: list[str] = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------
# CPU arm: the oracle port on host cores, bounded sample, extrapolated
# --------------------------------------------------------------------------------------
def cpu_sample(n_s: int, m_s: int, threads: int) -> dict:
    """One bounded sample: fit at N=n_s (single thread, as Eigen's LLT in the reference) and
    m_s sequential-per-thread queries fanned over `threads` host threads (tools::par)."""
    from limbo_b200 import synth
    from oracle import oracle as O
    X = synth.points(1234, n_s, DIM)
    y = synth.targets(X)
    Xq = synth.points(1235, m_s, DIM)
    g = O.OracleGP()
    g.set_data(X, (y - y.mean())[:, None])
    kid = {"SquaredExpARD": O.K_SE_ARD, "MaternFiveHalves": O.K_MATERN52}[KERNEL_NAME]
    g.set_kernel(kid, np.zeros(DIM + 1 if kid == O.K_SE_ARD else 2), NOISE)
    t0 = time.perf_counter()
    g.fit()
    t_fit = time.perf_counter() - t0
    t0 = time.perf_counter()
    mu, s2 = g.query(Xq, nthreads=threads)
    ucb = O.ucb(mu[:, 0] + y.mean(), s2, UCB_ALPHA)
    _ = int(np.argmax(ucb))
    t_q = time.perf_counter() - t0
    return {"t_fit": t_fit, "t_query": t_q}


def cpu_lapack_sample(n_s: int = 8192, m_s: int = 2048) -> dict | None:
    """NOT the reference's code path (which factors with single-threaded Eigen::LLT and predicts one point at a time): what a
    tuned host library does with the same mathematics - LAPACK dpotrf + ONE batched dtrsm over all candidates (scipy /
    OpenBLAS, all host cores).  Reported next to cpu_baseline so the GPU / CPU ratio can be read against a strong CPU too."""
    try:
        import scipy.linalg as sl
    except Exception:
        return None
    from limbo_b200 import synth
    X = synth.points(1234, n_s, DIM)
    y = synth.targets(X)
    Xq = synth.points(1235, m_s, DIM)
    if KERNEL_NAME != "SquaredExpARD":
        return None
    t0 = time.perf_counter()
    sq = (X ** 2).sum(1)
    K = np.exp(-0.5 * np.maximum(sq[:, None] + sq[None, :] - 2.0 * (X @ X.T), 0.0))
    K[np.diag_indices(n_s)] += NOISE + 1e-8
    t_k = time.perf_counter() - t0
    t0 = time.perf_counter()
    L = sl.cholesky(K, lower=True, overwrite_a=True, check_finite=False)
    alpha = sl.cho_solve((L, True), y - y.mean(), check_finite=False)
    t_fit = time.perf_counter() - t0
    t0 = time.perf_counter()
    sqq = (Xq ** 2).sum(1)
    Ks = np.exp(-0.5 * np.maximum(sq[:, None] + sqq[None, :] - 2.0 * (X @ Xq.T), 0.0))
    V = sl.solve_triangular(L, Ks, lower=True, check_finite=False)
    s2 = np.maximum(1.0 - (V ** 2).sum(0), 0.0) + NOISE
    ucb = Ks.T @ alpha + y.mean() + UCB_ALPHA * np.sqrt(s2)
    _ = int(np.argmax(ucb))
    t_q = time.perf_counter() - t0
    sec = t_k * (N_TRAIN / n_s) ** 2 + t_fit * (N_TRAIN / n_s) ** 3 + t_q * (M_CAND / m_s) * (N_TRAIN / n_s) ** 2
    return {"value": 1.0 / sec, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "host LAPACK, batched (not the reference's algorithm)",
            "sample": (f"numpy K build ({t_k:.2f} s) + scipy/OpenBLAS dpotrf ({t_fit:.2f} s) at N={n_s} + one dtrsm over {m_s} candidates "
                       f"({t_q:.2f} s), extrapolated to N={N_TRAIN}, M={M_CAND} by N^2 / N^3 / M*N^2 -> {sec:.1f} s per step")}


def cpu_extrapolate(s: dict, n_s: int, m_s: int) -> float:
    """fit ~ N^3, query ~ M N^2 -> seconds for the full workload."""
    t_fit = s["t_fit"] * (N_TRAIN / n_s) ** 3
    t_q = s["t_query"] * (M_CAND / m_s) * (N_TRAIN / n_s) ** 2
    return t_fit + t_q


def pick_cpu_sample(budget_s: float, threads: int):
    cal = cpu_sample(512, 64, threads)
    for n_s in (4096, 3072, 2048, 1536, 1024):
        est = cal["t_fit"] * (n_s / 512) ** 3 + cal["t_query"] * (256 / 64) * (n_s / 512) ** 2
        if est <= budget_s:
            return n_s, 256
    return 1024, 128


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    steps, warmup = args.steps, args.warmup
    per_step = max(1.0, min(12.0, 150.0 / max(1, steps + warmup)))
    n_s, m_s = pick_cpu_sample(per_step, threads)
    for _ in range(warmup):
        cpu_sample(n_s, m_s, threads)
    times = []
    t_all0 = time.perf_counter()
    for _ in range(steps):
        s = cpu_sample(n_s, m_s, threads)
        times.append(cpu_extrapolate(s, n_s, m_s))
    wall = time.perf_counter() - t_all0
    sec = float(np.mean(times))
    value = 1.0 / sec
    sample = (f"per step: oracle fit at N={n_s} (1 thread, like Eigen::LLT) + {m_s} queries over {threads} threads; "
              f"extrapolated to N={N_TRAIN}, M={M_CAND} by N^3 (fit) and M*N^2 (query); measured {wall / max(1, steps):.2f} s/sample")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (splitmix64 U[0,1)^6, Hartmann6 targets)", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------
PC = ["kbuild", "potf2", "trsm_panel", "syrk", "syrk_col", "trsv", "kstar", "qstep", "qreduce", "acq", "trtri", "lauum", "grad", "other"]


def syrk_flops_per_fit(n: int) -> float:
    """Algorithmic flops of the K=256 trailing updates (profile class "syrk") of one factorisation:
    panels are taken in pairs (k, k+1); the update of the matrix right of the pair has order
    m = (T - k - 2) * 128 and costs m (m + 1) * 256 flops (lower triangle incl. diagonal, 2 flops per MAC)."""
    nb, tot = 128, 0.0
    t = (n + nb - 1) // nb
    for k in range(0, t, 2):
        m = max(0, (t - k - 2)) * nb
        tot += float(m) * (m + 1) * (2 * nb)
    return tot


def fp64_tensor_peak() -> tuple[float, str]:
    p = os.path.join(ROOT, "profiles", "fp64_peaks.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            return float(j["dmma_tflops"]), f"measured DMMA peak ({j.get('source', 'tools/microbench.cu')})"
        except Exception:
            pass
    return 40.0, "nominal B200 fp64 tensor peak (40 TFLOP/s); MEASURED_PEAKS.json has no fp64 figure"


def ncu_traffic(kernel: str):
    """DRAM bytes of one captured launch (ncu --set full), from the committed profile summary."""
    try:
        j = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))[kernel]
        return j["dram_bytes"], j
    except Exception:
        return None, None


def hbm_peak() -> tuple[float, str]:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (limbo_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    from limbo_b200 import _lib, acqui, kernel, mean, model, synth
    from limbo_b200 import dist as lbdist

    lib = _lib.load()
    lib.lb_profile_enable.argtypes = [C.c_void_p, C.c_int]
    lib.lb_profile_read.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]

    steps, warmup = args.steps, max(args.warmup, 3)
    n, d, m = N_TRAIN, DIM, M_CAND
    X = synth.points(1234, n, d)
    y = synth.targets(X)
    Xq = synth.points(1235 + rank, m, d)  # each rank owns its shard of the global candidate batch

    kcls = getattr(kernel, KERNEL_NAME)
    kid_dev = kcls.kernel_id
    gp = model.GP(d, 1, kernel=kcls, mean=mean.Data, device=local_rank)
    h = gp._h
    # a real (non-default) stream: the library and the timing events must share it
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    gp.set_stream(stream.cuda_stream)

    # ---------------- device-resident leg ("value") ----------------
    dX = torch.from_numpy(X).to(dev)
    om = (y - y.mean())
    dY = torch.from_numpy(om).to(dev)
    dXq = torch.from_numpy(Xq).to(dev)
    dBest = torch.zeros(1, dtype=torch.float64, device=dev)
    dIdx = torch.zeros(1, dtype=torch.int64, device=dev)
    ap = np.array([UCB_ALPHA, 0.0])
    hp = np.zeros(d + 1 if KERNEL_NAME == "SquaredExpARD" else 2)
    mean_const = float(y.mean())
    gather_buf = [torch.zeros(2, dtype=torch.float64, device=dev) for _ in range(world)] if world > 1 else None

    def step_dev():
        _lib.check(lib.lb_set_data_dev(h, n, d, 1, dX.data_ptr(), dY.data_ptr()), "set_data_dev")
        _lib.check(lib.lb_set_kernel(h, kid_dev, hp.ctypes.data, hp.size, NOISE), "set_kernel")
        _lib.check(lib.lb_fit_async(h), "fit_async")
        _lib.check(lib.lb_acq_argmax_dev(h, 0, ap.ctypes.data, m, dXq.data_ptr(), None, mean_const, None, dBest.data_ptr(),
                                         dIdx.data_ptr()), "acq_argmax_dev")
        if world > 1:  # one collective: (value, global index) records, reduced locally
            rec = torch.stack([dBest[0], (dIdx[0] + rank * m).to(torch.float64)])
            dist.all_gather(gather_buf, rec)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for _ in range(warmup):
        step_dev()
    sync_all()
    _lib.check(lib.lb_check_info(h), "cholesky info")

    lib.lb_profile_enable(h, 1)
    ms0 = (C.c_double * len(PC))()
    cnt0 = (C.c_longlong * len(PC))()
    lib.lb_profile_read(h, ms0, cnt0, 1)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = gp.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record(stream)
    for _ in range(steps):
        step_dev()
    e1.record(stream)
    sync_all()
    t_ms = e0.elapsed_time(e1)
    launches = gp.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms = (C.c_double * len(PC))()
    cnt = (C.c_longlong * len(PC))()
    lib.lb_profile_read(h, ms, cnt, 1)
    lib.lb_profile_enable(h, 0)
    prof = {PC[i]: {"ms_total": ms[i], "launches": int(cnt[i])} for i in range(len(PC)) if cnt[i]}
    tt = torch.tensor([t_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_ms = float(tt.item())
    ms_per_step = t_ms / steps
    value = world / (ms_per_step * 1e-3)

    # ---------------- end-to-end leg through the public host API ----------------
    gp2 = model.GP(d, 1, kernel=kcls, mean=mean.Data, device=local_rank)
    gp2.set_stream(stream.cuda_stream)
    Xp = torch.from_numpy(X).pin_memory()
    yp = torch.from_numpy(y[:, None].copy()).pin_memory()
    Xqp = torch.from_numpy(Xq).pin_memory()
    Xl, yl = Xp.numpy(), yp.numpy()  # pinned host buffers, one point per row
    ucb = acqui.UCB(gp2)

    def step_e2e():
        gp2.compute(Xl, yl)                     # host samples/observations in, H2D inside
        best, idx = ucb.argmax_batch(Xqp.numpy())  # host candidates in, (value, index) out
        if world > 1:  # the one collective: (value, global index) records -> global argmax on every rank
            best, idx = lbdist.allgather_argmax(best, idx + rank * m, device=dev)
        return best, idx

    e2e_steps = max(2, min(steps, 5))
    step_e2e()
    sync_all()
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(e2e_steps):
        step_e2e()
    e1.record(stream)
    sync_all()
    t_e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)
    tt = torch.tensor([t_e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e_value = world / (float(tt.item()) / e2e_steps * 1e-3)
    h2d = n * d * 8 + n * 8 + m * d * 8 + 2 * 8
    d2h = 8 + 8 + 8

    if rank == 0:
        peak_t, peak_t_src = fp64_tensor_peak()
        peak_h, peak_h_src = hbm_peak()
        roof = {}
        if "syrk" in prof:
            per_launch_flops = syrk_flops_per_fit(n) * steps / prof["syrk"]["launches"]
            avg_ms = prof["syrk"]["ms_total"] / prof["syrk"]["launches"]
            ach = per_launch_flops / (avg_ms * 1e-3) / 1e12
            roof = {"kernel": "syrk_kernel (Cholesky trailing update, fp64 DMMA)", "bound": "tensor", "achieved": ach, "peak": peak_t,
                    "unit": "TFLOP/s", "frac": ach / peak_t, "traffic": ncu_traffic("syrk_kernel")[0],
                    "traffic_capture": ncu_traffic("syrk_kernel")[1], "peak_source": peak_t_src,
                    "share_of_step": prof["syrk"]["ms_total"] / t_ms, "avg_launch_ms": avg_ms,
                    "algorithmic_flops_per_launch": per_launch_flops}
        roof_k = {}
        if "kbuild" in prof:
            byts = 8.0 * n * n + 8.0 * n * d
            avg_ms = prof["kbuild"]["ms_total"] / prof["kbuild"]["launches"]
            ach = byts / (avg_ms * 1e-3) / 1e9
            roof_k = {"kernel": "kbuild_kernel (N x N kernel matrix)", "bound": "hbm", "achieved": ach, "peak": peak_h, "unit": "GB/s",
                      "frac": ach / peak_h, "traffic": ncu_traffic("kbuild_kernel")[0], "peak_source": peak_h_src, "avg_launch_ms": avg_ms,
                      "algorithmic_bytes_per_launch": byts}
        # CPU baseline on a bounded sample (rank 0, N=1 only)
        cpu = None
        if world == 1 and not args.no_cpu:
            threads = os.cpu_count() or 1
            n_s, m_s = pick_cpu_sample(12.0, threads)
            s = cpu_sample(n_s, m_s, threads)
            sec = cpu_extrapolate(s, n_s, m_s)
            cpu = {"value": 1.0 / sec, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": (f"oracle fit at N={n_s} (1 thread: {s['t_fit']:.2f} s) + {m_s} queries over {threads} threads "
                              f"({s['t_query']:.2f} s), extrapolated to N={n}, M={m} by N^3 / M*N^2 -> {sec:.0f} s per step")}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (splitmix64 U[0,1)^6, Hartmann6 targets, random-free deterministic)",
            "config": workload_config(world), "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "api": "limbo_b200.model.GP.compute + acqui.UCB.argmax_batch (host buffers)"},
            "gpu_launches": int(launches), "roofline": roof, "roofline_kbuild": roof_k, "cpu_baseline": cpu,
            "cpu_lapack_batched": (cpu_lapack_sample() if (world == 1 and not args.no_cpu) else None),
            "stage_ms_per_step": {k: v["ms_total"] / steps for k, v in prof.items()},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline sample")
    ap.add_argument("--workload", default="n16384_se_ard", choices=sorted(WORKLOADS))
    args = ap.parse_args()
    global N_TRAIN, DIM, M_CAND, KERNEL_NAME, METRIC
    N_TRAIN, DIM, M_CAND, KERNEL_NAME = WORKLOADS[args.workload]
    if args.workload != "n16384_se_ard":
        METRIC = f"GP fit+query/s at N={N_TRAIN},D={DIM} ({KERNEL_NAME}; fit = K-build+Cholesky+alpha, query = {M_CAND} UCB candidates+argmax, fp64)"
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
