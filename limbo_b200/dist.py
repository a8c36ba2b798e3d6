"""Multi-GPU plumbing for the parts of the path that shard: one process per GPU,
torch.distributed for the collective.

 * acquisition over a candidate batch: candidates are split into `world` contiguous ranges, every
   rank evaluates its range against its own (replicated) fitted model, and ONE collective — an
   all_gather of 16-byte (value, global index) records — yields the global argmax on every rank.
   Ties resolve to the lowest global index, like the reference's sequential scan.
 * hyper-parameter restarts (opt::ParallelRepeater): restart r runs on rank r % world; the winner is
   picked with the same record exchange.
The reference has no inter-process communication at all (tools::par is TBB shared memory only), so
there is no reference collective to mirror; NCCL is used on GPUs, gloo in the CPU tests."""
from __future__ import annotations

import numpy as np


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [lo, hi) of item indices owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_records(values: np.ndarray, indices: np.ndarray) -> tuple[float, int]:
    """argmax over (value, index) records with the lowest index winning ties; NaN never wins."""
    best_v, best_i = -np.inf, -1
    for v, i in zip(values, indices):
        if np.isnan(v) or i < 0:
            continue
        if v > best_v or (v == best_v and i < best_i) or best_i < 0:
            best_v, best_i = float(v), int(i)
    return best_v, best_i


def allgather_argmax(local_value: float, local_global_index: int, device=None, group=None) -> tuple[float, int]:
    """One collective: gather every rank's (value, global index) and reduce locally."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(local_value), int(local_global_index)
    world = dist.get_world_size(group)
    # the index travels as float64 (exact below 2^53)
    rec = torch.tensor([float(local_value), float(local_global_index)], dtype=torch.float64, device=device)
    out = [torch.zeros_like(rec) for _ in range(world)]
    dist.all_gather(out, rec, group=group)
    arr = torch.stack(out).cpu().numpy()
    return reduce_records(arr[:, 0], arr[:, 1].astype(np.int64))


def sharded_acq_argmax(acq, Xq_global: np.ndarray, rank: int, world: int, device=None, group=None) -> tuple[float, int]:
    """Evaluate `acq` (limbo_b200.acqui.UCB / EI / GP_UCB bound to this rank's model) on this rank's
    range of the global candidate batch and return the global (best value, best index) on every rank."""
    lo, hi = shard_range(len(Xq_global), rank, world)
    if hi > lo:
        v, i = acq.argmax_batch(Xq_global[lo:hi])
        i = i + lo
    else:
        v, i = -np.inf, -1
    return allgather_argmax(v, i, device=device, group=group)


def restart_owner(restart: int, world: int) -> int:
    return restart % world


class ShardedRepeater:
    """opt::ParallelRepeater (opt/parallel_repeater.hpp:76-107) with the restarts spread over the ranks:
    restart r runs on rank r % world (each rank on its own GPU and its own copy of the GP), then one
    all_gather of (value, restart index) picks the winner and its parameters are broadcast from the owner.
    Every rank draws the SAME perturbations (common seed), so the result does not depend on `world`."""

    def __init__(self, params=None, optimizer=None, seed: int = 2000, device=None, group=None):
        from . import opt as _opt
        from .params import get
        self._params = params
        self._optimizer = optimizer if optimizer is not None else _opt.Rprop(params)
        self._seed, self._device, self._group = seed, device, group
        self._get = get

    def __call__(self, f, init, bounded: bool):
        import torch
        import torch.distributed as dist
        from . import opt as _opt
        repeats = int(self._get(self._params, "opt_parallelrepeater", "repeats"))
        epsilon = float(self._get(self._params, "opt_parallelrepeater", "epsilon"))
        init = np.asarray(init, dtype=np.float64)
        on = dist.is_available() and dist.is_initialized()
        rank = dist.get_rank(self._group) if on else 0
        world = dist.get_world_size(self._group) if on else 1
        best_v, best_val, best_r = init.copy(), -float(np.finfo(np.float32).max), -1
        for r in range(repeats):
            dev = np.random.default_rng(self._seed + r).random(init.size) * 2.0 * epsilon - epsilon
            if restart_owner(r, world) != rank:
                continue
            v = self._optimizer(f, init + dev, bounded)
            val = _opt.eval(f, v)
            if val > best_val:
                best_v, best_val, best_r = v, val, r
        if not on or world == 1:
            return best_v
        val, win = allgather_argmax(best_val, best_r, device=self._device, group=self._group)
        owner = restart_owner(win, world) if win >= 0 else 0
        buf = torch.tensor(best_v if rank == owner else np.zeros_like(init), dtype=torch.float64, device=self._device)
        dist.broadcast(buf, src=owner, group=self._group)
        return buf.cpu().numpy()
