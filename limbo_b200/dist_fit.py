"""Distributed fit of ONE GP over the ranks (strong scaling of `GP::compute`, model/gp.hpp:88-116, for the headline
workload): the block-cyclic panel factorisation of config 5 (dist_chol.py) computes the factor, and every rank assembles
the COMPLETE factor in its own regular handle from the panel messages that are broadcast anyway (the message of a pair now
carries the pair's diagonal block and its two diagonal-block inverses in front of the rows below it).  After `fit()` each rank
holds exactly the state lb_fit would have produced (bit-identical factor, alpha solved locally), so prediction and acquisition
shard over the ranks with no further exchange (limbo_b200.dist.sharded_acq_argmax).

    fit time  ~  N^3 / 3 / G  of trailing update per GPU  +  the serial panel chain (potf2 -> trsm -> potf2 -> trsm -> pack ->
                 broadcast -> look-ahead update) once the per-GPU update is shorter than that chain (N = 16384: from G = 4)

There is no CPU path: one process per GPU, NCCL through torch.distributed; world == 1 degenerates to the same kernels on
one GPU (used by the single-GPU test)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .dist_chol import PAIR, TILE, DistCholesky, schedule

HEAD = 256 * 256 + 2 * 128 * 128  # LB_DCHOL_HEAD


class DistFit(DistCholesky):
    """fit(gp) for a limbo_b200.model.GP whose data and kernel are already in place (gp.compute(..., compute_kernel=False)
    or a previous compute): all ranks call it with the same samples and hyper-parameters."""

    def __init__(self, gp, rank: int, world: int, device, group=None):
        X = gp._sample_matrix()
        super().__init__(X, gp.kernel_function(), rank, world, device, group=group)
        import torch
        lib = self._lib
        vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
        lib.lb_dchol_pack_head.argtypes = [vp, vp, i64, i32, vp, vp]
        lib.lb_dchol_adopt_begin.argtypes = [vp, i64]
        lib.lb_dchol_unpack.argtypes = [vp, vp, i64, i32, vp]
        lib.lb_dchol_adopt_end.argtypes = [vp, i32]
        for f in ("lb_dchol_pack_head", "lb_dchol_adopt_begin", "lb_dchol_unpack", "lb_dchol_adopt_end"):
            getattr(lib, f).restype = i32
        # messages: head + rows below (double buffered), replacing the plain panels of the base class
        self.msgs = [torch.empty(HEAD + PAIR * max(self.Nd - PAIR, 1), dtype=torch.float64, device=self.device) for _ in range(2)]
        self.panels = [m[HEAD:] for m in self.msgs]
        self.aux = torch.cuda.Stream(self.device)  # unpacking runs beside the trailing update

    def supported(self, gp) -> bool:
        """the handle's padded order (multiple of 128) must equal the distributed one (multiple of 256)"""
        return gp.nb_samples() == self.N and max(TILE, (self.N + TILE - 1) // TILE * TILE) == self.Nd

    def set_kernel(self, kernel_fn) -> None:
        """new hyper-parameters for the column generator (the target handle gets them through gp._push_kernel())"""
        own = np.ascontiguousarray(kernel_fn.params(), dtype=np.float64)
        _lib.check(self._lib.lb_set_kernel(self._h_main, kernel_fn.kernel_id, own.ctypes.data, own.size, kernel_fn.noise()), "lb_set_kernel")

    def set_points(self, X: np.ndarray) -> None:
        """new samples for the column generator (same N and D as at construction)"""
        Xc = np.ascontiguousarray(X, dtype=np.float64)
        assert Xc.shape == (self.N, self.D)
        _lib.check(self._lib.lb_dchol_set_points(self._h_main, self.N, self.D, Xc.ctypes.data), "lb_dchol_set_points")

    def fit(self, gp, push: bool = True) -> int:
        """Distributed K -> L -> alpha for `gp` (every rank ends with the complete fitted model).  Returns the LAPACK-style
        info, or -5 when the handle's padded order differs from the distributed one (the caller then uses the replicated
        lb_fit).  push = False: samples / obs_mean / kernel are already on the device (lb_set_data[_dev] + lb_set_kernel done by
        the caller, generator points unchanged since construction)."""
        import torch.distributed as dist
        torch = self._torch
        lib = self._lib
        Ev = torch.cuda.Event
        h = gp._h
        if push:
            gp._push_data()
            gp._push_kernel()
            self.set_points(gp._sample_matrix())
        self.set_kernel(gp.kernel_function())
        # the target handle works on this object's main stream for the duration of the fit
        prev_stream = getattr(gp, "_stream_ptr", 0)
        _lib.check(lib.lb_set_stream(h, self.main.cuda_stream), "lb_set_stream")
        rc = lib.lb_dchol_adopt_begin(h, self.Nd)
        if rc == -5:  # padded orders differ (N mod 256 in (0, 128]): the caller falls back to the replicated lb_fit
            _lib.check(lib.lb_set_stream(h, C.c_void_p(prev_stream)), "lb_set_stream")
            return rc
        _lib.check(rc, "lb_dchol_adopt_begin")
        self.build()
        ev_a, ev_panel, ev_bcast = Ev(), Ev(), Ev()
        ev_free = [None, None]   # message buffer last read by an update
        ev_unpk = [None, None]   # ... and by an unpack
        with torch.cuda.stream(self.main):
            self.info.zero_()
        ev_a.record(self.main)
        self.aux.wait_event(ev_a)
        have_a = True
        for act in schedule(self.npairs, self.rank, self.world):
            kind, p = act[0], act[1]
            msg = self.msgs[p % 2]
            if kind == "panel":
                lp = p // self.world
                cols = self.L.data_ptr() + 8 * lp * PAIR * self.Nd
                if have_a:
                    self.side.wait_event(ev_a)
                for e in (ev_free[p % 2], ev_unpk[p % 2]):
                    if e is not None:
                        self.side.wait_event(e)
                _lib.check(lib.lb_dchol_panel(self._h_side, cols, self.Nd, 2 * p, self.invD.data_ptr(), self.info.data_ptr(),
                                              msg.data_ptr() + 8 * HEAD), "lb_dchol_panel")
                _lib.check(lib.lb_dchol_pack_head(self._h_side, cols, self.Nd, 2 * p, self.invD.data_ptr(), msg.data_ptr()), "lb_dchol_pack_head")
                ev_panel.record(self.side)
                if p == self.npairs - 1:  # the last pair has nothing below it and is not broadcast by the schedule: ship its head
                    self._bcast_last(p, msg, ev_panel, h)
            elif kind == "bcast":
                owner = act[2]
                n = HEAD + PAIR * self._ldp(p)
                with torch.cuda.stream(self.comm):
                    if owner == self.rank:
                        self.comm.wait_event(ev_panel)
                    else:
                        for e in (ev_free[p % 2], ev_unpk[p % 2]):
                            if e is not None:
                                self.comm.wait_event(e)
                    if self.world > 1:
                        dist.broadcast(msg[:n], src=owner, group=self.group)
                    ev_bcast.record(self.comm)
                self.main.wait_event(ev_bcast)
                self.aux.wait_event(ev_bcast)
                _lib.check(lib.lb_dchol_unpack(h, msg.data_ptr(), self.Nd, 2 * p, C.c_void_p(self.aux.cuda_stream)), "lb_dchol_unpack")
                e = Ev()
                e.record(self.aux)
                ev_unpk[p % 2] = e
            else:
                _, _, l0, l1, tag = act
                _lib.check(lib.lb_dchol_update(self._h_main, self.L.data_ptr(), self.Nd, self.panels[p % 2].data_ptr(), 2 * p, l0, l1,
                                               self.rank, self.world), "lb_dchol_update")
                if tag == "a":
                    ev_a.record(self.main)
                e = Ev()
                e.record(self.main)
                ev_free[p % 2] = e
        if self.npairs - 1 not in self.pairs and self.world > 1:
            self._bcast_last(self.npairs - 1, self.msgs[(self.npairs - 1) % 2], None, h, receive_only=True, ev_free=ev_free, ev_unpk=ev_unpk)
        self.main.wait_stream(self.side)
        self.main.wait_stream(self.comm)
        self.main.wait_stream(self.aux)
        with torch.cuda.stream(self.main):
            if self.world > 1:  # one small collective: the first failing pivot, if any
                inf = self.info[:1].to(torch.int64)
                out = [torch.zeros_like(inf) for _ in range(self.world)]
                dist.all_gather(out, inf, group=self.group)
                infos = torch.cat(out)
            else:
                infos = self.info[:1].to(torch.int64)
        self.main.synchronize()
        bad = infos[infos > 0]
        info = int(bad.min().item()) if bad.numel() else 0
        rc = lib.lb_dchol_adopt_end(h, info)
        _lib.check(lib.lb_set_stream(h, C.c_void_p(prev_stream)), "lb_set_stream")  # synchronises the fit's stream first
        if rc < 0:
            _lib.check(rc, "lb_dchol_adopt_end")
        gp._chol_info = rc
        gp._inv_kernel_updated = False
        return rc

    def _bcast_last(self, p, msg, ev_panel, h, receive_only=False, ev_free=None, ev_unpk=None):
        """The last pair only has a head (diagonal block + inverses): one extra small broadcast so that every rank's factor is
        complete."""
        import torch.distributed as dist
        torch = self._torch
        owner = p % self.world
        with torch.cuda.stream(self.comm):
            if not receive_only:
                self.comm.wait_event(ev_panel)
            else:
                for e in ((ev_free or [None, None])[p % 2], (ev_unpk or [None, None])[p % 2]):
                    if e is not None:
                        self.comm.wait_event(e)
            if self.world > 1:
                dist.broadcast(msg[:HEAD], src=owner, group=self.group)
            ev = torch.cuda.Event()
            ev.record(self.comm)
        self.aux.wait_event(ev)
        _lib.check(self._lib.lb_dchol_unpack(h, msg.data_ptr(), self.Nd, 2 * p, C.c_void_p(self.aux.cuda_stream)), "lb_dchol_unpack")
