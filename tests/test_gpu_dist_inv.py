"""Inversion of the factor by column tiles (lb_dinv_* in include/limbo_b200_dist.h, limbo_b200/dist_inv.py) on ONE GPU: the C ABI is
driven for G = 1, 2, 3 ranks chunk by chunk on the same device, the chunks are concatenated as the all_gather would, and the
reduced-precision variances scored against the assembled copy of L^-1 (model/gp.hpp:618-624) are compared with those of the
replicated inversion (lb_tf32_prepare) and with the fp64 path.  The NCCL version runs in tests/test_gpu_multirank.py."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _models(n, d, precision):
    from limbo_b200 import kernel, mean, model, synth
    X = synth.points(77, n, d)
    y = synth.targets(X)
    a = model.GP(d, 1, kernel=kernel.SquaredExpARD, mean=mean.Data, precision=precision)
    b = model.GP(d, 1, kernel=kernel.SquaredExpARD, mean=mean.Data, precision=precision)
    a.compute(X, y[:, None])
    b.compute(X, y[:, None])
    return a, b, synth.points(78, 3000, d)


def _assemble(gp, G):
    """what DistInverse.prepare does on G ranks, on one device"""
    import torch
    from limbo_b200 import _lib, dist_inv
    lib = dist_inv.bind(_lib.load())
    h = gp._h
    nbytes = int(lib.lb_dinv_chunk_bytes(h, G))
    assert nbytes > 0
    mx = C.c_double(0.0)
    amax = 0.0
    for r in range(G):
        _lib.check(lib.lb_dinv_columns(h, r, G, C.byref(mx)), "lb_dinv_columns")
        amax = max(amax, mx.value)
    everything = torch.empty(G * nbytes, dtype=torch.uint8, device="cuda:0")
    for r in range(G):
        _lib.check(lib.lb_dinv_columns(h, r, G, C.byref(mx)), "lb_dinv_columns")
        _lib.check(lib.lb_dinv_pack(h, r, G, amax, everything.data_ptr() + r * nbytes), "lb_dinv_pack")
    _lib.check(lib.lb_dinv_adopt(h, G, everything.data_ptr(), amax), "lb_dinv_adopt")
    _lib.check(lib.lb_sync(h), "lb_sync")
    return everything


@pytest.mark.parametrize("precision,tol", [("tf32", 1e-6), ("fp16", 1e-6), ("fp16x3", 1e-7)])
@pytest.mark.parametrize("n,G", [(700, 1), (2500, 2), (2500, 3), (4352, 8)])
def test_column_inverse_matches_replicated_inverse(precision, tol, n, G):
    a, b, Xq = _models(n, 5, precision)
    keep = _assemble(a, G)
    mu_a, s2_a = a.query_batch(Xq)
    mu_b, s2_b = b.query_batch(Xq)  # lb_tf32_prepare: recursive inverse of the whole factor + cast
    assert np.array_equal(mu_a, mu_b)
    assert np.abs(s2_a - s2_b).max() <= tol
    del keep


def test_column_inverse_against_fp64_variances():
    from limbo_b200 import kernel, mean, model, synth
    n, d = 2500, 5
    a, _, Xq = _models(n, d, "fp16x3")
    X = synth.points(77, n, d)
    y = synth.targets(X)
    ref = model.GP(d, 1, kernel=kernel.SquaredExpARD, mean=mean.Data)
    ref.compute(X, y[:, None])
    _assemble(a, 3)
    s2 = a.query_batch(Xq)[1]
    s2_ref = ref.query_batch(Xq)[1]
    assert np.abs(s2 - s2_ref).max() <= 2e-5  # the split mode's bar at N <= 4096 (tests/test_gpu_tf32.py)


def test_dist_inverse_world_one_and_refit_invalidates():
    from limbo_b200 import dist_inv
    a, b, Xq = _models(1500, 4, "fp16")
    di = dist_inv.DistInverse(a, 0, 1, "cuda:0")
    assert di.supported(a)
    di.prepare(a)
    s2_a = a.query_batch(Xq)[1]
    s2_b = b.query_batch(Xq)[1]
    assert np.abs(s2_a - s2_b).max() <= 1e-6
    # new hyper-parameters: the copy is invalidated by the fit and rebuilt (here by the replicated path)
    for g in (a, b):
        g.kernel_function().set_h_params(g.kernel_function().h_params() - 0.3)
        g.recompute(False)
    assert np.abs(a.query_batch(Xq)[1] - b.query_batch(Xq)[1]).max() <= 1e-6
    di.prepare(a)
    assert np.abs(a.query_batch(Xq)[1] - b.query_batch(Xq)[1]).max() <= 1e-6
    di.close()


def test_dinv_argument_checks():
    from limbo_b200 import _lib, dist_inv, kernel, mean, model, synth
    lib = dist_inv.bind(_lib.load())
    X = synth.points(1, 300, 3)
    g64 = model.GP(3, 1, kernel=kernel.SquaredExpARD, mean=mean.Data)
    g64.compute(X, synth.targets(X)[:, None])
    assert lib.lb_dinv_chunk_bytes(g64._h, 2) == -5  # LB_ERR_UNSUPPORTED: fp64 handles score in fp64
    g16 = model.GP(3, 1, kernel=kernel.SquaredExpARD, mean=mean.Data, precision="fp16")
    mx = C.c_double(0.0)
    assert lib.lb_dinv_columns(g16._h, 0, 2, C.byref(mx)) == -3  # LB_ERR_STATE: not fitted
    g16.compute(X, synth.targets(X)[:, None])
    assert lib.lb_dinv_columns(g16._h, 2, 2, C.byref(mx)) == -1  # LB_ERR_ARG: rank out of range
    g17 = model.GP(3, 1, kernel=kernel.SquaredExpARD, mean=mean.Data, precision="fp16")
    g17.compute(X, synth.targets(X)[:, None])
    assert lib.lb_dinv_pack(g17._h, 0, 2, 1.0, 1) == -3          # LB_ERR_STATE: lb_dinv_columns first
