"""Handle memory semantics behind the reference's value-semantics GP:
 * add_sample really takes the incremental Cholesky path (gp.hpp:126-152, 573-603) - including the growth of the padded
   capacity by one 128-tile - instead of silently refitting (round-1 ADVICE finding);
 * lb_clone (the copy constructor of kernel_lf_opt.hpp:79) shares buffers copy-on-write: writes through either handle
   never show through the other, and a likelihood evaluation on a warm pool performs no cudaMalloc."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _data(n, d, seed=11):
    from limbo_b200 import synth
    X = synth.points(seed, n, d)
    return X, synth.targets(X) if d == 6 else np.cos(3 * X).sum(1)


@pytest.mark.parametrize("kname", ["SquaredExpARD", "MaternFiveHalves"])
def test_add_sample_takes_incremental_path(kname, oracle_mod):
    from limbo_b200 import kernel, mean, model
    O = oracle_mod
    X, y = _data(133, 3)
    gp = model.GP(3, 1, kernel=getattr(kernel, kname), mean=mean.Data)
    gp.compute(X[:125], y[:125, None])
    assert gp.append_count() == 0
    for i in range(125, 133):  # crosses the 128 boundary: the padded capacity grows by one tile, the factor is kept
        gp.add_sample(X[i], y[i:i + 1])
    assert gp.append_count() == 8, "add_sample fell back to a full refit"
    full = model.GP(3, 1, kernel=getattr(kernel, kname), mean=mean.Data)
    full.compute(X, y[:, None])
    Xq = _data(300, 3, seed=5)[0]
    m1, s1 = gp.query_batch(Xq)
    m2, s2 = full.query_batch(Xq)
    assert np.abs(m1 - m2).max() <= 1e-10 and np.abs(s1 - s2).max() <= 1e-10
    assert np.abs(gp.matrixL() - full.matrixL()).max() <= 1e-10
    og = O.OracleGP()  # and against the oracle's incremental path (test_gp.cpp:513-635)
    og.set_data(X, (y - y.mean())[:, None])
    kid = {"SquaredExpARD": O.K_SE_ARD, "MaternFiveHalves": O.K_MATERN52}[kname]
    og.set_kernel(kid, np.zeros(4 if kid == O.K_SE_ARD else 2), 0.01)
    og.fit()
    mo, so = og.query(Xq)
    assert np.abs(m1 - (mo + y.mean())).max() <= 1e-10 and np.abs(s1 - so).max() <= 1e-10
    # changing the h-params invalidates the factor: the next add_sample must refit, not append to a stale factor
    gp.kernel_function().set_h_params(gp.kernel_function().h_params() + 0.1)
    before = gp.append_count()
    Xn, yn = _data(1, 3, seed=77)
    gp.add_sample(Xn[0], yn[:1])
    assert gp.append_count() == before
    full.kernel_function().set_h_params(gp.kernel_function().h_params())
    full.compute(np.vstack([X, Xn]), np.concatenate([y, yn])[:, None])
    m1, s1 = gp.query_batch(Xq)
    m2, s2 = full.query_batch(Xq)
    assert np.abs(m1 - m2).max() <= 1e-10 and np.abs(s1 - s2).max() <= 1e-10


def test_clone_is_copy_on_write():
    from limbo_b200 import kernel, mean, model
    X, y = _data(300, 4)
    Xq = _data(200, 4, seed=3)[0]
    gp = model.GP(4, 1, kernel=kernel.SquaredExpARD, mean=mean.Data)
    gp.compute(X, y[:, None])
    m0, s0 = gp.query_batch(Xq)
    c = gp.copy()
    mc, sc = c.query_batch(Xq)  # the clone predicts from the shared factor
    assert np.array_equal(mc, m0) and np.array_equal(sc, s0)
    c.kernel_function().set_h_params(np.array([0.3, -0.2, 0.1, 0.0, 0.2]))
    c.recompute(False)          # writes L / alpha: the clone takes private buffers, the source is untouched
    m1, s1 = gp.query_batch(Xq)
    assert np.array_equal(m1, m0) and np.array_equal(s1, s0)
    assert np.abs(c.query_batch(Xq)[1] - s0).max() > 1e-6
    c2 = gp.copy()
    xn = np.full(4, 0.5)
    c2.add_sample(xn, np.array([1.0]))  # incremental update on a shared factor: copy, then write
    assert c2.append_count() == 1 and c2.nb_samples() == 301 and gp.nb_samples() == 300
    m2, s2 = gp.query_batch(Xq)
    assert np.array_equal(m2, m0) and np.array_equal(s2, s0)
    ref = model.GP(4, 1, kernel=kernel.SquaredExpARD, mean=mean.Data)
    ref.compute(np.vstack([X, xn]), np.append(y, 1.0)[:, None])
    assert np.abs(c2.query_batch(Xq)[1] - ref.query_batch(Xq)[1]).max() <= 1e-10
    # the source refits while a clone still holds the old factor
    keep = gp.copy()
    gp.kernel_function().set_h_params(np.array([0.1, 0.1, 0.1, 0.1, 0.0]))
    gp.recompute(False)
    mk, sk = keep.query_batch(Xq)
    assert np.array_equal(mk, m0) and np.array_equal(sk, s0)
    del c, c2, keep


def test_likelihood_evaluations_do_not_allocate(lib):
    """KernelLFOpt copies the GP per evaluation (kernel_lf_opt.hpp:79): after the first evaluation has warmed the pool,
    clone -> recompute -> log-lik -> gradient -> destroy performs no cudaMalloc."""
    from limbo_b200 import kernel, mean, model
    from limbo_b200.model.hp_opt import _KernelLFOptimization
    X, y = _data(1500, 6)
    gp = model.GP(6, 1, kernel=kernel.SquaredExpARD, mean=mean.Data)
    gp.compute(X, y[:, None])
    f = _KernelLFOptimization(gp)
    p = gp.kernel_function().h_params()
    v0, g0 = f(p, True)
    f(p + 0.01, True)
    n0 = lib.lb_debug_pool_mallocs()
    vals = []
    for k in range(6):
        v, g = f(p + 0.01 * k, True)
        vals.append(v)
    assert lib.lb_debug_pool_mallocs() == n0, "a likelihood evaluation allocated device memory"
    assert lib.lb_debug_pool_hits() > 0
    assert vals[0] == v0 and np.all(np.isfinite(vals))
