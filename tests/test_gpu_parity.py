"""GPU parity: the CUDA path (through the C ABI / the limbo_b200 GP mirror) against the
CPU oracle on identical seeded inputs.  fp64 bar from BASELINE.json: |d| <= 1e-10 on
K, mu, sigma^2; alpha / log-lik / gradient are checked relative (they scale with
cond(K) between any two correct fp64 orderings) AND against the
long-double oracle so a miss can be attributed."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL_ABS = 1e-10
KERNELS = ["SquaredExpARD", "MaternFiveHalves", "MaternThreeHalves", "Exp"]


def _make(kname, N, D, P=1, noise=0.01, hp=None, seed=1234):
    from limbo_b200 import kernel, mean, model, synth
    from oracle import oracle as O

    class Prm:
        class kernel:
            pass
    Prm.kernel.noise = noise
    X = synth.points(seed, N, D)
    y = synth.targets(X)
    Y = np.stack([y * (p + 1) + 0.1 * p for p in range(P)], axis=1)
    kcls = getattr(kernel, kname)
    gp = model.GP(D, P, params=Prm, kernel=kcls, mean=mean.Data)
    if hp is not None:
        gp.kernel_function().set_h_params(np.asarray(hp, dtype=float))
    gp.compute(list(X), list(Y))
    kid = {"SquaredExpARD": O.K_SE_ARD, "MaternFiveHalves": O.K_MATERN52, "MaternThreeHalves": O.K_MATERN32, "Exp": O.K_EXP}[kname]
    og = O.OracleGP()
    og.set_data(X, Y - Y.mean(axis=0))
    og.set_kernel(kid, gp.kernel_function().params(), noise)
    og.fit()
    return gp, og, X, Y


@pytest.mark.parametrize("kname", KERNELS)
@pytest.mark.parametrize("N,D", [(3, 1), (8, 2), (50, 1), (129, 6), (300, 6), (640, 12)])
def test_fit_matches_oracle(kname, N, D, oracle_mod):
    gp, og, X, Y = _make(kname, N, D)
    K, L, A = gp.kernel_matrix(), gp.matrixL(), gp.alpha()
    Ko, Lo, Ao = og.get(0), og.get(1), og.get(2)
    assert np.abs(K - Ko).max() <= TOL_ABS
    assert np.allclose(K, K.T, rtol=0, atol=0)  # mirrored exactly (gp.hpp:560-562)
    assert np.abs(L - Lo).max() <= 1e-9 * max(1.0, np.abs(Lo).max())
    assert np.all(np.triu(L, 1) == 0.0)
    rel_a = np.abs(A - Ao).max() / np.abs(Ao).max()
    assert rel_a <= 1e-10, rel_a
    # residual form: K alpha = obs_mean
    assert np.abs(Ko @ A - (Y - Y.mean(axis=0))).max() <= 1e-9


@pytest.mark.parametrize("kname", KERNELS)
@pytest.mark.parametrize("N,D,M", [(3, 1, 5), (50, 1, 257), (300, 6, 1000), (640, 12, 130)])
def test_query_matches_oracle(kname, N, D, M, oracle_mod):
    from limbo_b200 import synth
    gp, og, X, Y = _make(kname, N, D)
    Xq = synth.points(1235, M, D)
    mu, s2 = gp.query_batch(Xq)
    mu_o, s2_o = og.query(Xq)
    mu_o = mu_o + Y.mean(axis=0)
    assert np.abs(mu - mu_o).max() <= TOL_ABS
    assert np.abs(s2 - s2_o).max() <= TOL_ABS
    # single-point API agrees with the batch bit for bit (test_gp.cpp:506-507: mu(v) == query(v).mu)
    m1, s1 = gp.query(Xq[0])
    # batches of >= 256 candidates take the panel path, a single point the slab kernel: same mathematics, different summation
    # order -> equal to rounding; below 256 the batch runs on the slab kernel too and the values are bit-identical
    if M >= 256:
        assert np.abs(m1 - mu[0]).max() <= 1e-12 * max(1.0, np.abs(mu).max()) and abs(s1 - s2[0]) <= 1e-12
    else:
        assert np.array_equal(m1, mu[0]) and s1 == s2[0]
    assert np.array_equal(gp.mu(Xq[0]), m1) and gp.sigma(Xq[0]) == s1


def test_query_at_training_points_clamps(oracle_mod):
    # tiny noise: sigma^2 at a training point underflows the DBL_EPSILON clamp (gp.hpp:623)
    gp, og, X, Y = _make("SquaredExpARD", 20, 2, noise=1e-12)
    mu, s2 = gp.query_batch(X)
    mu_o, s2_o = og.query(X)
    assert np.abs(s2 - s2_o).max() <= TOL_ABS
    assert np.abs(mu - (mu_o + Y.mean(axis=0))).max() <= 1e-6


def test_multi_output(oracle_mod):
    from limbo_b200 import synth
    gp, og, X, Y = _make("SquaredExpARD", 200, 6, P=3)
    Xq = synth.points(77, 300, 6)
    mu, s2 = gp.query_batch(Xq)
    mu_o, s2_o = og.query(Xq)
    assert mu.shape == (300, 3)
    assert np.abs(mu - (mu_o + Y.mean(axis=0))).max() <= TOL_ABS
    assert np.abs(s2 - s2_o).max() <= TOL_ABS
    assert abs(gp.compute_log_lik() - og.log_lik()) <= 1e-11 * abs(og.log_lik())
    g, go = gp.compute_kernel_grad_log_lik(), og.grad()
    assert np.abs(g - go).max() <= 1e-9 * np.abs(go).max()


@pytest.mark.parametrize("kname", KERNELS)
@pytest.mark.parametrize("N,D", [(40, 4), (300, 6), (513, 3)])
def test_loglik_and_gradient(kname, N, D, oracle_mod):
    rng = np.random.default_rng(5)
    nh = D + 1 if kname == "SquaredExpARD" else 2
    hp = rng.uniform(-0.5, 0.5, nh)
    gp, og, X, Y = _make(kname, N, D, hp=hp)
    ll, llo = gp.compute_log_lik(), og.log_lik()
    assert abs(ll - llo) <= 1e-11 * abs(llo)
    g, go = gp.compute_kernel_grad_log_lik(), og.grad()
    assert np.abs(g - go).max() <= 1e-9 * max(1.0, np.abs(go).max())
    assert gp.inv_kernel_computed()
    Kinv, Kinv_o = gp.inv_kernel(), og.get(3)
    assert np.abs(Kinv - Kinv_o).max() <= 1e-9 * np.abs(Kinv_o).max()
    assert np.allclose(Kinv, Kinv.T, rtol=0, atol=0)


def test_gradient_with_noise_param(oracle_mod):
    from limbo_b200 import kernel, mean, model, synth
    O = oracle_mod

    class Prm:
        class kernel:
            noise = 0.02
            optimize_noise = True
    X = synth.points(3, 150, 4)
    y = synth.targets(X)[:, None]
    gp = model.GP(4, 1, params=Prm, kernel=kernel.SquaredExpARD, mean=mean.Data)
    gp.compute(list(X), list(y))
    g = gp.compute_kernel_grad_log_lik()
    assert g.size == 6
    og = O.OracleGP()
    og.set_data(X, y - y.mean())
    og.set_kernel(O.K_SE_ARD, np.zeros(5), 0.02)
    og.fit()
    go = og.grad(optimize_noise=True)
    assert np.abs(g - go).max() <= 1e-9 * np.abs(go).max()


def test_add_sample_matches_full_and_oracle(oracle_mod):
    """test_gp.cpp:513-635: incremental Cholesky vs recompute vs fresh compute."""
    from limbo_b200 import kernel, mean, model, synth
    O = oracle_mod
    N0, D, extra = 120, 3, 20  # crosses the 128 padding boundary
    X = synth.points(9, N0 + extra, D)
    y = synth.targets(X)[:, None]
    gp = model.GP(D, 1, kernel=kernel.MaternFiveHalves, mean=mean.Data)
    gp.compute(list(X[:N0]), list(y[:N0]))
    og = O.OracleGP()
    og.set_data(X[:N0], y[:N0] - y[:N0].mean())
    og.set_kernel(O.K_MATERN52, gp.kernel_function().params(), 0.01)
    og.fit()
    for i in range(N0, N0 + extra):
        gp.add_sample(X[i], y[i])
        og.append(X[i], y[: i + 1] - y[: i + 1].mean())
    gp2 = model.GP(D, 1, kernel=kernel.MaternFiveHalves, mean=mean.Data)
    gp2.compute(list(X), list(y))
    L, L2, Lo = gp.matrixL(), gp2.matrixL(), og.get(1)
    assert np.abs(L - L2).max() <= 1e-9
    assert np.abs(L - Lo).max() <= 1e-9
    Xq = synth.points(10, 64, D)
    m1, s1 = gp.query_batch(Xq)
    m2, s2 = gp2.query_batch(Xq)
    mo, so = og.query(Xq)
    assert np.abs(m1 - m2).max() <= TOL_ABS and np.abs(s1 - s2).max() <= TOL_ABS
    assert np.abs(m1 - (mo + y.mean())).max() <= TOL_ABS and np.abs(s1 - so).max() <= TOL_ABS


def test_add_sample_from_empty_and_prior(oracle_mod):
    """gp.hpp:161-163 prior path, test_gp.cpp:697-758 prior variance."""
    from limbo_b200 import kernel, mean, model

    class Prm:
        class kernel_squared_exp_ard:
            sigma_sq = 10.0
    gp = model.GP(2, 1, params=Prm, kernel=kernel.SquaredExpARD, mean=mean.NullFunction)
    mu, s2 = gp.query(np.array([0.3, 0.4]))
    assert mu[0] == 0.0 and abs(s2 - (10.0 + 0.01)) <= 1e-12
    gp.add_sample(np.array([0.1, 0.2]), np.array([1.0]))
    gp.add_sample(np.array([0.5, 0.6]), np.array([2.0]))
    assert gp.nb_samples() == 2
    mu, s2 = gp.query(np.array([0.1, 0.2]))
    assert abs(mu[0] - 1.0) < 0.1


def test_inv_kernel_flag_state_machine():
    """test_gp.cpp:382-446"""
    from limbo_b200 import kernel, mean, model, synth
    X = synth.points(2, 30, 2)
    y = synth.targets(X)[:, None]
    gp = model.GP(2, 1, kernel=kernel.SquaredExpARD, mean=mean.Data)
    gp.compute(list(X), list(y))
    assert not gp.inv_kernel_computed()
    gp.compute_kernel_grad_log_lik()
    assert gp.inv_kernel_computed()
    gp.recompute(True, True)
    assert not gp.inv_kernel_computed()
    gp.compute_inv_kernel()
    assert gp.inv_kernel_computed()
    gp.add_sample(np.array([0.5, 0.5]), np.array([0.2]))
    assert not gp.inv_kernel_computed()


def test_not_positive_definite_reports_pivot(lib):
    """LAPACK-style info instead of the reference's silent NaNs."""
    import ctypes as C
    from limbo_b200 import _lib
    h = C.c_void_p()
    _lib.check(lib.lb_create(C.byref(h), 0, 0), "create")
    X = np.zeros((4, 2))  # four identical points and zero noise -> singular K
    Y = np.zeros((4, 1))
    _lib.check(lib.lb_set_data(h, 4, 2, 1, X.ctypes.data, Y.ctypes.data), "set_data")
    hp = np.zeros(3)
    _lib.check(lib.lb_set_kernel(h, 0, hp.ctypes.data, 3, -1e-8), "set_kernel")
    rc = lib.lb_fit(h)
    assert rc == 2, rc
    lib.lb_destroy(h)


def test_abi_argument_errors(lib):
    import ctypes as C
    h = C.c_void_p()
    assert lib.lb_create(C.byref(h), 0, 7) == -5
    assert lib.lb_create(C.byref(h), 99, 0) == -1
    assert lib.lb_create(C.byref(h), 0, 0) == 0
    assert lib.lb_fit(h) == -3
    hp = np.zeros(3)
    assert lib.lb_set_kernel(h, 0, hp.ctypes.data, 3, 0.01) == -3  # no data yet
    X = np.zeros((4, 2)); Y = np.zeros((4, 1))
    assert lib.lb_set_data(h, 4, 2, 1, X.ctypes.data, Y.ctypes.data) == 0
    assert lib.lb_set_kernel(h, 0, hp.ctypes.data, 2, 0.01) == -1
    assert lib.lb_set_kernel(h, 9, hp.ctypes.data, 3, 0.01) == -5
    out = C.c_double()
    assert lib.lb_log_lik(h, C.addressof(out)) == -3
    lib.lb_destroy(h)


def test_acquisitions_match_oracle(oracle_mod):
    from limbo_b200 import acqui, synth
    O = oracle_mod
    gp, og, X, Y = _make("MaternFiveHalves", 300, 6)
    Xq = synth.points(1235, 5000, 6)
    mu_o, s2_o = og.query(Xq)
    mu_o = mu_o[:, 0] + Y.mean()
    best, idx, vals = acqui.UCB(gp).argmax_batch(Xq, return_values=True)
    ref = O.ucb(mu_o, s2_o, 0.5)
    assert np.abs(vals - ref).max() <= TOL_ABS
    assert idx == int(np.argmax(ref)) and abs(best - ref.max()) <= TOL_ABS
    # scalar path == reference contract (ucb.hpp:83-90)
    v0 = acqui.UCB(gp)(Xq[17])[0]
    assert abs(v0 - ref[17]) <= TOL_ABS
    # EI: f_max over the training points (ei.hpp:100-108)
    mu_tr, _ = og.query(X)
    f_max = float((mu_tr[:, 0] + Y.mean()).max())
    ei = acqui.EI(gp)
    best, idx, vals = ei.argmax_batch(Xq, return_values=True)
    ref = O.ei(mu_o, s2_o, f_max, 0.0)
    assert abs(ei._f_max - f_max) <= TOL_ABS
    assert np.abs(vals - ref).max() <= TOL_ABS
    assert abs(vals[idx] - ref.max()) <= TOL_ABS
    assert abs(ei(Xq[3])[0] - ref[3]) <= TOL_ABS


def test_kernel_lf_opt_rprop_matches_oracle(oracle_mod):
    """KernelLFOpt + Rprop (kernel_lf_opt.hpp:59-92, rprop.hpp:84-144): same trajectory as the oracle."""
    from limbo_b200 import kernel, mean, model, opt, synth
    O = oracle_mod

    class Prm:
        class opt_rprop:
            iterations = 12
            eps_stop = 0.0
    X = synth.points(21, 200, 3)
    y = synth.targets(X)[:, None]
    gp = model.GP(3, 1, params=Prm, kernel=kernel.SquaredExpARD, mean=mean.Data, hp_opt=model.KernelLFOpt(Prm, opt.Rprop(Prm)))
    gp.compute(list(X), list(y))
    gp.optimize_hyperparams()
    og = O.OracleGP()
    og.set_data(X, y - y.mean())
    og.set_kernel(O.K_SE_ARD, np.zeros(4), 0.01)
    og.fit()
    best, ne = og.rprop_lml(np.zeros(4), 12)
    assert ne == 12
    assert np.abs(gp.kernel_function().h_params() - best).max() <= 1e-9
    ll_o, _ = og.lml_eval(best, False)
    assert abs(gp.get_log_lik() - ll_o) <= 1e-10 * abs(ll_o)


def test_concurrent_const_queries(oracle_mod):
    """query() is const and called from several host threads in the reference (SURVEY.md §8b threading)."""
    import threading
    from limbo_b200 import synth
    gp, og, X, Y = _make("SquaredExpARD", 200, 6)
    Xq = synth.points(99, 400, 6)
    ref = [gp.query_batch(Xq[i * 100:(i + 1) * 100]) for i in range(4)]  # same batch shape -> same kernel path as the threads
    ref_mu, ref_s2 = np.concatenate([r[0] for r in ref]), np.concatenate([r[1] for r in ref])
    errs = []

    def work(lo, hi):
        for _ in range(5):
            m, s = gp.query_batch(Xq[lo:hi])
            if not (np.array_equal(m, ref_mu[lo:hi]) and np.array_equal(s, ref_s2[lo:hi])):
                errs.append((lo, hi))
    th = [threading.Thread(target=work, args=(i * 100, (i + 1) * 100)) for i in range(4)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs


def test_fused_and_multilaunch_query_paths_agree(oracle_mod, lib):
    """The fused persistent slab kernel (D <= 16, P <= 4) and the multi-launch blocked TRSM path
    (any D / P) both match the oracle; large D / P route to the latter automatically."""
    import ctypes as C
    from limbo_b200 import synth
    lib.lb_debug_force_unfused_query.argtypes = [C.c_void_p, C.c_int]
    gp, og, X, Y = _make("SquaredExpARD", 700, 6)
    Xq = synth.points(5, 1111, 6)
    mu_p, s2_p = gp.query_batch(Xq)  # >= 256 candidates: panel path
    lib.lb_debug_set_query_panel_min(1 << 40)
    try:
        mu_f, s2_f = gp.query_batch(Xq)  # fused slab kernel
    finally:
        lib.lb_debug_set_query_panel_min(0)
    lib.lb_debug_force_unfused_query(gp._h, 1)
    mu_u, s2_u = gp.query_batch(Xq)  # multi-launch blocked TRSM
    lib.lb_debug_force_unfused_query(gp._h, 0)
    mu_o, s2_o = og.query(Xq)
    mu_o = mu_o + Y.mean(axis=0)
    for mu, s2 in ((mu_p, s2_p), (mu_f, s2_f), (mu_u, s2_u)):
        assert np.abs(mu - mu_o).max() <= TOL_ABS and np.abs(s2 - s2_o).max() <= TOL_ABS
    assert not np.array_equal(s2_p, s2_f) or np.array_equal(mu_p, mu_f)  # different kernels really ran (or agree exactly)
    # D = 20 (> 16) and P = 5 (> 4) take the multi-launch path
    gp, og, X, Y = _make("MaternFiveHalves", 260, 20, P=5)
    Xq = synth.points(6, 333, 20)
    mu, s2 = gp.query_batch(Xq)
    mu_o, s2_o = og.query(Xq)
    assert np.abs(mu - (mu_o + Y.mean(axis=0))).max() <= TOL_ABS and np.abs(s2 - s2_o).max() <= TOL_ABS


def test_query_slab_nine_tiles(oracle_mod):
    """M = 10000 candidates -> slabs of 8 and 9 n8-tiles per CTA (the benchmark shape).  Regression for the B-stage
    loader that skipped the 9th tile (caught by tests/test_gpu_fullsize.py)."""
    from limbo_b200 import _lib, synth
    gp, og, X, Y = _make("SquaredExpARD", 300, 6)
    Xq = synth.points(4321, 10000, 6)
    lib = _lib.load()
    lib.lb_debug_set_query_panel_min(1 << 40)  # keep this batch on the slab kernel (large batches default to the panel path)
    try:
        mu, s2 = gp.query_batch(Xq)
    finally:
        lib.lb_debug_set_query_panel_min(0)
    mu_o, s2_o = og.query(Xq, nthreads=8)
    assert np.abs(mu - (mu_o + Y.mean(axis=0))).max() <= TOL_ABS
    assert np.abs(s2 - s2_o).max() <= TOL_ABS
    mu_p, s2_p = gp.query_batch(Xq)  # and the panel path on the same batch
    assert np.abs(mu_p - (mu_o + Y.mean(axis=0))).max() <= TOL_ABS and np.abs(s2_p - s2_o).max() <= TOL_ABS


@pytest.mark.parametrize("kname,N,D,P,M", [("SquaredExpARD", 700, 6, 1, 4500), ("MaternFiveHalves", 2100, 3, 2, 5000), ("Exp", 4096, 12, 1, 4100),
                                             ("SquaredExpARD", 130, 20, 5, 4200), ("MaternThreeHalves", 300, 2, 1, 256), ("SquaredExpARD", 2100, 6, 1, 300)])
def test_query_panel_path_matches_oracle(kname, N, D, P, M, oracle_mod):
    """Batches >= 256 candidates take the panel path (blocked solve over 2048-row super-blocks: one, several and ragged
    super-blocks here, D > 16 and P > 4 included): against the oracle at the fp64 bar, against the slab kernel to 1e-12,
    deterministic, and a candidate's value does not depend on what else is in the batch."""
    from limbo_b200 import _lib, synth
    gp, og, X, Y = _make(kname, N, D, P=P)
    Xq = synth.points(777, M, D)
    mu, s2 = gp.query_batch(Xq)
    sub = np.arange(0, M, 23)
    mu_o, s2_o = og.query(Xq[sub], nthreads=8)
    assert np.abs(mu[sub] - (mu_o + Y.mean(axis=0))).max() <= TOL_ABS
    assert np.abs(s2[sub] - s2_o).max() <= TOL_ABS
    mu2, s22 = gp.query_batch(Xq)
    assert np.array_equal(mu, mu2) and np.array_equal(s2, s22)
    perm = np.random.default_rng(0).permutation(M)
    mu3, s23 = gp.query_batch(Xq[perm])
    assert np.array_equal(mu3, mu[perm]) and np.array_equal(s23, s2[perm])
    lib = _lib.load()
    lib.lb_debug_set_query_panel_min(1 << 40)
    try:
        mu_s, s2_s = gp.query_batch(Xq)
    finally:
        lib.lb_debug_set_query_panel_min(0)
    assert np.abs(mu - mu_s).max() <= 1e-12 * max(1.0, np.abs(mu).max()) and np.abs(s2 - s2_s).max() <= 1e-12
    # the factor changes: the cached diagonal-block inverses must follow
    gp.kernel_function().set_h_params(gp.kernel_function().h_params() - 0.2)
    gp.recompute(False)
    mu4, s24 = gp.query_batch(Xq)
    lib.lb_debug_set_query_panel_min(1 << 40)
    try:
        mu5, s25 = gp.query_batch(Xq)
    finally:
        lib.lb_debug_set_query_panel_min(0)
    assert np.abs(mu4 - mu5).max() <= 1e-12 * max(1.0, np.abs(mu4).max()) and np.abs(s24 - s25).max() <= 1e-12
    assert np.abs(s24 - s2).max() > 1e-6


def test_identical_samples_full_vs_incremental(oracle_mod):
    """test_gp.cpp:513-566: 10 identical samples (K is only regularised by the noise): full vs incremental Cholesky."""
    from limbo_b200 import kernel, mean, model
    O = oracle_mod
    X = np.tile(np.array([[0.3, 0.7]]), (10, 1))
    y = np.linspace(0.9, 1.1, 10)[:, None]
    gp = model.GP(2, 1, kernel=kernel.MaternFiveHalves, mean=mean.Data)
    gp.compute(X, y)
    gi = model.GP(2, 1, kernel=kernel.MaternFiveHalves, mean=mean.Data)
    gi.compute(X[:1], y[:1])
    for i in range(1, 10):
        gi.add_sample(X[i], y[i])
    K, L, Li = gp.kernel_matrix(), gp.matrixL(), gi.matrixL()
    assert np.abs(L @ L.T - K).max() <= 1e-12 and np.abs(Li @ Li.T - K).max() <= 1e-12  # isApprox(1e-5) there
    q = np.array([[0.3, 0.7], [0.31, 0.69]])
    (m1, s1), (m2, s2) = gp.query_batch(q), gi.query_batch(q)
    assert np.abs(m1 - m2).max() <= 1e-9 and np.abs(s1 - s2).max() <= 1e-10  # 1e-4 there
    og = O.OracleGP()
    og.set_data(X, y - y.mean())
    og.set_kernel(O.K_MATERN52, np.zeros(2), 0.01)
    og.fit()
    mo, so = og.query(q)
    assert np.abs(m1 - (mo + y.mean())).max() <= 1e-9 and np.abs(s1 - so).max() <= TOL_ABS


@pytest.mark.parametrize("N", [1, 2, 127, 128, 129, 256, 257])
def test_padding_boundaries(N, oracle_mod):
    """Sizes around the 128-padding quantum, single sample included."""
    from limbo_b200 import synth
    gp, og, X, Y = _make("SquaredExpARD", N, 3)
    Xq = synth.points(8, 9, 3)
    mu, s2 = gp.query_batch(Xq)
    mu_o, s2_o = og.query(Xq)
    assert np.abs(mu - (mu_o + Y.mean(axis=0))).max() <= TOL_ABS and np.abs(s2 - s2_o).max() <= TOL_ABS
    assert abs(gp.compute_log_lik() - og.log_lik()) <= 1e-11 * max(1.0, abs(og.log_lik()))
    g, go = gp.compute_kernel_grad_log_lik(), og.grad()
    assert np.abs(g - go).max() <= 1e-9 * max(1.0, np.abs(go).max())


def test_max_input_dimension_and_single_query(oracle_mod):
    """D = 64 (LB_MAX_D; four TMA passes of 16 dimensions) and a one-candidate batch."""
    from limbo_b200 import synth
    gp, og, X, Y = _make("SquaredExpARD", 200, 64, hp=np.concatenate([np.full(64, 1.2), [0.1]]))
    Xq = synth.points(3, 1, 64)
    mu, s2 = gp.query_batch(Xq)
    mu_o, s2_o = og.query(Xq)
    assert np.abs(mu - (mu_o + Y.mean(axis=0))).max() <= TOL_ABS and np.abs(s2 - s2_o).max() <= TOL_ABS
    assert np.abs(gp.kernel_matrix() - og.get(0)).max() <= TOL_ABS
    g, go = gp.compute_kernel_grad_log_lik(), og.grad()
    assert np.abs(g - go).max() <= 1e-9 * max(1.0, np.abs(go).max())
    from limbo_b200 import kernel, mean, model
    with pytest.raises(Exception):
        model.GP(65, 1, kernel=kernel.SquaredExpARD, mean=mean.Data).compute(np.zeros((4, 65)), np.zeros((4, 1)))


def test_empty_candidate_batch_and_reuse(oracle_mod):
    gp, og, X, Y = _make("Exp", 40, 2)
    mu, s2 = gp.query_batch(np.zeros((0, 2)))
    assert mu.shape == (0, 1) and s2.shape == (0,)
    # refit with different data of another size on the same handle (buffers are re-allocated)
    gp.compute(X[:17], Y[:17])
    assert gp.nb_samples() == 17 and gp.matrixL().shape == (17, 17)
