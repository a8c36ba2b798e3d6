"""The CPU oracle against everything the reference's own tests pin for this path
(SURVEY.md §4 / §8c): SE-ARD known answers, prior variance, kernel and LML gradients
vs finite differences, L L^T = K, incremental vs full Cholesky, EI/UCB formulas — plus
an independent SciPy/LAPACK opinion and the long-double restatement."""
import math

import numpy as np
import pytest
import scipy.linalg as sl
import scipy.special as sp

from limbo_b200 import synth


def test_se_ard_known_answers(oracle_mod):
    """src/tests/test_kernel.cpp:196-224"""
    O = oracle_mod
    hp = np.zeros(3)  # ell = (1,1), sigma_f^2 = 1
    v1 = np.array([1.0, 1.0])
    assert abs(O.kernel_eval(O.K_SE_ARD, hp, 0.01, v1, v1) - 1.0) < 1e-6
    v2 = np.array([0.0, 1.0])
    assert abs(O.kernel_eval(O.K_SE_ARD, hp, 0.01, v1, v2) - math.exp(-0.5)) < 1e-5
    # larger length scale on the differing axis -> larger k (test_kernel.cpp:213-217)
    hp2 = np.array([1.0, 0.0, 0.0])
    assert O.kernel_eval(O.K_SE_ARD, hp2, 0.01, v1, v2) > O.kernel_eval(O.K_SE_ARD, hp, 0.01, v1, v2)
    # noise + 1e-8 only when i == j (kernel.hpp:83)
    assert abs(O.kernel_eval(O.K_SE_ARD, hp, 0.01, v1, v1, same_index=True) - (1.0 + 0.01 + 1e-8)) < 1e-15


@pytest.mark.parametrize("kid", [0, 1, 2, 3])
def test_kernel_gradient_vs_fd(oracle_mod, kid):
    """src/tests/test_kernel.cpp:112-194: D = 1..10, h-params in [-3,3], x in [-5,5], eps 1e-?; bar 1e-5"""
    O = oracle_mod
    rng = np.random.default_rng(kid)
    for D in range(1, 11):
        nh = D + 1 if kid == 0 else 2
        for _ in range(10):
            hp = rng.uniform(-3, 3, nh)
            x1, x2 = rng.uniform(-5, 5, D), rng.uniform(-5, 5, D)
            g = O.kernel_grad(kid, hp, x1, x2)
            fd = np.zeros(nh)
            e = 1e-6
            for i in range(nh):
                a, b = hp.copy(), hp.copy()
                a[i] += e
                b[i] -= e
                fd[i] = (O.kernel_eval(kid, a, 0.0, x1, x2) - O.kernel_eval(kid, b, 0.0, x1, x2)) / (2 * e)
            assert np.linalg.norm(g - fd) < 1e-5 * max(1.0, np.linalg.norm(g))


@pytest.mark.parametrize("kid", [0, 1, 2, 3])
def test_prior_variance(oracle_mod, kid):
    """src/tests/test_gp.cpp:697-758: sigma(x) = sigma_f^2 (+ noise) with no samples"""
    O = oracle_mod
    g = O.OracleGP()
    g.set_data(np.zeros((0, 2)), np.zeros((0, 1)))
    nh = 3 if kid == 0 else 2
    hp = np.zeros(nh)
    hp[-1] = math.log(math.sqrt(10.0))
    g.set_kernel(kid, hp, 0.01)
    mu, s2 = g.query(np.array([[0.3, -0.2]]))
    assert mu[0, 0] == 0.0
    assert abs(s2[0] - 10.0) / 10.0 < 0.01


def _fit(O, kid, N, D, P=1, hp=None, noise=0.01, prec=0, seed=1234):
    X = synth.points(seed, N, D)
    y = synth.targets(X)
    Y = np.stack([y * (p + 1) for p in range(P)], axis=1)
    g = O.OracleGP(prec)
    g.set_data(X, Y - Y.mean(axis=0))
    nh = D + 1 if kid == 0 else 2
    g.set_kernel(kid, np.zeros(nh) if hp is None else hp, noise)
    assert g.fit() == -1
    return g, X, Y


@pytest.mark.parametrize("kid", [0, 1])
def test_cholesky_alpha_vs_lapack(oracle_mod, kid):
    O = oracle_mod
    for N in (3, 31, 33, 200, 515):
        g, X, Y = _fit(O, kid, N, 6)
        K, L, A = g.get(0), g.get(1), g.get(2)
        assert np.array_equal(K, K.T)
        L2 = sl.cholesky(K, lower=True)
        assert np.abs(L - L2).max() < 1e-12
        assert np.abs(L @ L.T - K).max() < 1e-12  # test_gp.cpp:550-555 (1e-5 there)
        A2 = sl.cho_solve((L2, True), Y - Y.mean(axis=0))
        assert np.abs(A - A2).max() <= 1e-11 * max(1.0, np.abs(A2).max())


def test_lml_gradient_vs_fd(oracle_mod):
    """src/tests/test_gp.cpp:131-193: N=40, D=4, P=2, eps=1e-4, bar sum < 100*eps (looser there)"""
    O = oracle_mod
    rng = np.random.default_rng(0)
    N, D = 40, 4
    for trial in range(10):
        hp = rng.uniform(-1, 1, D + 1)
        g, X, Y = _fit(O, 0, N, D, P=2, hp=hp)
        v, gr = g.lml_eval(hp, True)
        fd = np.zeros(D + 1)
        e = 1e-4
        for i in range(D + 1):
            a, b = hp.copy(), hp.copy()
            a[i] += e
            b[i] -= e
            fd[i] = (g.lml_eval(a, False)[0] - g.lml_eval(b, False)[0]) / (2 * e)
        assert np.linalg.norm(gr - fd) < 1e-4 * max(1.0, np.linalg.norm(gr))


def test_lml_gradient_with_noise_vs_fd(oracle_mod):
    """src/tests/test_gp.cpp:195-271 (optimize_noise = true)"""
    O = oracle_mod
    g, X, Y = _fit(O, 0, 40, 3)
    hp = np.array([0.1, -0.2, 0.3, 0.05, math.log(math.sqrt(0.02))])
    v, gr = g.lml_eval(hp, True, optimize_noise=True)
    e = 1e-5
    fd = np.zeros(5)
    for i in range(5):
        a, b = hp.copy(), hp.copy()
        a[i] += e
        b[i] -= e
        fd[i] = (g.lml_eval(a, False, True)[0] - g.lml_eval(b, False, True)[0]) / (2 * e)
    assert np.linalg.norm(gr - fd) < 1e-5 * max(1.0, np.linalg.norm(gr))


def test_incremental_vs_full(oracle_mod):
    """src/tests/test_gp.cpp:568-635: add_sample vs fresh compute, |dmu| < 1e-5, L approx 1e-5"""
    O = oracle_mod
    X = synth.points(5, 100, 1)
    y = synth.targets(X)[:, None]
    g = O.OracleGP()
    g.set_data(X[:50], y[:50] - y[:50].mean())
    g.set_kernel(O.K_MATERN52, np.zeros(2), 0.01)
    g.fit()
    for i in range(50, 100):
        g.append(X[i], y[: i + 1] - y[: i + 1].mean())
    g2, _, _ = None, None, None
    h = O.OracleGP()
    h.set_data(X, y - y.mean())
    h.set_kernel(O.K_MATERN52, np.zeros(2), 0.01)
    h.fit()
    assert np.abs(g.get(1) - h.get(1)).max() < 1e-9
    Xq = synth.points(6, 50, 1)
    (m1, s1), (m2, s2) = g.query(Xq), h.query(Xq)
    assert np.abs(m1 - m2).max() < 1e-9 and np.abs(s1 - s2).max() < 1e-9


def test_three_point_interpolation(oracle_mod):
    """src/tests/test_gp.cpp:448-511: |mu(x_i) - y_i| < 1, sigma^2 <= 2 (noise + 1e-8)"""
    O = oracle_mod
    X = np.array([[1.0], [2.0], [3.0]])
    y = np.array([[5.0], [10.0], [5.0]])
    g = O.OracleGP()
    g.set_data(X, y - y.mean())
    g.set_kernel(O.K_MATERN52, np.zeros(2), 0.01)
    g.fit()
    mu, s2 = g.query(X)
    assert np.all(np.abs(mu + y.mean() - y) < 1.0)
    assert np.all(s2 <= 2 * (0.01 + 1e-8))


def test_query_formulas_vs_numpy(oracle_mod):
    O = oracle_mod
    g, X, Y = _fit(O, 0, 150, 6)
    K, L, A = g.get(0), g.get(1), g.get(2)
    Xq = synth.points(1235, 40, 6)
    mu, s2 = g.query(Xq)
    d2 = ((X[:, None, :] - Xq[None, :, :]) ** 2).sum(-1)
    Ks = np.exp(-0.5 * d2)
    assert np.abs(mu[:, 0] - Ks.T @ A[:, 0]).max() < 1e-12
    V = sl.solve_triangular(L, Ks, lower=True)
    ref = 1.0 - (V * V).sum(0) + 0.01
    assert np.abs(s2 - ref).max() < 1e-12
    # multi-threaded candidate fan-out (tools::par) gives identical results
    mu2, s22 = g.query(Xq, nthreads=4)
    assert np.array_equal(mu, mu2) and np.array_equal(s2, s22)


def test_loglik_and_kinv_vs_numpy(oracle_mod):
    O = oracle_mod
    g, X, Y = _fit(O, 1, 120, 3, P=2)
    K, L, A = g.get(0), g.get(1), g.get(2)
    om = Y - Y.mean(axis=0)
    ll = -0.5 * np.trace(om.T @ A) - np.log(np.diag(L)).sum() - 0.5 * 120 * math.log(2 * math.pi)
    assert abs(g.log_lik() - ll) < 1e-11 * abs(ll)
    assert np.abs(g.get(3) - np.linalg.inv(K)).max() < 1e-8 * np.abs(np.linalg.inv(K)).max()


def test_acquisition_formulas(oracle_mod):
    """acqui/ucb.hpp:89, acqui/ei.hpp:92-115, acqui/gp_ucb.hpp:85-88"""
    O = oracle_mod
    rng = np.random.default_rng(1)
    mu, s2 = rng.normal(size=100), rng.uniform(1e-6, 2.0, 100)
    assert np.abs(O.ucb(mu, s2, 0.5) - (mu + 0.5 * np.sqrt(s2))).max() < 1e-15
    f_max, jit = 0.3, 0.01
    s = np.sqrt(s2)
    Z = (mu - f_max - jit) / s
    ref = (mu - f_max - jit) * 0.5 * sp.erfc(-Z / math.sqrt(2)) + s * np.exp(-0.5 * Z * Z) / math.sqrt(2 * math.pi)
    assert np.abs(O.ei(mu, s2, f_max, jit) - ref).max() < 1e-14
    assert O.ei(np.array([1.0]), np.array([1e-22]), 0.0)[0] == 0.0  # sigma < 1e-10 -> 0 (ei.hpp:95)
    beta = math.sqrt(2.0 * math.log(math.pow(7, 6 / 2.0 + 2.0) * math.pi ** 2 / 0.3))
    assert abs(O.load().lbo_gp_ucb_beta(7, 6, 0.1) - beta) < 1e-14


def test_double_vs_long_double(oracle_mod):
    """Distance between the reference's fp64 arithmetic and an 80-bit restatement of the
    same algorithm: the conditioning floor the 1e-10 bar has to be read against."""
    O = oracle_mod
    gd, X, Y = _fit(O, 0, 400, 6)
    gl, _, _ = _fit(O, 0, 400, 6, prec=O.PREC_LONG_DOUBLE)
    Xq = synth.points(1235, 100, 6)
    (m1, s1), (m2, s2) = gd.query(Xq), gl.query(Xq)
    assert np.abs(gd.get(0) - gl.get(0)).max() < 1e-15
    assert np.abs(m1 - m2).max() < 1e-10 and np.abs(s1 - s2).max() < 1e-12
    assert np.abs(gd.get(2) - gl.get(2)).max() < 1e-9 * np.abs(gl.get(2)).max()
    assert abs(gd.log_lik() - gl.log_lik()) < 1e-11 * abs(gl.log_lik())


def test_rprop_call_count_and_ascent(oracle_mod):
    """src/tests/test_optimizers.cpp:182-193: exactly `iterations` evaluations; LML does not decrease"""
    O = oracle_mod
    g, X, Y = _fit(O, 0, 60, 2)
    best, ne = g.rprop_lml(np.zeros(3), 15)
    assert ne == 15
    assert g.lml_eval(best, False)[0] >= g.lml_eval(np.zeros(3), False)[0]
