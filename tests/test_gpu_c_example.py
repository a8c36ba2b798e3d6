"""The C ABI driven from plain C (examples/c_abi_example.c, built by __graft_entry__.build()) against the CPU oracle on the same
inputs: this is the call sequence a cgo / JNI / N-API binding of include/limbo_b200.h makes."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "limbo_b200", "lib", "c_abi_example")
MASK = (1 << 64) - 1


def _stream(seed, n):
    out = np.empty(n)
    s = seed
    for i in range(n):
        s = (s + 0x9E3779B97F4A7C15) & MASK
        z = s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK
        z ^= z >> 31
        out[i] = (z >> 11) / 9007199254740992.0
    return out


def test_c_example_is_built_and_links(lib):
    assert os.path.exists(EXE), "run __graft_entry__.build()"
    r = subprocess.run(["ldd", EXE], capture_output=True, text=True)
    assert "liblimbo_b200.so" in r.stdout and "not found" not in r.stdout


@pytest.mark.gpu
def test_c_example_matches_the_oracle(lib, oracle_mod):
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    vals = {}
    for line in r.stdout.strip().splitlines():
        f = line.split()
        for k, v in zip(f[0::2], f[1::2]):
            vals[k] = float(v)
    N, D, M = 300, 2, 1000
    u = _stream(2024, N * D + M * D)
    X = u[: N * D].reshape(N, D)
    Xq = u[N * D:].reshape(M, D)
    y = np.cos(3.0 * (X[:, 0] + X[:, 1]))
    og = oracle_mod.OracleGP()
    og.set_data(X, (y - y.mean())[:, None])
    og.set_kernel(0, np.array([-0.5, -0.3, 0.1]), 0.01)
    og.fit()
    mu, s2 = og.query(Xq)
    mu = mu[:, 0] + y.mean()
    ucb = oracle_mod.ucb(mu, s2, 0.5)
    assert abs(vals["loglik"] - og.log_lik()) <= 1e-10 * abs(og.log_lik())
    assert abs(vals["mu0"] - mu[0]) <= 1e-10 and abs(vals["sigma2_0"] - s2[0]) <= 1e-10
    assert abs(vals["mu_last"] - mu[-1]) <= 1e-10 and abs(vals["sigma2_last"] - s2[-1]) <= 1e-10
    assert abs(vals["best"] - ucb.max()) <= 1e-10 and int(vals["idx"]) == int(np.argmax(ucb))
    assert int(vals["n"]) == N and vals["launches"] > 0
