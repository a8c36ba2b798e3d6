"""Deterministic synthetic inputs shared by the CUDA path, the oracle and bench.py
(SURVEY.md §8d): u(i) = (splitmix64(seed + i) >> 11) * 2^-53, X ~ U[0,1)^D,
targets = Hartmann6 (reference: src/benchmarks/regression/test_functions.hpp:343-367)."""
from __future__ import annotations

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64)
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def uniform(seed: int, n: int, offset: int = 0) -> np.ndarray:
    idx = np.arange(offset, offset + n, dtype=np.uint64) + np.uint64(seed)
    return (splitmix64(idx) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def points(seed: int, n: int, d: int) -> np.ndarray:
    """n x d row-major points in [0,1)^d; x[i, k] = u(i*d + k)."""
    return uniform(seed, n * d).reshape(n, d)


_A = np.array([[10, 3, 17, 3.5, 1.7, 8], [0.05, 10, 17, 0.1, 8, 14], [3, 3.5, 1.7, 10, 17, 8], [17, 8, 0.05, 10, 0.1, 14]])
_P = np.array([[0.1312, 0.1696, 0.5569, 0.0124, 0.8283, 0.5886], [0.2329, 0.4135, 0.8307, 0.3736, 0.1004, 0.9991],
               [0.2348, 0.1451, 0.3522, 0.2883, 0.3047, 0.665], [0.4047, 0.8828, 0.8732, 0.5743, 0.1091, 0.0381]])
_ALPHA = np.array([1.0, 1.2, 3.0, 3.2])


def hartmann6(x: np.ndarray) -> np.ndarray:
    x = np.atleast_2d(x)
    s = np.einsum("ij,nij->ni", _A, (x[:, None, :6] - _P[None]) ** 2)
    return (np.exp(-s) * _ALPHA).sum(axis=1)


def targets(x: np.ndarray) -> np.ndarray:
    """D=6 -> Hartmann6(x); D=12 -> H6(x[:6]) + H6(x[6:]); D=1 -> cos(4x-2) (tutorials/gp.cpp-like);
    other D: sum of cos over dimensions plus Hartmann6 on the first min(D,6) padded dims."""
    d = x.shape[1]
    if d == 6:
        return hartmann6(x)
    if d == 12:
        return hartmann6(x[:, :6]) + hartmann6(x[:, 6:])
    if d == 1:
        return np.cos(4.0 * x[:, 0] - 2.0)
    return np.cos(3.0 * x).sum(axis=1) + np.sin(5.0 * x[:, 0])
