"""Deterministic synthetic inputs for benchmarks and tests: numpy default_rng, fixed seeds."""
from __future__ import annotations

import numpy as np


def dense_sym_tbt(n: int, seed: int) -> np.ndarray:
    """N(0,1) tensor symmetrised over p<->q, r<->s and (pq)<->(rs), scaled by 1/N^2 (8-fold symmetric)."""
    rng = np.random.default_rng(seed)
    g = rng.standard_normal((n, n, n, n))
    g = g + g.transpose(1, 0, 2, 3)
    g = g + g.transpose(0, 1, 3, 2)
    g = g + g.transpose(2, 3, 0, 1)
    g *= 1.0 / (n * n)
    return g


def pairsym_tbt(n: int, seed: int) -> np.ndarray:
    """Only (pq)<->(rs) symmetric."""
    rng = np.random.default_rng(seed)
    g = rng.standard_normal((n, n, n, n))
    return (g + g.transpose(2, 3, 0, 1)) / (2.0 * n * n)


def general_tbt(n: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return rng.standard_normal((n, n, n, n)) / (n * n)


def random_x(n: int, seed: int, flavour: str = "CSA") -> np.ndarray:
    """x = [lambda ~ N(0,1)/N ; theta ~ U[0, 2 pi)]; theta matches the reference's x0 (decompose.jl:93)."""
    rng = np.random.default_rng(1000 + seed)
    cL, uL = n * (n + 1) // 2, n * (n - 1) // 2
    parts = []
    if flavour != "CSA":
        parts.append(rng.standard_normal(n) / n)
    parts.append(rng.standard_normal(cL) / n)
    parts.append(2 * np.pi * rng.random(uL))
    return np.concatenate(parts)
