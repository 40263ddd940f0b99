"""The synthetic workload of the benchmark and of the full-size parity tests: Latin
hypercube inputs on [0, 1]^d as ``oneshot_lhs`` draws them (exauq/core/designers.py:177-181 calls scipy's
``LatinHypercube``), a smooth deterministic target, Latin hypercube candidates.  ONE definition for both arms of
``bench.py`` (tests/test_host_logic.py checks that the oracle's copy produces the same arrays)."""

from __future__ import annotations

import math

import numpy as np


def synthetic_target(X: np.ndarray) -> np.ndarray:
    """sum_k sin(6 pi x_k) / k + x_1 x_2 - sqrt 2.  Three periods per coordinate keep the leave-one-out errors of
    a length-scale-0.5 GP rough, so the errors GP gets short length scales and the PEI product over N repulsion
    factors does not underflow to 0 at N = 4096."""
    d = X.shape[1]
    k = np.arange(1, d + 1)
    return np.sum(np.sin(6 * np.pi * X) / k, axis=1) + X[:, 0] * X[:, 1 % d] - math.sqrt(2.0)


def synthetic_problem(N: int, d: int, seed: int = 1):
    from scipy.stats.qmc import LatinHypercube

    X = LatinHypercube(d, seed=seed).random(N)
    return X, synthetic_target(X)


def synthetic_candidates(M: int, d: int, seed: int = 2):
    from scipy.stats.qmc import LatinHypercube

    return LatinHypercube(d, seed=seed).random(M)
