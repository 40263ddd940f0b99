import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np


def synthetic_problem(N, d, seed=1):
    """LHS inputs + the bench's deterministic target (same as oracle.gp_oracle.synthetic_problem, restated
    here because only tests/, smoke() and bench.py's CPU legs may import oracle/)."""
    import math
    from scipy.stats.qmc import LatinHypercube
    X = LatinHypercube(d, seed=seed).random(N)
    y = np.sum(np.sin(6 * np.pi * X) / np.arange(1, d + 1), axis=1) + X[:, 0] * X[:, 1 % d] - math.sqrt(2.0)
    return X, y


def synthetic_candidates(M, d, seed=2):
    from scipy.stats.qmc import LatinHypercube
    return LatinHypercube(d, seed=seed).random(M)

import exauq_toolbox_b200 as exq
from exauq_toolbox_b200 import _lib
_lib.init(0)
N, d, R = 4096, 8, 13
X, y = synthetic_problem(N, d)
g = exq.DeviceGP(X, y, "Matern52")
th = np.tile(np.concatenate([np.full(d, -2 * np.log(0.5)), [np.log(1.7)]]), (R, 1))
g.logpost_batch(th, 0.0)
lib = _lib.load()
lib.exq_debug_levels.restype = None
import io
print("--- after warm-up (discard) ---"); lib.exq_debug_levels()
for _ in range(3): g.logpost_batch(th, 0.0)
print("--- cumulative over 4 calls of logpost_batch R=13 ---"); lib.exq_debug_levels()
