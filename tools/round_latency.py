"""Time of ONE batched log-posterior + gradient evaluation (one L-BFGS-B round of the MAP fit) against the number
of systems in the batch, with the per-class kernel times: shows the latency floor of the recursion (what a round
costs when only one or two restarts are still active).  python tools/round_latency.py [N d] -> one JSON line per B."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np



def synthetic_problem(N, d, seed=1):
    """same as oracle.gp_oracle.synthetic_problem (tools/ may not import oracle/)"""
    import math
    from scipy.stats.qmc import LatinHypercube
    X = LatinHypercube(d, seed=seed).random(N)
    y = np.sum(np.sin(6 * np.pi * X) / np.arange(1, d + 1), axis=1) + X[:, 0] * X[:, 1 % d] - math.sqrt(2.0)
    return X, y


import exauq_toolbox_b200 as exq  # noqa: E402
from exauq_toolbox_b200 import _lib  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
d = int(sys.argv[2]) if len(sys.argv) > 2 else 8
_lib.init(0)
X, y = synthetic_problem(N, d)
g = exq.DeviceGP(X, y, "Matern52")
NAMES = ["gemm_predict", "gemm_factor", "leaf", "assemble", "stream", "gemm_i8", "slice", "fused", "lauum_grad"]
base = np.concatenate([np.full(d, -2 * np.log(0.5)), [np.log(1.7)]])
rows = []
for B in (1, 2, 3, 4, 8, 15, 30):
    th = np.tile(base, (B, 1)) + np.random.default_rng(B).uniform(-0.3, 0.3, (B, d + 1))
    for _ in range(2):
        g.logpost_batch(th, 0.0)
    _lib.prof_enable(0)
    reps = 5
    _lib.timer_start()
    for _ in range(reps):
        g.logpost_batch(th, 0.0)
    ms = _lib.timer_stop() / reps
    _lib.prof_enable(0x1FF)
    _lib.prof_reset()
    g.logpost_batch(th, 0.0)
    out = {"N": N, "d": d, "systems": B, "ms_per_round": round(ms, 3), "ms_per_system": round(ms / B, 3)}
    for c, n in enumerate(NAMES):
        p = _lib.prof_get(c)
        if p["launches"]:
            out[n] = {"ms": round(p["ms"], 3), "launches": p["launches"]}
    _lib.prof_enable(0)
    rows.append(out)
    print(json.dumps(out), flush=True)
