"""Development perf probe on a GPU box: per-class kernel times at benchmark sizes."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
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
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
d = int(sys.argv[2]) if len(sys.argv) > 2 else 8
M = int(sys.argv[3]) if len(sys.argv) > 3 else 1000000
R = int(sys.argv[4]) if len(sys.argv) > 4 else 8
_lib.init(0)
X, y = synthetic_problem(N, d)
Xc = synthetic_candidates(M, d)
corr_raw = np.full(d, -2 * np.log(0.5)); cov = 1.7
NAMES = ["gemm_predict", "gemm_factor", "leaf", "assemble", "stream"]
def report(tag, ms):
    out = {"tag": tag, "ms": ms}
    for c, n in enumerate(NAMES):
        p = _lib.prof_get(c)
        if p["launches"]:
            out[n] = {"ms": round(p["ms"], 3), "launches": p["launches"], "tflops": round(p["flops"] / max(p["ms"], 1e-9) / 1e9, 2), "gbs": round(p["bytes"] / max(p["ms"], 1e-9) / 1e6, 1)}
    print(json.dumps(out), flush=True)
g = exq.DeviceGP(X, y, "Matern52")
g.fit(corr_raw, cov, 0.0)  # warm
_lib.prof_enable(0b11111)
for rep in range(2):
    _lib.prof_reset(); _lib.timer_start(); g.fit(corr_raw, cov, 0.0); report("fit N=%d" % N, _lib.timer_stop())
_lib.prof_reset(); _lib.timer_start(); lm = g.loo(); report("loo", _lib.timer_stop())
th = np.tile(np.concatenate([corr_raw, [np.log(cov)]]), (R, 1)) + np.random.default_rng(0).uniform(-0.3, 0.3, (R, d + 1))
g.logpost_batch(th, 0.0)
for rep in range(2):
    _lib.prof_reset(); _lib.timer_start(); g.logpost_batch(th, 0.0); report("logpost_batch R=%d" % R, _lib.timer_stop())
cand = exq.DeviceCandidates(Xc)
rep_pts = np.random.default_rng(1).uniform(0, 1, (272, d))
cand.pei_eval(g, rep_pts, float(y.max()))
for rep in range(2):
    _lib.prof_reset(); _lib.timer_start(); cand.pei_eval(g, rep_pts, float(y.max())); report("pei_eval M=%d" % M, _lib.timer_stop())
_lib.prof_reset(); _lib.timer_start(); bi, bv = cand.batch(g, 16); report("pei_batch 16", _lib.timer_stop())
print(bi.tolist())
idx = np.arange(16)
g.loo_refit(idx)
_lib.prof_reset(); _lib.timer_start(); g.loo_refit(idx); report("loo_refit 16", _lib.timer_stop())
