"""Small program for ncu captures: one fit (124 factor GEMMs + 32 leaves), one PEI evaluation
over 131072 candidates (4 chunks: cross assembly + variance GEMM + epilogue each), one batched
log-posterior over R restarts (default 4; NCU_TARGET_R=15 is the bench's batch)."""
import os, sys, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import exauq_toolbox_b200 as exq
from exauq_toolbox_b200 import _lib
from scipy.stats.qmc import LatinHypercube
_lib.init(0)
N, d, M = 4096, 8, 131072
X = LatinHypercube(d, seed=1).random(N)
y = np.sum(np.sin(6 * np.pi * X) / np.arange(1, d + 1), axis=1)
Xc = LatinHypercube(d, seed=2).random(M)
corr_raw = np.full(d, -2 * math.log(0.5))
g = exq.DeviceGP(X, y, "Matern52").fit(corr_raw, 1.7, 0.0)
cand = exq.DeviceCandidates(Xc)
cand.pei_eval(g, np.random.default_rng(0).uniform(0, 1, (272, d)), float(y.max()))
print(cand.batch(g, 2))
th = np.tile(np.concatenate([corr_raw, [math.log(1.7)]]), (int(os.environ.get("NCU_TARGET_R", "4")), 1))
print(g.logpost_batch(th, 0.0)[0])
