"""Small program for ncu captures: one fit (factor GEMMs + fused sub-tree launches), one PEI evaluation over 131072
candidates (4 chunks: fused cross-covariance tiles + int8 variance product + per-candidate finish each), one batched
log-posterior + gradient over R restarts (default 4; NCU_TARGET_R=15 is the bench's batch).
NCU_TARGET_PARTS = comma list of fit,pei,logpost (default all): a capture of the int8 GEMMs of ONE evaluation round uses
"logpost" with NCU_TARGET_WARM=0 so that the 13 launches ncu sees are exactly that round's."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from scipy.stats.qmc import LatinHypercube

import exauq_toolbox_b200 as exq
from exauq_toolbox_b200 import _lib

parts = os.environ.get("NCU_TARGET_PARTS", "fit,pei,logpost").split(",")
_lib.init(0)
N, d, M = 4096, 8, 131072
X = LatinHypercube(d, seed=1).random(N)
y = np.sum(np.sin(6 * np.pi * X) / np.arange(1, d + 1), axis=1)
corr_raw = np.full(d, -2 * math.log(0.5))
g = exq.DeviceGP(X, y, "Matern52")
if "fit" in parts or "pei" in parts:
    g.fit(corr_raw, 1.7, 0.0)
if "pei" in parts:
    Xc = LatinHypercube(d, seed=2).random(M)
    cand = exq.DeviceCandidates(Xc)
    cand.pei_eval(g, np.random.default_rng(0).uniform(0, 1, (272, d)), float(y.max()))
    print(cand.batch(g, 2))
if "logpost" in parts:
    R = int(os.environ.get("NCU_TARGET_R", "4"))
    th = np.tile(np.concatenate([corr_raw, [math.log(1.7)]]), (R, 1))
    th = th + np.random.default_rng(3).uniform(-0.2, 0.2, th.shape)
    print(g.logpost_batch(th, 0.0)[0])
