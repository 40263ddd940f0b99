import os, sys, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import exauq_toolbox_b200 as exq
from exauq_toolbox_b200 import _lib
from exauq_toolbox_b200.design import design_batch, pseudopoints
from scipy.stats.qmc import LatinHypercube
_lib.init(0)
N, d, M = int(sys.argv[1]), 8, 200000
k = np.arange(1, d + 1)
X = LatinHypercube(d, seed=1).random(N)
variant = sys.argv[2] if len(sys.argv) > 2 else "f1"
if variant == "f1":
    y = np.sum(np.sin(2 * np.pi * X) / k, axis=1) + X[:, 0] * X[:, 1 % d] - math.sqrt(2.0)
elif variant == "f2":
    y = np.sum(np.sin(6 * np.pi * X) / k, axis=1) + X[:, 0] * X[:, 1 % d] - math.sqrt(2.0)
elif variant == "f3":
    y = np.sum(np.sin(2 * np.pi * X) / k, axis=1) + np.sin(4 * np.pi * X[:, 0] * X[:, 1]) + np.sin(5 * np.pi * X[:, 2] * X[:, 3]) + X[:, 0] * X[:, 1 % d] - math.sqrt(2.0)
elif variant == "f4":
    y = np.sum(np.sin(2 * np.pi * X) / k, axis=1) + 2.0 * np.exp(-40.0 * np.sum((X[:, :3] - 0.3) ** 2, axis=1)) + np.sin(4 * np.pi * X[:, 0] * X[:, 1]) - math.sqrt(2.0)
Xc = LatinHypercube(d, seed=2).random(M)
cand = exq.DeviceCandidates(Xc)
st = {}
pts, gidx, vals, theta = design_batch(X, y, "Matern52", [0.5] * d, 1.7, 0.0, [(0.0, 1.0)] * d, cand, batch_size=4, n_tries=15, seed=12345, stats=st)
print("theta errors gp: ell", np.exp(-0.5 * theta[:d]), "cov", math.exp(theta[d]), st)
print("chosen", gidx, vals)
pei = cand.download("pei"); m = cand.download("mean"); v = cand.download("var")
print("pei max", pei.max(), "nonzero", (pei > 0).sum(), "nan", np.isnan(pei).sum())
print("mean range", m.min(), m.max(), "var range", v.min(), v.max())
g = exq.DeviceGP(X, y, "Matern52").fit(np.full(d, -2 * math.log(0.5)), 1.7, 0.0)
lm, lv, nes = g.loo()
print("nes range", nes.min(), nes.max(), np.median(nes), "loo var range", lv.min(), lv.max())
