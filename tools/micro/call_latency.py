"""Per-call latency of the drop-in class's scalar entry points (what the unchanged designers' DE loop pays)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "baseline", "_ref")]
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

from exauq.core.modelling import Input, TrainingDatum, GaussianProcessHyperparameters, SimulatorDomain
from exauq.core.designers import PEICalculator
from exauq_toolbox_b200.gp import B200GaussianProcess
for N, d in [(50, 2), (1024, 8), (4096, 8)]:
    X, y = synthetic_problem(N, d)
    data = [TrainingDatum(Input(*r), float(t)) for r, t in zip(X, y)]
    gp = B200GaussianProcess(kernel="Matern52")
    gp.fit(data, hyperparameters=GaussianProcessHyperparameters([0.5] * d, 1.7, 0.0))
    xs = [Input(*r) for r in synthetic_candidates(100, d)]
    train_inputs = [dt.input for dt in data]
    gp.predict(xs[0])
    t0 = time.perf_counter()
    for x in xs: gp.predict(x)
    t1 = time.perf_counter()
    for x in xs[:20]: gp.covariance_matrix([x])
    t2 = time.perf_counter()
    print(f"N={N} d={d}: predict {1e3*(t1-t0)/100:.3f} ms/call, covariance_matrix([x]) {1e3*(t2-t1)/20:.3f} ms/call", flush=True)
    if N <= 1024:
        dom = SimulatorDomain([(0.0, 1.0)] * d)
        t0 = time.perf_counter(); pei = PEICalculator(dom, gp); t1 = time.perf_counter()
        for x in xs[:20]: pei.compute(x)
        t2 = time.perf_counter()
        print(f"   PEICalculator ctor {t1-t0:.2f} s, compute {1e3*(t2-t1)/20:.3f} ms/call", flush=True)
