"""Wall-clock of the reference's UNCHANGED compute_single_level_loo_samples(gp, domain, batch_size=16) on a
B200GaussianProcess at N = 4096, d = 8 (VERDICT r01 "next round" item 8: make "drop-in" true at benchmark size):

  stock      exauq's SimulatorDomain + exauq's scalar maximise        (only --stock-n points, it takes minutes)
  fast       FastSimulatorDomain + accelerate_designers()             (the whole batch)

Run on the GPU box:  python tools/dropin_timing.py > gpurun_out/dropin_timing.json
"""

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "baseline", "_ref")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=4096)
    ap.add_argument("--d", type=int, default=8)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--stock-batch", type=int, default=0, help="design points through the stock scalar path (0: skip)")
    ap.add_argument("--profile", action="store_true", help="cProfile of the fast path (top 30 by cumulative time on stderr)")
    ap.add_argument("--multi-level", default="", help="sizes 'n1,n2,n3': time the unchanged compute_multi_level_loo_samples "
                    "(d = --d, costs 1 / 10 / 100, batch --batch) on B200 GPs inside accelerate_designers() instead")
    a = ap.parse_args()
    import exauq.core.designers as designers
    from exauq.core.modelling import GaussianProcessHyperparameters, Input, SimulatorDomain, TrainingDatum

    from exauq_toolbox_b200 import _lib
    from exauq_toolbox_b200.accel import FastSimulatorDomain, accelerate_designers
    from exauq_toolbox_b200.gp import B200GaussianProcess
    from exauq_toolbox_b200.workload import synthetic_problem

    _lib.init(0)
    if a.multi_level:
        return multi_level(a)
    X, y = synthetic_problem(a.n, a.d)
    t0 = time.perf_counter()
    data = [TrainingDatum(Input(*row), float(v)) for row, v in zip(X, y)]
    out = {"N": a.n, "d": a.d, "batch_size": a.batch, "build_training_data_s": time.perf_counter() - t0}
    bounds = [(0.0, 1.0)] * a.d
    gp = B200GaussianProcess(kernel="Matern52")
    t0 = time.perf_counter()
    gp.fit(data, hyperparameters=GaussianProcessHyperparameters([0.5] * a.d, 1.7, 0.0))
    out["fit_with_hyperparameters_s"] = time.perf_counter() - t0

    t0 = time.perf_counter()
    fast = FastSimulatorDomain(bounds)
    out["fast_domain_ctor_s"] = time.perf_counter() - t0
    np.random.seed(5)
    l0 = _lib.launch_count()
    t0 = time.perf_counter()
    prof = None
    if a.profile:
        import cProfile
        prof = cProfile.Profile()
        prof.enable()
    with accelerate_designers():
        pts = designers.compute_single_level_loo_samples(gp, fast, batch_size=a.batch, seed=11)
    if prof is not None:
        import pstats
        prof.disable()
        pstats.Stats(prof, stream=sys.stderr).sort_stats("cumulative").print_stats(30)
    out["fast"] = {"compute_single_level_loo_samples_s": time.perf_counter() - t0,
                   "gpu_launches": _lib.launch_count() - l0,
                   "points": [[float(c) for c in p] for p in pts]}
    # the pieces of the stock path that make it minutes
    t0 = time.perf_counter()
    stock = SimulatorDomain(bounds)
    out["stock_domain_ctor_s"] = time.perf_counter() - t0
    inputs = [d.input for d in data]
    t0 = time.perf_counter()
    pp_fast = fast.calculate_pseudopoints(inputs)
    out["fast_calculate_pseudopoints_s"] = time.perf_counter() - t0
    sub = inputs[: max(64, a.n // 32)]
    t0 = time.perf_counter()
    stock.calculate_pseudopoints(sub)
    dt = time.perf_counter() - t0
    out["stock_calculate_pseudopoints_s_extrapolated"] = dt * a.n / len(sub)
    out["stock_calculate_pseudopoints_sample"] = f"{len(sub)} of {a.n} inputs, scaled linearly"
    out["n_pseudopoints"] = len(pp_fast)
    if a.stock_batch > 0:
        np.random.seed(5)
        t0 = time.perf_counter()
        pts_s = designers.compute_single_level_loo_samples(gp, fast, batch_size=a.stock_batch, seed=11)
        out["scalar_maximise"] = {"batch_size": a.stock_batch, "compute_single_level_loo_samples_s": time.perf_counter() - t0,
                                  "note": "FastSimulatorDomain but the reference's scalar maximise (one predict + one "
                                          "covariance_matrix device round trip per objective evaluation)",
                                  "points": [[float(c) for c in p] for p in pts_s]}
    print(json.dumps(out))


def multi_level(a):
    """BASELINE.json configs[3] through the reference's own compute_multi_level_loo_samples (designers.py:1184-1378)."""
    import exauq.core.designers as designers
    from exauq.core.modelling import (GaussianProcessHyperparameters, Input, MultiLevel, MultiLevelGaussianProcess,
                                      TrainingDatum)

    from exauq_toolbox_b200.accel import FastSimulatorDomain, accelerate_designers
    from exauq_toolbox_b200.gp import B200GaussianProcess
    from exauq_toolbox_b200.workload import synthetic_problem

    sizes = [int(v) for v in a.multi_level.split(",")]
    bounds = [(0.0, 1.0)] * a.d
    data, hps = {}, {}
    for lv, n in enumerate(sizes, start=1):
        X, y = synthetic_problem(n, a.d, seed=lv)
        data[lv] = [TrainingDatum(Input(*row), float(v) / lv) for row, v in zip(X, y)]
        hps[lv] = GaussianProcessHyperparameters([0.5] * a.d, 1.7 / lv, 0.0)
    mlgp = MultiLevelGaussianProcess([B200GaussianProcess(kernel="Matern52") for _ in sizes])
    t0 = time.perf_counter()
    mlgp.fit(MultiLevel(data), hyperparameters=MultiLevel(hps))
    out = {"sizes": sizes, "d": a.d, "batch_size": a.batch, "fit_with_hyperparameters_s": time.perf_counter() - t0}
    costs = MultiLevel([10.0 ** k for k in range(len(sizes))])
    domain = FastSimulatorDomain(bounds)
    np.random.seed(5)
    prof = None
    if a.profile:
        import cProfile
        prof = cProfile.Profile()
        prof.enable()
    t0 = time.perf_counter()
    with accelerate_designers():
        pts = designers.compute_multi_level_loo_samples(mlgp, domain, costs, batch_size=a.batch,
                                                        seeds=MultiLevel([11 + k for k in range(len(sizes))]))
    out["compute_multi_level_loo_samples_s"] = time.perf_counter() - t0
    if prof is not None:
        import pstats
        prof.disable()
        pstats.Stats(prof, stream=sys.stderr).sort_stats("cumulative").print_stats(30)
    out["level_chosen"] = [int(lv) for lv in pts.levels if len(pts[lv])]
    out["points"] = {int(lv): [[float(c) for c in p] for p in pts[lv]] for lv in pts.levels}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
