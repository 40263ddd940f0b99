"""Times BASELINE.json's configs on this GPU (device-resident inputs, CUDA events on the library
stream) and writes gpurun_out/config_timings.json.  Evidence for SURVEY.md section 8 coverage; the
bench line itself is bench.py."""
import json, math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from scipy.stats.qmc import LatinHypercube
import exauq_toolbox_b200 as exq
from exauq_toolbox_b200 import _lib
from exauq_toolbox_b200.design import design_batch, multi_level_design_batch

_lib.init(0)
out = {}


def target(X):
    d = X.shape[1]
    return np.sum(np.sin(6 * np.pi * X) / np.arange(1, d + 1), axis=1) + X[:, 0] * X[:, 1 % d] - math.sqrt(2.0)


def timed(fn, reps=2):
    fn()
    best = 1e30
    for _ in range(reps):
        _lib.timer_start()
        r = fn()
        best = min(best, _lib.timer_stop())
    return best, r


# configs[1]: N=2048, d=8, batch 16, 1e6 candidates
N, d, M = 2048, 8, 1_000_000
X = LatinHypercube(d, seed=1).random(N); y = target(X)
cand = exq.DeviceCandidates(LatinHypercube(d, seed=2).random(M))
st = {}
ms, r = timed(lambda: design_batch(X, y, "Matern52", [0.5] * d, 1.7, 0.0, [(0.0, 1.0)] * d, cand, batch_size=16, n_tries=15, seed=12345, stats=st))
out["config1_design_batch_N2048_d8_M1e6_batch16"] = {"ms": ms, "candidates_per_s": M / ms * 1e3, "phases_s": {k: st[k] for k in ("loo_s", "map_s", "pei_s")}, "map_evals": st["map_evals"]}
cand.close()

# configs[2]: LOO sweep N=8192, d=10 (closed form; plus explicit batched refits of 16 left-out points)
N, d = 8192, 10
X = LatinHypercube(d, seed=1).random(N); y = target(X)
g = exq.DeviceGP(X, y, "Matern52")
raw = np.full(d, -2 * math.log(0.6))
ms, _ = timed(lambda: (g.fit(raw, 1.7, 0.0), g.loo()))
ms_r, _ = timed(lambda: g.loo_refit(np.arange(16)), reps=1)
out["config2_loo_sweep_N8192_d10"] = {"fit_plus_closed_form_sweep_ms": ms, "loo_predictions_per_s": N / ms * 1e3,
                                      "explicit_refit_16_systems_ms": ms_r, "explicit_refits_per_s": 16 / ms_r * 1e3,
                                      "explicit_refit_tflops": 16 * (2.0 / 3.0) * 8191.0 ** 3 / (ms_r * 1e-3) / 1e12}
g.close()

# configs[3]: 3-level, N=4096/1024/256, d=6, costs 1/10/100, batch 8 (1e5 candidates)
d = 6
levels = {}
for lv, n, ell, cov in [(1, 4096, 0.6, 2.0), (2, 1024, 0.5, 0.7), (3, 256, 0.7, 0.2)]:
    Xl = LatinHypercube(d, seed=10 + lv).random(n)
    levels[lv] = dict(X=Xl, y=target(Xl) / lv, kernel="Matern52", corr_length_scales=[ell] * d, process_var=cov, nugget=0.0)
Xc = LatinHypercube(d, seed=2).random(100_000)
st = {}
t0 = time.perf_counter()
chosen = multi_level_design_batch(levels, {1: 1.0, 2: 1.0, 3: 1.0}, {1: 1.0, 2: 10.0, 3: 100.0}, [(0.0, 1.0)] * d, Xc, batch_size=8, n_tries=15, seed=2, stats=st)
out["config3_multilevel_N4096_1024_256_d6_batch8_M1e5"] = {"wall_s": time.perf_counter() - t0, "chosen_levels": [c[0] for c in chosen], "map_evals": st.get("map_evals")}

# configs[4]: 128 restarts... 16 per GPU: one batched log-posterior + gradient evaluation at N=4096, d=20
N, d, R = 4096, 20, 16
X = LatinHypercube(d, seed=1).random(N); y = target(X)
g = exq.DeviceGP(X, y, "Matern52")
th = np.random.default_rng(3).uniform(-2.5, 2.5, (R, d + 1)); th[:, :d] = np.clip(th[:, :d], -1.0, 2.0)
ms, r = timed(lambda: g.logpost_batch(th, 0.0, adaptive=True))
out["config4_logpost_grad_N4096_d20_R16_per_gpu"] = {"ms": ms, "evals_per_s": R / ms * 1e3, "tflops_N3_per_eval": R * 4096.0 ** 3 / (ms * 1e-3) / 1e12, "failed": int(np.sum(r[2] != 0))}
print(json.dumps(out, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "config_timings.json"), "w"), indent=1)
