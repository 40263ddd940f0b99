"""Design-batch time and GEMM rates as a function of N (d=8, 1e6 candidates, batch 16, 15 restarts) on one GPU."""
import json, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from scipy.stats.qmc import LatinHypercube
import exauq_toolbox_b200 as exq
from exauq_toolbox_b200 import _lib
from exauq_toolbox_b200.design import design_batch
_lib.init(0)
d, M = 8, 1_000_000
cand = exq.DeviceCandidates(LatinHypercube(d, seed=2).random(M))
rows = []
for N in (512, 1024, 2048, 4096, 8192):
    X = LatinHypercube(d, seed=1).random(N)
    y = np.sum(np.sin(6 * np.pi * X) / np.arange(1, d + 1), axis=1) + X[:, 0] * X[:, 1] - math.sqrt(2.0)
    st = {}
    run = lambda: design_batch(X, y, "Matern52", [0.5] * d, 1.7, 0.0, [(0.0, 1.0)] * d, cand, batch_size=16, n_tries=15, seed=12345, stats=st)
    run()
    _lib.prof_enable(0b100011); _lib.prof_reset()
    st.clear(); _lib.timer_start(); run(); ms = _lib.timer_stop()
    pg, fg, ig = _lib.prof_get(0), _lib.prof_get(1), _lib.prof_get(5)
    _lib.prof_enable(0)
    rows.append({"N": N, "design_batch_ms": ms, "candidates_per_s": M / ms * 1e3, "loo_s": st["loo_s"], "map_s": st["map_s"], "pei_s": st["pei_s"],
                 "map_evals": st["map_evals"], "variance_gemm_tflops": pg["flops"] / max(pg["ms"], 1e-9) / 1e9,
                 "factor_gemm_tflops": fg["flops"] / max(fg["ms"], 1e-9) / 1e9,
                 "factor_gemm_int8_fp64_equivalent_tflops": ig["flops"] / max(ig["ms"], 1e-9) / 1e9 if ig["launches"] else None,
                 "variance_digits": _lib.variance_slices(), "factor_digits": _lib.factor_slices()})
    print(json.dumps(rows[-1]), flush=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "size_sweep.json"), "w"), indent=1)
