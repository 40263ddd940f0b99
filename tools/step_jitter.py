"""Per-step times of the benchmark's design batch (resident candidates, then host candidates), to look for
outliers; run next to `nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.active --format=csv -lms 100`."""
import gc, math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
from scipy.stats.qmc import LatinHypercube
import exauq_toolbox_b200 as exq
from exauq_toolbox_b200 import _lib
from exauq_toolbox_b200.design import design_batch
_lib.init(0)
N, d, M = 4096, 8, 1_000_000
X = LatinHypercube(d, seed=1).random(N)
y = np.sum(np.sin(6 * np.pi * X) / np.arange(1, d + 1), axis=1) + X[:, 0] * X[:, 1] - math.sqrt(2.0)
Xc = LatinHypercube(d, seed=2).random(M)
res = exq.DeviceCandidates(Xc)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
gc_log = []
def _gc_cb(phase, info, _t=[0.0]):
    if phase == "start":
        _t[0] = time.perf_counter()
    else:
        gc_log.append((info["generation"], time.perf_counter() - _t[0]))
gc.callbacks.append(_gc_cb)
for mode, cand in (("resident", res), ("host", Xc)):
    for s in range(n):
        st = {}
        t0 = time.time()
        _lib.timer_start()
        design_batch(X, y, "Matern52", [0.5] * d, 1.7, 0.0, [(0.0, 1.0)] * d, cand, batch_size=16, n_tries=15, seed=12345, stats=st)
        t_ret = time.perf_counter()
        ms = _lib.timer_stop()
        t_stop = time.perf_counter()
        print("%s step %2d: %.1f ms  map %.3f (gpu %.3f) pei %.3f (eval call %.3f) loo %.3f | after phases: return %.3f s, timer_stop %.3f s | gc %s" % (
            mode, s, ms, st["map_s"], st["map_gpu_s"], st["pei_s"], st["pei_eval_call_s"], st["loo_s"],
            t_ret - st["t_end"], t_stop - t_ret, [(g_, round(t_, 3)) for g_, t_ in gc_log if t_ > 0.005]), flush=True)
        gc_log.clear()
