"""Errors of the int8 (digit-sliced tcgen05) paths against the CPU oracle over kernels of increasing
conditioning -- the measurement behind the thresholds of the conditioning guard (csrc/exq_api.cu:
guard_widen / guard_fallback) and behind tests/test_gpu_illcond.py.

For every case the oracle (numpy/scipy restatement of mogp, oracle/gp_oracle.py) gives the negative
log-likelihood, predictions at 3000 points (half of them a hair away from training points), the
closed-form LOO sweep, the gradient and K^-1; the GPU evaluates the same in four arithmetic modes:

  dmma        FP64 DMMA everywhere                     (exq_set_int8_digits(0, 0))
  i8_6_7      variance 6 digits, factorisation 7, LAUUM 6, guard off
  i8_7_7      variance 7 digits, factorisation 7, LAUUM 7 (guard forced to "widen"), no fallback
  default     library defaults: guard on, one step of iterative refinement of alpha after an int8 factorisation
  (the three modes above run with the refinement off, as in round 1)

Run on the GPU box:  python tools/guard_survey.py > gpurun_out/guard_survey.json
"""

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [
    # kernel, N, d, length scale, nugget (float or "adaptive")
    ("Matern52", 2304, 8, 0.5, 0.0),
    ("SquaredExponential", 2304, 8, 0.3, 0.0),
    ("Matern52", 2304, 8, 2.0, 0.0),
    ("SquaredExponential", 2304, 8, 0.6, 1e-8),
    ("Matern52", 2304, 5, 1.0, 0.0),
    ("ProductMat52", 2304, 4, 0.5, 0.0),
    ("Matern52", 2304, 3, 0.5, 0.0),
    ("SquaredExponential", 2304, 3, 0.2, 1e-8),
    ("SquaredExponential", 2304, 8, 0.6, "adaptive"),
    ("Matern52", 4096, 8, 0.5, 0.0),
    ("Matern52", 4096, 8, 2.0, 0.0),
]


def rel(a, b, floor=0.0):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor if floor else 1e-300)))


def main():
    import exauq_toolbox_b200 as exq
    from oracle import gp_oracle as orc

    lib = exq._lib
    lib.init(0)
    out = []
    for kernel, N, d, ell, nug in CASES:
        X, y = orc.synthetic_problem(N, d)
        rng = np.random.default_rng(3)
        Xc = np.vstack([orc.synthetic_candidates(1500, d),
                        np.clip(X[rng.integers(0, N, 1500)] + rng.normal(0, 1e-3, (1500, d)), 0, 1)])
        raw = np.full(d, -2 * np.log(ell))
        t0 = time.time()
        ref = orc.fit(X, y, kernel, raw, 1.7, nug)
        o = {"nll": orc.neg_log_likelihood(ref)}
        o["mean"], o["var"] = orc.predict(ref, Xc)
        nug_used = float(ref.theta.nugget) if ref.theta.nugget is not None else 0.0
        K = 1.7 * orc.corr(kernel, raw, X, X) + nug_used * np.eye(N)
        Kinv = np.linalg.inv(K)
        alpha = Kinv @ y
        o["loo_mean"] = y - alpha / np.diag(Kinv)
        o["loo_var"] = 1.0 / np.diag(Kinv)
        th = np.concatenate([raw, [np.log(1.7)]])
        o["val"], o["grad"] = orc.neg_log_likelihood_grad(X, y, kernel, th, nug)
        o["kinv"] = orc.kinv(ref, kernel) if N <= 2304 else None
        ev = np.linalg.eigvalsh(K)
        cond = float(ev[-1] / max(ev[0], 1e-300))
        t_or = time.time() - t0
        row = {"kernel": kernel, "N": N, "d": d, "ell": ell, "nugget": nug, "nugget_used_oracle": nug_used, "cond2": cond,
               "oracle_s": round(t_or, 1), "modes": {}}
        for mode in ("dmma", "i8_6_7", "i8_7_7", "default"):
            lib.set_refine_steps(1 if mode == "default" else 0)
            if mode == "dmma":
                lib.set_int8_digits(0, 0); lib.set_guard(1e300, 1e300)
            elif mode == "i8_6_7":
                lib.set_int8_digits(6, 7); lib.set_guard(1e300, 1e300)
            elif mode == "i8_7_7":
                lib.set_int8_digits(6, 7); lib.set_guard(1.0, 1e300)
            else:
                lib.set_int8_digits(6, 7); lib.set_guard()
            w0, f0 = lib.guard_counts()
            dev = exq.DeviceGP(X, y, kernel).fit(raw, 1.7, 0.0 if nug == "adaptive" else nug, nug == "adaptive")
            kappa, on64 = dev.kappa()
            m, v = dev.predict(Xc)
            lm, lv, _ = dev.loo()
            val, grad, info, nu = dev.logpost_batch(th[None, :], 0.0 if nug == "adaptive" else nug, nug == "adaptive")
            r = {"kappa": kappa, "fit_on_fp64": on64, "nugget_used": dev.nugget_used,
                 "nll_rel": abs(dev.neg_log_likelihood - o["nll"]) / abs(o["nll"]),
                 "mean_err": float(np.max(np.abs(m - o["mean"]) / np.maximum(np.abs(o["mean"]), 1.0))),
                 "var_abs": float(np.max(np.abs(v - o["var"]))),
                 "var_rel_or_abs": float(np.max(np.abs(v - o["var"]) / np.maximum(np.abs(o["var"]), 1e-1))),
                 "loo_mean_err": float(np.max(np.abs(lm - o["loo_mean"]) / np.maximum(np.abs(o["loo_mean"]), 1.0))),
                 "loo_var_rel": rel(lv, o["loo_var"]),
                 "val_rel": abs(val[0] - o["val"]) / abs(o["val"]),
                 "grad_rel_max": float(np.max(np.abs(grad[0] - o["grad"])) / np.max(np.abs(o["grad"]))),
                 "info": int(info[0])}
            if o["kinv"] is not None:
                try:
                    ki = dev.kinv()
                    r["kinv_rel_max"] = float(np.max(np.abs(ki - o["kinv"])) / np.max(np.abs(o["kinv"])))
                except Exception as exc:  # singular by the library's estimate
                    r["kinv_error"] = str(exc)[:120]
            w1, f1 = lib.guard_counts()
            r["guard_widened"], r["guard_fallbacks"] = w1 - w0, f1 - f0
            row["modes"][mode] = r
            dev.close()
        out.append(row)
        print(json.dumps(row), file=sys.stderr, flush=True)
    lib.set_int8_digits(6, 7)
    lib.set_guard()
    lib.set_refine_steps(1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
