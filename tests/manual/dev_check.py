"""Manual parity dump run on a GPU box: every entry point of the CUDA path vs the oracle on small
seeded cases (python tests/manual/dev_check.py).  Lives under tests/ because it imports the oracle."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import gp_oracle as orc
import exauq_toolbox_b200 as exq
from exauq_toolbox_b200 import _lib

def rel(a, b):
    a = np.asarray(a); b = np.asarray(b)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))

def relabs(a, b, tol=1e-9):
    a = np.asarray(a); b = np.asarray(b)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0)))

_lib.init(0)
for kernel in ["Matern52", "SquaredExponential"]:
    for (N, d) in [(50, 2), (128, 3), (300, 8), (700, 5)]:
        X, y = orc.synthetic_problem(N, d)
        ell = 0.5 if kernel == "Matern52" else 0.25
        corr_raw = np.full(d, -2 * np.log(ell)); cov = 1.7; nug = 1e-8 if kernel == "SquaredExponential" else 0.0
        t0 = time.time()
        C = _lib.corr(kernel, corr_raw, X[:40], X)
        print(kernel, N, d, "corr maxabs", np.max(np.abs(C - orc.corr(kernel, corr_raw, X[:40], X))))
        g = exq.DeviceGP(X, y, kernel).fit(corr_raw, cov, nug)
        o = orc.fit(X, y, kernel, corr_raw, cov, nug)
        print("   nll", g.neg_log_likelihood, orc.neg_log_likelihood(o), "rel", rel(g.neg_log_likelihood, orc.neg_log_likelihood(o)))
        Xc = orc.synthetic_candidates(1000, d)
        m, v = g.predict(Xc)
        mo, vo = orc.predict(o, Xc)
        print("   predict mean relabs", relabs(m, mo), "var relabs", relabs(v, vo), "var rel", rel(v, vo))
        lm, lv, ln = g.loo()
        idx = np.arange(0, N, max(1, N // 8))
        om, ov, on = orc.loo_refit(X, y, kernel, corr_raw, cov, nug, idx)
        print("   loo closed-form vs oracle refit: mean", relabs(lm[idx], om), "var", rel(lv[idx], ov), "nes", rel(ln[idx], on))
        rm, rv, rn = g.loo_refit(idx)
        print("   loo gpu-refit  vs oracle refit: mean", relabs(rm, om), "var", rel(rv, ov), "nes", rel(rn, on))
        ki = g.kinv()
        ko = orc.kinv(o, kernel)
        print("   kinv rel-fro", np.linalg.norm(ki - ko) / np.linalg.norm(ko))
        # logpost batch
        rng = np.random.default_rng(3)
        th = np.concatenate([np.tile(corr_raw, (4, 1)) + rng.uniform(-0.5, 0.5, (4, d)), np.log(cov) + rng.uniform(-0.5, 0.5, (4, 1))], axis=1)
        val, grad, info, nu = g.logpost_batch(th, nug)
        for r in range(4):
            ov_, og = orc.neg_log_likelihood_grad(X, y, kernel, th[r], nug)
            print("   logpost r%d" % r, "val rel", rel(val[r], ov_), "grad relmax", float(np.max(np.abs(grad[r] - og)) / np.max(np.abs(og))), info[r])
        # PEI
        rep = rng.uniform(0, 1, (20, d))
        cand = exq.DeviceCandidates(Xc)
        tmax = float(np.max(y))
        cand.pei_eval(g, rep, tmax)
        p = cand.download("pei")
        po = orc.pei(o, kernel, Xc, rep, tmax)
        print("   pei relmax", float(np.max(np.abs(p - po)) / np.max(np.abs(po))), "argmax", cand.argmax()[0], int(np.argmax(po)))
        bi, bv = cand.batch(g, 5)
        oi, ovv, _ = orc.pei_batch_argmax(o, kernel, Xc, rep, tmax, 5)
        print("   batch idx", bi.tolist(), oi.tolist(), "vals rel", rel(bv, ovv))
        print("   elapsed", time.time() - t0)
print("launches", _lib.launch_count())
