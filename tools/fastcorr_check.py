import os, sys, math, subprocess, json
import numpy as np
sys.path.insert(0, os.getcwd())
def run():
    import exauq_toolbox_b200 as exq
    from exauq_toolbox_b200 import _lib
    from scipy.stats.qmc import LatinHypercube
    _lib.init(0)
    out = {}
    for kern in ("Matern52", "SquaredExponential", "ProductMat52"):
        d = 8
        A = LatinHypercube(d, seed=1).random(3000); B = LatinHypercube(d, seed=2).random(2000)
        raw = np.linspace(-1.0, 3.0, d)
        c = _lib.corr(kern, raw, A, B)
        out[kern] = c
    # timing of the cross assembly + gradient at bench size
    N, d, M = 4096, 8, 131072
    X = LatinHypercube(d, seed=1).random(N); y = np.sum(np.sin(6*np.pi*X)/np.arange(1, d+1), axis=1)
    g = exq.DeviceGP(X, y, "Matern52").fit(np.full(d, -2*math.log(0.5)), 1.7, 0.0)
    cand = exq.DeviceCandidates(LatinHypercube(d, seed=2).random(M))
    cand.pei_eval(g, np.zeros((0, d)), float(y.max()))
    th = np.tile(np.concatenate([np.full(d, -2*math.log(0.5)), [math.log(1.7)]]), (15, 1))
    g.logpost_batch(th, 0.0)
    _lib.prof_enable(0b1011000); _lib.prof_reset()
    cand.pei_eval(g, np.zeros((0, d)), float(y.max()))
    val, grad, info, _ = g.logpost_batch(th, 0.0)
    pa, ps = _lib.prof_get(3), _lib.prof_get(4)
    _lib.prof_enable(0)
    out["t_assemble_ms"] = pa["ms"]; out["n_assemble"] = pa["launches"]; out["t_stream_ms"] = ps["ms"]
    out["val"] = val; out["grad"] = grad
    m, v = cand.download("mean"), cand.download("var")
    out["mean"] = m; out["var"] = v
    return out
if len(sys.argv) > 1:
    o = run(); np.savez("/tmp/fc_%s.npz" % sys.argv[1], **o); sys.exit(0)
for tag, env in (("slow", "0"), ("fast", "1")):
    subprocess.run([sys.executable, __file__, tag], env=dict(os.environ, EXQ_FAST_CORR=env), check=True)
a, b = np.load("/tmp/fc_slow.npz"), np.load("/tmp/fc_fast.npz")
for k in ("Matern52", "SquaredExponential", "ProductMat52", "mean", "var", "val", "grad"):
    den = np.maximum(np.abs(a[k]), 1e-300)
    print(k, "max abs diff %.3e  max rel diff %.3e" % (np.max(np.abs(a[k]-b[k])), np.max(np.abs(a[k]-b[k])/den)))
print("assemble ms: slow %.3f fast %.3f (%d launches); stream-class ms (incl. grad kernel, same in both): %.3f / %.3f" % (a["t_assemble_ms"], b["t_assemble_ms"], int(a["n_assemble"]), a["t_stream_ms"], b["t_stream_ms"]))
