"""Parity of the CUDA path (through the C ABI) with the oracle and the golden vectors.
Tolerances are BASELINE.json's: mean / variance / LOO errors 1e-8 relative (the
reference's own rel-OR-abs 1e-9 rule near zero, exauq/core/numerics.py:31-73),
log-likelihood 1e-9 relative, identical selected indices."""

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu

FIXED = ["fixed_matern_d3_n40.json", "fixed_sqexp_d2_n30.json", "fixed_matern_d8_n150.json"]


@pytest.fixture(scope="module")
def exq():
    import exauq_toolbox_b200 as e

    e._lib.init()
    return e


@pytest.fixture(scope="module")
def orc():
    from oracle import gp_oracle

    return gp_oracle


def close(a, b, rel=1e-8, abs_=1e-9):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return bool(np.all(np.abs(a - b) <= np.maximum(rel * np.maximum(np.abs(a), np.abs(b)), abs_)))


def relerr(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def _device_fit(exq, g):
    hp = g["hyperparameters"]
    X, y = np.array(g["X"]), np.array(g["y"])
    corr_raw = -2 * np.log(np.array(hp["corr_length_scales"]))
    dev = exq.DeviceGP(X, y, g["kernel"]).fit(corr_raw, hp["process_var"], hp["nugget"])
    return X, y, corr_raw, dev


@pytest.mark.parametrize("name", FIXED)
def test_golden_fixed_hyperparameters(exq, name):
    """Every quantity the reference produced for this case, recomputed on the GPU."""
    g = load_golden(name)
    X, y, corr_raw, dev = _device_fit(exq, g)
    Xq = np.array(g["query"])
    # correlation (exauq/core/emulators.py:495-567)
    c = exq._lib.corr(g["kernel"], corr_raw, Xq[:5], X)
    assert np.max(np.abs(c - np.array(g["correlation_first5"]))) < 1e-13
    # predict (emulators.py:612-667)
    m, v = dev.predict(Xq)
    ref = np.array(g["predict"])
    assert close(m, ref[:, 0]) and close(v, ref[:, 1])
    # LOO sweep (designers.py:263-272): closed form and explicit batched refits
    loo = np.array(g["loo"])
    lm, lv, ln = dev.loo()
    assert close(lm, loo[:, 0]) and close(lv, loo[:, 1]) and close(ln, loo[:, 2])
    idx = np.arange(len(y))
    rm, rv, rn = dev.loo_refit(idx)
    assert close(rm, loo[:, 0]) and close(rv, loo[:, 1]) and close(rn, loo[:, 2])
    # kinv (modelling.py:1095-1113)
    ki = dev.kinv()
    assert np.allclose(np.diag(ki), g["kinv_diag"], rtol=1e-8)
    assert np.allclose(ki[0], g["kinv_row0"], rtol=1e-7, atol=1e-7 * np.max(np.abs(g["kinv_row0"])))
    assert np.allclose(ki, ki.T)
    # EI, repulsion, PEI (designers.py:522-706)
    cand = exq.DeviceCandidates(Xq)
    cand.pei_eval(dev, np.array(g["other_repulsion_points"]), g["max_target"])
    pei = np.array(g["pei"])[:, 2]
    got = cand.download("pei")
    assert np.allclose(got, pei, rtol=1e-6, atol=1e-14 * np.max(pei))
    assert cand.argmax()[0] == int(np.argmax(pei))
    # log-posterior and gradient (mogp_fitting.py:318-324) with the product's priors
    from exauq_toolbox_b200.priors import MapPriors

    pri = MapPriors(X)
    th = np.array(g["logpost_thetas"])
    val, grad, info, _ = dev.logpost_batch(th, g["hyperparameters"]["nugget"])
    assert np.all(info == 0)
    for r in range(th.shape[0]):
        ref_val, ref_grad = g["logpost"][r]
        assert abs((val[r] - pri.logp(th[r])) - ref_val) <= 1e-9 * abs(ref_val)
        gr = grad[r] - pri.dlogp_draw(th[r])
        assert np.allclose(gr, ref_grad, rtol=1e-7, atol=1e-8 * np.max(np.abs(ref_grad)))


def test_tutorial_prediction(exq):
    """docs/designers/tutorials/training_gp_tutorial.md:259-263 at the stored hyperparameters."""
    g = load_golden("tutorial_gp.json")
    f = g["fit"]
    X, y = np.array(g["inputs"]), np.array(g["outputs"])
    dev = exq.DeviceGP(X, y, "Matern52").fit(-2 * np.log(np.array(f["corr_length_scales"])), f["process_var"],
                                            f["nugget"])
    m, v = dev.predict(np.array([g["predict_at"]]))
    assert close(m[0], g["estimate"]) and close(v[0], g["variance"])
    assert abs(m[0] - g["published"]["estimate"]) < 5e-7 and abs(v[0] - g["published"]["variance"]) < 5e-7


@pytest.mark.parametrize("kernel,N,d", [("Matern52", 1, 2), ("Matern52", 2, 1), ("Matern52", 127, 3),
                                        ("Matern52", 128, 3), ("Matern52", 129, 4), ("SquaredExponential", 385, 6),
                                        ("Matern52", 1000, 8), ("ProductMat52", 200, 3)])
def test_seeded_sizes_against_oracle(exq, orc, kernel, N, d):
    """Ragged sizes around the 128 tile, one-point and two-point GPs."""
    X, y = orc.synthetic_problem(N, d) if N > 1 else (np.array([[0.3] * d]), np.array([0.7]))
    ell, cov = (0.5, 1.7) if kernel != "SquaredExponential" else (0.2, 0.9)
    nug = 0.0 if kernel != "SquaredExponential" else 1e-8
    corr_raw = np.full(d, -2 * np.log(ell))
    dev = exq.DeviceGP(X, y, kernel).fit(corr_raw, cov, nug)
    o = orc.fit(X, y, kernel, corr_raw, cov, nug)
    assert relerr(dev.neg_log_likelihood, orc.neg_log_likelihood(o)) < 1e-9
    Xc = orc.synthetic_candidates(333, d)
    m, v = dev.predict(Xc)
    mo, vo = orc.predict(o, Xc)
    assert close(m, mo) and close(v, vo)
    if N >= 2:
        lm, lv, ln = dev.loo()
        idx = np.unique(np.linspace(0, N - 1, 6).astype(int))
        om, ov, on = orc.loo_refit(X, y, kernel, corr_raw, cov, nug, idx)
        assert close(lm[idx], om) and close(lv[idx], ov) and close(ln[idx], on)
        rm, rv, rn = dev.loo_refit(idx)
        assert close(rm, om) and close(rv, ov) and close(rn, on)
    if True:
        th = np.concatenate([corr_raw, [np.log(cov)]])[None, :] + np.random.default_rng(5).uniform(-0.3, 0.3, (3, d + 1))
        val, grad, info, _ = dev.logpost_batch(th, nug)
        for r in range(3):
            ov_, og = orc.neg_log_likelihood_grad(X, y, kernel, th[r], nug)
            assert relerr(val[r], ov_) < 1e-9
            assert np.allclose(grad[r], og, rtol=1e-7, atol=1e-8 * np.max(np.abs(og)))


def test_fitted_nugget_logposterior_and_estimation(exq, orc):
    """nugget='fit' (mogp: the nugget is a (d+2)-th raw parameter with an inverse-gamma prior):
    value and gradient incl. d/d ln(nugget) vs the oracle, then a full MAP fit through the
    drop-in class vs the reference's MogpEmulator(nugget='fit') with the same numpy seed."""
    X, y = orc.synthetic_problem(120, 3)
    y = y + 0.05 * np.random.default_rng(0).standard_normal(y.shape)
    dev = exq.DeviceGP(X, y, "Matern52")
    rng = np.random.default_rng(8)
    th = np.concatenate([rng.uniform(-0.5, 1.5, (4, 3)), rng.uniform(-0.5, 0.5, (4, 1)), rng.uniform(-9, -3, (4, 1))], axis=1)
    val, grad, info, nug = dev.logpost_batch(th, fit_nugget=True)
    for r in range(4):
        ov, og = orc.neg_log_likelihood_grad(X, y, "Matern52", th[r], "fit")
        assert info[r] == 0 and abs(val[r] - ov) <= 1e-9 * abs(ov)
        assert np.allclose(grad[r], og, rtol=1e-7, atol=1e-8 * np.max(np.abs(og)))
        assert np.isclose(nug[r], np.exp(th[r, -1]))
    pytest.importorskip("exauq")
    from exauq.core.emulators import MogpEmulator
    from exauq.core.modelling import Input, TrainingDatum

    from exauq_toolbox_b200.gp import B200GaussianProcess

    data = [TrainingDatum(Input(*x), float(t)) for x, t in zip(X, y)]
    np.random.seed(2)
    ref = MogpEmulator(kernel="Matern52", nugget="fit")
    ref.fit(data)
    np.seterr(all="warn")
    np.random.seed(2)
    gp = B200GaussianProcess(kernel="Matern52", nugget="fit")
    gp.fit(data)
    a, b = gp.fit_hyperparameters, ref.fit_hyperparameters
    assert np.allclose(a.corr_length_scales, b.corr_length_scales, rtol=1e-4)
    assert np.isclose(a.process_var, b.process_var, rtol=1e-4) and np.isclose(a.nugget, b.nugget, rtol=1e-3)


def test_pei_batch_selects_same_indices(exq, orc):
    """Identical selected design-point indices over a batch (designers.py:830-835)."""
    N, d, M = 500, 4, 20000
    X, y = orc.synthetic_problem(N, d)
    corr_raw = np.full(d, -2 * np.log(0.4))
    dev = exq.DeviceGP(X, y, "Matern52").fit(corr_raw, 1.3, 0.0)
    o = orc.fit(X, y, "Matern52", corr_raw, 1.3, 0.0)
    Xc = orc.synthetic_candidates(M, d)
    rep = np.random.default_rng(2).uniform(0, 1, (24, d))
    cand = exq.DeviceCandidates(Xc)
    cand.pei_eval(dev, rep, float(y.max()))
    idx, val = cand.batch(dev, 8)
    oi, ov, ofinal = orc.pei_batch_argmax(o, "Matern52", Xc, rep, float(y.max()), 8)
    assert idx.tolist() == oi.tolist()
    assert np.allclose(val, ov, rtol=1e-8)
    assert np.allclose(cand.download("pei"), ofinal, rtol=1e-7, atol=1e-12 * np.max(ov))
    # host-driven variant: add points by coordinates (what the multi-GPU path does)
    cand2 = exq.DeviceCandidates(Xc)
    cand2.pei_eval(dev, rep, float(y.max()))
    picked = []
    for _ in range(8):
        i, _v = cand2.argmax()
        picked.append(i)
        cand2.add_point(dev, Xc[i])
    assert picked == oi.tolist()


def test_adaptive_nugget_and_not_positive_definite(exq, orc):
    """mogp's adaptive jitter: 0 when the factorisation succeeds; duplicated points make a
    fixed-nugget fit fail (LinAlgError) and an adaptive one add mean(diag) * 1e-6 * 10^k."""
    X, y = orc.synthetic_problem(60, 2)
    corr_raw = np.full(2, -2 * np.log(0.5))
    dev = exq.DeviceGP(X, y, "Matern52").fit(corr_raw, 1.7, 0.0, adaptive=True)
    assert dev.nugget_used == 0.0
    Xd = np.vstack([X, X[:3]])
    yd = np.concatenate([y, y[:3]])
    with pytest.raises(np.linalg.LinAlgError):
        exq.DeviceGP(Xd, yd, "Matern52").fit(corr_raw, 1.7, 0.0)
    dev = exq.DeviceGP(Xd, yd, "Matern52").fit(corr_raw, 1.7, 0.0, adaptive=True)
    o = orc.fit(Xd, yd, "Matern52", corr_raw, 1.7, "adaptive")
    assert dev.nugget_used > 0 and np.isclose(dev.nugget_used, o.theta.nugget)
    val, grad, info, nug = dev.logpost_batch(np.array([np.concatenate([corr_raw, [np.log(1.7)]])]), 0.0, adaptive=False)
    assert info[0] != 0 and np.isnan(val[0])


def test_full_size_properties(exq, orc):
    """N = 4096, d = 8 (BASELINE.json's metric size): size-independent properties.
    (1) interpolation: predicting at training inputs returns the targets with ~0 variance
    (tests/unit/test_emulators.py:811-821); (2) the closed-form LOO sweep agrees with explicit
    batched refits; (3) linearity of the predictive mean in y; (4) log-likelihood vs numpy."""
    N, d = 4096, 8
    X, y = orc.synthetic_problem(N, d)
    corr_raw = np.full(d, -2 * np.log(0.5))
    dev = exq.DeviceGP(X, y, "Matern52").fit(corr_raw, 1.7, 0.0)
    m, v = dev.predict(X[:512])
    assert np.allclose(m, y[:512], rtol=0, atol=1e-8) and np.all(v < 1e-8) and np.all(v >= 0)
    idx = np.array([0, 1, 127, 128, 2047, 4095])
    lm, lv, ln = dev.loo()
    rm, rv, rn = dev.loo_refit(idx)
    assert close(lm[idx], rm) and close(lv[idx], rv) and close(ln[idx], rn)
    Xc = orc.synthetic_candidates(2000, d)
    m1, v1 = dev.predict(Xc)
    y2 = np.cos(3 * X[:, 0]) - X[:, 2]
    dev2 = exq.DeviceGP(X, y2, "Matern52").fit(corr_raw, 1.7, 0.0)
    m2, v2 = dev2.predict(Xc)
    dev3 = exq.DeviceGP(X, 2.0 * y - 3.0 * y2, "Matern52").fit(corr_raw, 1.7, 0.0)
    m3, v3 = dev3.predict(Xc)
    assert np.allclose(m3, 2.0 * m1 - 3.0 * m2, rtol=1e-9, atol=1e-9)
    assert np.allclose(v1, v2, rtol=1e-12) and np.allclose(v1, v3, rtol=1e-12)
    o = orc.fit(X, y, "Matern52", corr_raw, 1.7, 0.0)
    assert relerr(dev.neg_log_likelihood, orc.neg_log_likelihood(o)) < 1e-9
    mo, vo = orc.predict(o, Xc[:200])
    assert close(m1[:200], mo) and close(v1[:200], vo)


def test_randomised_shapes_and_hyperparameters(exq, orc):
    """Seeded sweep over N (1..600, incl. non-multiples of the 64/128 tiles), d (1..11, odd and even),
    kernels, length scales and nuggets: fit / predict / LOO / log-likelihood+gradient / PEI vs the oracle."""
    rng = np.random.default_rng(2024)
    kernels = ["Matern52", "SquaredExponential", "ProductMat52"]
    for case in range(14):
        N = int(rng.choice([2, 3, 17, 63, 64, 65, 129, 200, 257, 383, 600]))
        d = int(rng.integers(1, 12))
        kernel = kernels[case % 3]
        X = rng.uniform(0, 1, (N, d))
        y = np.sin(3 * X).sum(axis=1) + 0.1 * rng.standard_normal(N)
        ell = rng.uniform(0.15, 0.6, d) if kernel != "SquaredExponential" else rng.uniform(0.08, 0.25, d)
        cov = float(rng.uniform(0.3, 3.0))
        nug = float(rng.choice([0.0, 1e-8, 1e-4])) if kernel != "SquaredExponential" else 1e-6
        corr_raw = -2 * np.log(ell)
        dev = exq.DeviceGP(X, y, kernel).fit(corr_raw, cov, nug)
        o = orc.fit(X, y, kernel, corr_raw, cov, nug)
        tag = f"case {case}: N={N} d={d} {kernel} nug={nug}"
        assert relerr(dev.neg_log_likelihood, orc.neg_log_likelihood(o)) < 1e-9, tag
        Xc = rng.uniform(0, 1, (int(rng.integers(1, 300)), d))
        m, v = dev.predict(Xc)
        mo, vo = orc.predict(o, Xc)
        assert close(m, mo) and close(v, vo), tag
        if N >= 2:
            idx = np.unique(rng.integers(0, N, 4))
            lm, lv, ln = dev.loo()
            om, ov, on = orc.loo_refit(X, y, kernel, corr_raw, cov, nug, idx)
            assert close(lm[idx], om) and close(lv[idx], ov) and close(ln[idx], on, rel=1e-7), tag
        th = np.concatenate([corr_raw, [np.log(cov)]])[None, :] + rng.uniform(-0.2, 0.2, (2, d + 1))
        val, grad, info, _ = dev.logpost_batch(th, nug)
        for r in range(2):
            ov_, og = orc.neg_log_likelihood_grad(X, y, kernel, th[r], nug)
            assert info[r] == 0 and relerr(val[r], ov_) < 1e-9, tag
            assert np.allclose(grad[r], og, rtol=1e-6, atol=1e-7 * max(1.0, np.max(np.abs(og)))), tag
        rep = rng.uniform(0, 1, (int(rng.integers(0, 9)), d))
        cand = exq.DeviceCandidates(Xc)
        cand.pei_eval(dev, rep, float(y.max()))
        po = orc.pei(o, kernel, Xc, rep, float(y.max()))
        assert np.allclose(cand.download("pei"), po, rtol=1e-6, atol=1e-13 * max(np.max(np.abs(po)), 1e-300)), tag
