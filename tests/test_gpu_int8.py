"""The int8 tcgen05 (digit-sliced) arithmetic against the FP64 DMMA path on identical inputs.

Both compute the same FP64 quantities of the reference's path (mogp fit / predict / logposterior /
logpost_deriv behind exauq/core/emulators.py:418-450, 612-667 and
exauq/utilities/mogp_fitting.py:289-346); the int8 kernels rebuild every FP64 product exactly from
base-256 digit products and must stay far inside the 1e-8 / 1e-9 parity tolerances of
BASELINE.json, so they are compared (a) with the DMMA kernels at full size and (b) with the CPU
oracle at a size the oracle finishes in seconds."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def exq():
    import exauq_toolbox_b200 as m

    m._lib.init(0)
    yield m
    m._lib.set_int8_digits(6, 7)   # library defaults


@pytest.fixture(scope="module")
def orc():
    from oracle import gp_oracle

    return gp_oracle


def _relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def _evaluate(exq, X, y, Xc, corr_raw, thetas):
    dev = exq.DeviceGP(X, y, "Matern52").fit(corr_raw, 1.7, 0.0)
    out = {"nll": dev.neg_log_likelihood}
    out["mean"], out["var"] = dev.predict(Xc)
    out["loo"] = dev.loo()
    out["val"], out["grad"], out["info"], _ = dev.logpost_batch(thetas, 0.0)
    return out


def test_defaults_are_int8(exq):
    assert exq._lib.variance_slices() == 6
    assert exq._lib.factor_slices() == 7


@pytest.mark.parametrize("n,d,r", [(4096, 8, 6), (2048, 5, 16)])
def test_int8_matches_dmma_at_full_size(exq, orc, n, d, r):
    """Fit, predict, LOO sweep and a batch of log-posterior + gradient evaluations: int8 digits
    (variance 6 / factorisation 7, and the 7 / 6 combination) against FP64 DMMA."""
    X, y = orc.synthetic_problem(n, d)
    Xc = orc.synthetic_candidates(20000, d)
    corr_raw = np.full(d, -2 * np.log(0.5))
    rng = np.random.default_rng(5)
    thetas = np.concatenate([corr_raw + rng.uniform(-0.7, 0.7, (r, d)), np.log(1.7) + rng.uniform(-0.5, 0.5, (r, 1))], axis=1)

    exq._lib.set_int8_digits(0, 0)
    ref = _evaluate(exq, X, y, Xc, corr_raw, thetas)
    assert np.all(ref["info"] == 0)
    # defaults (variance 6 digits, factorisation 7) far inside BASELINE.json's tolerances; 6 digits in the
    # factorisation only just meet them (log-likelihood 1e-10, mean 5e-10), which is why 7 is the default
    # (with the nodes of 1024 rows on the int8 kernel too -- three levels of 7-digit products at n = 4096 -- the
    # log-likelihood sits 3e-12 from the FP64 DMMA value; BASELINE.json asks for 1e-9)
    for vd, fd, tol_nll, tol_mean, tol_var, tol_grad in [(6, 7, 1e-11, 1e-9, 1e-9, 1e-8), (7, 6, 1e-9, 1e-8, 1e-8, 1e-6)]:
        exq._lib.set_int8_digits(vd, fd)
        got = _evaluate(exq, X, y, Xc, corr_raw, thetas)
        assert np.all(got["info"] == 0)
        assert abs(got["nll"] - ref["nll"]) <= tol_nll * abs(ref["nll"]), (vd, fd, got["nll"], ref["nll"])
        assert _relerr(got["val"], ref["val"]) <= tol_nll, (vd, fd)
        assert np.all(np.abs(got["mean"] - ref["mean"]) <= tol_mean * np.maximum(np.abs(ref["mean"]), 1.0)), (vd, fd)
        assert np.all(np.abs(got["var"] - ref["var"]) <= np.maximum(tol_var * ref["var"], 1e-11)), (
            vd, fd, float(np.max(np.abs(got["var"] - ref["var"]))))
        for a, b in zip(got["loo"], ref["loo"]):
            assert np.all(np.abs(a - b) <= np.maximum(tol_mean * np.abs(b), 1e-9)), (vd, fd)
        gscale = np.max(np.abs(ref["grad"]), axis=1, keepdims=True)
        assert np.all(np.abs(got["grad"] - ref["grad"]) <= tol_grad * gscale), (
            vd, fd, float(np.max(np.abs(got["grad"] - ref["grad"]) / gscale)))


def test_int8_batch_matches_oracle(exq, orc):
    """Np = 2048 with 16 systems is the smallest batch that reaches the int8 GEMMs; the oracle's
    log-posterior and gradient (mogp logposterior / logpost_deriv) are the reference."""
    n, d, r = 2048, 3, 16
    X, y = orc.synthetic_problem(n, d)
    corr_raw = np.full(d, -2 * np.log(0.3))
    rng = np.random.default_rng(11)
    thetas = np.concatenate([corr_raw + rng.uniform(-0.5, 0.5, (r, d)), np.log(1.2) + rng.uniform(-0.3, 0.3, (r, 1))], axis=1)
    exq._lib.set_int8_digits(6, 7)
    dev = exq.DeviceGP(X, y, "Matern52")
    val, grad, info, _ = dev.logpost_batch(thetas, 1e-8)
    assert np.all(info == 0)
    for k in (0, 7, 15):
        oval, ograd = orc.neg_log_likelihood_grad(X, y, "Matern52", thetas[k], 1e-8)
        assert abs(val[k] - oval) <= 1e-9 * abs(oval)
        assert np.all(np.abs(grad[k] - ograd) <= 1e-7 * np.max(np.abs(ograd)))


def test_int8_variance_near_training_points(exq, orc):
    """Candidates a hair away from training points: the variance is ~1e-5 of the prior and is the
    difference of two nearly equal numbers; the reference's rule (1e-8 relative or 1e-9 absolute,
    exauq/core/numerics.py:31-73) must hold against the oracle."""
    n, d = 1500, 4
    X, y = orc.synthetic_problem(n, d)
    rng = np.random.default_rng(2)
    Xc = np.clip(X[rng.integers(0, n, 3000)] + rng.normal(0, 1e-3, (3000, d)), 0, 1)
    corr_raw = np.full(d, -2 * np.log(0.4))
    exq._lib.set_int8_digits(6, 7)
    dev = exq.DeviceGP(X, y, "Matern52").fit(corr_raw, 1.7, 0.0)
    m, v = dev.predict(Xc)
    ref = orc.fit(X, y, "Matern52", corr_raw, 1.7, 0.0)
    mo, vo = orc.predict(ref, Xc)
    assert np.all(np.abs(m - mo) <= np.maximum(1e-8 * np.abs(mo), 1e-9))
    assert np.all(np.abs(v - vo) <= np.maximum(1e-8 * np.abs(vo), 1e-9))
