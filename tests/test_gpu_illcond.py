"""The DEFAULT arithmetic (int8 tcgen05 digit products: factorisation 7 digits, variance / LAUUM 6, conditioning
guard on, one refinement step of alpha) against the CPU ORACLE -- not against the FP64 DMMA path -- where the
int8 path can break: squared-exponential and long-length-scale kernels, a 1e-8 nugget rescuing a near-singular
K, the adaptive nugget, the product kernel, at sizes where a single fit reaches the int8 GEMMs (Np >= 2304) and
the variance / LAUUM kernels (any Np).   VERDICT r01 "next round" item 1 / ADVICE r01 (int8 exactness).

Tolerances.  BASELINE.json: predictive mean / variance and LOO within 1e-8 relative (or 1e-9 absolute near zero,
the reference's own rule exauq/core/numerics.py:31-73), log-likelihood 1e-9 relative, gradient 1e-7 of its
largest component (the reference compares estimated hyperparameters at 1e-5).  Two FP64 implementations of the
same formulas differ by ~cond(K) * 1e-16, so for cond(K) >~ 1e8 no implementation can meet 1e-8 against another;
there the bound is 4 x the error of this library's own all-FP64 (DMMA) path against the same oracle, measured
in the same test -- i.e. "int8 is as good as FP64 on this chip", which is the claim being tested.
The measurements behind the guard thresholds: tools/guard_survey.py -> profiles/guard_survey_r02*.json.
"""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

# kernel, N, d, length scale, nugget, regime
CASES = [
    ("Matern52", 2304, 8, 2.0, 0.0),                    # cond 4e7, kappa 8e3
    ("SquaredExponential", 2304, 8, 0.6, 1e-8),         # class default kernel, small fixed nugget, cond 8e6
    ("SquaredExponential", 2304, 8, 0.6, "adaptive"),   # class defaults: kernel AND nugget handling
    ("ProductMat52", 2304, 4, 0.5, 0.0),                # cond 2e9, kappa 5e5: guard widens the 6-digit products
    ("Matern52", 2304, 3, 0.5, 0.0),                    # cond 1e10, kappa 6e6
    ("SquaredExponential", 2304, 3, 0.2, 1e-8),         # nugget-rescued near-singular K: cond 3e10, kappa 1e8
]


@pytest.fixture(scope="module")
def exq():
    import exauq_toolbox_b200 as m

    m._lib.init(0)
    yield m
    m._lib.set_int8_digits(6, 7)
    m._lib.set_guard()
    m._lib.set_refine_steps(1)


@pytest.fixture(scope="module")
def orc():
    from oracle import gp_oracle

    return gp_oracle


def _errors(dev, o, Xc, th, nug):
    adaptive = nug == "adaptive"
    m, v = dev.predict(Xc)
    lm, lv, _ = dev.loo()
    val, grad, info, _ = dev.logpost_batch(th[None, :], 0.0 if adaptive else nug, adaptive)
    assert info[0] == 0
    return {
        "nll": abs(dev.neg_log_likelihood - o["nll"]) / abs(o["nll"]),
        "val": abs(val[0] - o["val"]) / abs(o["val"]),
        # the reference's closeness rule: relative, or 1e-9 absolute near zero -> error in units of the tolerance
        "mean": float(np.max(np.abs(m - o["mean"]) / np.maximum(np.abs(o["mean"]), 0.1))),
        "var": float(np.max(np.abs(v - o["var"]) / np.maximum(np.abs(o["var"]), 0.1))),
        "loo_mean": float(np.max(np.abs(lm - o["loo_mean"]) / np.maximum(np.abs(o["loo_mean"]), 0.1))),
        "loo_var": float(np.max(np.abs(lv - o["loo_var"]) / np.abs(o["loo_var"]))),
        "grad": float(np.max(np.abs(grad[0] - o["grad"])) / np.max(np.abs(o["grad"]))),
    }


BASE_TOL = {"nll": 1e-9, "val": 1e-9, "mean": 1e-8, "var": 1e-8, "loo_mean": 1e-8, "loo_var": 1e-8, "grad": 1e-7}


@pytest.mark.parametrize("kernel,N,d,ell,nug", CASES)
def test_default_int8_arithmetic_against_oracle(exq, orc, kernel, N, d, ell, nug):
    lib = exq._lib
    X, y = orc.synthetic_problem(N, d)
    rng = np.random.default_rng(3)
    Xc = np.vstack([orc.synthetic_candidates(1500, d),
                    np.clip(X[rng.integers(0, N, 1500)] + rng.normal(0, 1e-3, (1500, d)), 0, 1)])
    raw = np.full(d, -2 * np.log(ell))
    th = np.concatenate([raw, [np.log(1.7)]])
    ref = orc.fit(X, y, kernel, raw, 1.7, nug)
    nug_used = float(ref.theta.nugget)
    o = {"nll": orc.neg_log_likelihood(ref)}
    o["mean"], o["var"] = orc.predict(ref, Xc)
    Kinv = np.linalg.inv(1.7 * orc.corr(kernel, raw, X, X) + nug_used * np.eye(N))
    o["loo_mean"] = y - (Kinv @ y) / np.diag(Kinv)      # closed form of the reference's N refits (tests/test_oracle_golden.py)
    o["loo_var"] = 1.0 / np.diag(Kinv)
    o["val"], o["grad"] = orc.neg_log_likelihood_grad(X, y, kernel, th, nug)

    adaptive = nug == "adaptive"
    fit = lambda: exq.DeviceGP(X, y, kernel).fit(raw, 1.7, 0.0 if adaptive else nug, adaptive)  # noqa: E731
    lib.set_int8_digits(0, 0)
    dev = fit()
    e64 = _errors(dev, o, Xc, th, nug)
    dev.close()
    lib.set_int8_digits(6, 7)
    lib.set_guard()
    lib.set_refine_steps(1)
    dev = fit()
    assert abs(dev.nugget_used - nug_used) <= 1e-12 * max(nug_used, 1e-300)      # same jitter decision as mogp
    kappa, on64 = dev.kappa()
    assert not on64 and kappa < 1e12                    # none of these cases is sent to the FP64 fallback
    e8 = _errors(dev, o, Xc, th, nug)
    dev.close()
    for q, tol in BASE_TOL.items():
        bound = max(tol, 4.0 * e64[q])
        assert e8[q] <= bound, (q, e8[q], e64[q], kappa)


def test_batched_evaluation_keeps_ill_conditioned_systems_on_the_top_levels(exq, orc):
    """Inside a batched log-posterior evaluation every level of the recursion down to the 1024-row nodes runs on 7-digit
    int8 products (and batches that do not fill the chip too).  At kappa ~ 1e8 that leaves the value 5e-6 from the
    oracle where FP64 is at 2e-8, so systems whose kappa reaches the guard's widening threshold are re-factored with
    the int8 kernel on the top levels only.  A batch of three such systems and three benign ones: the guarded values
    meet the tolerance, the unguarded ones of the ill-conditioned systems do not, and the guard counts its work."""
    lib = exq._lib
    kernel, N, d, nug = "SquaredExponential", 2304, 3, 1e-8
    X, y = orc.synthetic_problem(N, d)
    rng = np.random.default_rng(12)
    ells = np.array([0.2, 0.21, 0.19, 0.04, 0.045, 0.05])
    thetas = np.column_stack([np.tile(-2 * np.log(ells)[:, None], (1, d)) + rng.uniform(-0.02, 0.02, (6, d)),
                              np.log(1.7) + rng.uniform(-0.1, 0.1, 6)])
    checked = (0, 3)                                           # one ill-conditioned, one benign system against the oracle
    want = {b: orc.neg_log_likelihood_grad(X, y, kernel, thetas[b], nug) for b in checked}

    def errors(val, grad):
        return {b: (abs(val[b] - want[b][0]) / abs(want[b][0]),
                    float(np.max(np.abs(grad[b] - want[b][1])) / np.max(np.abs(want[b][1])))) for b in checked}

    lib.set_int8_digits(0, 0)                                  # (the digit setting applies to GPs created afterwards)
    dev = exq.DeviceGP(X, y, kernel)
    val64, grad64, info64, _ = dev.logpost_batch(thetas, nug)
    dev.close()
    e64 = errors(val64, grad64)
    lib.set_int8_digits(6, 7)
    lib.set_guard()
    dev = exq.DeviceGP(X, y, kernel)
    w0, f0 = lib.guard_counts()
    val, grad, info, _ = dev.logpost_batch(thetas, nug)
    w1, f1 = lib.guard_counts()
    e8 = errors(val, grad)
    lib.set_guard(1e300, 1e300)                                # guard off: deep 7-digit levels for every system
    val_off, _, _, _ = dev.logpost_batch(thetas, nug)
    lib.set_guard()
    dev.close()
    assert np.all(info64 == 0) and np.all(info == 0)
    assert f1 == f0 and w1 - w0 >= 3                           # three systems re-factored on the top levels (+ the LAUUM digits)
    for b in checked:
        assert e8[b][0] <= max(1e-9, 4.0 * e64[b][0]), (b, e8[b], e64[b])
        assert e8[b][1] <= max(1e-7, 4.0 * e64[b][1]), (b, e8[b], e64[b])
    # the benign systems never left the deep path: identical with the guard off; the ill-conditioned ones moved
    assert np.array_equal(val_off[3:], val[3:])
    off0 = abs(val_off[0] - want[0][0]) / abs(want[0][0])
    assert off0 > 5.0 * max(e8[0][0], 1e-9), (off0, e8[0])


def test_refinement_is_what_keeps_the_mean(exq, orc):
    """Without the refinement step the int8 factorisation's predictive mean is ~50 x worse than the FP64 path at
    cond(K) = 4e7 (tools/guard_survey.py); with it the two are level.  Guards the fix, not just the outcome."""
    lib = exq._lib
    kernel, N, d, ell = "Matern52", 2304, 8, 2.0
    X, y = orc.synthetic_problem(N, d)
    Xc = orc.synthetic_candidates(2000, d)
    raw = np.full(d, -2 * np.log(ell))
    mo, _ = orc.predict(orc.fit(X, y, kernel, raw, 1.7, 0.0), Xc)
    errs = {}
    for steps in (0, 1):
        lib.set_int8_digits(6, 7)
        lib.set_refine_steps(steps)
        dev = exq.DeviceGP(X, y, kernel).fit(raw, 1.7, 0.0)
        m, _ = dev.predict(Xc)
        errs[steps] = float(np.max(np.abs(m - mo) / np.maximum(np.abs(mo), 0.1)))
        dev.close()
    lib.set_refine_steps(1)
    assert errs[1] <= 1e-8
    assert errs[1] < errs[0]


def test_guard_widens_six_digit_products(exq, orc):
    """kappa >= guard_widen: the variance product runs with 7 digits.  Forced by setting the threshold to 1."""
    lib = exq._lib
    X, y = orc.synthetic_problem(1500, 4)
    Xc = orc.synthetic_candidates(4000, 4)
    raw = np.full(4, -2 * np.log(0.4))
    lib.set_int8_digits(7, 7)
    lib.set_guard(1e300, 1e300)
    dev = exq.DeviceGP(X, y, "Matern52").fit(raw, 1.7, 0.0)
    _, v7 = dev.predict(Xc)
    dev.close()
    lib.set_int8_digits(6, 7)
    dev = exq.DeviceGP(X, y, "Matern52").fit(raw, 1.7, 0.0)
    _, v6 = dev.predict(Xc)
    w0, f0 = lib.guard_counts()
    lib.set_guard(1.0, 1e300)                      # every kappa >= 1: widen
    _, vw = dev.predict(Xc)
    w1, f1 = lib.guard_counts()
    dev.close()
    lib.set_guard()
    assert w1 == w0 + 1 and f1 == f0
    assert np.array_equal(vw, v7)                  # the guarded call IS the 7-digit product
    assert not np.array_equal(v6, v7) and np.max(np.abs(v6 - v7)) < 1e-10


def test_guard_falls_back_to_fp64(exq, orc):
    """kappa >= guard_fallback: an int8 factorisation is redone on FP64 DMMA and the system stays there (variance
    included).  Forced by setting both thresholds to 1: results must be those of the all-FP64 configuration, bit
    for bit, in a single fit AND inside a batched log-posterior evaluation."""
    lib = exq._lib
    N, d = 2304, 5
    X, y = orc.synthetic_problem(N, d)
    Xc = orc.synthetic_candidates(3000, d)
    raw = np.full(d, -2 * np.log(0.7))
    rng = np.random.default_rng(8)
    thetas = np.concatenate([raw + rng.uniform(-0.4, 0.4, (16, d)), np.log(1.7) + rng.uniform(-0.3, 0.3, (16, 1))], axis=1)
    lib.set_int8_digits(0, 0)
    dev = exq.DeviceGP(X, y, "Matern52").fit(raw, 1.7, 0.0)
    nll64 = dev.neg_log_likelihood
    m64, v64 = dev.predict(Xc)
    val64, grad64, _, _ = dev.logpost_batch(thetas, 0.0)
    dev.close()
    lib.set_int8_digits(6, 7)
    lib.set_guard(1.0, 1.0)
    w0, f0 = lib.guard_counts()
    dev = exq.DeviceGP(X, y, "Matern52").fit(raw, 1.7, 0.0)
    kappa, on64 = dev.kappa()
    m, v = dev.predict(Xc)
    val, grad, info, _ = dev.logpost_batch(thetas, 0.0)
    w1, f1 = lib.guard_counts()
    dev.close()
    lib.set_guard()
    assert on64 and kappa >= 1.0
    assert f1 - f0 == 1 + 16                       # the fit and each of the 16 batched systems
    assert np.all(info == 0)
    assert np.array_equal(m, m64) and np.array_equal(v, v64)
    assert np.array_equal(val, val64)
    assert np.allclose(grad, grad64, rtol=1e-13, atol=0)
    # and with the guard at its defaults the same systems stay on the int8 path
    dev = exq.DeviceGP(X, y, "Matern52").fit(raw, 1.7, 0.0)
    assert not dev.kappa()[1]
    assert abs(dev.neg_log_likelihood - nll64) <= 1e-11 * abs(nll64)
    dev.close()


def test_kinv_full_matrix_at_2048(exq, orc):
    """K^-1 at N = 2048 goes through the int8 LAUUM (7 digits): the whole matrix against numpy's inverse, 1e-8 of
    its largest entry (the unchanged multi-level path multiplies it into 1e-8-tolerance quantities,
    exauq/core/designers.py:1029)."""
    N, d = 2048, 6
    X, y = orc.synthetic_problem(N, d)
    raw = np.full(d, -2 * np.log(0.5))
    exq._lib.set_int8_digits(6, 7)
    dev = exq.DeviceGP(X, y, "Matern52").fit(raw, 1.7, 1e-6)     # kinv excludes the nugget (modelling.py:1095-1113)
    ki = dev.kinv()
    ref = np.linalg.inv(1.7 * orc.corr("Matern52", raw, X, X))
    assert np.max(np.abs(ki - ref)) <= 1e-8 * np.max(np.abs(ref))
    assert np.array_equal(ki, ki.T)
    # the product the reference forms with it: k(x)^T kinv y
    xq = orc.synthetic_candidates(5, d)
    k = 1.7 * orc.corr("Matern52", raw, xq, X)
    assert np.allclose(k @ ki @ y, k @ ref @ y, rtol=1e-8, atol=1e-9)
    dev.close()


def test_kinv_refuses_what_the_reference_refuses(exq, orc):
    """np.linalg.cond(K) >= 1/eps => ValueError in the reference (exauq/core/modelling.py:1082-1093).  A Matern-5/2
    K over 600 points in 2-D with length scale 8 has cond_2 = 5e17 but factorises, with a diagonal ratio
    (max L_ii / min L_ii)^2 ~ 1e13 -- the case round 1's diagonal-ratio bound let through; the 2-norm estimate
    (power iterations with K and with Linv^T Linv) refuses it.  With length scale 3 over 300 points cond_2 = 1e14:
    the reference accepts, and so must the library."""
    eps = np.finfo(float).eps
    X, y = orc.synthetic_problem(600, 2)
    raw = np.full(2, -2 * np.log(8.0))
    K = 1.7 * orc.corr("Matern52", raw, X, X)
    assert np.linalg.cond(K) >= 1 / eps
    dg = np.diag(np.linalg.cholesky(K))
    assert (dg.max() / dg.min()) ** 2 < 1 / eps              # the old criterion would have accepted it
    dev = exq.DeviceGP(X, y, "Matern52").fit(raw, 1.7, 1e-6)
    with pytest.raises(np.linalg.LinAlgError):
        dev.kinv()
    dev.close()
    X, y = orc.synthetic_problem(300, 2)
    raw = np.full(2, -2 * np.log(3.0))
    K = 1.7 * orc.corr("Matern52", raw, X, X)
    assert np.linalg.cond(K) < 0.1 / eps
    dev = exq.DeviceGP(X, y, "Matern52").fit(raw, 1.7, 1e-6)
    ki = dev.kinv()
    assert np.all(np.isfinite(ki)) and np.array_equal(ki, ki.T)
    # an inverse at cond 1e14: residual against K at the level cond * eps allows
    assert np.max(np.abs(ki @ K - np.eye(300))) < 0.3
    dev.close()
    raw = np.full(2, -2 * np.log(0.05))
    K = 1.7 * orc.corr("Matern52", raw, X, X)
    dev = exq.DeviceGP(X, y, "Matern52").fit(raw, 1.7, 0.0)
    ref = np.linalg.inv(K)
    assert np.max(np.abs(dev.kinv() - ref)) <= 1e-10 * np.max(np.abs(ref))
    dev.close()
