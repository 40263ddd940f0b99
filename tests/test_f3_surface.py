"""SURVEY.md section 8f.3 -- the remaining mogp surface ``MogpEmulator`` passes through (exauq/core/emulators.py
:80-82, 129-137): ``nugget="pivot"`` (pivoted Cholesky, mogp ``pivot_cholesky`` -> LAPACK ``dpstrf``) and
user-supplied ``priors``.  CPU tests pin the oracle's restatement; GPU tests compare ``B200GaussianProcess`` with
``MogpEmulator`` on the oracle, through the reference's own classes."""

import numpy as np
import pytest

from conftest import needs_exauq


def _toy(n, d, seed=0):
    from scipy.stats.qmc import LatinHypercube

    X = LatinHypercube(d, seed=seed).random(n)
    y = np.sum(np.sin(4.0 * X), axis=1) + X[:, 0] * X[:, -1]
    return X, y


# ------------------------------------------------------------------------------------------
# oracle (CPU)
# ------------------------------------------------------------------------------------------
def test_oracle_pivot_cholesky_full_rank_and_rank_deficient():
    from mogp_emulator.GaussianProcess import pivot_cholesky
    from mogp_emulator.Kernel import Matern52, SquaredExponential

    X, _ = _toy(40, 2)
    K = 1.3 * Matern52().kernel_f(X, X, np.array([1.0, 1.5]))
    L, P = pivot_cholesky(K)
    assert sorted(P.tolist()) == list(range(40))
    assert np.allclose(L @ L.T, K[np.ix_(P, P)], rtol=0, atol=1e-13)
    assert np.all(np.diff(np.diag(L)) <= 1e-15)                        # decreasing diagonal
    assert abs(2 * np.sum(np.log(np.diag(L))) - np.linalg.slogdet(K)[1]) < 1e-9
    # an exactly repeated row: rank n - 1, last diagonal entry = L[r-1, r-1] / n (mogp's rule)
    Xd = np.vstack([X, X[7:8]])
    Kd = 1.3 * Matern52().kernel_f(Xd, Xd, np.array([1.0, 1.5]))
    Ld, Pd = pivot_cholesky(Kd)
    assert Pd[-1] in (7, 40)
    assert np.isclose(Ld[-1, -1], Ld[-2, -2] / 41.0, rtol=1e-14)
    # smooth kernel, long length scales: numerically rank deficient well before n
    Ks = SquaredExponential().kernel_f(X, X, np.array([-3.0, -3.0]))
    Ls, Ps = pivot_cholesky(Ks)
    rank = int(np.sum(np.diag(Ls) > np.diag(Ls)[0] * 1e-9))
    assert 3 < rank < 40


@needs_exauq
def test_oracle_pivot_emulator_equals_zero_nugget_when_well_conditioned():
    from exauq.core.emulators import MogpEmulator, MogpHyperparameters
    from exauq.core.modelling import Input, TrainingDatum

    X, y = _toy(30, 2, seed=3)
    data = [TrainingDatum(Input(*x), float(t)) for x, t in zip(X, y)]
    hp = MogpHyperparameters([0.4, 0.5], 1.1)
    a = MogpEmulator(kernel="Matern52", nugget="pivot")
    a.fit(data, hyperparameters=hp)
    b = MogpEmulator(kernel="Matern52", nugget=0.0)
    b.fit(data, hyperparameters=hp)
    assert a.fit_hyperparameters.nugget is None
    for q in [(0.1, 0.9), (0.5, 0.5), (0.77, 0.2)]:
        pa, pb = a.predict(Input(*q)), b.predict(Input(*q))
        assert abs(pa.estimate - pb.estimate) < 1e-10 and abs(pa.variance - pb.variance) < 1e-10


def test_product_priors_equal_oracle_priors():
    """Same densities, same gradients on the raw scale, same draws from the same seed (host logic, no GPU)."""
    from mogp_emulator import Priors as OP
    from mogp_emulator.GPParams import GPParams

    from exauq_toolbox_b200 import priors as PP

    spec = [("InvGammaPrior", 2.5, 0.7), ("GammaPrior", 3.0, 0.4), ("LogNormalPrior", 0.6, 1.8), None]
    mk = lambda mod, s: None if s is None else getattr(mod, s[0])(s[1], s[2])
    for fit_nugget in (False, True):
        ntype = "fit" if fit_nugget else "adaptive"
        o = OP.GPPriors(corr=[mk(OP, s) for s in spec[:3]], cov=mk(OP, spec[2]), nugget=mk(OP, spec[0]), nugget_type=ntype)
        p = PP.CustomMapPriors(PP.GPPriors(corr=[mk(PP, s) for s in spec[:3]], cov=mk(PP, spec[2]), nugget=mk(PP, spec[0])),
                               3, fit_nugget=fit_nugget)
        # mogp's own objects are accepted as they are (duck typing), and so is a dict
        q = PP.CustomMapPriors(o, 3, fit_nugget=fit_nugget)
        r = PP.CustomMapPriors({"corr": [mk(PP, s) for s in spec[:3]], "cov": mk(PP, spec[2]), "nugget": mk(PP, spec[0])},
                               3, fit_nugget=fit_nugget)
        rng = np.random.default_rng(5)
        for _ in range(5):
            raw = rng.uniform(-1.5, 1.5, 4 + fit_nugget)
            if fit_nugget:
                raw[-1] = rng.uniform(-16, -12)
            theta = GPParams(n_mean=0, n_corr=3, nugget="fit" if fit_nugget else "adaptive")
            theta.set_data(raw)
            for c in (p, q, r):
                assert np.isclose(c.logp(raw), o.logp(theta), rtol=1e-14, atol=1e-14)
                assert np.allclose(c.dlogp_draw(raw), o.dlogpdtheta(theta), rtol=1e-13, atol=1e-14)
        np.random.seed(11)
        so = [o.sample() for _ in range(3)]
        np.random.seed(11)
        sp = [p.sample() for _ in range(3)]
        assert np.array_equal(np.array(so), np.array(sp))
    with pytest.raises(ValueError):
        PP.CustomMapPriors(PP.GPPriors(n_corr=2), 3)
    with pytest.raises(TypeError):
        PP.CustomMapPriors(3.0, 3)
    assert isinstance(PP.make_priors(None, np.random.default_rng(0).uniform(size=(10, 3))), PP.MapPriors)


# ------------------------------------------------------------------------------------------
# GPU
# ------------------------------------------------------------------------------------------
def _close(a, b, rel=1e-8, abs_=1e-9):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return bool(np.all(np.abs(a - b) <= np.maximum(rel * np.maximum(np.abs(a), np.abs(b)), abs_)))


@pytest.mark.gpu
@pytest.mark.parametrize("kernel,n,d,raw", [("Matern52", 60, 2, 0.5), ("SquaredExponential", 48, 3, -2.5),
                                            ("Matern52", 333, 4, 1.0)])
def test_gpu_pivot_order_matches_dpstrf(kernel, n, d, raw):
    """Device pivoted Cholesky against scipy's dpstrf (n <= 64: LAPACK runs the same unblocked algorithm, so pivot
    order and rank agree exactly; above, the blocked variant may break near-ties differently -- compare the factor)."""
    from scipy.linalg import lapack

    import exauq_toolbox_b200 as exq

    X, y = _toy(n, d, seed=n)
    rawv = np.full(d, raw)
    dev = exq.DeviceGP(X, y, kernel)
    piv, rank = dev.pivot_order(rawv, 1.3)
    K = 1.3 * exq._lib.corr(kernel, rawv, X, X)
    L, P, r, info = lapack.dpstrf(np.ascontiguousarray(K), lower=1)
    P = P - 1
    assert sorted(piv.tolist()) == list(range(n))
    if n <= 64:
        assert rank == r
        assert piv[:rank].tolist() == P[:rank].tolist()
    else:
        assert abs(rank - r) <= 1
        agree = int(np.argmax(piv != P)) if np.any(piv != P) else n
        assert agree >= 10
    dev.close()


@pytest.mark.gpu
@needs_exauq
@pytest.mark.parametrize("kernel,n,d,ells", [("Matern52", 120, 3, [0.4, 0.5, 0.3]), ("SquaredExponential", 90, 2, [0.12, 0.15]),
                                             ("ProductMat52", 700, 4, [0.6, 0.5, 0.7, 0.4])])
def test_gpu_pivot_emulator_matches_reference(kernel, n, d, ells):
    """B200GaussianProcess(nugget="pivot") against MogpEmulator(nugget="pivot") on the oracle, through the reference's
    own classes: predictions, covariance ordering, kinv, LOO served in the caller's order."""
    from exauq.core.emulators import MogpEmulator, MogpHyperparameters
    from exauq.core.modelling import GaussianProcessHyperparameters, Input, TrainingDatum

    from exauq_toolbox_b200.gp import B200GaussianProcess

    X, y = _toy(n, d, seed=2)
    data = [TrainingDatum(Input(*x), float(t)) for x, t in zip(X, y)]
    ref = MogpEmulator(kernel=kernel, nugget="pivot")
    ref.fit(data, hyperparameters=MogpHyperparameters(ells, 1.4))
    gp = B200GaussianProcess(kernel=kernel, nugget="pivot")
    gp.fit(data, hyperparameters=GaussianProcessHyperparameters(ells, 1.4))
    assert gp.fit_hyperparameters.nugget is None and ref.fit_hyperparameters.nugget is None
    st = gp._real_state()
    assert st.dev.rank == n and sorted(st.dev.piv.tolist()) == list(range(n))
    assert st.dev.piv.tolist() != list(range(n))                       # the data really were reordered
    qs = [Input(*q) for q in np.random.default_rng(1).uniform(0, 1, (12, d))]
    for q in qs:
        a, b = gp.predict(q), ref.predict(q)
        assert _close(a.estimate, b.estimate) and _close(a.variance, b.variance)
    assert np.allclose(gp.covariance_matrix(qs), ref.covariance_matrix(qs), rtol=1e-12, atol=1e-15)
    kr = ref.kinv
    assert np.allclose(gp.kinv, kr, rtol=1e-6, atol=1e-7 * np.max(np.abs(kr)))
    lm, lv, _ = gp.loo_predictions()
    for i in (0, 5, n - 1):
        sub = MogpEmulator(kernel=kernel, nugget="pivot")
        sub.fit(data[:i] + data[i + 1:], hyperparameters=MogpHyperparameters(ells, 1.4))
        p = sub.predict(data[i].input)
        assert _close(lm[i], p.estimate, 1e-7) and _close(lv[i], p.variance, 1e-7)


@pytest.mark.gpu
def test_gpu_pivot_rank_deficient_follows_mogp_rule():
    """Rank-deficient covariance matrices (n <= 64, so that LAPACK's dpstrf is the unblocked algorithm restated on the
    device).  (i) Two exactly repeated rows: same rank, same pivots, the skipped rows get mogp's decreasing diagonal,
    log-determinant and predictions agree with the oracle.  (ii) Smooth kernel with long length scales, ~40 skipped
    rows: same rank and pivots; the log-determinant inherits the few digits of the last accepted pivot, and the
    predictions are NaN in the oracle too (the factorially decreasing diagonal overflows the solve)."""
    import mogp_emulator
    from mogp_emulator.GaussianProcess import pivot_cholesky

    import exauq_toolbox_b200 as exq

    X, y = _toy(50, 2, seed=9)
    Xd, yd = np.vstack([X, X[[7, 31]]]), np.concatenate([y, y[[7, 31]]])
    n = 52
    raw = np.array([1.0, 0.5])
    ref = mogp_emulator.GaussianProcess(Xd, yd, kernel="Matern52", nugget="pivot")
    ref.fit(np.concatenate([raw, [np.log(1.3)]]))
    g = exq._lib.PivotedDeviceGP(Xd, yd, "Matern52").fit(raw, 1.3)
    assert g.rank == 50
    assert g.piv[:50].tolist() == ref.P[:50].tolist()
    ref_logdet = 2.0 * np.sum(np.log(np.diag(ref.L)))
    assert abs(g.logdet - ref_logdet) <= 1e-9 * abs(ref_logdet)
    Xq = np.random.default_rng(0).uniform(0, 1, (20, 2))
    m, v = g.predict(Xq)
    pr = ref.predict(Xq)
    assert np.all(np.isfinite(m)) and np.all(v >= 0.0)
    assert _close(m, pr.mean, 1e-6, 1e-8) and _close(v, pr.unc, 1e-6, 1e-8)
    g.close()

    n = 56
    X, y = _toy(n, 2, seed=9)
    raw = np.array([-3.0, -3.0])
    K = exq._lib.corr("SquaredExponential", raw, X, X)
    L, P = pivot_cholesky(K)
    g = exq._lib.PivotedDeviceGP(X, y, "SquaredExponential").fit(raw, 1.0)
    r = g.rank
    assert 3 < r < n - 10
    assert g.piv[:r].tolist() == P[:r].tolist()
    ref_logdet = 2.0 * np.sum(np.log(np.diag(L)))
    assert abs(g.logdet - ref_logdet) <= 2.0 * (n - r) * 0.02 + 1e-9 * abs(ref_logdet)
    g.close()


@pytest.mark.gpu
@needs_exauq
def test_gpu_custom_priors_map_fit_matches_reference():
    """MAP estimation with user-supplied priors: same seed, same optimiser, same optimum as MogpEmulator on the oracle
    (the comparison the reference's own tests make for the default priors, tests/unit/test_emulators.py:489-516)."""
    from exauq.core.emulators import MogpEmulator
    from exauq.core.modelling import Input, TrainingDatum
    from mogp_emulator import Priors as OP

    from exauq_toolbox_b200 import priors as PP
    from exauq_toolbox_b200.gp import B200GaussianProcess

    X, y = _toy(70, 2, seed=4)
    data = [TrainingDatum(Input(*x), float(t)) for x, t in zip(X, y)]
    o_pri = OP.GPPriors(corr=[OP.InvGammaPrior(3.0, 1.2), OP.LogNormalPrior(0.5, 0.6)], cov=OP.GammaPrior(2.0, 1.5),
                        nugget_type="adaptive")
    p_pri = PP.GPPriors(corr=[PP.InvGammaPrior(3.0, 1.2), PP.LogNormalPrior(0.5, 0.6)], cov=PP.GammaPrior(2.0, 1.5))
    np.random.seed(7)
    ref = MogpEmulator(kernel="Matern52", priors=o_pri)
    ref.fit(data)
    np.random.seed(7)
    gp = B200GaussianProcess(kernel="Matern52", priors=p_pri)
    gp.fit(data)
    a, b = gp.fit_hyperparameters, ref.fit_hyperparameters
    assert np.allclose(a.corr_length_scales, b.corr_length_scales, rtol=1e-5)
    assert np.isclose(a.process_var, b.process_var, rtol=1e-5)
    # and the default priors give a different optimum: the custom ones were really used
    np.random.seed(7)
    dflt = B200GaussianProcess(kernel="Matern52")
    dflt.fit(data)
    assert not np.allclose(dflt.fit_hyperparameters.corr_length_scales, a.corr_length_scales, rtol=1e-3)
    q = Input(0.3, 0.6)
    pa, pb = gp.predict(q), ref.predict(q)
    assert _close(pa.estimate, pb.estimate, 1e-5) and _close(pa.variance, pb.variance, 1e-4)
    with pytest.raises(RuntimeError):
        B200GaussianProcess(kernel="Matern52", priors=3.0)
    with pytest.raises(RuntimeError):
        B200GaussianProcess(kernel="Matern52", mean="x[0]")


@pytest.mark.gpu
@needs_exauq
def test_gpu_pivot_hyperparameter_estimation_matches_reference():
    """fit() without hyperparameters under nugget="pivot": the restarts run on the batched recursion (a full-rank
    covariance has the same log-posterior with or without pivoting), the final fit is the pivoted one."""
    from exauq.core.emulators import MogpEmulator
    from exauq.core.modelling import Input, TrainingDatum

    from exauq_toolbox_b200.gp import B200GaussianProcess

    X, y = _toy(80, 2, seed=6)
    data = [TrainingDatum(Input(*x), float(t)) for x, t in zip(X, y)]
    np.random.seed(3)
    ref = MogpEmulator(kernel="Matern52", nugget="pivot")
    ref.fit(data)
    np.random.seed(3)
    gp = B200GaussianProcess(kernel="Matern52", nugget="pivot")
    gp.fit(data)
    a, b = gp.fit_hyperparameters, ref.fit_hyperparameters
    assert a.nugget is None and b.nugget is None
    # the posterior of this data set is flat around its optimum (length scales 2.5, variance 40): the restarts see
    # objective values that differ in the 10th digit between the pivoted and the plain factorisation, and L-BFGS-B's
    # stopping rule (ftol 2.2e-9) turns that into 1e-5 .. 1e-4 of the optimum
    assert np.allclose(a.corr_length_scales, b.corr_length_scales, rtol=2e-4)
    assert np.isclose(a.process_var, b.process_var, rtol=2e-4)
    q = Input(0.41, 0.73)
    pa, pb = gp.predict(q), ref.predict(q)
    assert _close(pa.estimate, pb.estimate, 1e-5, 1e-6) and _close(pa.variance, pb.variance, 1e-3, 1e-8)
