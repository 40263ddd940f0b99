"""B200GaussianProcess used by the reference's UNCHANGED designers, compared with the
reference's MogpEmulator (running on the oracle's mogp restatement) on the same inputs."""

import copy

import numpy as np
import pytest

from conftest import load_golden, needs_exauq

pytestmark = [pytest.mark.gpu, needs_exauq]


@pytest.fixture(scope="module")
def api():
    import exauq.core.designers as designers
    from exauq.core.emulators import MogpEmulator, MogpHyperparameters
    from exauq.core.modelling import (GaussianProcessHyperparameters, Input, MultiLevel,
                                      MultiLevelGaussianProcess, SimulatorDomain, TrainingDatum)

    from exauq_toolbox_b200.gp import B200GaussianProcess

    class NS:
        pass

    ns = NS()
    ns.__dict__.update(locals())
    return ns


def close(a, b, rel=1e-8, abs_=1e-9):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return bool(np.all(np.abs(a - b) <= np.maximum(rel * np.maximum(np.abs(a), np.abs(b)), abs_)))


def _data(api, g):
    return [api.TrainingDatum(api.Input(*x), y) for x, y in zip(g["X"] if "X" in g else g["inputs"],
                                                               g["y"] if "y" in g else g["outputs"])]


def test_fit_predict_correlation_match_reference(api):
    g = load_golden("fixed_matern_d3_n40.json")
    data = _data(api, g)
    hp = g["hyperparameters"]
    ref = api.MogpEmulator(kernel="Matern52")
    ref.fit(data, hyperparameters=api.MogpHyperparameters(hp["corr_length_scales"], hp["process_var"], hp["nugget"]))
    gp = api.B200GaussianProcess(kernel="Matern52")
    assert gp.training_data == () and gp.fit_hyperparameters is None and gp.kinv.size == 0
    gp.fit(data, hyperparameters=api.GaussianProcessHyperparameters(hp["corr_length_scales"], hp["process_var"],
                                                                   hp["nugget"]))
    assert gp.training_data == tuple(data)
    assert gp.fit_hyperparameters == api.GaussianProcessHyperparameters(hp["corr_length_scales"],
                                                                        hp["process_var"], hp["nugget"])
    qs = [api.Input(*q) for q in g["query"]]
    for q in qs:
        a, b = gp.predict(q), ref.predict(q)
        assert close(a.estimate, b.estimate) and close(a.variance, b.variance)
    train_inputs = [d.input for d in data]
    assert np.max(np.abs(gp.correlation(qs, train_inputs) - ref.correlation(qs, train_inputs))) < 1e-13
    assert np.allclose(gp.covariance_matrix(qs), ref.covariance_matrix(qs), rtol=1e-13)
    assert np.allclose(gp.kinv, ref.kinv, rtol=1e-6, atol=1e-7 * np.max(np.abs(ref.kinv)))
    assert gp.correlation([], train_inputs).size == 0
    with pytest.raises(TypeError):
        gp.predict([0.1, 0.2, 0.3])
    with pytest.raises(ValueError):
        gp.predict(api.Input(0.1, 0.2))
    with pytest.raises(ValueError):
        gp.fit(data + [data[0]])


def test_unchanged_loo_designers_are_served_from_one_sweep(api):
    """compute_loo_gp / compute_loo_prediction (exauq/core/designers.py:295-370, 958-998)
    called exactly as the reference's compute_loo_errors_gp does."""
    g = load_golden("fixed_matern_d8_n150.json")
    data = _data(api, g)
    hp = g["hyperparameters"]
    gp = api.B200GaussianProcess(kernel="Matern52")
    gp.fit(data, hyperparameters=api.GaussianProcessHyperparameters(hp["corr_length_scales"], hp["process_var"],
                                                                   hp["nugget"]))
    from exauq_toolbox_b200 import _lib

    loo_gp = copy.deepcopy(gp)
    n0 = _lib.launch_count()
    got = []
    for i, datum in enumerate(gp.training_data):
        api.designers.compute_loo_gp(gp, i, loo_gp=loo_gp)
        assert len(loo_gp.training_data) == len(data) - 1
        pr = loo_gp.predict(datum.input)
        got.append([pr.estimate, pr.variance, pr.nes_error(datum.output)])
    assert _lib.launch_count() - n0 < 20, "the N refits must be served by one LOO sweep"
    got, ref = np.array(got), np.array(g["loo"])
    assert close(got[:, 0], ref[:, 0]) and close(got[:, 1], ref[:, 1]) and close(got[:, 2], ref[:, 2])
    # asking a LOO GP anything else triggers a real refit of the N-1 system
    api.designers.compute_loo_gp(gp, 3, loo_gp=loo_gp)
    q = api.Input(*g["query"][0])
    refgp = api.MogpEmulator(kernel="Matern52")
    refgp.fit(data[:3] + data[4:], hyperparameters=api.MogpHyperparameters(hp["corr_length_scales"],
                                                                          hp["process_var"], hp["nugget"]))
    a, b = loo_gp.predict(q), refgp.predict(q)
    assert close(a.estimate, b.estimate) and close(a.variance, b.variance)
    # and the user-visible state of a deep copy is independent
    loo_gp.custom_attribute = [1, 2]
    c = copy.deepcopy(loo_gp)
    c.custom_attribute.append(3)
    assert loo_gp.custom_attribute == [1, 2]


def test_config1_single_level_loo_samples(api):
    """BASELINE.json configs[0]: N=50 LHS on the 2-D toy simulator, Matern-5/2, then
    compute_single_level_loo_samples(batch_size=1) through the unchanged designers; golden
    values were produced by the reference's MogpEmulator with the same numpy seeds."""
    g = load_golden("slas_config1.json")
    domain = api.SimulatorDomain([tuple(b) for b in g["domain"]])
    data = _data(api, g)
    np.random.seed(g["gp_fit_seed"])
    gp = api.B200GaussianProcess(kernel="Matern52")
    gp.fit(data)
    hp, ref = gp.fit_hyperparameters, g["gp_hyperparameters"]
    # the reference's own tests compare estimated hyperparameters at 1e-5 (test_emulators.py:497-516)
    assert np.allclose(hp.corr_length_scales, ref["corr_length_scales"], rtol=1e-5)
    assert np.isclose(hp.process_var, ref["process_var"], rtol=1e-5)
    assert hp.nugget == ref["nugget"]
    np.random.seed(g["errors_gp_fit_seed"])
    gp_e = api.designers.compute_loo_errors_gp(gp, domain)
    nes = np.array([d.output for d in gp_e.training_data])
    assert np.allclose(nes, g["loo_nes_errors"], rtol=1e-4)
    hpe, refe = gp_e.fit_hyperparameters, g["errors_gp_hyperparameters"]
    assert np.allclose(hpe.corr_length_scales, refe["corr_length_scales"], rtol=1e-4)
    assert np.isclose(hpe.process_var, refe["process_var"], rtol=1e-4)
    np.random.seed(g["errors_gp_fit_seed"])
    new = api.designers.compute_single_level_loo_samples(gp, domain, batch_size=1, seed=g["de_seed"])
    assert np.allclose(np.array(new[0]), g["new_design_point"], atol=2e-3)


def test_multi_level_gp_and_loo(api):
    """MultiLevelGaussianProcess (exauq/core/modelling.py:1352-1706) and the multi-level LOO
    prediction (exauq/core/designers.py:840-922) with B200 GPs vs reference GPs."""
    rng = np.random.default_rng(4)
    d = 2
    X1, X2 = rng.uniform(0, 1, (24, d)), rng.uniform(0, 1, (10, d))
    f1 = lambda x: x[1] + x[0] ** 2 + x[1] ** 2 - np.sqrt(2)  # noqa: E731
    f2 = lambda x: np.sin(2 * np.pi * x[0]) + np.sin(4 * np.pi * x[0] * x[1])  # noqa: E731
    data = api.MultiLevel([[api.TrainingDatum(api.Input(*x), float(f1(x))) for x in X1],
                           [api.TrainingDatum(api.Input(*x), float(f2(x))) for x in X2]])
    hps = {1: ([0.6, 0.9], 2.0, 1e-9), 2: ([0.2, 0.3], 1.5, 1e-9)}
    ours = api.MultiLevelGaussianProcess([api.B200GaussianProcess(kernel="Matern52") for _ in range(2)])
    refs = api.MultiLevelGaussianProcess([api.MogpEmulator(kernel="Matern52") for _ in range(2)])
    ours.fit(data, hyperparameters=api.MultiLevel({k: api.GaussianProcessHyperparameters(*v) for k, v in hps.items()}))
    refs.fit(data, hyperparameters=api.MultiLevel({k: api.MogpHyperparameters(*v) for k, v in hps.items()}))
    x = api.Input(0.5, 0.7)
    a, b = ours.predict(x), refs.predict(x)
    assert close(a.estimate, b.estimate) and close(a.variance, b.variance)
    for level, i in [(1, 0), (1, 7), (2, 3)]:
        a = api.designers.compute_multi_level_loo_prediction(ours, level, i)
        b = api.designers.compute_multi_level_loo_prediction(refs, level, i)
        assert close(a.estimate, b.estimate, rel=1e-7, abs_=1e-8) and close(a.variance, b.variance, rel=1e-7)


def test_accelerated_loo_errors_gp_equals_the_unchanged_loop(api):
    """accel.loo_errors_gp (patched in by accelerate_designers) against the reference's compute_loo_errors_gp on the
    same B200 GP: the same error training data and, with the same numpy seed, the same fitted errors GP; the same
    ValueError for a training input outside the domain."""
    from exauq_toolbox_b200.accel import accelerate_designers

    g = load_golden("fixed_matern_d3_n40.json")
    data = _data(api, g)
    domain = api.SimulatorDomain([(0.0, 1.0)] * 3)
    gp = api.B200GaussianProcess(kernel="Matern52")
    gp.fit(data, hyperparameters=api.GaussianProcessHyperparameters([0.4, 0.5, 0.6], 1.3, 0.0))
    np.random.seed(3)
    plain = api.designers.compute_loo_errors_gp(gp, domain)
    np.random.seed(3)
    with accelerate_designers():
        fast = api.designers.compute_loo_errors_gp(gp, domain)
    assert len(plain.training_data) == len(fast.training_data) == len(data)
    for a, b in zip(plain.training_data, fast.training_data):
        assert a.input == b.input and close(a.output, b.output, rel=1e-12, abs_=1e-14)
    ha, hb = plain.fit_hyperparameters, fast.fit_hyperparameters
    assert close(ha.corr_length_scales, hb.corr_length_scales, rel=1e-6) and close(ha.process_var, hb.process_var, rel=1e-6)
    small = api.SimulatorDomain([(0.0, 0.5)] * 3)
    for ctx in (False, True):
        with pytest.raises(ValueError, match="belong to the domain"):
            if ctx:
                with accelerate_designers():
                    api.designers.compute_loo_errors_gp(gp, small)
            else:
                api.designers.compute_loo_errors_gp(gp, small)


def test_accelerated_multi_level_loo_error_data_equals_the_unchanged_path(api):
    """Inside accelerate_designers() the reference's compute_multi_level_loo_error_data (designers.py:1045-1105) --
    zero-mean predictions through ``kinv @ y`` formed once, cross-level repetition check by a sort -- returns the
    errors of the untouched code path on the same B200 GPs, and both agree with reference GPs."""
    from exauq_toolbox_b200.accel import accelerate_designers, zero_mean_prediction

    rng = np.random.default_rng(8)
    d = 3
    sizes = (40, 17, 9)
    data = api.MultiLevel([[api.TrainingDatum(api.Input(*x), float(np.sin(3 * x[0]) + x[1] * x[2] / (lv + 1)))
                            for x in rng.uniform(0, 1, (n, d))] for lv, n in enumerate(sizes)])
    hps = {1: ([0.5, 0.6, 0.7], 2.0, 0.0), 2: ([0.3, 0.3, 0.4], 1.5, 0.0), 3: ([0.6, 0.2, 0.5], 0.7, 0.0)}
    ours = api.MultiLevelGaussianProcess([api.B200GaussianProcess(kernel="Matern52") for _ in range(3)])
    refs = api.MultiLevelGaussianProcess([api.MogpEmulator(kernel="Matern52") for _ in range(3)])
    ours.fit(data, hyperparameters=api.MultiLevel({k: api.GaussianProcessHyperparameters(*v) for k, v in hps.items()}))
    refs.fit(data, hyperparameters=api.MultiLevel({k: api.MogpHyperparameters(*v) for k, v in hps.items()}))
    plain = api.designers.compute_multi_level_loo_error_data(ours)
    with accelerate_designers():
        fast = api.designers.compute_multi_level_loo_error_data(ours)
    want = api.designers.compute_multi_level_loo_error_data(refs)
    for lv in plain.levels:
        a = np.array([t.output for t in plain[lv]])
        b = np.array([t.output for t in fast[lv]])
        c = np.array([t.output for t in want[lv]])
        assert all(p.input == q.input for p, q in zip(plain[lv], fast[lv]))
        assert close(a, b, rel=1e-10, abs_=1e-12) and close(b, c, rel=1e-6, abs_=1e-8)
    # the patched function alone, and its error wrapping for an unfitted GP
    x = api.Input(0.3, 0.4, 0.5)
    p, q = zero_mean_prediction(ours[1], x), api.designers.compute_zero_mean_prediction(ours[1], x)
    assert close(p.estimate, q.estimate, rel=1e-11, abs_=1e-13) and p.variance == q.variance
    with pytest.raises(ValueError, match="hasn't been trained on any data"):
        zero_mean_prediction(api.B200GaussianProcess(kernel="Matern52"), x)
    # a repeated input across levels is reported as the reference reports it
    bad = api.MultiLevel({1: data[1], 2: tuple(data[2]) + (api.TrainingDatum(data[1][4].input, 0.0),), 3: data[3]})
    ours2 = api.MultiLevelGaussianProcess([api.B200GaussianProcess(kernel="Matern52") for _ in range(3)])
    ours2.fit(bad, hyperparameters=api.MultiLevel({k: api.GaussianProcessHyperparameters(*v) for k, v in hps.items()}))
    msgs = []
    for ctx in (None, accelerate_designers()):
        try:
            if ctx is None:
                api.designers.compute_multi_level_loo_prediction(ours2, 1, 0)
            else:
                with ctx:
                    api.designers.compute_multi_level_loo_prediction(ours2, 1, 0)
        except ValueError as e:
            msgs.append(str(e))
    assert len(msgs) == 2 and msgs[0] == msgs[1] and "levels 1, 2" in msgs[0]
    with accelerate_designers():
        with pytest.raises(ValueError) as ei:
            api.designers.compute_multi_level_loo_error_data(ours2)
    assert str(ei.value) == msgs[0]


def test_pei_candidates_matches_reference_peicalculator(api):
    """B200GaussianProcess.pei_candidates vs the reference's PEICalculator evaluated point
    by point, including the batch loop's repulsion updates."""
    g = load_golden("fixed_matern_d3_n40.json")
    data = _data(api, g)
    hp = g["hyperparameters"]
    gp = api.B200GaussianProcess(kernel="Matern52")
    gp.fit(data, hyperparameters=api.GaussianProcessHyperparameters(hp["corr_length_scales"], hp["process_var"],
                                                                   hp["nugget"]))
    domain = api.SimulatorDomain([(0.0, 1.0)] * 3)
    pei = api.designers.PEICalculator(domain, gp)     # unchanged reference class on our GP
    Xc = np.random.default_rng(9).uniform(0, 1, (300, 3))
    rep = np.array([list(x) for x in pei._other_repulsion_points])
    idx, val, cand = gp.pei_candidates(Xc, rep, batch_size=3)
    ref_vals = np.array([pei.compute(api.Input(*x)) for x in Xc])
    chosen = []
    for _ in range(3):
        i = int(np.argmax(ref_vals))
        chosen.append(i)
        pei.add_repulsion_points([api.Input(*Xc[i])])
        ref_vals = np.array([pei.compute(api.Input(*x)) for x in Xc])
    assert idx.tolist() == chosen
    assert np.allclose(cand.download("pei"), ref_vals, rtol=1e-7, atol=1e-13 * np.max(val))


def test_kinv_raises_value_error_for_singular_covariance(api):
    """exauq/core/modelling.py:1082-1113: a (numerically) singular covariance makes kinv raise
    ValueError('Cannot compute covariance inverse: ...'); here at first access (kinv is lazy)."""
    rng = np.random.default_rng(0)
    X = rng.uniform(0, 1, (30, 2))
    data = [api.TrainingDatum(api.Input(*x), float(x[0])) for x in X]
    gp = api.B200GaussianProcess(kernel="SquaredExponential", nugget=1e-6)
    # very long length scales: cov * C is singular to working precision without the nugget
    gp.fit(data, hyperparameters=api.GaussianProcessHyperparameters([50.0, 50.0], 1.0, 1e-6))
    with pytest.raises(ValueError, match="Cannot compute covariance inverse"):
        _ = gp.kinv
    good = api.B200GaussianProcess(kernel="Matern52")
    good.fit(data, hyperparameters=api.GaussianProcessHyperparameters([0.3, 0.3], 1.0, 0.0))
    k = good.covariance_matrix([d.input for d in data])
    assert np.allclose(good.kinv @ k, np.eye(30), atol=1e-8)


def test_unchanged_multi_level_loo_samples(api):
    """compute_multi_level_loo_samples (exauq/core/designers.py:1184-1378), unchanged, on a
    MultiLevelGaussianProcess of B200 GPs vs the same call on MogpEmulators (same numpy / DE seeds):
    same level chosen and the same design point to optimiser tolerance."""
    rng = np.random.default_rng(11)
    d = 2
    domain = api.SimulatorDomain([(0.0, 1.0)] * d)
    X1, X2 = rng.uniform(0.05, 0.95, (12, d)), rng.uniform(0.05, 0.95, (7, d))
    f1 = lambda x: x[1] + x[0] ** 2 + x[1] ** 2 - np.sqrt(2)  # noqa: E731
    f2 = lambda x: np.sin(2 * np.pi * x[0]) + np.sin(4 * np.pi * x[0] * x[1])  # noqa: E731
    data = api.MultiLevel([[api.TrainingDatum(api.Input(*x), float(f1(x))) for x in X1],
                           [api.TrainingDatum(api.Input(*x), float(f2(x))) for x in X2]])
    hps = {1: ([0.5, 0.6], 2.0, 0.0), 2: ([0.25, 0.3], 1.5, 0.0)}
    costs = api.MultiLevel([1.0, 11.0])
    seeds = api.MultiLevel([3, 4])
    results = []
    for make, hp_cls in ((lambda: api.B200GaussianProcess(kernel="Matern52"), api.GaussianProcessHyperparameters),
                         (lambda: api.MogpEmulator(kernel="Matern52"), api.MogpHyperparameters)):
        mlgp = api.MultiLevelGaussianProcess([make() for _ in range(2)])
        mlgp.fit(data, hyperparameters=api.MultiLevel({k: hp_cls(*v) for k, v in hps.items()}))
        np.random.seed(21)
        out = api.designers.compute_multi_level_loo_samples(mlgp, domain, costs, batch_size=1, seeds=seeds)
        np.seterr(all="warn")
        results.append(out)
    ours, ref = results
    assert ours.levels == ref.levels
    for lv in ref.levels:
        assert len(ours[lv]) == len(ref[lv])
        for a, b in zip(ours[lv], ref[lv]):
            assert a in domain
            assert np.allclose(np.array(a), np.array(b), atol=5e-3)


def test_reference_emulator_behaviours(api):
    """Backend-agnostic behaviours the reference's own tests pin for MogpEmulator
    (tests/unit/test_emulators.py), checked on B200GaussianProcess."""
    import warnings

    rng = np.random.default_rng(5)
    X = rng.uniform(0, 1, (18, 2))
    f = lambda x: np.sin(3 * x[0]) + x[1] ** 2  # noqa: E731
    data = [api.TrainingDatum(api.Input(*x), float(f(x))) for x in X]
    HP = api.GaussianProcessHyperparameters

    # interpolation with zero predictive variance at the training inputs (:811-821)
    gp = api.B200GaussianProcess(kernel="Matern52")
    gp.fit(data, hyperparameters=HP([0.4, 0.5], 1.3, 0.0))
    for d in data[:6]:
        p = gp.predict(d.input)
        assert abs(p.estimate - d.output) < 1e-8 and 0.0 <= p.variance < 1e-8

    # correlation depends only on the length scales, covariance = process_var * correlation (:320-407)
    gp2 = api.B200GaussianProcess(kernel="Matern52")
    gp2.fit(data, hyperparameters=HP([0.4, 0.5], 7.7, 1e-3))
    xs = [api.Input(0.1, 0.9), api.Input(0.6, 0.2)]
    tr = [d.input for d in data]
    assert np.array_equal(gp.correlation(xs, tr), gp2.correlation(xs, tr))
    assert np.allclose(gp2.covariance_matrix(xs), 7.7 * gp2.correlation(xs, tr), rtol=1e-14)
    c = gp.correlation(tr, tr)
    assert np.allclose(np.diag(c), 1.0) and np.allclose(c, c.T) and np.all((c > 0) & (c <= 1))

    # nugget semantics (emulators.py:427-445): supplied nugget wins; else the constructor's float; else the method
    g1 = api.B200GaussianProcess(kernel="Matern52", nugget=0.25)
    g1.fit(data, hyperparameters=HP([0.4, 0.5], 1.3))
    assert g1.fit_hyperparameters.nugget == 0.25
    g1.fit(data, hyperparameters=HP([0.4, 0.5], 1.3, 0.5))
    assert g1.fit_hyperparameters.nugget == 0.5
    assert gp.fit_hyperparameters.nugget == 0.0                   # adaptive, no jitter needed
    v_small, v_big = gp.predict(xs[0]).variance, g1.predict(xs[0]).variance
    assert v_big > v_small                                        # predictive variance includes the nugget
    with pytest.raises(ValueError, match="'nugget'='fit'"):
        api.B200GaussianProcess(kernel="Matern52", nugget="fit").fit(data, hyperparameters=HP([0.4, 0.5], 1.3))

    # estimation respects hyperparameter bounds (:551-583) and refits replace the data
    np.random.seed(4)
    gb = api.B200GaussianProcess(kernel="Matern52")
    gb.fit(data, hyperparameter_bounds=[(0.3, 0.6), (0.2, None), (0.5, 4.0)])
    hp = gb.fit_hyperparameters
    assert 0.3 - 1e-9 <= hp.corr_length_scales[0] <= 0.6 + 1e-9 and hp.corr_length_scales[1] >= 0.2 - 1e-9
    assert 0.5 - 1e-9 <= hp.process_var <= 4.0 + 1e-9
    gb.fit(data[:9], hyperparameters=hp)
    assert len(gb.training_data) == 9

    # update() adds data in front of the old data and refits (modelling.py:988-1040)
    gu = api.B200GaussianProcess(kernel="Matern52")
    gu.fit(data[:10], hyperparameters=HP([0.4, 0.5], 1.3, 0.0))
    gu.update(training_data=data[10:14], hyperparameters=HP([0.4, 0.5], 1.3, 0.0))
    assert len(gu.training_data) == 14 and gu.training_data[:4] == tuple(data[10:14])
    with pytest.warns(UserWarning, match="No arguments were passed"):
        gu.update()

    # fewer training points than hyperparameters: warning (:303-307)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        api.B200GaussianProcess(kernel="Matern52").fit(data[:2], hyperparameters=HP([0.4, 0.5], 1.3, 0.0))
        assert any("Fewer training points" in str(x.message) for x in w)


def test_multi_level_tutorial_end_to_end(api):
    """The two-level tutorial (training_multi_level_gp_tutorial.md) with B200 GPs: MAP estimation on the
    GPU at both levels (same numpy seed as the golden run) and MultiLevelGaussianProcess.predict."""
    g = load_golden("tutorial_mlgp.json")
    data = api.MultiLevel([[api.TrainingDatum(api.Input(*x), y) for x, y in zip(g["level" + lv]["inputs"], g["level" + lv]["outputs"])]
                           for lv in ("1", "2")])
    mlgp = api.MultiLevelGaussianProcess([api.B200GaussianProcess(kernel="Matern52") for _ in range(2)])
    np.random.seed(g["numpy_seed"])
    mlgp.fit(data)
    for lv in (1, 2):
        hp, ref = mlgp[lv].fit_hyperparameters, g["fit"][str(lv)]
        assert np.allclose(hp.corr_length_scales, ref["corr_length_scales"], rtol=1e-4)
        assert np.isclose(hp.process_var, ref["process_var"], rtol=1e-4)
    x = api.Input(*g["predict_at"])
    pub = g["published"]
    for level, key, ref in ((1, "level1", g["predictions"][0]), (2, "level2", g["predictions"][1])):
        p = mlgp.predict(x, level)
        assert abs(p.estimate - ref["estimate"]) < 1e-6 and abs(p.variance - ref["variance"]) < 1e-6
        assert abs(p.estimate - pub[key]["estimate"]) < 1e-5 and abs(p.variance - pub[key]["variance"]) < 1e-5


def test_single_level_tutorial_end_to_end(api):
    """training_gp_tutorial.md:259-263, 300-302: MAP estimation on the GPU (same numpy seed as the golden
    run) reproduces the published prediction and NES error at (0.5, 1.5)."""
    g = load_golden("tutorial_gp.json")
    data = [api.TrainingDatum(api.Input(*x), y) for x, y in zip(g["inputs"], g["outputs"])]
    np.random.seed(g["numpy_seed"])
    gp = api.B200GaussianProcess(kernel="Matern52")
    gp.fit(data)
    hp, ref = gp.fit_hyperparameters, g["fit"]
    assert np.allclose(hp.corr_length_scales, ref["corr_length_scales"], rtol=1e-5)
    assert np.isclose(hp.process_var, ref["process_var"], rtol=1e-5) and hp.nugget == ref["nugget"]
    p = gp.predict(api.Input(*g["predict_at"]))
    pub = g["published"]
    assert abs(p.estimate - g["estimate"]) < 1e-6 and abs(p.variance - g["variance"]) < 1e-6
    assert abs(p.estimate - pub["estimate"]) < 1e-6 and abs(p.variance - pub["variance"]) < 1e-6
    assert abs(p.nes_error(pub["observed"]) - pub["nes_error"]) < 1e-6


def test_csv_in_design_points_out(api, tmp_path):
    """The steps either side of the path, with the reference's own code:
    simulator outputs arrive as csv files per level (TrainingDatum.read_from_csv,
    exauq/core/modelling.py:459-532), are turned into successive-level differences with
    create_data_for_multi_level_loo_sampling / compute_delta_coefficients
    (exauq/core/designers.py:1381-1553), the multi-level GP is fitted on B200 GPs (hyperparameters estimated) and
    compute_multi_level_loo_samples returns the next design point, written back as csv; the same
    pipeline on MogpEmulators must agree."""
    import csv

    rng = np.random.default_rng(5)
    d = 2
    domain = api.SimulatorDomain([(0.0, 1.0)] * d)
    X2 = rng.uniform(0.05, 0.95, (8, d))
    X1 = np.vstack([X2, rng.uniform(0.05, 0.95, (14, d))])     # level-2 inputs are also run at level 1
    sims = {1: lambda x: x[1] + x[0] ** 2 + x[1] ** 2 - np.sqrt(2),
            2: lambda x: x[1] + x[0] ** 2 + x[1] ** 2 - np.sqrt(2) + 0.3 * np.sin(2 * np.pi * x[0])}
    paths = {}
    for lv, X in ((1, X1), (2, X2)):
        paths[lv] = tmp_path / f"level{lv}.csv"
        with open(paths[lv], "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["x1", "x2", "y"])
            for x in X:
                w.writerow([repr(float(x[0])), repr(float(x[1])), repr(float(sims[lv](x)))])
    raw = api.MultiLevel({lv: api.TrainingDatum.read_from_csv(paths[lv], header=True) for lv in (1, 2)})
    data = api.designers.create_data_for_multi_level_loo_sampling(raw, correlations=1)
    coeffs = api.designers.compute_delta_coefficients(2, correlations=1)
    assert len(data[2]) == 8 and len(data[1]) == 14          # shared inputs live at the top level only
    outs = []
    for make in (lambda: api.B200GaussianProcess(kernel="Matern52"), lambda: api.MogpEmulator(kernel="Matern52")):
        mlgp = api.MultiLevelGaussianProcess([make() for _ in range(2)], coefficients=coeffs)
        np.random.seed(77)
        mlgp.fit(data)
        np.random.seed(78)
        new = api.designers.compute_multi_level_loo_samples(mlgp, domain, api.MultiLevel([1.0, 8.0]), batch_size=1,
                                                            seeds=api.MultiLevel([5, 6]))
        np.seterr(all="warn")
        outs.append((mlgp, new))
    (ours_gp, ours), (ref_gp, ref) = outs
    for lv in (1, 2):
        a, b = ours_gp[lv].fit_hyperparameters, ref_gp[lv].fit_hyperparameters
        assert np.allclose(a.corr_length_scales, b.corr_length_scales, rtol=1e-3)
        assert np.isclose(a.process_var, b.process_var, rtol=1e-3)
    assert ours.levels == ref.levels
    out_csv = tmp_path / "new_design_points.csv"
    with open(out_csv, "w", newline="") as f:
        w = csv.writer(f)
        for lv in ours.levels:
            for x in ours[lv]:
                assert x in domain
                w.writerow([lv, *np.array(x)])
    rows = list(csv.reader(open(out_csv)))
    assert len(rows) == sum(len(ours[lv]) for lv in ours.levels) >= 1
    for lv in ref.levels:
        for a, b in zip(ours[lv], ref[lv]):
            assert np.allclose(np.array(a), np.array(b), atol=5e-3)


def test_unchanged_single_level_designer_with_fast_domain_and_batched_de(api):
    """The reference's compute_single_level_loo_samples, UNCHANGED, given a FastSimulatorDomain and run inside
    accelerate_designers() (differential-evolution populations evaluated per generation on the device).  The PEI
    surface is multi-modal and deferred-update DE does not follow the scalar run's trajectory, so the point is
    not compared with the golden scalar-path point but (a) with a direct call of the same batched optimiser on
    the same errors GP -- the hook must add nothing -- and (b) by value: the reference's own PEICalculator must
    rate it at least as high as the golden point of the scalar path (config 1, tests/golden/slas_config1.json)."""
    from exauq_toolbox_b200 import _lib
    from exauq_toolbox_b200.accel import FastSimulatorDomain, accelerate_designers
    from exauq_toolbox_b200.design import maximise_pei_de, pseudopoints

    g = load_golden("slas_config1.json")
    bounds = [tuple(b) for b in g["domain"]]
    domain = FastSimulatorDomain(bounds)
    data = _data(api, g)
    np.random.seed(g["gp_fit_seed"])
    gp = api.B200GaussianProcess(kernel="Matern52")
    gp.fit(data)
    np.random.seed(g["errors_gp_fit_seed"])
    gp_e = api.designers.compute_loo_errors_gp(gp, domain)
    n0 = _lib.launch_count()
    np.random.seed(g["errors_gp_fit_seed"])          # the designer fits its own errors GP: same restarts as gp_e above
    with accelerate_designers():
        new = api.designers.compute_single_level_loo_samples(gp, domain, batch_size=1, seed=g["de_seed"])
    assert _lib.launch_count() > n0 and new[0] in domain
    # (b) by value, through the reference's calculator on the errors GP
    pei = api.designers.PEICalculator(domain, gp_e)
    assert pei.compute(new[0]) >= 0.999 * pei.compute(api.Input(*g["new_design_point"]))
    # (a) the same optimiser called directly
    X = np.array([list(d.input) for d in gp_e.training_data])
    nes = np.array([d.output for d in gp_e.training_data])
    x_direct, v_direct = maximise_pei_de(gp_e._real_state().dev, bounds, pseudopoints(X, bounds), float(nes.max()),
                                         seed=api.designers.generate_seeds(g["de_seed"], 1)[0])
    assert np.allclose(np.array(new[0]), x_direct, atol=1e-6), (np.array(new[0]), x_direct)
    # a batch of 3: distinct points, all inside the domain
    np.random.seed(g["errors_gp_fit_seed"])
    with accelerate_designers():
        batch = api.designers.compute_single_level_loo_samples(gp, domain, batch_size=3, seed=g["de_seed"])
    assert len(batch) == 3 and all(x in domain for x in batch)
    assert len({tuple(np.round(np.array(x), 6)) for x in batch}) == 3


def test_batched_de_pei_equals_peicalculator_at_the_optimum(api):
    """batched_maximise returns (x, PEI(x)) like maximise: the value must be the reference PEICalculator's own
    value at that point.  (Which local maximum of the multi-modal PEI surface a differential-evolution run ends in
    depends on its update rule -- the deferred-update population run and the reference's immediate-update scalar
    run differ there, in either direction -- so the two optima are not compared.)"""
    from exauq_toolbox_b200.accel import FastSimulatorDomain, batched_maximise

    g = load_golden("fixed_matern_d3_n40.json")
    data = _data(api, g)
    hp = g["hyperparameters"]
    gp = api.B200GaussianProcess(kernel="Matern52")
    gp.fit(data, hyperparameters=api.GaussianProcessHyperparameters(hp["corr_length_scales"], hp["process_var"],
                                                                   hp["nugget"]))
    lo = np.min(np.array(g["X"]), axis=0) - 0.05
    hi = np.max(np.array(g["X"]), axis=0) + 0.05
    domain = FastSimulatorDomain([(float(a), float(b)) for a, b in zip(lo, hi)])
    pei = api.designers.PEICalculator(domain, gp)
    x, val = batched_maximise(lambda x: pei.compute(x), domain, seed=7)
    assert x in domain
    assert val > 0.0 and close(val, pei.compute(x), rel=1e-7, abs_=1e-15)
