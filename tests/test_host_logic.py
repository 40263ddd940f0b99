"""Host-side logic that needs no GPU: duplicate detection, priors, bounds, pseudopoints, the
MAP restart driver's batching, the drop-in class's argument validation, and the design batch
on one rank (with the oracle standing in for the device)."""

import numpy as np
import pytest

from conftest import load_golden, needs_exauq
from fakes import OracleCandidates, OracleGP


def test_find_duplicate_rows_matches_brute_force():
    pytest.importorskip("exauq")
    from exauq_toolbox_b200.gp import find_duplicate_rows

    rng = np.random.default_rng(0)
    X = rng.uniform(0, 1, (300, 3))
    assert find_duplicate_rows(X) is None
    X[217] = X[40] + np.array([5e-10, -5e-10, 0.0])       # within abs tolerance 1e-9
    assert find_duplicate_rows(X) == (40, 217)
    X[217, 1] += 1e-6
    assert find_duplicate_rows(X) is None
    Y = np.array([[1e6, 2.0], [1e6 * (1 + 5e-10), 2.0]])   # within relative tolerance
    assert find_duplicate_rows(Y) == (0, 1)


@needs_exauq
def test_pseudopoints_and_bounds_match_reference():
    from exauq.core.designers import _compute_loo_error_bounds
    from exauq.core.modelling import GaussianProcessHyperparameters as HP
    from exauq.core.modelling import Input, SimulatorDomain

    from exauq_toolbox_b200.design import loo_error_raw_bounds, pseudopoints

    rng = np.random.default_rng(1)
    for bounds, n in [([(0, 1), (-0.5, 1.5)], 8), ([(0, 1)] * 3, 20), ([(0, 2), (1, 3), (0, 1), (-1, 1)], 30)]:
        lo, hi = np.array([b[0] for b in bounds], float), np.array([b[1] for b in bounds], float)
        X = lo + (hi - lo) * rng.uniform(0, 1, (n, len(bounds)))
        X[0] = lo                      # a corner input must be skipped (modelling.py:2212)
        X[1, 0] = lo[0]                # already on a face
        dom = SimulatorDomain(bounds)
        ref = np.array([list(p) for p in dom.calculate_pseudopoints([Input(*r) for r in X])])
        assert np.array_equal(ref, pseudopoints(X, bounds))
        # raw bounds: (None, -2 ln(lower)) per length scale, (None, None) for the variance
        raw = loo_error_raw_bounds(bounds)
        for (lo_b, _), (rlo, rhi) in zip(_compute_loo_error_bounds(dom)[:-1], raw[:-1]):
            assert rlo is None and np.isclose(rhi, HP.transform_corr(lo_b))
        assert raw[-1] == (None, None)


def test_map_driver_batches_and_matches_sequential_reference_fit():
    """run_map_restarts = the restart loop of exauq/utilities/mogp_fitting.py:309-346: same
    starting points (same numpy seed), same optimiser => same optimum as the oracle's
    sequential fit_GP_MAP; evaluations arrive in batches."""
    pytest.importorskip("exauq")
    import mogp_emulator as mogp

    from exauq_toolbox_b200.gp import run_map_restarts
    from exauq_toolbox_b200.priors import MapPriors

    g = load_golden("fixed_matern_d3_n40.json")
    X, y = np.array(g["X"]), np.array(g["y"])
    np.random.seed(3)
    ref = mogp.fit_GP_MAP(mogp.GaussianProcess(X, y, kernel="Matern52"), n_tries=6)
    np.seterr(all="warn")
    pri = MapPriors(X)
    np.random.seed(3)
    starts = [pri.sample() for _ in range(6)]
    dev = OracleGP(X, y, "Matern52")
    stats = {}
    best = run_map_restarts(dev, pri, starts, None, 0.0, True, stats=stats)
    assert np.allclose(best, ref.theta.get_data(), rtol=1e-6, atol=1e-6)
    assert stats["evals"] > stats["batches"] >= 1
    assert dev.n_batches == stats["batches"]
    # every restart failing => None (the caller raises RuntimeError("GP fitting failed"))

    class Broken(OracleGP):
        def logpost_batch(self, theta_raw, nugget=0.0, adaptive=False, want_grad=True):
            R = np.atleast_2d(theta_raw).shape[0]
            return np.full(R, np.nan), np.full((R, self.d + 1), np.nan), np.ones(R, dtype=np.int32), np.zeros(R)

    assert run_map_restarts(Broken(X, y, "Matern52"), pri, starts[:2], None, 0.0, True) is None


@needs_exauq
def test_dropin_argument_validation_without_gpu():
    """Checks that run before any device call mirror MogpEmulator (exauq/core/emulators.py:129-190,
    268-282, 350-395, 466-468, 527-544, 633-641)."""
    from exauq.core.modelling import AbstractGaussianProcess, Input, TrainingDatum

    from exauq_toolbox_b200.gp import B200GaussianProcess

    with pytest.raises(ValueError, match="not a supported kernel function"):
        B200GaussianProcess(kernel="Cubic")
    with pytest.raises(RuntimeError):
        B200GaussianProcess(foo=1)
    gp = B200GaussianProcess(kernel="Matern52", inputs=[[1.0]], targets=[2.0])   # inputs/targets ignored
    assert isinstance(gp, AbstractGaussianProcess)
    assert gp.training_data == () and gp.fit_hyperparameters is None and gp.kinv.size == 0
    assert gp.fit([]) is None and gp.fit(None) is None and gp.training_data == ()
    with pytest.raises(TypeError, match="finite collection of TrainingDatum"):
        gp.fit([1, 2, 3])
    data = [TrainingDatum(Input(0.1, 0.2), 1.0), TrainingDatum(Input(0.3, 0.4), 2.0)]
    with pytest.raises(TypeError, match="Expected 'hyperparameters' to be None or of type"):
        gp.fit(data, hyperparameters=(1.0, 2.0))
    with pytest.raises(ValueError, match="not unique within tolerance"):
        gp.fit(data + [TrainingDatum(Input(0.1, 0.2 + 1e-10), 3.0)])
    with pytest.raises(ValueError, match="Invalid bound"):
        gp.fit(data, hyperparameter_bounds=[(2.0, 1.0), (None, None), (None, None)])
    with pytest.raises(TypeError, match="Invalid bound"):
        gp.fit(data, hyperparameter_bounds=[("a", 1.0), (None, None), (None, None)])
    with pytest.raises(ValueError, match="Upper bounds must be positive numbers"):
        B200GaussianProcess._compute_raw_param_bounds([(None, -1.0), (None, None)])
    raw = B200GaussianProcess._compute_raw_param_bounds([(0.5, 2.0), (None, 3.0)])
    assert np.allclose(raw[0], (-2 * np.log(2.0), -2 * np.log(0.5))) and raw[1][0] is None
    with pytest.raises(TypeError):
        gp.predict([0.1, 0.2])
    with pytest.raises(RuntimeError, match="has not been trained"):
        gp.predict(Input(0.1, 0.2))
    assert gp.correlation([], [Input(1.0)]).size == 0
    with pytest.raises(TypeError):
        gp.correlation(1, 2)
    with pytest.raises(TypeError):
        gp.correlation([1.0], [Input(1.0)])
    with pytest.raises(AssertionError):
        gp.correlation([Input(1.0)], [Input(1.0)])
    assert gp.covariance_matrix([Input(1.0, 2.0)]).size == 0


def test_design_batch_single_rank_on_oracle_device():
    """design.design_batch end to end on one rank with the oracle as the device: the chosen
    candidates are exactly the oracle's PEI batch arg max for the fitted errors GP."""
    pytest.importorskip("exauq")
    from oracle import gp_oracle as orc

    from exauq_toolbox_b200.design import design_batch, pseudopoints

    X, y = orc.synthetic_problem(40, 2)
    Xc = orc.synthetic_candidates(400, 2)
    bounds = [(0.0, 1.0)] * 2
    stats = {}
    pts, gidx, vals, theta = design_batch(X, y, "Matern52", [0.5, 0.5], 1.7, 0.0, bounds, Xc, batch_size=3, n_tries=4,
                                          seed=5, gp_factory=OracleGP, cand_factory=OracleCandidates, stats=stats)
    assert len(set(gidx.tolist())) == 3 and np.allclose(pts, Xc[gidx])
    _, _, nes = OracleGP(X, y, "Matern52").fit(np.full(2, -2 * np.log(0.5)), 1.7, 0.0).loo()
    ge = orc.fit(X, nes, "Matern52", theta[:2], float(np.exp(theta[2])), "adaptive")
    oi, ov, _ = orc.pei_batch_argmax(ge, "Matern52", Xc, pseudopoints(X, bounds), float(nes.max()), 3)
    assert gidx.tolist() == oi.tolist() and np.allclose(vals, ov, rtol=1e-10)
    assert stats["map_evals"] > 0


@needs_exauq
def test_multi_level_loo_errors_match_reference_on_oracle_device():
    """design.multi_level_loo_errors vs the reference's compute_multi_level_loo_error_data
    (exauq/core/designers.py:1045-1105) run with MogpEmulator on the restated mogp."""
    from exauq.core.designers import compute_multi_level_loo_error_data
    from exauq.core.emulators import MogpEmulator, MogpHyperparameters
    from exauq.core.modelling import Input, MultiLevel, MultiLevelGaussianProcess, TrainingDatum

    from exauq_toolbox_b200.design import find_cross_level_duplicate, multi_level_loo_errors

    rng = np.random.default_rng(4)
    d = 2
    Xs = {1: rng.uniform(0, 1, (14, d)), 2: rng.uniform(0, 1, (9, d)), 3: rng.uniform(0, 1, (6, d))}
    fs = {1: lambda x: x[1] + x[0] ** 2, 2: lambda x: np.sin(3 * x[0]) * x[1], 3: lambda x: 0.1 * np.cos(5 * x[0] + x[1])}
    hps = {1: ([0.6, 0.9], 2.0, 0.0), 2: ([0.3, 0.4], 1.5, 1e-8), 3: ([0.5, 0.5], 0.2, 0.0)}
    data = MultiLevel({lv: [TrainingDatum(Input(*x), float(fs[lv](x))) for x in Xs[lv]] for lv in Xs})
    ref = MultiLevelGaussianProcess([MogpEmulator(kernel="Matern52") for _ in range(3)], coefficients=[1.0, 0.8, 1.3])
    ref.fit(data, hyperparameters=MultiLevel({k: MogpHyperparameters(*v) for k, v in hps.items()}))
    ref_err = compute_multi_level_loo_error_data(ref)
    levels = {lv: dict(X=Xs[lv], y=np.array([fs[lv](x) for x in Xs[lv]]), kernel="Matern52",
                       corr_length_scales=hps[lv][0], process_var=hps[lv][1], nugget=hps[lv][2]) for lv in Xs}
    got, _ = multi_level_loo_errors(levels, {lv: ref.coefficients[lv] for lv in Xs}, gp_factory=OracleGP)
    for lv in Xs:
        assert np.allclose(got[lv], [dt.output for dt in ref_err[lv]], rtol=1e-7)
    assert find_cross_level_duplicate(Xs) is None
    Xs[3][2] = Xs[1][5] + 3e-10
    rep = find_cross_level_duplicate(Xs)
    assert rep is not None and {rep[1], rep[2]} == {1, 3}


def test_multi_level_design_batch_on_oracle_device():
    """The level/point chosen each round maximises PEI_l / cost_l over levels and candidates, and
    only the chosen level's repulsion set grows (exauq/core/designers.py:1357-1376)."""
    pytest.importorskip("exauq")
    from oracle import gp_oracle as orc

    from exauq_toolbox_b200.design import multi_level_design_batch, multi_level_loo_errors, pseudopoints

    rng = np.random.default_rng(7)
    d = 2
    bounds = [(0.0, 1.0)] * d
    levels = {}
    for lv, n, f in [(1, 16, lambda X: X[:, 1] + X[:, 0] ** 2), (2, 10, lambda X: np.sin(6 * X[:, 0]) * X[:, 1])]:
        X = rng.uniform(0, 1, (n, d))
        levels[lv] = dict(X=X, y=f(X), kernel="Matern52", corr_length_scales=[0.4, 0.5], process_var=1.2, nugget=0.0)
    coeff, costs = {1: 1.0, 2: 1.0}, {1: 1.0, 2: 11.0}
    Xc = orc.synthetic_candidates(300, d)
    chosen = multi_level_design_batch(levels, coeff, costs, bounds, Xc, batch_size=4, n_tries=3, seed=11,
                                      gp_factory=OracleGP, cand_factory=OracleCandidates)
    assert len(chosen) == 4 and all(c[0] in (1, 2) for c in chosen)
    assert len({(c[0], c[1]) for c in chosen}) == 4          # a (level, candidate) pair is never chosen twice
    for lv, gi, pt, v in chosen:
        assert np.allclose(pt, Xc[gi]) and v >= 0
    vals = [c[3] for c in chosen]
    assert vals[0] == max(vals)                                # the first round sees the largest PEI / cost


@needs_exauq
def test_vectorised_de_maximiser_finds_the_reference_maximum():
    """design.maximise_pei_de (population evaluated per generation, oracle device) vs the
    reference's maximise(pei.compute) (exauq/utilities/optimisation.py:14-104) on the same PEI."""
    from exauq.core.designers import PEICalculator
    from exauq.core.emulators import MogpEmulator, MogpHyperparameters
    from exauq.core.modelling import Input, SimulatorDomain, TrainingDatum
    from exauq.utilities.optimisation import maximise

    from exauq_toolbox_b200.design import maximise_pei_de, pseudopoints

    g = load_golden("fixed_matern_d3_n40.json")
    X, y = np.array(g["X"])[:, :2], np.array(g["y"])
    bounds = [(0.0, 1.0)] * 2
    hp = ([0.4, 0.4], 1.3, 0.0)
    ref = MogpEmulator(kernel="Matern52")
    ref.fit([TrainingDatum(Input(*x), float(t)) for x, t in zip(X, y)], hyperparameters=MogpHyperparameters(*hp))
    pei = PEICalculator(SimulatorDomain(bounds), ref)
    x_ref, v_ref = maximise(lambda x: pei.compute(x), SimulatorDomain(bounds), seed=3)
    dev = OracleGP(X, y, "Matern52").fit(-2 * np.log(np.array(hp[0])), hp[1], hp[2])
    x, v = maximise_pei_de(dev, bounds, pseudopoints(X, bounds), float(y.max()), seed=3, cand_factory=OracleCandidates)
    assert abs(v - v_ref) <= 1e-6 * abs(v_ref)
    assert np.allclose(x, np.array(x_ref), atol=1e-4)


@needs_exauq
def test_vectorised_repulsion_point_collection_equals_the_reference_loop():
    """accel.add_repulsion_points_vectorised (patched in by accelerate_designers) keeps exactly the points, in the order,
    that PEICalculator._add_repulsion_points keeps (exauq/core/designers.py:601-608): pseudopoints at construction,
    then points equal within tolerance to a training input / an earlier point / each other, and new ones."""
    from exauq.core.designers import PEICalculator
    from exauq.core.emulators import MogpEmulator, MogpHyperparameters
    from exauq.core.modelling import Input, SimulatorDomain, TrainingDatum

    from exauq_toolbox_b200.accel import accelerate_designers

    g = load_golden("fixed_matern_d3_n40.json")
    X, y = np.array(g["X"]), np.array(g["y"])
    X[3] = [0.0, 1.0, 0.5]                                    # a training input on an edge: its pseudopoints coincide
    X[7] = [1.0, 1.0, 1.0]                                    # ... and one in a corner
    bounds = [(0.0, 1.0)] * 3
    ref = MogpEmulator(kernel="Matern52")
    ref.fit([TrainingDatum(Input(*x), float(t)) for x, t in zip(X, y)], hyperparameters=MogpHyperparameters([0.4] * 3, 1.3, 0.0))
    domain = SimulatorDomain(bounds)
    extra = [Input(0.2, 0.3, 0.4), Input(*X[5]), Input(*(X[6] + 5e-10)), Input(0.2 + 1e-12, 0.3, 0.4), Input(0.9, 0.1, 0.5),
             Input(0.9, 0.1, 0.5)]
    stock = PEICalculator(domain, ref, additional_repulsion_pts=extra[:2])
    stock.add_repulsion_points(extra[2:])
    with accelerate_designers():
        fast = PEICalculator(domain, ref, additional_repulsion_pts=extra[:2])
        fast.add_repulsion_points(extra[2:])
    a, b = stock._other_repulsion_points, fast._other_repulsion_points
    assert len(a) == len(b) and all(p is q or tuple(p.value) == tuple(q.value) for p, q in zip(a, b))
    assert Input(0.2, 0.3, 0.4) in b and sum(1 for p in b if p == Input(0.9, 0.1, 0.5)) == 1
    assert all(not (p == Input(*X[5])) for p in b)             # a training input is never an "other" repulsion point
    # restored on exit
    import exauq.core.designers as designers
    assert designers.PEICalculator._add_repulsion_points.__name__ == "_add_repulsion_points"


@needs_exauq
def test_cross_level_repetition_check_returns_what_the_reference_returns():
    """accel.find_input_repetition_across_levels against exauq.core.designers._find_input_repetition_across_levels
    (designers.py:925-955): the same (input object, level, level) -- the FIRST equal pair of the reference's
    product / combinations enumeration -- over random multi-level data with planted repetitions (exact, within
    tolerance, several at once, between any two levels)."""
    from exauq.core.designers import _find_input_repetition_across_levels as reference
    from exauq.core.modelling import Input, MultiLevel, TrainingDatum

    from exauq_toolbox_b200.accel import find_input_repetition_across_levels as fast

    rng = np.random.default_rng(11)
    n_hits = 0
    for trial in range(120):
        n_levels = int(rng.integers(1, 4))
        sizes = [int(rng.integers(1, 6)) for _ in range(n_levels)]
        level_ids = list(rng.permutation([1, 2, 3, 5])[:n_levels])
        pts = [rng.uniform(0, 1, (n, 2)) for n in sizes]
        for _ in range(int(rng.integers(0, 3))):                # plant up to two repetitions
            if n_levels < 2:
                break
            a, b = rng.choice(n_levels, 2, replace=False)
            pts[b][rng.integers(sizes[b])] = pts[a][rng.integers(sizes[a])] + rng.choice([0.0, 4e-10, -4e-10, 5e-9])
        data = MultiLevel({int(lv): tuple(TrainingDatum(Input(*x), 1.0) for x in p) for lv, p in zip(level_ids, pts)})
        want, got = reference(data), fast(data)
        assert (want is None) == (got is None), trial
        if want is not None:
            n_hits += 1
            assert got[0] is want[0] and got[1:] == want[1:], (trial, want, got)
        assert fast(data) is got or fast(data) == got          # remembered answer for the same tuples
    assert n_hits >= 30
    assert fast(MultiLevel({1: (TrainingDatum(Input(0.5), 1.0),), 2: ()})) is None


def test_batched_polish_equals_scipys_scalar_polish():
    """design._polish_batched: the L-BFGS-B polish of scipy's differential evolution with the d + 1 points of each
    finite-difference gradient in ONE objective call -- same iterates as scipy's own scalar 2-point scheme, incl. at a bound."""
    from scipy.optimize import minimize

    from exauq_toolbox_b200.design import _polish_batched

    c = np.array([0.3, 2.5, 0.55])            # the second coordinate's optimum lies beyond the upper bound
    calls = []

    def f_scalar(x):
        x = np.asarray(x, dtype=float)
        return float(np.sum((x - c) ** 4) + 0.1 * np.sum(np.cos(5 * x)) + 0.05 * x[0] * x[2])

    def f_batched(xt):                        # (d, S) like a vectorised DE objective
        pts = np.asarray(xt, dtype=float).T
        calls.append(pts.shape[0])
        return np.array([f_scalar(p) for p in pts])

    bounds = [(0.0, 1.0)] * 3
    x0 = np.array([0.9, 0.8, 0.1])
    ref = minimize(f_scalar, x0, method="L-BFGS-B", bounds=bounds)
    x, fx = _polish_batched(f_batched, x0, f_scalar(x0), bounds)
    assert set(calls) == {4}                  # every objective call carries the point and its d neighbours
    assert len(calls) <= ref.nfev // 4 + 1    # scipy counts each of the d + 1 scalar evaluations
    assert np.allclose(x, ref.x, atol=1e-7) and abs(fx - ref.fun) <= 1e-12 * max(1.0, abs(ref.fun))
    assert x[1] == 1.0                        # stays on the bound
    # never worse than the start, and a start that is given a better value than it has is returned unchanged
    x2, f2 = _polish_batched(f_batched, x, fx, bounds)
    assert f2 <= fx and np.allclose(x2, x, atol=1e-4)
    x3, f3 = _polish_batched(f_batched, x, -1e9, bounds)
    assert np.array_equal(x3, x) and f3 == -1e9


def test_lockstep_lbfgsb_stepper_reproduces_scipy_minimize():
    """_LbfgsbStepper drives scipy's own L-BFGS-B core: iterates, optimum and evaluation count equal
    scipy.optimize.minimize(method='L-BFGS-B') bit for bit, with and without bounds."""
    pytest.importorskip("exauq")
    from scipy.optimize import minimize

    from exauq_toolbox_b200.gp import _get_setulb, _LbfgsbStepper

    setulb = _get_setulb()
    if setulb is None:
        pytest.skip("this scipy does not expose the reverse-communication core in the expected form")

    def f(x):
        return float(np.sum((x - np.array([1.0, -2.0, 0.5])) ** 4) + np.sum(np.cos(3 * x)) + x[0] * x[1])

    def g(x):
        return 4 * (x - np.array([1.0, -2.0, 0.5])) ** 3 - 3 * np.sin(3 * x) + np.array([x[1], x[0], 0.0])

    for bounds in (None, [(None, 0.8), (-1.5, None), (0.0, 1.0)]):
        x0 = np.array([0.3, 0.2, 2.0])
        ref = minimize(f, x0, method="L-BFGS-B", jac=g, bounds=bounds)
        st = _LbfgsbStepper(setulb, x0, bounds)
        evals = 0
        while st.advance() == "FG":
            st.feed(f(st.x), g(st.x))
            evals += 1
        assert np.array_equal(st.x, ref.x) and float(st.f) == float(ref.fun)
        assert st.n_iterations == ref.nit


def test_map_driver_lockstep_and_threaded_agree():
    """Both drivers (lock-step stepper, one thread per minimize) give the same optimum and the same
    per-restart results on the oracle device."""
    pytest.importorskip("exauq")
    from exauq_toolbox_b200.gp import run_map_restarts
    from exauq_toolbox_b200.priors import MapPriors

    g = load_golden("fixed_matern_d3_n40.json")
    X, y = np.array(g["X"]), np.array(g["y"])
    pri = MapPriors(X)
    np.random.seed(9)
    starts = [pri.sample() for _ in range(5)]
    bounds = [(None, 3.0)] * 3 + [(None, None)]
    s1, s2 = {}, {}
    a = run_map_restarts(OracleGP(X, y, "Matern52"), pri, starts, bounds, 0.0, True, stats=s1, return_all=True, lockstep=True)
    b = run_map_restarts(OracleGP(X, y, "Matern52"), pri, starts, bounds, 0.0, True, stats=s2, return_all=True, lockstep=False)
    for ra, rb in zip(a, b):
        assert (ra is None) == (rb is None)
        if ra is not None:
            assert ra[0] == rb[0] and np.array_equal(ra[1], rb[1])
    assert s1["evals"] <= s2["evals"] and s1["batches"] >= 1


def test_map_priors_fast_paths_are_bit_identical_to_the_scipy_formulas():
    """priors.py evaluates the default inverse-gamma priors through vectorised expressions and a direct
    gammaincc call; both must give exactly the doubles of the per-dimension formulas / scipy.stats.invgamma
    they replace, otherwise the L-BFGS-B trajectories of the reference are not reproduced."""
    import scipy.stats
    from scipy.optimize import root
    from scipy.stats.qmc import LatinHypercube

    from exauq_toolbox_b200 import priors as P

    rng = np.random.default_rng(0)
    with np.errstate(all="ignore"):
        for _ in range(3000):
            a, b, x = np.exp(rng.uniform(-30, 30)), np.exp(rng.uniform(-30, 30)), 10 ** rng.uniform(-8, 2)
            r, f = scipy.stats.invgamma.cdf(x, a, scale=b), P._invgamma_cdf(x, a, b)
            assert r == f or (np.isnan(r) and np.isnan(f))
        for a, b, x in [(np.inf, 1.0, 0.1), (1.0, np.inf, 0.1), (0.0, 1.0, 0.1), (1.0, 0.0, 0.1), (1.0, 1e-320, 0.1)]:
            r, f = scipy.stats.invgamma.cdf(x, a, scale=b), P._invgamma_cdf(x, a, b)
            assert r == f or (np.isnan(r) and np.isnan(f))

    def solve_with_scipy_stats(lo, hi):
        def resid(z):
            a, b = np.exp(z[0]), np.exp(z[1])
            cdf = scipy.stats.invgamma.cdf
            return np.array([cdf(lo, a, scale=b) - 0.005, cdf(hi, a, scale=b) - 0.995])

        with np.errstate(all="ignore"):
            sol = root(resid, np.zeros(2))
        return (float(np.exp(sol["x"][0])), float(np.exp(sol["x"][1]))) if sol["success"] else None

    X = LatinHypercube(4, seed=3).random(200)
    X[:, 2] = 0.25                                   # a degenerate dimension gets a flat prior
    for col in X.T:
        u = np.unique(col)
        if u.size > 2:
            lo, hi = float(np.median(np.diff(u))), float(u[-1] - u[0])
            assert P._solve_invgamma(lo, hi) == solve_with_scipy_stats(lo, hi)
    for fit_nugget in (False, True):
        pri = P.MapPriors(X, fit_nugget=fit_nugget)
        if fit_nugget:
            pri.nugget = (2.5, 3e-7)                 # exercise the nugget term whatever the solver returned
        for _ in range(2000):
            th = rng.uniform(-4, 8, 5 + fit_nugget)
            if fit_nugget:
                th[5] = rng.uniform(-20, -10)
            lengths = np.exp(-0.5 * th[:4])
            ref = sum(pri._logp(x, prm) for x, prm in zip(lengths, pri.corr))
            refg = np.zeros_like(th)
            for k, (x, prm) in enumerate(zip(lengths, pri.corr)):
                refg[k] = pri._dlogp_dx(x, prm) * (-0.5 * x)
            if fit_nugget:
                nu = float(np.exp(th[5]))
                ref += pri._logp(nu, pri.nugget)
                refg[5] = pri._dlogp_dx(nu, pri.nugget) * nu
            assert pri.logp(th) == ref
            assert np.array_equal(pri.dlogp_draw(th), refg)


def test_committed_bench_lines_follow_the_contract():
    """The bench lines committed under profiles/ carry every key the measurement contract names (so a change of
    bench.py that drops one is caught without a GPU)."""
    import json
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    line = json.load(open(os.path.join(root, "profiles", "bench_r01_1gpu_final.json")))
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in line, key
    assert line["dtype"] == "f64" and line["data"] == "synthetic" and "workload" in line["config"]
    assert line["gpu_launches"] > 0 and line["value"] > 0
    for key in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert key in line["e2e"], key
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["value"] != line["value"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in line["roofline"], key
    assert abs(line["roofline"]["frac"] - line["roofline"]["achieved"] / line["roofline"]["peak"]) < 1e-12
    for key in ("value", "unit", "cores", "kind", "sample"):
        assert key in line["cpu_baseline"], key
    for key in ("sm_mhz", "sm_max_mhz", "reasons"):
        assert key in line["clocks"], key
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(line["clocks"]["reasons"])
    ref = json.load(open(os.path.join(root, "profiles", "bench_r01_reference_arm.json")))
    assert ref["impl"] == "reference" and ref["metric"] == line["metric"] and ref["unit"] == line["unit"]
    assert ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["cpu_baseline"]["kind"] in ("port", "reference")


def test_bench_workload_is_the_oracles_workload():
    """bench.py's two arms share exauq_toolbox_b200.workload; the oracle keeps its own copy for the tests -- same arrays."""
    import exauq_toolbox_b200.workload as w
    from oracle import gp_oracle as orc

    Xa, ya = w.synthetic_problem(64, 5)
    Xb, yb = orc.synthetic_problem(64, 5)
    assert np.array_equal(Xa, Xb) and np.array_equal(ya, yb)
    assert np.array_equal(w.synthetic_candidates(100, 5, seed=7), orc.synthetic_candidates(100, 5, seed=7))


@needs_exauq
def test_fast_simulator_domain_equals_the_reference_domain():
    """FastSimulatorDomain (vectorised corners / pseudopoints) against exauq's SimulatorDomain: same points in the
    same order, same errors, usable wherever a SimulatorDomain is expected (isinstance, deepcopy, scale, in)."""
    import copy

    from exauq.core.modelling import Input, SimulatorDomain

    from exauq_toolbox_b200.accel import FastSimulatorDomain

    rng = np.random.default_rng(0)
    for bounds in ([(0, 1), (0, 1)], [(-1, 2), (0, 0.5), (3, 4)], [(0, 1), (2, 2), (0, 1)], [(0, 1)] * 5):
        ref, fast = SimulatorDomain(bounds), FastSimulatorDomain(bounds)
        assert isinstance(fast, SimulatorDomain) and fast.dim == ref.dim and fast.bounds == ref.bounds
        assert fast.corners == ref.corners
        d = len(bounds)
        lo = np.array([b[0] for b in bounds], float)
        hi = np.array([b[1] for b in bounds], float)
        X = lo + rng.uniform(0, 1, (40, d)) * (hi - lo)
        X[3] = lo                                             # a corner among the inputs
        X[7, 0] = hi[0]                                       # an input on a face
        X[9] = X[8]                                           # duplicates
        inputs = [Input(*row) for row in X]
        assert fast.calculate_pseudopoints(inputs) == ref.calculate_pseudopoints(inputs)
        assert fast.closest_boundary_points(inputs) == ref.closest_boundary_points(inputs)
        assert fast.calculate_pseudopoints([]) == ref.calculate_pseudopoints([])
        assert fast.closest_boundary_points([ref.corners[0]]) == ref.closest_boundary_points([ref.corners[0]])
        c = copy.deepcopy(fast)
        assert c.calculate_pseudopoints(inputs) == ref.calculate_pseudopoints(inputs)
        assert (Input(*X[0]) in fast) and fast.scale([0.5] * d) == ref.scale([0.5] * d)
        outside = [Input(*(hi + 1.0))]
        for dom in (ref, fast):
            with pytest.raises(ValueError, match="within the domain bounds"):
                dom.calculate_pseudopoints(outside)
            with pytest.raises(ValueError, match="same dimensionality"):
                dom.calculate_pseudopoints([Input(*([0.5] * (d + 1)))])


@needs_exauq
def test_batched_maximise_falls_back_for_foreign_objectives():
    """accelerate_designers(): objectives that are not the PEI of a B200 GP reach the reference's maximise
    unchanged (same optimum, same seed), and the module attribute is restored afterwards."""
    import exauq.core.designers as designers
    from exauq.core.modelling import SimulatorDomain
    from exauq.utilities.optimisation import maximise

    from exauq_toolbox_b200.accel import accelerate_designers, batched_maximise

    dom = SimulatorDomain([(0, 1), (0, 1)])
    f = lambda x: -((x[0] - 0.3) ** 2 + (x[1] - 0.6) ** 2)  # noqa: E731
    a = maximise(f, dom, seed=3)
    b = batched_maximise(f, dom, seed=3)
    assert a[0] == b[0] and a[1] == b[1]
    before = designers.maximise
    with accelerate_designers():
        assert designers.maximise is not before
        c = designers.maximise(f, dom, seed=3)
    assert designers.maximise is before and c[0] == a[0]
    with pytest.raises(TypeError):
        batched_maximise(f, "not a domain")
