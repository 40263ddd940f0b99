"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference
(/root/reference/exauq) on top of the oracle's mogp restatement (oracle/mogp_emulator).

Run here (the dev container, where /root/reference exists):  python tests/golden/make_golden.py
The GPU box has no /root/reference: tests only read the JSON files this script wrote.
"""

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.environ.get("EXAUQ_REFERENCE", "/root/reference"))

import copy  # noqa: E402

import mogp_emulator as mogp  # noqa: E402
from exauq.core.designers import (  # noqa: E402
    PEICalculator,
    compute_loo_errors_gp,
    compute_loo_gp,
    compute_single_level_loo_samples,
    oneshot_lhs,
)
from exauq.core.emulators import MogpEmulator, MogpHyperparameters  # noqa: E402
from exauq.core.modelling import Input, SimulatorDomain, TrainingDatum  # noqa: E402
from scipy.stats.qmc import LatinHypercube  # noqa: E402


def sim_func(x):
    # docs/designers/tutorials/training_gp_tutorial.md:115-119
    return (x[1] + x[0] ** 2 + x[1] ** 2 - np.sqrt(2) + np.sin(2 * np.pi * x[0])
            + np.sin(4 * np.pi * x[0] * x[1]))


def smooth(X):
    d = X.shape[1]
    return np.sum(np.sin(2 * np.pi * X) / np.arange(1, d + 1), axis=1) + X[:, 0] * X[:, 1 % d] - np.sqrt(2.0)


def dump(name, obj):
    with open(os.path.join(HERE, name), "w") as f:
        json.dump(obj, f, indent=1)
    print("wrote", name)


def tutorial_case():
    domain = SimulatorDomain([(0, 1), (-0.5, 1.5)])
    lhs = oneshot_lhs(domain, 20, seed=1)
    data = [TrainingDatum(x, sim_func(x)) for x in lhs]
    np.random.seed(0)
    gp = MogpEmulator(kernel="Matern52")
    gp.fit(data)
    x = Input(0.5, 1.5)
    p = gp.predict(x)
    hp = gp.fit_hyperparameters
    return {
        "source": "docs/designers/tutorials/training_gp_tutorial.md:198-205,259-263,300-302",
        "published": {"estimate": 2.9826684809655353, "variance": 0.3252150944485681,
                      "nes_error": 0.7480506048975407, "observed": 2.5857864376269055,
                      "first_inputs": [[0.2744089188, 1.0049536304], [0.5927920194, -0.3948649447]],
                      "first_outputs": [1.3460506974, -2.0511268894]},
        "inputs": [list(map(float, d.input)) for d in data],
        "outputs": [float(d.output) for d in data],
        "numpy_seed": 0,
        "fit": {"corr_length_scales": list(map(float, hp.corr_length_scales)),
                "process_var": float(hp.process_var), "nugget": float(hp.nugget),
                "neg_log_posterior": float(gp.gp.current_logpost)},
        "predict_at": [0.5, 1.5],
        "estimate": float(p.estimate), "variance": float(p.variance),
        "nes_error": float(p.nes_error(sim_func(x))),
    }


def fixed_hp_case(kernel, N, d, ell, cov, nugget, seed):
    X = LatinHypercube(d, seed=seed).random(N)
    y = smooth(X)
    domain = SimulatorDomain([(0.0, 1.0)] * d)
    data = [TrainingDatum(Input(*row), float(t)) for row, t in zip(X, y)]
    hp = MogpHyperparameters([ell] * d, cov, nugget)
    gp = MogpEmulator(kernel=kernel)
    gp.fit(data, hyperparameters=hp)
    Xq = LatinHypercube(d, seed=seed + 100).random(12)
    q_inputs = [Input(*row) for row in Xq]
    preds = [gp.predict(x) for x in q_inputs]
    train_inputs = [dt.input for dt in data]
    corr = gp.correlation(q_inputs[:5], train_inputs)
    # the loop of compute_loo_errors_gp, exauq/core/designers.py:263-272
    loo_gp = copy.deepcopy(gp)
    loo = []
    for i, datum in enumerate(gp.training_data):
        compute_loo_gp(gp, i, loo_gp=loo_gp)
        pr = loo_gp.predict(datum.input)
        loo.append([float(pr.estimate), float(pr.variance), float(pr.nes_error(datum.output))])
    pei = PEICalculator(domain, gp)
    other = [list(map(float, x)) for x in pei._other_repulsion_points]
    pei_vals = [[float(pei.expected_improvement(x)), float(pei.repulsion(x)), float(pei.compute(x))]
                for x in q_inputs]
    # negative log-posterior and gradient (default priors) at a few raw parameter vectors
    m = mogp.GaussianProcess(X, y, kernel=kernel, nugget=nugget)
    rng = np.random.default_rng(seed)
    thetas = np.concatenate([np.full((3, d), -2 * np.log(ell)) + rng.uniform(-0.4, 0.4, (3, d)),
                             np.log(cov) + rng.uniform(-0.4, 0.4, (3, 1))], axis=1)
    lp = [[float(m.logposterior(t)), list(map(float, m.logpost_deriv(t)))] for t in thetas]
    priors = [[float(p.shape), float(p.scale)] if hasattr(p, "shape") else None for p in m.priors.corr]
    return {
        "kernel": kernel, "N": N, "d": d, "lhs_seed": seed,
        "X": X.tolist(), "y": y.tolist(),
        "hyperparameters": {"corr_length_scales": [ell] * d, "process_var": cov, "nugget": nugget},
        "query": Xq.tolist(),
        "predict": [[float(p.estimate), float(p.variance)] for p in preds],
        "correlation_first5": np.asarray(corr).tolist(),
        "kinv_diag": np.diag(gp.kinv).tolist(),
        "kinv_row0": gp.kinv[0].tolist(),
        "loo": loo,
        "max_target": float(pei._max_targets),
        "other_repulsion_points": other,
        "pei": pei_vals,
        "logpost_thetas": thetas.tolist(),
        "logpost": lp,
        "corr_priors": priors,
    }


def slas_case():
    """Config 1 (BASELINE.json configs[0]): N=50 oneshot_lhs on the 2-D toy simulator, Matern-5/2
    fit, then compute_single_level_loo_samples(batch_size=1)."""
    domain = SimulatorDomain([(0, 1), (-0.5, 1.5)])
    lhs = oneshot_lhs(domain, 50, seed=1)
    data = [TrainingDatum(x, sim_func(x)) for x in lhs]
    np.random.seed(0)
    gp = MogpEmulator(kernel="Matern52")
    gp.fit(data)
    np.random.seed(1)
    gp_e = compute_loo_errors_gp(gp, domain)
    np.random.seed(1)
    new_pts = compute_single_level_loo_samples(gp, domain, batch_size=1, seed=7)
    hp, hpe = gp.fit_hyperparameters, gp_e.fit_hyperparameters
    pei = PEICalculator(domain, gp_e)
    return {
        "domain": [[0, 1], [-0.5, 1.5]], "lhs_seed": 1, "N": 50,
        "inputs": [list(map(float, d.input)) for d in data],
        "outputs": [float(d.output) for d in data],
        "gp_fit_seed": 0,
        "gp_hyperparameters": {"corr_length_scales": list(map(float, hp.corr_length_scales)),
                               "process_var": float(hp.process_var), "nugget": float(hp.nugget)},
        "loo_nes_errors": [float(d.output) for d in gp_e.training_data],
        "errors_gp_fit_seed": 1,
        "errors_gp_hyperparameters": {"corr_length_scales": list(map(float, hpe.corr_length_scales)),
                                      "process_var": float(hpe.process_var), "nugget": float(hpe.nugget)},
        "de_seed": 7,
        "new_design_point": list(map(float, new_pts[0])),
        "pei_at_new_point": float(pei.compute(new_pts[0])),
    }


def mlgp_tutorial_case():
    """docs/designers/tutorials/training_multi_level_gp_tutorial.md: 2-level GP, LHS seed 1 with 20 / 8 points,
    level 1 = cheap simulator, level 2 = difference to the expensive one; predictions at (0.5, 1)."""
    from exauq.core.modelling import MultiLevel, MultiLevelGaussianProcess

    def sim_level1(x):
        return x[1] + x[0] ** 2 + x[1] ** 2 - np.sqrt(2)

    def sim_delta(x):
        return sim_func(x) - sim_level1(x)

    domain = SimulatorDomain([(0, 1), (-0.5, 1.5)])
    d1 = [TrainingDatum(x, sim_level1(x)) for x in oneshot_lhs(domain, 20, seed=1)]
    d2 = [TrainingDatum(x, sim_delta(x)) for x in oneshot_lhs(domain, 8, seed=1)]
    mlgp = MultiLevelGaussianProcess([MogpEmulator(kernel="Matern52"), MogpEmulator(kernel="Matern52")])
    np.random.seed(0)
    mlgp.fit(MultiLevel([d1, d2]))
    x = Input(0.5, 1)
    preds = [mlgp.predict(x, level) for level in mlgp.levels]
    hp = {lv: mlgp[lv].fit_hyperparameters for lv in mlgp.levels}
    return {
        "source": "docs/designers/tutorials/training_multi_level_gp_tutorial.md:339-349",
        "published": {"level1": {"estimate": 0.8353095922897182, "variance": 1.168910654314459e-05},
                      "level2": {"estimate": 0.9008758897257811, "variance": 0.6237762865894648}},
        "numpy_seed": 0,
        "level1": {"inputs": [list(map(float, d.input)) for d in d1], "outputs": [float(d.output) for d in d1]},
        "level2": {"inputs": [list(map(float, d.input)) for d in d2], "outputs": [float(d.output) for d in d2]},
        "fit": {str(lv): {"corr_length_scales": list(map(float, hp[lv].corr_length_scales)),
                          "process_var": float(hp[lv].process_var), "nugget": float(hp[lv].nugget)} for lv in hp},
        "predict_at": [0.5, 1.0],
        "predictions": [{"estimate": float(p.estimate), "variance": float(p.variance)} for p in preds],
    }


if __name__ == "__main__":
    dump("tutorial_mlgp.json", mlgp_tutorial_case())
    dump("tutorial_gp.json", tutorial_case())
    dump("fixed_matern_d3_n40.json", fixed_hp_case("Matern52", 40, 3, 0.5, 1.7, 0.0, 11))
    dump("fixed_sqexp_d2_n30.json", fixed_hp_case("SquaredExponential", 30, 2, 0.3, 0.8, 1e-6, 12))
    dump("fixed_matern_d8_n150.json", fixed_hp_case("Matern52", 150, 8, 0.5, 1.7, 0.0, 13))
    dump("slas_config1.json", slas_case())
