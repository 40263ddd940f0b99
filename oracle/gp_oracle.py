"""numpy/scipy restatement of the reference's hot path, one function per SURVEY.md
section 8a row.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

All GP arithmetic goes through ``oracle/mogp_emulator`` (the restated mogp-emulator 0.7.2);
the functions here restate what the *reference's own* code does around it and cite the
reference file:line they follow.  Parity pinning: see oracle/mogp_emulator/__init__.py --
the tutorial outputs and the reference's unit tests pin this; bit-level agreement with the
unavailable mogp wheel is unpinned.
"""

from __future__ import annotations

import math
import os
import sys

import numpy as np
from scipy.stats import norm

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)

import mogp_emulator as mogp  # noqa: E402  (the restated package next to this file)
from mogp_emulator.GPParams import GPParams  # noqa: E402
from mogp_emulator.Priors import GPPriors  # noqa: E402

KERNELS = {
    "Matern52": mogp.Kernel.Matern52,
    "SquaredExponential": mogp.Kernel.SquaredExponential,
    "ProductMat52": mogp.Kernel.ProductMat52,
}


def corr(kernel: str, corr_raw, A, B) -> np.ndarray:
    """exauq/core/emulators.py:495-567 (line 547): ``self._kernel(inputs1, inputs2, corr_raw)``."""
    return KERNELS[kernel]().kernel_f(np.atleast_2d(A), np.atleast_2d(B), np.asarray(corr_raw, float))


def fit(X, y, kernel: str, corr_raw, cov: float, nugget=0.0, flat_priors: bool = True):
    """exauq/core/emulators.py:418-450: ``GaussianProcess(inputs, targets, **kwargs)`` then
    ``gp.fit(hyperparameters.to_mogp_gp_params(nugget_type))``.  ``nugget`` is a float
    (type 'fixed') or the string 'adaptive'."""
    X = np.asarray(X, float)
    d = X.shape[1]
    kw = {}
    if flat_priors:
        kw["priors"] = GPPriors(n_corr=d, nugget_type="adaptive" if nugget == "adaptive" else "fixed")
    gp = mogp.GaussianProcess(X, np.asarray(y, float), kernel=kernel, nugget=nugget, **kw)
    params = GPParams(n_corr=d, nugget=nugget)
    params.set_data(np.concatenate([np.asarray(corr_raw, float), [math.log(cov)]]))
    gp.fit(params)
    return gp


def predict(gp, Xc):
    """exauq/core/emulators.py:612-667 (line 649) for many inputs at once: mean, variance."""
    res = gp.predict(np.atleast_2d(np.asarray(Xc, float)))
    return res.mean, res.unc


def neg_log_likelihood(gp) -> float:
    """0.5 (log|K| + y^T K^-1 y + n ln 2 pi): mogp ``logposterior`` without the prior term,
    as minimised at exauq/utilities/mogp_fitting.py:318-324."""
    return float(0.5 * (2.0 * np.sum(np.log(np.diag(gp.L))) + gp.targets @ gp.Kinv_t
                        + gp.n * np.log(2.0 * np.pi)))


def neg_log_likelihood_grad(X, y, kernel: str, theta_raw, nugget=0.0):
    """mogp ``logposterior`` / ``logpost_deriv`` with flat priors at raw parameters
    ``theta_raw = (corr_raw..., ln cov)`` (exauq/utilities/mogp_fitting.py:318-324)."""
    X = np.asarray(X, float)
    d = X.shape[1]
    gp = mogp.GaussianProcess(X, np.asarray(y, float), kernel=kernel, nugget=nugget,
                              priors=GPPriors(n_corr=d, nugget_type=nugget if isinstance(nugget, str) else "fixed"))
    theta_raw = np.asarray(theta_raw, float)
    val = gp.logposterior(theta_raw)
    grad = gp.logpost_deriv(theta_raw)
    return float(val), np.array(grad)


def nes_error(mean, var, observed):
    """exauq/core/modelling.py:754-825 vectorised; 0/0 -> 0 and x/0 -> inf as documented there."""
    mean, var, observed = np.broadcast_arrays(np.asarray(mean, float), np.asarray(var, float),
                                              np.asarray(observed, float))
    sq = (mean - observed) ** 2
    num = var + sq
    den = np.sqrt(2.0 * var**2 + 4.0 * var * sq)
    out = np.empty_like(num)
    zero = den == 0.0
    out[~zero] = num[~zero] / den[~zero]
    out[zero] = np.where(num[zero] == 0.0, 0.0, np.inf)
    return out


def loo_refit(X, y, kernel: str, corr_raw, cov: float, nugget: float, idx=None):
    """The loop of ``compute_loo_errors_gp`` (exauq/core/designers.py:263-272): for each
    left-out index refit on the remaining data with the parent's hyperparameters
    (``compute_loo_gp`` :295-370), predict at the left-out input, take the NES error."""
    X = np.asarray(X, float)
    y = np.asarray(y, float)
    idx = range(len(y)) if idx is None else idx
    means, vars_ = [], []
    for i in idx:
        keep = np.arange(len(y)) != i
        gp = fit(X[keep], y[keep], kernel, corr_raw, cov, nugget)
        m, v = predict(gp, X[i : i + 1])
        means.append(m[0])
        vars_.append(v[0])
    means, vars_ = np.array(means), np.array(vars_)
    return means, vars_, nes_error(means, vars_, y[list(idx)])


def expected_improvement(mean, var, tmax: float):
    """exauq/core/designers.py:610-658: 0 where sd is within 1e-9 of 0, else
    (m - tmax) Phi(u) + sd phi(u), u = (m - tmax) / sd."""
    mean = np.asarray(mean, float)
    sd = np.sqrt(np.asarray(var, float))
    out = np.zeros_like(mean)
    ok = np.abs(sd) > 1e-9
    u = (mean[ok] - tmax) / sd[ok]
    out[ok] = (mean[ok] - tmax) * norm.cdf(u) + sd[ok] * norm.pdf(u)
    return out


def repulsion(gp, kernel: str, Xc, other_pts):
    """exauq/core/designers.py:660-706: prod_j (1 - cov(x, x_j)/process_var) over training
    inputs times prod_p (1 - corr(x, r_p)) over the other repulsion points."""
    Xc = np.atleast_2d(np.asarray(Xc, float))
    cov = gp.theta.cov
    cov_matrix = cov * corr(kernel, gp.theta.corr_raw, Xc, gp.inputs)
    inputs_term = np.prod(1 - cov_matrix / cov, axis=1)
    if other_pts is None or len(other_pts) == 0:
        return inputs_term
    other = np.prod(1 - corr(kernel, gp.theta.corr_raw, Xc, np.atleast_2d(other_pts)), axis=1)
    return inputs_term * other


def pei(gp, kernel: str, Xc, other_pts, tmax: float):
    """exauq/core/designers.py:522-553: expected_improvement(x) * repulsion(x)."""
    m, v = predict(gp, Xc)
    return expected_improvement(m, v, tmax) * repulsion(gp, kernel, Xc, other_pts)


def pei_batch_argmax(gp, kernel: str, Xc, other_pts, tmax: float, batch_size: int):
    """The loop at exauq/core/designers.py:830-835 over a FIXED candidate set: pick the
    arg max, add it to the repulsion points (no refit), repeat."""
    Xc = np.atleast_2d(np.asarray(Xc, float))
    vals = pei(gp, kernel, Xc, other_pts, tmax)
    chosen, chosen_vals = [], []
    for _ in range(batch_size):
        i = int(np.argmax(vals))
        chosen.append(i)
        chosen_vals.append(vals[i])
        vals = vals * (1 - corr(kernel, gp.theta.corr_raw, Xc, Xc[i : i + 1])[:, 0])
    return np.array(chosen), np.array(chosen_vals), vals


def kinv(gp, kernel: str):
    """exauq/core/modelling.py:1095-1113: inverse of cov * correlation(train, train) (no nugget)."""
    k = gp.theta.cov * corr(kernel, gp.theta.corr_raw, gp.inputs, gp.inputs)
    return np.linalg.inv(k)


# ---- synthetic workload shared by tests and bench --------------------
def synthetic_target(X: np.ndarray) -> np.ndarray:
    """Deterministic test function: sum_k sin(6 pi x_k)/k + x_1 x_2 - sqrt 2.  Three periods per
    coordinate keep the leave-one-out errors of a length-scale-0.5 GP rough, so that the errors GP
    gets short length scales and the PEI product over N repulsion factors does not underflow to 0
    at N=4096 (with one period per coordinate every candidate's PEI is exactly 0.0 -- in the
    reference's arithmetic too -- and the arg max is meaningless)."""
    d = X.shape[1]
    k = np.arange(1, d + 1)
    return np.sum(np.sin(6 * np.pi * X) / k, axis=1) + X[:, 0] * X[:, 1 % d] - math.sqrt(2.0)


def synthetic_problem(N: int, d: int, seed: int = 1):
    from scipy.stats.qmc import LatinHypercube

    X = LatinHypercube(d, seed=seed).random(N)
    return X, synthetic_target(X)


def synthetic_candidates(M: int, d: int, seed: int = 2):
    from scipy.stats.qmc import LatinHypercube

    return LatinHypercube(d, seed=seed).random(M)
