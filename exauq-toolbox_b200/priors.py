"""Default hyperparameter priors used by MAP estimation (host side, O(d) work).

The reference estimates hyperparameters by minimising mogp's negative log-POSTERIOR
(exauq/utilities/mogp_fitting.py:289-346); with ``priors=None`` -- the only case
``MogpEmulator`` reaches (exauq/core/emulators.py:129-137) -- mogp-emulator 0.7.2 puts, per
input dimension, an inverse-gamma prior on the correlation length whose cdf is 0.005 at the
median gap between sorted unique coordinates and 0.995 at their range, a flat prior on the
process variance, and (only when the nugget is fitted) an inverse-gamma prior on the nugget
with the same construction on (1e-8, 1e-6).  Densities act on the linear-scale parameter;
no Jacobian term.  Starting points are drawn from these priors (flat ones from U(-2.5, 2.5)
on the raw scale) through numpy's global RNG, in parameter order, so that seeding
``np.random`` reproduces the reference's restarts.
"""

from __future__ import annotations

from typing import Optional

import numpy as np
import scipy.stats
from scipy.optimize import root
from scipy.special import gammaincc, gammaln


def _invgamma_cdf(x: float, a: float, b: float) -> float:
    """``scipy.stats.invgamma.cdf(x, a, scale=b)`` without the generic distribution machinery (35 ms -> 2 ms
    per training set): the same expression ``gammaincc(a, 1 / ((x - 0) / b))`` and the same handling of
    invalid parameters and of the support, so the values -- and the root finder's path -- are identical."""
    with np.errstate(all="ignore"):
        xs = np.float64(x - 0.0) / np.float64(b)
    if not (a > 0.0 and b > 0.0) or np.isnan(xs):
        return float("nan")
    if xs >= np.inf:
        return 1.0
    if not xs > 0.0:
        return 0.0
    return float(gammaincc(a, 1.0 / xs))


def _solve_invgamma(lo: float, hi: float) -> Optional[tuple[float, float]]:
    """(shape, scale) with cdf(lo) = 0.005 and cdf(hi) = 0.995, or None if the solver fails."""

    def resid(z):
        a, b = np.exp(z[0]), np.exp(z[1])
        return np.array([_invgamma_cdf(lo, a, b) - 0.005, _invgamma_cdf(hi, a, b) - 0.995])

    with np.errstate(all="ignore"):
        sol = root(resid, np.zeros(2))
    if not sol["success"]:
        return None
    return float(np.exp(sol["x"][0])), float(np.exp(sol["x"][1]))


# ---------------------------------------------------------------------------------------------------
# user-supplied priors (the ``priors=`` keyword MogpEmulator passes through, exauq/core/emulators.py:129-137)
# ---------------------------------------------------------------------------------------------------
class WeakPrior:
    """Improper flat prior (mogp ``WeakPrior``): contributes nothing; starting points from U(-2.5, 2.5) on the
    raw scale."""

    def logp(self, x):
        return 0.0

    def dlogpdx(self, x):
        return 0.0


class InvGammaPrior(WeakPrior):
    """Inverse-gamma density on the linear-scale parameter (mogp ``InvGammaPrior(shape, scale)``)."""

    def __init__(self, shape, scale):
        if not (shape > 0.0 and scale > 0.0):
            raise ValueError("shape and scale must be positive")
        self.shape, self.scale = float(shape), float(scale)

    def logp(self, x):
        return float(self.shape * np.log(self.scale) - gammaln(self.shape) - (self.shape + 1.0) * np.log(x) - self.scale / x)

    def dlogpdx(self, x):
        return float(-(self.shape + 1.0) / x + self.scale / x**2)

    def sample_x(self):
        return float(scipy.stats.invgamma.rvs(self.shape, scale=self.scale))


class GammaPrior(InvGammaPrior):
    """Gamma density (mogp ``GammaPrior(shape, scale)``)."""

    def logp(self, x):
        return float(-self.shape * np.log(self.scale) - gammaln(self.shape) + (self.shape - 1.0) * np.log(x) - x / self.scale)

    def dlogpdx(self, x):
        return float((self.shape - 1.0) / x - 1.0 / self.scale)

    def sample_x(self):
        return float(scipy.stats.gamma.rvs(self.shape, scale=self.scale))


class LogNormalPrior(InvGammaPrior):
    """Log-normal density (mogp ``LogNormalPrior(shape, scale)``: shape = sigma of ln x, scale = exp(mean of ln x))."""

    def logp(self, x):
        return float(-0.5 * (np.log(x / self.scale) / self.shape) ** 2 - 0.5 * np.log(2.0 * np.pi) - np.log(x)
                     - np.log(self.shape))

    def dlogpdx(self, x):
        return float(-np.log(x / self.scale) / self.shape**2 / x - 1.0 / x)

    def sample_x(self):
        return float(scipy.stats.lognorm.rvs(self.shape, scale=self.scale))


class GPPriors:
    """Container mirroring mogp ``GPPriors(corr=..., cov=..., nugget=...)`` for the zero-mean case: one prior per
    correlation length, one on the process variance, one on the nugget (used only when the nugget is fitted).
    ``None`` entries are flat priors."""

    def __init__(self, corr=None, cov=None, nugget=None, n_corr=None, nugget_type="fit", mean=None):
        if mean is not None:
            raise ValueError("mean-function priors are not supported: the designers assume a zero mean")
        if corr is None:
            if n_corr is None:
                raise ValueError("n_corr is needed when no correlation-length priors are given")
            corr = [None] * int(n_corr)
        self.corr = [WeakPrior() if c is None else c for c in corr]
        self.cov = WeakPrior() if cov is None else cov
        self.nugget = WeakPrior() if nugget is None else nugget
        self.nugget_type = nugget_type

    @property
    def n_corr(self):
        return len(self.corr)


def _is_flat(dist) -> bool:
    return dist is None or not hasattr(dist, "sample_x")


class CustomMapPriors:
    """User-supplied priors behind the interface the MAP driver uses (``logp``, ``dlogp_draw``, ``sample``): any
    object with ``corr`` (sequence), ``cov`` and ``nugget`` attributes whose entries offer ``logp(x)`` /
    ``dlogpdx(x)`` / ``sample_x()`` on the linear-scale parameter -- this module's classes or mogp's own -- or a dict
    with those keys.  Densities act on the linear scale without a Jacobian term and starting points are drawn
    in parameter order through numpy's global RNG, exactly as for the default priors."""

    def __init__(self, priors, d: int, fit_nugget: bool = False):
        if isinstance(priors, dict):
            extra = set(priors) - {"corr", "cov", "nugget", "n_corr", "nugget_type", "mean"}
            if extra:
                raise ValueError(f"unknown prior specification keys: {sorted(extra)}")
            kw = dict(priors)
            kw.setdefault("n_corr", d)
            priors = GPPriors(**kw)
        if not (hasattr(priors, "corr") and hasattr(priors, "cov")):
            raise TypeError("priors must be a GPPriors-like object or a dict")
        if getattr(priors, "mean", None) is not None and getattr(getattr(priors, "mean"), "n_params", 0):
            raise ValueError("mean-function priors are not supported: the designers assume a zero mean")
        self.corr = list(priors.corr)
        if len(self.corr) != d:
            raise ValueError("bad number of correlation lengths in new GPPriors object")
        self.cov = priors.cov
        self.nugget = getattr(priors, "nugget", None)
        self.d, self.fit_nugget = d, fit_nugget

    @staticmethod
    def _lp(dist, x):
        return 0.0 if dist is None else float(dist.logp(x))

    @staticmethod
    def _dl(dist, x):
        return 0.0 if dist is None else float(dist.dlogpdx(x))

    def logp(self, theta_raw: np.ndarray) -> float:
        d = self.d
        total = 0.0
        for k in range(d):
            total += self._lp(self.corr[k], float(np.exp(-0.5 * theta_raw[k])))
        total += self._lp(self.cov, float(np.exp(theta_raw[d])))
        if self.fit_nugget:
            total += self._lp(self.nugget, float(np.exp(theta_raw[d + 1])))
        return total

    def dlogp_draw(self, theta_raw: np.ndarray) -> np.ndarray:
        d = self.d
        out = np.zeros_like(theta_raw, dtype=float)
        for k in range(d):
            x = float(np.exp(-0.5 * theta_raw[k]))
            out[k] = self._dl(self.corr[k], x) * (-0.5 * x)
        x = float(np.exp(theta_raw[d]))
        out[d] = self._dl(self.cov, x) * x
        if self.fit_nugget:
            x = float(np.exp(theta_raw[d + 1]))
            out[d + 1] = self._dl(self.nugget, x) * x
        return out

    @staticmethod
    def _draw(dist, to_raw) -> float:
        if _is_flat(dist):
            return float(5.0 * (np.random.rand() - 0.5))
        return float(to_raw(dist.sample_x()))

    def sample(self) -> np.ndarray:
        pt = [self._draw(c, lambda x: -2.0 * np.log(x)) for c in self.corr]
        pt.append(self._draw(self.cov, np.log))
        if self.fit_nugget:
            pt.append(self._draw(self.nugget, np.log))
        return np.array(pt, dtype=float)


def make_priors(priors, X: np.ndarray, fit_nugget: bool = False):
    """``priors=None``: mogp's data-driven defaults (``MapPriors``); otherwise the user's (``CustomMapPriors``)."""
    if priors is None:
        return MapPriors(X, fit_nugget=fit_nugget)
    return CustomMapPriors(priors, np.asarray(X).shape[1], fit_nugget=fit_nugget)


class MapPriors:
    """Priors on (corr lengths..., process variance[, nugget]) for one training set."""

    def __init__(self, X: np.ndarray, fit_nugget: bool = False):
        X = np.asarray(X, dtype=float)
        self.d = X.shape[1]
        self.corr = []  # per dimension: (shape, scale) or None for a flat prior
        for col in X.T:
            u = np.unique(col)
            gap = float(np.median(np.diff(u))) if u.size > 2 else 0.0
            span = float(u[-1] - u[0]) if u.size > 1 else 0.0
            self.corr.append(_solve_invgamma(gap, span) if gap > 0.0 and span > 0.0 else None)
        self.nugget = _solve_invgamma(1.0e-8, 1.0e-6) if fit_nugget else None
        self.fit_nugget = fit_nugget
        # vectorised form of the per-dimension densities (same operations in the same order as _logp /
        # _dlogp_dx, so the values are bit-identical); flat priors contribute exact zeros
        self._has = np.array([prm is not None for prm in self.corr])
        self._a = np.array([prm[0] if prm is not None else 0.0 for prm in self.corr])
        self._b = np.array([prm[1] if prm is not None else 0.0 for prm in self.corr])
        self._c0 = np.array([float(prm[0] * np.log(prm[1]) - gammaln(prm[0])) if prm is not None else 0.0
                             for prm in self.corr])

    # -- log density and its gradient w.r.t. the raw parameters ---------------------------
    @staticmethod
    def _logp(x: float, prm) -> float:
        if prm is None:
            return 0.0
        a, b = prm
        return float(a * np.log(b) - gammaln(a) - (a + 1.0) * np.log(x) - b / x)

    @staticmethod
    def _dlogp_dx(x: float, prm) -> float:
        if prm is None:
            return 0.0
        a, b = prm
        return float(-(a + 1.0) / x + b / x**2)

    def logp(self, theta_raw: np.ndarray) -> float:
        lengths = np.exp(-0.5 * theta_raw[: self.d])
        with np.errstate(all="ignore"):
            terms = np.where(self._has, self._c0 - (self._a + 1.0) * np.log(lengths) - self._b / lengths, 0.0)
        total = sum(terms.tolist())          # sequential, as the reference's sum over its prior objects
        if self.fit_nugget:
            total += self._logp(float(np.exp(theta_raw[self.d + 1])), self.nugget)
        return total

    def dlogp_draw(self, theta_raw: np.ndarray) -> np.ndarray:
        out = np.zeros_like(theta_raw, dtype=float)
        lengths = np.exp(-0.5 * theta_raw[: self.d])
        with np.errstate(all="ignore"):
            # squares through the scalar pow() the per-dimension formula x**2 uses (numpy's array power differs from it
            # in the last bit about once in 10^4 values, enough to alter an L-BFGS-B trajectory)
            sq = np.array([x**2 for x in lengths])
            ddx = -(self._a + 1.0) / lengths + self._b / sq
            out[: self.d] = np.where(self._has, ddx * (-0.5 * lengths), 0.0)   # d length / d raw = -length / 2
        if self.fit_nugget:
            nu = float(np.exp(theta_raw[self.d + 1]))
            out[self.d + 1] = self._dlogp_dx(nu, self.nugget) * nu
        return out

    # -- starting points ------------------------------------------------------------------
    @staticmethod
    def _draw(prm, to_raw) -> float:
        if prm is None:
            return float(5.0 * (np.random.rand() - 0.5))
        a, b = prm
        return float(to_raw(scipy.stats.invgamma.rvs(a, scale=b)))

    def sample(self) -> np.ndarray:
        pt = [self._draw(prm, lambda x: -2.0 * np.log(x)) for prm in self.corr]
        pt.append(self._draw(None, None))                     # process variance: flat
        if self.fit_nugget:
            pt.append(self._draw(self.nugget, np.log))
        return np.array(pt, dtype=float)
