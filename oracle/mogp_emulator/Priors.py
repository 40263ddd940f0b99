"""Hyperparameter priors of mogp-emulator 0.7.2 (restated): the defaults reached when ``priors=None`` and the
user-supplied case (``priors=`` is passed through by ``MogpEmulator``, exauq/core/emulators.py:129-137):
``GPPriors`` of ``InvGammaPrior`` / ``GammaPrior`` / ``LogNormalPrior`` / ``WeakPrior`` entries or a dict of its
keyword arguments.  The densities are the textbook ones in mogp's (shape, scale) parametrisation; without the
wheel their agreement with mogp is unpinned beyond that.

Used by the reference at exauq/utilities/mogp_fitting.py:316 (``gp.priors.sample()``) and
inside ``logposterior``/``logpost_deriv`` (:318-324).

Rules: per input dimension an inverse-gamma prior on the
correlation length whose cdf is 0.005 at the median spacing of the sorted unique
coordinates and 0.995 at their range (solved with ``scipy.optimize.root`` from log-params
0); flat ("weak") prior on the process variance; inverse-gamma prior on the nugget only
when it is fitted.  Priors act on the linear-scale parameter with no Jacobian term.
"""

import numpy as np
import scipy.stats
from scipy.optimize import root
from scipy.special import gammaln


class CorrTransform(object):
    "scaled = exp(-raw/2)"

    @staticmethod
    def inv_transform(x):
        return -2.0 * np.log(x)

    @staticmethod
    def transform(x):
        return np.exp(-0.5 * x)

    @staticmethod
    def dscaled_draw(x):
        return -0.5 * np.exp(-0.5 * x)


class CovTransform(object):
    "scaled = exp(raw)"

    @staticmethod
    def inv_transform(x):
        return np.log(x)

    @staticmethod
    def transform(x):
        return np.exp(x)

    @staticmethod
    def dscaled_draw(x):
        return np.exp(x)


class WeakPrior(object):
    "Improper flat prior"

    def logp(self, x):
        return 0.0

    def dlogpdx(self, x):
        return 0.0

    def dlogpdtheta(self, x, transform):
        return 0.0

    def sample(self, transform=None):
        return float(5.0 * (np.random.rand() - 0.5))


def min_spacing(values):
    "median spacing between sorted unique values (0 if too few)"
    values = np.unique(np.array(values))
    if len(values) <= 2:
        return 0.0
    return np.median(np.diff(np.sort(values)))


def max_spacing(values):
    "range of the unique values (0 if too few)"
    values = np.unique(np.array(values))
    if len(values) <= 1:
        return 0.0
    return float(np.max(values) - np.min(values))


class PriorDist(WeakPrior):
    "A proper prior on the linear-scale parameter"

    def dlogpdtheta(self, x, transform):
        return float(self.dlogpdx(transform.transform(x)) * transform.dscaled_draw(x))

    def sample(self, transform=None):
        return float(transform.inv_transform(self.sample_x()))

    @classmethod
    def default_prior(cls, min_val, max_val):
        assert min_val > 0.0 and max_val > 0.0 and min_val < max_val

        def f(x):
            dist = cls(np.exp(x[0]), np.exp(x[1]))
            return np.array([dist.cdf(min_val) - 0.005, dist.cdf(max_val) - 0.995])

        with np.errstate(all="ignore"):
            result = root(f, np.zeros(2))
        if not result["success"]:
            print("Prior solver failed to converge")
            return WeakPrior()
        return cls(np.exp(result["x"][0]), np.exp(result["x"][1]))

    @classmethod
    def default_prior_corr(cls, inputs):
        min_val = min_spacing(inputs)
        max_val = max_spacing(inputs)
        if min_val == 0.0 or max_val == 0.0:
            print("Too few unique inputs; defaulting to flat priors")
            return WeakPrior()
        return cls.default_prior(min_val, max_val)

    @classmethod
    def default_prior_nugget(cls, min_val=1.0e-8, max_val=1.0e-6):
        return cls.default_prior(min_val, max_val)


class InvGammaPrior(PriorDist):
    def __init__(self, shape, scale):
        assert shape > 0.0 and scale > 0.0
        self.shape = float(shape)
        self.scale = float(scale)

    def logp(self, x):
        return float(
            self.shape * np.log(self.scale)
            - gammaln(self.shape)
            - (self.shape + 1.0) * np.log(x)
            - self.scale / x
        )

    def dlogpdx(self, x):
        return float(-(self.shape + 1.0) / x + self.scale / x**2)

    def cdf(self, x):
        return scipy.stats.invgamma.cdf(x, self.shape, scale=self.scale)

    def sample_x(self):
        return float(scipy.stats.invgamma.rvs(self.shape, scale=self.scale))


class GammaPrior(InvGammaPrior):
    def logp(self, x):
        return float(
            -self.shape * np.log(self.scale) - gammaln(self.shape) + (self.shape - 1.0) * np.log(x) - x / self.scale
        )

    def dlogpdx(self, x):
        return float((self.shape - 1.0) / x - 1.0 / self.scale)

    def cdf(self, x):
        return scipy.stats.gamma.cdf(x, self.shape, scale=self.scale)

    def sample_x(self):
        return float(scipy.stats.gamma.rvs(self.shape, scale=self.scale))


class LogNormalPrior(InvGammaPrior):
    def logp(self, x):
        return float(
            -0.5 * (np.log(x / self.scale) / self.shape) ** 2 - 0.5 * np.log(2.0 * np.pi) - np.log(x) - np.log(self.shape)
        )

    def dlogpdx(self, x):
        return float(-np.log(x / self.scale) / self.shape**2 / x - 1.0 / x)

    def cdf(self, x):
        return scipy.stats.lognorm.cdf(x, self.shape, scale=self.scale)

    def sample_x(self):
        return float(scipy.stats.lognorm.rvs(self.shape, scale=self.scale))


class GPPriors(object):
    def __init__(self, mean=None, corr=None, cov=None, nugget=None, n_corr=None, nugget_type="fit"):
        assert mean is None, "only the zero mean function is restated in the oracle"
        if corr is None:
            assert n_corr is not None
            corr = [WeakPrior() for _ in range(n_corr)]
        corr = [WeakPrior() if c is None else c for c in corr]
        self.corr = list(corr)
        self.cov = cov if cov is not None else WeakPrior()
        self.nugget_type = nugget_type
        if nugget_type == "fit":
            self.nugget = nugget if nugget is not None else WeakPrior()
        else:
            self.nugget = None

    @property
    def n_corr(self):
        return len(self.corr)

    @classmethod
    def default_priors(cls, inputs, n_corr, nugget_type="fit"):
        inputs = np.array(inputs)
        assert inputs.ndim == 2 and inputs.shape[1] == n_corr
        corr = [InvGammaPrior.default_prior_corr(col) for col in np.transpose(inputs)]
        nugget = InvGammaPrior.default_prior_nugget() if nugget_type == "fit" else None
        return cls(corr=corr, cov=WeakPrior(), nugget=nugget, nugget_type=nugget_type)

    def logp(self, theta):
        lp = 0.0
        for dist, val in zip(self.corr, theta.corr):
            lp += dist.logp(val)
        lp += self.cov.logp(theta.cov)
        if self.nugget_type == "fit":
            lp += self.nugget.logp(theta.nugget)
        return lp

    def dlogpdtheta(self, theta):
        raw = theta.get_data()
        partials = [
            dist.dlogpdtheta(r, CorrTransform) for dist, r in zip(self.corr, theta.corr_raw)
        ]
        partials.append(self.cov.dlogpdtheta(raw[theta.n_corr], CovTransform))
        if self.nugget_type == "fit":
            partials.append(self.nugget.dlogpdtheta(raw[-1], CovTransform))
        return np.array(partials, dtype=float)

    def sample(self):
        pt = [dist.sample(CorrTransform) for dist in self.corr]
        pt.append(self.cov.sample(CovTransform))
        if self.nugget_type == "fit":
            pt.append(self.nugget.sample(CovTransform))
        return np.array(pt, dtype=float)
