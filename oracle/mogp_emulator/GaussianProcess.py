"""Single-output zero-mean Gaussian process of mogp-emulator 0.7.2 (restated).

Call sites in the reference: exauq/core/emulators.py:183 (empty ctor), :414,:447 (ctor
with data), :448 (``fit(GPParams)``), :649 (``predict``), :299-309 (``theta``, ``n_params``);
exauq/utilities/mogp_fitting.py:316-324 (``priors.sample``, ``logposterior``,
``logpost_deriv``), :344 (``fit(ndarray)``).

Arithmetic:
  K = cov * kernel_f(X, X, corr_raw) (+ nugget I when fixed/fit);  L = chol(K) (adaptive:
  retry with jitter mean(diag K) 1e-6 * 10^k, k < 5);  alpha = K^-1 y;
  neg-log-posterior = 0.5 (log|K| + y^T alpha + n ln 2 pi) - log prior;
  predict: mean = k^T alpha;  var = max(cov + nugget - k^T K^-1 k, 0).
"""

import copy

import numpy as np
from scipy import linalg
from scipy.linalg import lapack

from . import Kernel
from .GPParams import GPParams
from .Priors import GPPriors


class PredictResult(dict):
    "dict with attribute access; also indexable as (mean, unc, deriv)"

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __getitem__(self, key):
        if isinstance(key, int):
            return (dict.__getitem__(self, "mean"), dict.__getitem__(self, "unc"),
                    dict.__getitem__(self, "deriv"))[key]
        return dict.__getitem__(self, key)


def jit_cholesky(A, maxtries=5):
    "Cholesky with adaptive jitter; returns (L, jitter_used)"
    A = np.ascontiguousarray(A)
    L, info = lapack.dpotrf(A, lower=1)
    if info == 0:
        return L, 0.0
    diagA = np.diag(A)
    if np.any(diagA <= 0.0):
        raise linalg.LinAlgError("not pd: non-positive diagonal elements")
    jitter = diagA.mean() * 1e-6
    num_tries = 1
    while num_tries <= maxtries and np.isfinite(jitter):
        try:
            L = linalg.cholesky(A + np.eye(A.shape[0]) * jitter, lower=True)
            return L, jitter
        except Exception:
            jitter *= 10
        finally:
            num_tries += 1
    raise linalg.LinAlgError("not positive definite, even with jitter.")


def pivot_cholesky(A):
    """Pivoted Cholesky of a possibly singular covariance matrix (mogp-emulator 0.7.2, linalg/cholesky.py,
    reached by ``GaussianProcess.fit`` when ``nugget == "pivot"``; option passed through at
    exauq/core/emulators.py:129-137).  LAPACK ``dpstrf`` (lower, default tolerance n * eps * max diag) interchanges
    rows / columns so that the diagonal of the factor decreases and stops at the numerical rank; the diagonal
    entries of the skipped rows are then replaced by a decreasing sequence
    ``L[rank-1, rank-1] / ((rank+1) (rank+2) ... (i+1))`` so that ``2 sum log diag L`` stays defined.  Returns
    (L, P) with P the 0-based pivot order: ``A[P][:, P] = L L^T`` on the leading ``rank`` block.
    NOTE: below the rank, what ``dpstrf`` leaves in the trailing block depends on LAPACK's block size, so values for
    numerically rank-deficient matrices are not reproducible between LAPACK builds (parity there is unpinned)."""
    A = np.ascontiguousarray(A)
    assert A.ndim == 2 and A.shape[0] == A.shape[1], "A must be a square matrix"
    L, P, rank, info = lapack.dpstrf(A, lower=1)
    L = np.tril(L)
    if info < 0:
        raise linalg.LinAlgError("Illegal value in covariance matrix")
    n = A.shape[0]
    idx = np.arange(rank, n)
    divs = np.cumprod(np.arange(rank + 1, n + 1, dtype=np.float64))
    L[idx, idx] = L[rank - 1, rank - 1] / divs
    return L, P - 1


def _pivot_transpose(P):
    "inverse of the permutation P"
    out = np.empty(len(P), dtype=np.int32)
    out[np.asarray(P)] = np.arange(len(P), dtype=np.int32)
    return out


class GaussianProcessBase(object):
    pass


class GaussianProcess(GaussianProcessBase):
    def __init__(self, inputs, targets, mean=None, kernel="SquaredExponential", priors=None,
                 nugget="adaptive", inputdict={}, use_patsy=True):
        inputs = np.array(inputs)
        if inputs.ndim == 1:
            inputs = np.reshape(inputs, (-1, 1))
        assert inputs.ndim == 2, "inputs must be 2D"
        targets = np.array(targets)
        assert targets.ndim == 1, "targets must be 1D"
        assert targets.shape[0] == inputs.shape[0], "inputs/targets length mismatch"
        if mean is not None:
            raise NotImplementedError("only the zero mean function is restated in the oracle")

        self._inputs = np.array(inputs, dtype=float) if inputs.size else inputs.astype(float)
        self._targets = np.array(targets, dtype=float) if targets.size else targets.astype(float)

        if isinstance(kernel, str):
            if kernel not in ("SquaredExponential", "Matern52", "ProductMat52"):
                raise ValueError("provided kernel '{}' not a supported kernel type".format(kernel))
            kernel = getattr(Kernel, kernel)()
        if not isinstance(kernel, Kernel.KernelBase):
            raise ValueError("provided kernel is not a subclass of KernelBase")
        self.kernel = kernel

        self._nugget_type = None
        self._nugget_value = None
        self._set_nugget(nugget)

        self._theta = GPParams(n_mean=0, n_corr=self.n_corr, nugget=self._nugget_arg())
        self.priors = priors

        self.K = None
        self.L = None
        self.P = None
        self.Kinv_t = None
        self.current_logpost = None

    # -- construction helpers -------------------------------------------------
    def _set_nugget(self, nugget):
        if isinstance(nugget, str):
            if nugget not in ("adaptive", "fit", "pivot"):
                raise ValueError("nugget must be 'adaptive', 'pivot', 'fit' or a float")
            self._nugget_type = nugget
        else:
            try:
                value = float(nugget)
            except TypeError:
                raise TypeError("nugget parameter must be a string or a non-negative float")
            if value < 0.0:
                raise ValueError("nugget parameter must be non-negative")
            self._nugget_type = "fixed"
            self._nugget_value = value

    def _nugget_arg(self):
        return self._nugget_value if self._nugget_type == "fixed" else self._nugget_type

    # -- read-only views ----------------------------------------------------
    @property
    def inputs(self):
        return self._inputs

    @property
    def targets(self):
        return self._targets

    @property
    def n(self):
        return self._inputs.shape[0]

    @property
    def D(self):
        return self._inputs.shape[1]

    @property
    def n_corr(self):
        return self._inputs.shape[1]

    @property
    def n_params(self):
        return self._theta.n_params

    @property
    def nugget_type(self):
        return self._nugget_type

    @property
    def nugget(self):
        return self._theta.nugget

    @property
    def priors(self):
        return self._priors

    @priors.setter
    def priors(self, newpriors):
        if newpriors is None:
            newpriors = GPPriors.default_priors(self._inputs, self.n_corr, self._nugget_type)
        if isinstance(newpriors, dict):
            kw = dict(newpriors)
            kw.setdefault("n_corr", self.n_corr)
            kw["nugget_type"] = self._nugget_type
            newpriors = GPPriors(**kw)
        if not isinstance(newpriors, GPPriors):
            raise TypeError("priors must be a GPPriors object in the oracle")
        if self.n > 0 and newpriors.n_corr != self.n_corr:      # (an empty GP, as MogpEmulator first builds, has no shape yet)
            raise ValueError("bad number of correlation lengths in new GPPriors object")
        self._priors = newpriors

    @property
    def theta(self):
        return self._theta

    @theta.setter
    def theta(self, theta):
        if theta is None:
            self._theta = GPParams(n_mean=0, n_corr=self.n_corr, nugget=self._nugget_arg())
            self.K = self.L = self.Kinv_t = self.current_logpost = None
            return
        if isinstance(theta, GPParams):
            assert self._theta.same_shape(theta), "bad shape for hyperparameters"
            self._theta = copy.deepcopy(theta)
        else:
            self._theta.set_data(theta)
        self._refit()

    # -- arithmetic ---------------------------------------------------------
    def get_K_matrix(self):
        "cov * correlation(train, train), no nugget"
        return self._theta.cov * self.kernel.kernel_f(self._inputs, self._inputs,
                                                      self._theta.corr_raw)

    def get_cov_matrix(self, other_inputs):
        "cov * correlation(train, other) of shape (n, n_other)"
        return self._theta.cov * self.kernel.kernel_f(self._inputs, other_inputs,
                                                      self._theta.corr_raw)

    def _refit(self):
        self.K = self.get_K_matrix()
        if self._nugget_type == "adaptive":
            self.L, jitter = jit_cholesky(self.K)
            self._theta.nugget = jitter
        elif self._nugget_type == "pivot":
            self.L, self.P = pivot_cholesky(self.K)
        else:
            Kn = self.K + np.eye(self.n) * self._theta.nugget
            L, info = lapack.dpotrf(np.ascontiguousarray(Kn), lower=1)
            if info != 0:
                raise linalg.LinAlgError("covariance matrix is not positive definite")
            self.L = L
        self.Kinv_t = self._kinv_solve(self._targets)
        self.current_logpost = 0.5 * (
            2.0 * np.sum(np.log(np.diag(self.L)))
            + np.dot(self._targets, self.Kinv_t)
            + self.n * np.log(2.0 * np.pi)
        )
        self.current_logpost -= self._priors.logp(self._theta)

    def _kinv_solve(self, b):
        "K^-1 b through the factor of the last fit (ChoInv / ChoInvPivot.solve in mogp)"
        if self._nugget_type == "pivot":
            return linalg.cho_solve((self.L, True), b[self.P])[_pivot_transpose(self.P)]
        return linalg.cho_solve((self.L, True), b)

    def fit(self, theta):
        self.theta = theta

    def _same_theta(self, theta):
        data = self._theta.get_data()
        return data is not None and np.allclose(theta, data, rtol=1.0e-10, atol=1.0e-10)

    def logposterior(self, theta):
        if not self._same_theta(theta):
            self.fit(theta)
        return self.current_logpost

    def logpost_deriv(self, theta):
        theta = np.array(theta)
        assert theta.shape == (self.n_params,), "bad shape for new parameters"
        if not self._same_theta(theta):
            self.fit(theta)
        partials = np.zeros(self.n_params)
        dKdtheta = self._theta.cov * self.kernel.kernel_deriv(
            self._inputs, self._inputs, self._theta.corr_raw)
        alpha = self.Kinv_t
        for d in range(self.D):
            trace = np.trace(self._kinv_solve(dKdtheta[d]))
            partials[d] = -0.5 * (np.dot(alpha, np.dot(dKdtheta[d], alpha)) - trace)
        trace = np.trace(self._kinv_solve(self.K))
        partials[self.D] = -0.5 * (np.dot(alpha, np.dot(self.K, alpha)) - trace)
        if self._nugget_type == "fit":
            nugget = self._theta.nugget
            Kinv = self._kinv_solve(np.eye(self.n))
            partials[-1] = 0.5 * nugget * (np.trace(Kinv) - np.dot(alpha, alpha))
        partials -= self._priors.dlogpdtheta(self._theta)
        return partials

    def predict(self, testing, unc=True, deriv=False, include_nugget=True):
        if self._theta.get_data() is None:
            raise ValueError("hyperparameters have not been fit for this Gaussian Process")
        testing = np.array(testing, dtype=float)
        if testing.ndim == 1:
            testing = np.reshape(testing, (1, len(testing)))
        assert testing.ndim == 2, "testing must be a 2D array"
        assert testing.shape[1] == self.D, "second dimension of testing must be the same as D"

        Ktest = self.get_cov_matrix(testing)
        mu = np.dot(Ktest.T, self.Kinv_t)
        var = None
        if unc:
            sigma_2 = self._theta.cov
            if include_nugget and self._nugget_type != "pivot":
                sigma_2 = sigma_2 + self._theta.nugget
            var = np.maximum(
                sigma_2 - np.sum(Ktest * self._kinv_solve(Ktest), axis=0), 0.0)
        return PredictResult(mean=mu, unc=var, deriv=None)

    def __call__(self, testing):
        return self.predict(testing, unc=False)[0]

    def __str__(self):
        return "Gaussian Process with {} training examples and {} input variables".format(
            self.n, self.D)
