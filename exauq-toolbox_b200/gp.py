"""``B200GaussianProcess``: drop-in ``AbstractGaussianProcess`` backed by libexauq_b200.so.

Mirrors ``exauq.core.emulators.MogpEmulator`` (exauq/core/emulators.py:69-667) -- same
constructor keywords, properties, argument checks and error types -- so that the
reference's own ``exauq/core/designers.py`` (``compute_loo_gp``, ``compute_loo_errors_gp``,
``PEICalculator``, ``compute_single_level_loo_samples``, ``compute_multi_level_loo_samples``)
and ``MultiLevelGaussianProcess`` use it unchanged.  All arithmetic runs on the GPU through
the C ABI (``_lib``); there is no CPU path.

What differs from the reference, on purpose:

* ``kinv`` is computed lazily on first access (the reference recomputes an SVD condition
  number and ``np.linalg.inv`` inside every ``fit``, exauq/core/modelling.py:1082-1113);
  a singular covariance therefore raises ``ValueError`` at first ``kinv`` access.
* Fits with supplied hyperparameters on "parent data minus one point" -- what
  ``compute_loo_gp`` does N times (exauq/core/designers.py:263-272, 365-369) -- are served
  from ONE leave-one-out sweep of the parent on the GPU (``exq_gp_loo``); anything other
  than predicting at the left-out input triggers a real refit.
* Batched entry points the reference does not have: ``predict_batch``, ``loo_predictions``,
  ``pei_candidates``, ``logpost_batch``.
"""

from __future__ import annotations

import copy
import itertools
import math
import threading
import time
from collections.abc import Collection, Sequence
from numbers import Real
from typing import Any, Optional
from warnings import warn

import numpy as np
from scipy.linalg import LinAlgError
from scipy.optimize import minimize

from . import _lib
from ._lib import DeviceCandidates, DeviceGP, PivotedDeviceGP
from .priors import MapPriors, make_priors
from .refapi import (
    AbstractGaussianProcess,
    GaussianProcessHyperparameters,
    GaussianProcessPrediction,
    Input,
    TrainingDatum,
)

SUPPORTED_KERNELS = ("Matern52", "SquaredExponential", "ProductMat52")
TOL = 1e-9  # exauq.core.numerics.FLOAT_TOLERANCE


def inputs_to_array(inputs: Sequence[Input], dim: Optional[int] = None) -> np.ndarray:
    """(n, d) float64 array of the coordinates of a sequence of ``Input``."""
    rows = [np.atleast_1d(np.asarray(x.value, dtype=float)) if len(x) else np.empty(0) for x in inputs]
    if dim is not None:
        for r in rows:
            if r.shape[0] != dim:
                raise ValueError(
                    f"Expected inputs to have dimension equal to {dim}, but "
                    f"received input of dimension {r.shape[0]}."
                )
    return np.array(rows, dtype=float).reshape(len(rows), -1)


def find_duplicate_rows(X: np.ndarray) -> Optional[tuple[int, int]]:
    """First pair of rows equal within tolerance (rel 1e-9 or abs 1e-9 per coordinate, the rule
    of exauq/core/numerics.py:31-73), found by a sort on the first coordinate instead of the
    reference's O(N^2) Python loop (exauq/core/emulators.py:316-327)."""
    n = X.shape[0]
    if n < 2:
        return None
    order = np.argsort(X[:, 0], kind="stable")
    x0 = X[order, 0]
    slack = TOL * np.maximum(1.0, np.abs(x0)) * (1.0 + 1e-6)
    hi = np.searchsorted(x0, x0 + 2.0 * slack, side="right")
    for i in range(n):
        j1 = hi[i]
        if j1 <= i + 1:
            continue
        a = X[order[i]]
        b = X[order[i + 1 : j1]]
        close = np.abs(b - a) <= np.maximum(TOL * np.maximum(np.abs(a), np.abs(b)), TOL)
        hit = np.nonzero(np.all(close, axis=1))[0]
        if hit.size:
            p, q = int(order[i]), int(order[i + 1 + hit[0]])
            return (min(p, q), max(p, q))
    return None


class _FitState:
    """Immutable result of one fit: training data, hyperparameters and the device handle.
    Shared (not copied) between deep copies of a ``B200GaussianProcess``."""

    def __init__(self, training_data, X, y, kernel, hyperparameters, corr_raw, dev: DeviceGP):
        self.training_data = training_data
        self.X, self.y = X, y
        self.N, self.d = X.shape
        self.kernel = kernel
        self.hp = hyperparameters
        self.corr_raw = corr_raw
        self.dev = dev
        self.ids = np.fromiter(map(id, training_data), dtype=np.int64, count=len(training_data))
        self._loo = None
        self._kinv = None
        self._kinv_y = None
        self._lock = threading.Lock()

    def loo(self):
        """(mean, var, nes) of the leave-one-out sweep; one GPU call, cached."""
        with self._lock:
            if self._loo is None:
                self._loo = self.dev.loo()
            return self._loo

    def kinv(self) -> np.ndarray:
        with self._lock:
            if self._kinv is None:
                try:
                    self._kinv = self.dev.kinv()
                except LinAlgError:
                    raise ValueError(
                        "Cannot compute covariance inverse: Covariance Matrix is, or too close to singular."
                    ) from None
            return self._kinv


    def kinv_y(self) -> np.ndarray:
        """``kinv @ y`` as a column, formed once: the weights of the zero-mean prediction
        ``covariance_matrix([x]) @ kinv @ y`` (exauq/core/designers.py:1029) that the multi-level designers evaluate for
        every left-out input of every other level."""
        kinv = self.kinv()
        with self._lock:
            if self._kinv_y is None:
                self._kinv_y = kinv @ self.y.reshape(-1, 1)
            return self._kinv_y


class _LooView:
    """State of a GP 'fitted' to a parent's data minus one point with the parent's
    hyperparameters, served from the parent's LOO sweep until something else is asked."""

    def __init__(self, parent: _FitState, idx: int, training_data):
        self.parent, self.idx, self.training_data = parent, idx, training_data
        self.real: Optional[_FitState] = None


class B200GaussianProcess(AbstractGaussianProcess):
    """Gaussian process emulator computed on an NVIDIA B200.

    Parameters
    ----------
    kernel : {"SquaredExponential", "Matern52", "ProductMat52"}
        Covariance function (default "SquaredExponential", as ``MogpEmulator``).
    nugget : {"adaptive", "fit", "pivot"} or float
        Nugget handling (default "adaptive": 0 unless the Cholesky factorisation fails).
        "pivot": no nugget, pivoted Cholesky factorisation (mogp ``pivot_cholesky`` / LAPACK ``dpstrf``) that
        tolerates a numerically rank-deficient covariance matrix.
    n_tries : int
        Number of L-BFGS-B restarts in hyperparameter estimation (mogp default 15).
    **kwargs :
        ``inputs`` / ``targets`` are ignored as in ``MogpEmulator``; ``priors`` may be ``None`` (mogp's data-driven
        defaults), a ``GPPriors``-like object (``exauq_toolbox_b200.priors.GPPriors`` or mogp's own) or a dict with
        ``corr`` / ``cov`` / ``nugget`` entries; ``mean=None``, ``inputdict``, ``use_patsy`` are accepted; anything
        else -- including a non-zero mean function, which the designers do not support
        (docs/designers/tutorials/training_gp_tutorial.md:146-148) -- raises ``RuntimeError``
        (exauq/core/emulators.py:175-190).
    """

    def __init__(self, **kwargs):
        kwargs = {k: v for k, v in kwargs.items() if k not in ("inputs", "targets")}
        kernel = kwargs.pop("kernel", "SquaredExponential")
        if kernel not in SUPPORTED_KERNELS:
            raise ValueError(
                f"Could not initialise {self.__class__.__name__} with kernel = {kernel}: not a "
                "supported kernel function."
            )
        nugget = kwargs.pop("nugget", "adaptive")
        self._n_tries = int(kwargs.pop("n_tries", 15))
        mean = kwargs.pop("mean", None)
        priors = kwargs.pop("priors", None)
        kwargs.pop("inputdict", None)
        kwargs.pop("use_patsy", None)
        bad_nugget = (isinstance(nugget, str) and nugget not in ("adaptive", "fit", "pivot")) or (
            not isinstance(nugget, str) and (not isinstance(nugget, Real) or nugget < 0)
        )
        if kwargs or mean is not None or bad_nugget:
            raise RuntimeError(
                f"Could not construct the device Gaussian process during initialisation of "
                f"{self.__class__.__name__}"
            )
        if priors is not None:
            try:                     # shape is checked against the data at fit time
                ok = isinstance(priors, dict) or (hasattr(priors, "corr") and hasattr(priors, "cov"))
            except Exception:
                ok = False
            if not ok:
                raise RuntimeError(
                    f"Could not construct the device Gaussian process during initialisation of "
                    f"{self.__class__.__name__}"
                )
        self._priors = priors
        self._kernel_name = kernel
        self._nugget = nugget
        self._state: Optional[_FitState | _LooView] = None
        self._loo_base: Optional[_FitState] = None

    # ------------------------------------------------------------------ properties
    @property
    def kernel(self) -> str:
        return self._kernel_name

    @property
    def training_data(self) -> tuple[TrainingDatum]:
        return self._state.training_data if self._state is not None else tuple()

    @property
    def fit_hyperparameters(self) -> Optional[GaussianProcessHyperparameters]:
        if self._state is None:
            return None
        return self._state.parent.hp if isinstance(self._state, _LooView) else self._state.hp

    @property
    def kinv(self) -> np.ndarray:
        if self._state is None:
            return np.array([])
        return self._real_state().kinv()

    def __deepcopy__(self, memo):
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k in ("_state", "_loo_base"):
                setattr(new, k, v)  # immutable fit results (incl. device buffers) are shared
            else:
                setattr(new, k, copy.deepcopy(v, memo))
        return new

    # ------------------------------------------------------------------ fitting
    def fit(
        self,
        training_data: Collection[TrainingDatum],
        hyperparameters: Optional[GaussianProcessHyperparameters] = None,
        hyperparameter_bounds: Optional[Sequence] = None,
    ) -> None:
        """Fit to data; signature, checks and errors as ``MogpEmulator.fit``
        (exauq/core/emulators.py:220-313)."""
        training_data, loo_idx = self._parse_for_fit(training_data)
        if not training_data:
            return None

        if not (hyperparameters is None or isinstance(hyperparameters, GaussianProcessHyperparameters)):
            raise TypeError(
                "Expected 'hyperparameters' to be None or of type "
                f"{GaussianProcessHyperparameters.__name__}, but received {type(hyperparameters)} instead."
            )

        view = self._match_loo(training_data, hyperparameters, loo_idx)
        if view is not None:
            self._validate_hyperparameter_bounds(hyperparameter_bounds)
            self._state = view
            return None

        X = inputs_to_array([datum.input for datum in training_data])
        y = np.array([datum.output for datum in training_data], dtype=float)
        dup = find_duplicate_rows(X)
        if dup is not None:
            a, b = training_data[dup[0]].input, training_data[dup[1]].input
            raise ValueError(
                f"Points {np.round(a, 9)} and {np.round(b, 9)}" " in 'TrainingDatum' are not unique within tolerance."
            )
        self._validate_hyperparameter_bounds(hyperparameter_bounds)

        if hyperparameters is None:
            state = self._fit_with_estimation(training_data, X, y, hyperparameter_bounds)
        elif self._nugget == "fit" and hyperparameters.nugget is None:
            raise ValueError(
                "The underlying Gaussian process was created with 'nugget'='fit', "
                "but the nugget supplied during fitting is "
                f"{hyperparameters.nugget}, when it should instead be a float."
            )
        else:
            state = self._fit_with_hyperparameters(training_data, X, y, hyperparameters)

        n_params = X.shape[1] + 1 + (1 if self._nugget == "fit" else 0)
        if len(training_data) < n_params:
            warn(
                f"Fewer training points ({len(training_data)}) than hyperparameters ({n_params}) being "
                f"estimated. Estimates may be unreliable."
            )
        self._state = state
        self._loo_base = state
        return None

    def _loo_index(self, data) -> Optional[int]:
        """Index of the element ``data`` leaves out of the LOO base's training data (by object identity), or None."""
        base = self._loo_base
        if base is None or base.N < 2 or len(data) != base.N - 1:
            return None
        ids = np.fromiter(map(id, data), dtype=np.int64, count=len(data))
        diff = np.nonzero(ids != base.ids[:-1])[0]
        idx = int(diff[0]) if diff.size else base.N - 1
        if not np.array_equal(ids[idx:], base.ids[idx + 1 :]):
            return None
        return idx

    def _parse_for_fit(self, training_data):
        """``_parse_training_data`` for ``fit``: returns (items, leave-one-out index or None).  The N refits of
        ``compute_loo_errors_gp`` (exauq/core/designers.py:263-272) pass the base's own objects minus one: recognised by
        identity first, they skip the per-element type scan -- those objects were checked when the base was fitted
        (the scan was 0.9 of the 1.1 ms such a refit cost at N = 4096)."""
        if training_data is not None and isinstance(training_data, (tuple, list)) and self._loo_base is not None:
            items = tuple(training_data)
            idx = self._loo_index(items)
            if idx is not None:
                return items, idx
        return self._parse_training_data(training_data), None

    def _match_loo(self, data, hp, idx: Optional[int] = None) -> Optional[_LooView]:
        """Recognise ``parent.training_data`` minus one element refitted with the parent's
        hyperparameters (exauq/core/designers.py:365-369)."""
        base = self._loo_base
        if base is None or hp is None or len(data) != base.N - 1 or base.N < 2:
            return None
        if not (hp is base.hp or hp == base.hp):
            return None
        if idx is None:
            idx = self._loo_index(data)
        if idx is None:
            return None
        return _LooView(base, idx, data)

    def _nugget_args(self, hp: GaussianProcessHyperparameters):
        """(value, adaptive?) following exauq/core/emulators.py:427-445; a supplied numeric
        nugget becomes this emulator's nugget setting from then on (the reference mutates
        its kwargs in place at :433)."""
        if hp.nugget is not None:
            self._nugget = float(hp.nugget)
            return float(hp.nugget), False
        if isinstance(self._nugget, Real):
            return float(self._nugget), False
        return 0.0, True  # "adaptive" ("pivot" is dispatched before this is used)

    def _fit_with_hyperparameters(self, training_data, X, y, hp) -> _FitState:
        if len(hp.corr_length_scales) != X.shape[1]:
            raise ValueError(
                f"Expected {X.shape[1]} correlation length scales but received {len(hp.corr_length_scales)}."
            )
        corr_raw = np.array([GaussianProcessHyperparameters.transform_corr(c) for c in hp.corr_length_scales])
        nugget, adaptive = self._nugget_args(hp)
        pivot = self._nugget == "pivot"          # only when no numeric nugget was supplied (emulators.py:427-445)
        dev = PivotedDeviceGP(X, y, self._kernel_name) if pivot else DeviceGP(X, y, self._kernel_name)
        dev.fit(corr_raw, float(hp.process_var), nugget, adaptive)
        fitted = GaussianProcessHyperparameters(
            corr_length_scales=np.array(hp.corr_length_scales, dtype=float),
            process_var=float(hp.process_var),
            nugget=None if pivot else float(dev.nugget_used),      # mogp leaves theta.nugget unset for "pivot"
        )
        return _FitState(training_data, X, y, self._kernel_name, fitted, corr_raw, dev)

    def _fit_with_estimation(self, training_data, X, y, bounds) -> _FitState:
        """MAP estimation: multi-restart L-BFGS-B on the negative log-posterior, as
        ``fit_GP_MAP`` / ``_fit_single_GP_MAP`` (exauq/utilities/mogp_fitting.py:51-346), with
        objective and gradient evaluated on the GPU for all restarts at once."""
        raw_bounds = self._compute_raw_param_bounds(bounds) if bounds is not None else None
        d = X.shape[1]
        fit_nugget = self._nugget == "fit"
        if fit_nugget and raw_bounds is not None:
            raw_bounds = tuple(raw_bounds) + ((None, None),)      # the nugget itself is unbounded
        pivot = self._nugget == "pivot"
        dev = PivotedDeviceGP(X, y, self._kernel_name) if pivot else DeviceGP(X, y, self._kernel_name)
        priors = make_priors(self._priors, X, fit_nugget=fit_nugget)
        nugget = 0.0 if isinstance(self._nugget, str) else float(self._nugget)
        adaptive = self._nugget == "adaptive"
        starts = [priors.sample() for _ in range(self._n_tries)]
        best = run_map_restarts(dev, priors, starts, raw_bounds, nugget, adaptive, fit_nugget=fit_nugget)
        if best is None:
            raise RuntimeError("GP fitting failed")
        theta = best
        if fit_nugget:
            dev.fit(theta[:d], float(np.exp(theta[d])), float(np.exp(theta[d + 1])), False)
        else:
            dev.fit(theta[:d], float(np.exp(theta[d])), nugget, adaptive)
        fitted = GaussianProcessHyperparameters(
            corr_length_scales=np.exp(-0.5 * theta[:d]),
            process_var=float(np.exp(theta[d])),
            nugget=None if pivot else float(dev.nugget_used),
        )
        return _FitState(training_data, X, y, self._kernel_name, fitted, np.array(theta[:d]), dev)

    # ---- argument validation, as the reference ---------------------------------------------
    @staticmethod
    def _parse_training_data(training_data: Any) -> tuple[TrainingDatum]:
        """``None`` -> empty; anything that is not a sized collection of ``TrainingDatum`` -> TypeError
        (same message as ``MogpEmulator._parse_training_data``, exauq/core/emulators.py:330-348)."""
        if training_data is None:
            return ()
        bad = TypeError(
            f"Expected 'training_data' to be of type finite collection of TrainingDatum, "
            f"but received {type(training_data)} instead."
        )
        if not hasattr(training_data, "__len__"):      # e.g. generators: no finite length
            raise bad
        items = tuple(training_data)
        if any(not isinstance(item, TrainingDatum) for item in items):
            raise bad
        return items

    def _validate_hyperparameter_bounds(self, hyperparameter_bounds) -> None:
        if hyperparameter_bounds is None:
            return None
        for i, bound in enumerate(hyperparameter_bounds):
            try:
                self._validate_bound_pair(bound)
            except (TypeError, ValueError) as e:
                raise e.__class__(f"Invalid bound {bound} at index {i} of 'hyperparameter_bounds': {e}")
        return None

    @staticmethod
    def _validate_bound_pair(bounds: Sequence):
        """A bound is a pair (lower, upper) of ``None`` or reals with lower <= upper
        (messages as exauq/core/emulators.py:368-395)."""
        if not hasattr(bounds, "__len__"):
            raise TypeError(f"Expected 'bounds' to be of type sequence, but received {type(bounds)} instead.")
        if len(bounds) != 2:
            raise ValueError(
                "Expected 'bounds' to be a sequence of length 2, but " f"got length {len(bounds)}."
            )
        lower, upper = bounds
        for value in (lower, upper):
            if value is not None and not isinstance(value, Real):
                raise TypeError(
                    f"Expected each 'bound' in {bounds} to be None or of type {Real} "
                    "but one or more elements were of an unexpected type."
                )
        if None not in (lower, upper) and upper < lower:
            raise ValueError(
                "Lower bound must be less than or equal to upper bound, but received "
                f"lower bound = {lower} and upper bound = {upper}."
            )

    @staticmethod
    def _compute_raw_param_bounds(bounds):
        """Linear-scale bounds -> raw bounds (exauq/core/emulators.py:453-485): correlation
        bounds swap because -2 ln is decreasing."""
        for _, upper in bounds:
            if upper is not None and upper <= 0:
                raise ValueError("Upper bounds must be positive numbers")
        tc = lambda v: GaussianProcessHyperparameters.transform_corr(v) if v is not None else None  # noqa: E731
        tv = lambda v: GaussianProcessHyperparameters.transform_cov(v) if v is not None else None  # noqa: E731
        raw = [(tc(b[1]), tc(b[0])) for b in bounds[:-1]] + [(tv(bounds[-1][0]), tv(bounds[-1][1]))]
        return tuple(raw)

    # ------------------------------------------------------------------ state helpers
    def _real_state(self) -> _FitState:
        """The fitted state, materialising a LOO view into a real (N-1)-point fit if needed."""
        st = self._state
        if isinstance(st, _LooView):
            if st.real is None:
                keep = np.arange(st.parent.N) != st.idx
                st.real = self._fit_with_hyperparameters(
                    st.training_data, st.parent.X[keep], st.parent.y[keep], st.parent.hp
                )
            return st.real
        return st

    def _input_dim(self) -> Optional[int]:
        st = self._state
        if st is None:
            return None
        return st.parent.d if isinstance(st, _LooView) else st.d

    # ------------------------------------------------------------------ prediction
    def predict(self, x: Input) -> GaussianProcessPrediction:
        """As ``MogpEmulator.predict`` (exauq/core/emulators.py:612-667)."""
        if not isinstance(x, Input):
            raise TypeError(f"Expected 'x' to be of type Input, but received {type(x)} instead.")
        if len(self.training_data) == 0:
            raise RuntimeError("Cannot make prediction because emulator has not been trained on any data.")
        if not len(x) == (expected_dim := self._input_dim()):
            raise ValueError(
                f"Expected 'x' to be an Input with {expected_dim} coordinates, but " f"it has {len(x)} instead."
            )
        st = self._state
        if isinstance(st, _LooView) and st.real is None:
            left_out = st.parent.training_data[st.idx].input
            if x is left_out or x == left_out:
                mean, var, _ = st.parent.loo()
                return GaussianProcessPrediction(estimate=float(mean[st.idx]), variance=float(var[st.idx]))
        mean, var = self._real_state().dev.predict(inputs_to_array([x]))
        return GaussianProcessPrediction(estimate=float(mean[0]), variance=float(var[0]))

    def predict_batch(self, X) -> tuple[np.ndarray, np.ndarray]:
        """Predictive means and variances at the rows of ``X`` (or a sequence of ``Input``)."""
        if len(self.training_data) == 0:
            raise RuntimeError("Cannot make prediction because emulator has not been trained on any data.")
        if len(X) and isinstance(X[0], Input):
            X = inputs_to_array(X, self._input_dim())
        X = np.asarray(X, dtype=float).reshape(-1, self._input_dim())
        return self._real_state().dev.predict(X)

    def loo_predictions(self) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
        """Means, variances and NES errors of the N leave-one-out refits (fixed
        hyperparameters), the quantities ``compute_loo_errors_gp`` collects
        (exauq/core/designers.py:263-272)."""
        if len(self.training_data) < 2:
            raise ValueError("Cannot compute leave one out error: fewer than 2 training points.")
        return self._real_state().loo()

    # ------------------------------------------------------------------ correlation
    def correlation(self, inputs1: Sequence[Input], inputs2: Sequence[Input]) -> np.ndarray:
        """As ``MogpEmulator.correlation`` (exauq/core/emulators.py:495-567)."""
        for seq in (inputs1, inputs2):
            if not hasattr(seq, "__len__"):
                raise TypeError(
                    "Expected 'inputs1' and 'inputs2' to be of type sequences of Input objects, "
                    f"but received {type(inputs1)} and {type(inputs2)} instead."
                )
        if min(len(inputs1), len(inputs2)) == 0:
            return np.array([])
        if any(not isinstance(x, Input) for x in itertools.chain(inputs1, inputs2)):
            raise TypeError(
                "Expected all 'inputs1' and 'inputs2' to be of type Input objects, "
                "but one or more elements were of an unexpected type."
            )
        assert self._state is not None, (
            f"Cannot calculate correlations for this instance of {self.__class__} "
            "because it hasn't yet been trained on data."
        )
        expected_dim = self._input_dim()
        wrong = [len(x) for x in itertools.chain(inputs1, inputs2) if len(x) != expected_dim]
        if wrong:
            raise ValueError(
                f"Expected inputs to have dimension equal to {expected_dim}, but "
                f"received input of dimension {wrong[0]}."
            )
        st = self._state
        corr_raw = st.parent.corr_raw if isinstance(st, _LooView) else st.corr_raw
        return _lib.corr(self._kernel_name, corr_raw, inputs_to_array(inputs1), inputs_to_array(inputs2))

    def covariance_matrix(self, inputs: Sequence[Input]) -> np.ndarray:
        """As ``MogpEmulator.covariance_matrix`` (exauq/core/emulators.py:569-610):
        ``process_var * correlation(inputs, training_inputs)`` with the same checks, computed
        against the training inputs already resident on the device (``exq_gp_cross_cov``)."""
        if not self.training_data:
            return np.array([])
        if not isinstance(inputs, Sequence) and not hasattr(inputs, "__len__"):
            raise TypeError(
                "Expected 'inputs' to be of type sequence of Input objects, but received " f"{type(inputs)} instead."
            )
        if len(inputs) == 0:
            return np.array([])
        if any(not isinstance(x, Input) for x in inputs):
            raise TypeError(
                "Expected all elements of 'inputs' to be of type Input objects, "
                "but one or more elements were of an unexpected type."
            )
        return self._real_state().dev.cross_cov(inputs_to_array(inputs, self._input_dim()))

    # ------------------------------------------------------------------ batched extras
    def logpost_batch(self, theta_raw, with_prior: bool = True):
        """Negative log-posterior (or likelihood) and gradient at R raw parameter vectors
        ``(corr_raw..., ln cov)`` for the current training data, all on the GPU."""
        st = self._real_state()
        theta_raw = np.asarray(theta_raw, dtype=float).reshape(-1, st.d + 1)
        nugget = 0.0 if isinstance(self._nugget, str) else float(self._nugget)
        val, grad, info, _ = st.dev.logpost_batch(theta_raw, nugget, isinstance(self._nugget, str))
        if with_prior:
            pri = make_priors(self._priors, st.X)
            for r in range(theta_raw.shape[0]):
                val[r] -= pri.logp(theta_raw[r])
                grad[r] -= pri.dlogp_draw(theta_raw[r])
        return val, grad, info

    def pei_candidates(self, candidates, repulsion_points, batch_size: int = 1, max_target: Optional[float] = None):
        """Pseudo-expected improvement of this GP (as ``PEICalculator.compute``,
        exauq/core/designers.py:522-706) over a fixed candidate array, and ``batch_size``
        rounds of arg max + repulsion update (the loop at :830-835).  Returns the chosen
        indices, their PEI values and the ``DeviceCandidates`` (for downloading arrays)."""
        st = self._real_state()
        cand = candidates if isinstance(candidates, DeviceCandidates) else DeviceCandidates(candidates)
        tmax = float(np.max(st.y)) if max_target is None else float(max_target)
        rep = np.asarray(repulsion_points, dtype=float).reshape(-1, st.d) if len(repulsion_points) else None
        cand.pei_eval(st.dev, rep, tmax)
        idx, val = cand.batch(st.dev, batch_size)
        return idx, val, cand


# =========================================================================================
# MAP estimation driver
# =========================================================================================
class _Rendezvous:
    """Batches objective evaluations requested concurrently by the restart threads into
    ``exq_gp_logpost_batch`` calls (hard part 5).

    A dispatcher thread owns the device and launches a batch when every active restart is
    waiting.  With ``overlap=True`` it also launches as soon as half of them are, so that the
    host-side L-BFGS-B work of one half hides behind the device time of the other; measured on
    B200 at N=4096 with 15 restarts this loses more to the smaller batches (34 rounds instead of
    26, latency-bound leaves) than the ~70 ms of host time it hides, so it is off by default.
    Each restart's objective values do not depend on which other restarts share its batch, so
    the result is independent of the batching."""

    def __init__(self, dev: DeviceGP, n_workers: int, nugget: float, adaptive: bool, fit_nugget: bool = False,
                 overlap: bool = False):
        self.dev, self.nugget, self.adaptive, self.fit_nugget = dev, nugget, adaptive, fit_nugget
        self.active = n_workers
        self.overlap = overlap
        self.cv = threading.Condition()
        self.pending: dict[int, np.ndarray] = {}
        self.results: dict[int, tuple] = {}
        self.inflight = 0
        self.n_batches = 0
        self.n_evals = 0
        self.gpu_s = 0.0
        self._stop = False
        self._thread = threading.Thread(target=self._dispatch, daemon=True)
        self._thread.start()

    def _ready_locked(self) -> bool:
        if not self.pending:
            return False
        waiting = len(self.pending)
        computing = self.active - waiting - self.inflight
        if computing <= 0:
            return True
        return self.overlap and 2 * waiting >= self.active

    def _dispatch(self):
        while True:
            with self.cv:
                while not self._stop and not self._ready_locked():
                    self.cv.wait()
                if self._stop and not self.pending:
                    return
                keys = sorted(self.pending)
                thetas = np.array([self.pending[k] for k in keys])
                self.pending.clear()
                self.inflight = len(keys)
            t0 = time.perf_counter()
            try:
                if self.fit_nugget:
                    val, grad, info, _ = self.dev.logpost_batch(thetas, self.nugget, self.adaptive, fit_nugget=True)
                else:
                    val, grad, info, _ = self.dev.logpost_batch(thetas, self.nugget, self.adaptive)
                res = {k: (val[i], grad[i], int(info[i]), None) for i, k in enumerate(keys)}
            except Exception as exc:  # propagate to every waiting worker
                res = {k: (None, None, -1, exc) for k in keys}
            with self.cv:
                self.gpu_s += time.perf_counter() - t0
                self.n_batches += 1
                self.n_evals += len(keys)
                self.results.update(res)
                self.inflight = 0
                self.cv.notify_all()

    def evaluate(self, wid: int, theta: np.ndarray):
        with self.cv:
            self.pending[wid] = np.array(theta, dtype=float)
            self.cv.notify_all()
            while wid not in self.results:
                self.cv.wait()
            return self.results.pop(wid)

    def retire(self):
        with self.cv:
            self.active -= 1
            self.cv.notify_all()

    def close(self):
        with self.cv:
            self._stop = True
            self.cv.notify_all()
        self._thread.join()


class _LbfgsbStepper:
    """scipy's L-BFGS-B (``scipy.optimize._lbfgsb_py._minimize_lbfgsb`` with its default options, which
    is what ``minimize(..., method="L-BFGS-B")`` at exauq/utilities/mogp_fitting.py:318-324 runs) driven
    through its reverse-communication core ``_lbfgsb.setulb``: ``advance()`` runs the optimiser until it
    asks for the objective and gradient at ``self.x`` ("FG") or stops ("DONE"); ``feed(f, g)`` hands them
    over.  Same routine, same inputs, same iterates as ``minimize`` -- but the caller decides WHEN the
    objective is evaluated, so all restarts of a round share one batched device call without one
    Python thread per restart."""

    MAXCOR, FTOL, GTOL, MAXFUN, MAXITER, MAXLS = 10, 2.220446049250313e-09, 1e-05, 15000, 15000, 20

    def __init__(self, setulb, x0, bounds):
        n = len(x0)
        self._setulb = setulb
        self.m = self.MAXCOR
        self.factr = self.FTOL / np.finfo(float).eps
        int_dtype = np.int32
        try:
            from scipy.optimize._lbfgsb_py import HAS_ILP64

            int_dtype = np.int64 if HAS_ILP64 else np.int32
        except ImportError:
            pass
        self.nbd = np.zeros(n, dtype=int_dtype)
        self.low = np.zeros(n, dtype=np.float64)
        self.up = np.zeros(n, dtype=np.float64)
        x0 = np.asarray(x0, dtype=np.float64).ravel()
        if bounds is not None:
            if len(bounds) != n:
                raise ValueError("length of x0 != length of bounds")
            lb = np.array([-np.inf if b[0] is None else float(b[0]) for b in bounds])
            ub = np.array([np.inf if b[1] is None else float(b[1]) for b in bounds])
            if (lb > ub).any():
                raise ValueError("LBFGSB - one of the lower bounds is greater than an upper bound.")
            x0 = np.clip(x0, lb, ub)
            code = {(False, False): 0, (True, False): 1, (True, True): 2, (False, True): 3}
            for k in range(n):
                has_l, has_u = not np.isinf(lb[k]), not np.isinf(ub[k])
                if has_l:
                    self.low[k] = lb[k]
                if has_u:
                    self.up[k] = ub[k]
                self.nbd[k] = code[(has_l, has_u)]
        m = self.m
        self.x = np.array(x0, dtype=np.float64)
        self.f = np.array(0.0, dtype=np.float64)
        self.g = np.zeros((n,), dtype=np.float64)
        self.wa = np.zeros(2 * m * n + 5 * n + 11 * m * m + 8 * m, np.float64)
        self.iwa = np.zeros(3 * n, dtype=int_dtype)
        self.task = np.zeros(2, dtype=int_dtype)
        self.ln_task = np.zeros(2, dtype=int_dtype)
        self.lsave = np.zeros(4, dtype=int_dtype)
        self.isave = np.zeros(44, dtype=int_dtype)
        self.dsave = np.zeros(29, dtype=np.float64)
        self.n_iterations = 0
        self.nfev = 0

    def advance(self) -> str:
        while True:
            self.g = np.asarray(self.g, dtype=np.float64)
            self._setulb(self.m, self.x, self.low, self.up, self.nbd, self.f, self.g, self.factr, self.GTOL, self.wa,
                         self.iwa, self.task, self.lsave, self.isave, self.dsave, self.MAXLS, self.ln_task)
            if self.task[0] == 3:
                return "FG"
            if self.task[0] == 1:
                self.n_iterations += 1
                if self.n_iterations >= self.MAXITER:
                    self.task[0], self.task[1] = 5, 504
                elif self.nfev > self.MAXFUN:
                    self.task[0], self.task[1] = 5, 502
                continue
            return "DONE"

    def feed(self, f: float, g: np.ndarray):
        self.f = np.float64(f)
        self.g = np.array(g, dtype=np.float64)
        self.nfev += 1


def _get_setulb():
    """scipy's private reverse-communication entry point, or None when this scipy lays it out differently."""
    try:
        import inspect

        from scipy.optimize import _lbfgsb
        from scipy.optimize._lbfgsb_py import _minimize_lbfgsb

        if "ln_task" not in inspect.getsource(_minimize_lbfgsb):
            return None
        return _lbfgsb.setulb
    except Exception:
        return None


def _evaluate_shared(dev, thetas: np.ndarray, nugget, adaptive, fit_nugget, comm, cap: int, stats: dict):
    """Objective + gradient at this rank's ``thetas`` -- evaluated by ALL ranks together.

    Every round each rank contributes the raw parameter rows its own restarts ask for (at most ``cap``); one
    all-gather makes the global list known everywhere, rank r evaluates rows r, r + G, r + 2G, ... of it in one
    batched device call, and a second all-gather returns (value, gradient, info) for every row.  Rounds are
    therefore global and every rank evaluates ceil(total / G) systems per round, whoever owns the restarts:
    no rank waits in the post-fit all-gather for another one's stragglers (round 1: 0.10 s of a 0.77 s step at
    2-8 GPUs), and a strong-scaling split of 15 restarts over 8 ranks keeps all 8 busy until fewer than 8 restarts
    are left.  A restart's values do not depend on which systems share its batch, so the optimiser trajectories
    are those of the single-rank run.  Returns (val, grad, info) for this rank's rows and the global row count;
    ranks with no rows still take part (collectives) until the global count is zero."""
    p = thetas.shape[1] if thetas.size else (dev.d + 1 + (1 if fit_nugget else 0))
    G, r = comm.size, comm.rank
    k = thetas.shape[0]
    msg = np.zeros((cap, p + 1))
    msg[:k, 0] = 1.0
    msg[:k, 1:] = thetas
    t0 = time.perf_counter()
    allm = comm.allgather(msg)                                   # (G, cap, p + 1)
    stats["map_comm_s"] = stats.get("map_comm_s", 0.0) + time.perf_counter() - t0
    valid = allm[:, :, 0] > 0.5
    owners, slots = np.nonzero(valid)                            # rank-major order: identical on every rank
    total = owners.shape[0]
    if total == 0:
        return np.empty(0), np.empty((0, p)), np.empty(0, dtype=np.int32), 0
    rows = allm[owners, slots, 1:]
    # row g goes to rank (g + rot) % G; the rotation moves the ranks that get the extra row of an uneven split
    rot = stats.get("rounds", 0) % G
    stats["rounds"] = stats.get("rounds", 0) + 1
    mine = np.arange((r - rot) % G, total, G)
    share = (total + G - 1) // G
    res = np.zeros((share, p + 2))
    res[:, p + 1] = -1.0                                          # info of padding rows
    if mine.size:
        if fit_nugget:
            val, grad, info, _ = dev.logpost_batch(rows[mine], nugget, adaptive, fit_nugget=True)
        else:
            val, grad, info, _ = dev.logpost_batch(rows[mine], nugget, adaptive)
        res[: mine.size, 0] = val
        res[: mine.size, 1 : p + 1] = grad
        res[: mine.size, p + 1] = info
        stats["evals_here"] = stats.get("evals_here", 0) + int(mine.size)
    t0 = time.perf_counter()
    allr = comm.allgather(res)                                   # (G, share, p + 2)
    stats["map_comm_s"] = stats.get("map_comm_s", 0.0) + time.perf_counter() - t0
    # global row g was evaluated by rank (g + rot) % G as its (g // G)-th row
    gidx = np.nonzero(owners == r)[0]
    got = allr[(gidx + rot) % G, gidx // G]
    return got[:, 0].copy(), got[:, 1 : p + 1].copy(), got[:, p + 1].astype(np.int32), total


def _run_map_restarts_lockstep(setulb, dev, priors, starts, raw_bounds, nugget, adaptive, fit_nugget, stats,
                               comm=None, cap=None):
    """All restarts advance in lock step in ONE thread; each round is one batched device call (shared with the
    other ranks when ``comm`` has more than one: see ``_evaluate_shared``)."""
    n = len(starts)
    shared = comm is not None and comm.size > 1
    cap = max(n, 1) if cap is None else cap
    sstats: dict = {}
    steppers: list = [None] * n
    out: list = [None] * n
    active = []
    for i, s0 in enumerate(starts):
        try:
            steppers[i] = _LbfgsbStepper(setulb, s0, raw_bounds)
            active.append(i)
        except ValueError:
            raise
    n_batches = n_evals = 0
    gpu_s = 0.0
    last_eval: dict[int, bytes] = {}
    while True:
        need = []
        for i in list(active):
            st = steppers[i]
            status = st.advance()
            # the first request repeats x0, which minimize() has already evaluated once (ScalarFunction
            # caches it); evaluating it here once is the same single evaluation
            if status == "FG":
                key = st.x.tobytes()
                if last_eval.get(i) == key:      # same point asked twice: feed the cached values back
                    st.feed(st.f, st.g)
                    need_again = True
                    while need_again:
                        status = st.advance()
                        need_again = status == "FG" and last_eval.get(i) == st.x.tobytes()
                        if need_again:
                            st.feed(st.f, st.g)
                    if status == "DONE":
                        out[i] = (float(st.f), np.array(st.x))
                        active.remove(i)
                        continue
                need.append(i)
            else:
                out[i] = (float(st.f), np.array(st.x))
                active.remove(i)
        ok = []
        for i in need:
            with np.errstate(over="ignore"):
                finite = np.all(np.isfinite(np.exp(steppers[i].x)))
            if finite:
                ok.append(i)
            else:                                     # FloatingPointError in the reference: restart skipped
                out[i] = None
                active.remove(i)
        if not shared:
            if not need:
                break
            if not ok:
                continue
        npar = len(starts[0]) if n else (dev.d + 1 + (1 if fit_nugget else 0))
        thetas = np.array([steppers[i].x for i in ok]).reshape(len(ok), npar)
        t0 = time.perf_counter()
        if shared:
            val, grad, info, total = _evaluate_shared(dev, thetas, nugget, adaptive, fit_nugget, comm, cap, sstats)
            if total == 0:                            # no rank has anything left to evaluate
                break
        elif fit_nugget:
            val, grad, info, _ = dev.logpost_batch(thetas, nugget, adaptive, fit_nugget=True)
        else:
            val, grad, info, _ = dev.logpost_batch(thetas, nugget, adaptive)
        gpu_s += time.perf_counter() - t0
        n_batches += 1
        n_evals += len(ok)
        for k, i in enumerate(ok):
            if info[k] != 0 or not np.isfinite(val[k]):   # LinAlgError in the reference: restart skipped
                out[i] = None
                active.remove(i)
                continue
            theta = steppers[i].x
            steppers[i].feed(float(val[k] - priors.logp(theta)), grad[k] - priors.dlogp_draw(theta))
            last_eval[i] = theta.tobytes()
    if stats is not None:
        stats.update(batches=n_batches, evals=n_evals, gpu_s=gpu_s, **sstats)
    return out


def run_map_restarts(dev, priors: MapPriors, starts, raw_bounds, nugget: float, adaptive: bool,
                     stats: Optional[dict] = None, return_all: bool = False, fit_nugget: bool = False,
                     lockstep: bool = True, comm=None, cap: Optional[int] = None):
    """One scipy L-BFGS-B minimisation per starting point, as the restart loop of
    ``_fit_single_GP_MAP`` (exauq/utilities/mogp_fitting.py:309-335), with all objective/gradient
    requests of a round evaluated in one batched GPU call.  Default: the optimisers are stepped in
    lock step through scipy's reverse-communication core (``_LbfgsbStepper``); if this scipy does not
    expose it in the expected form (or ``lockstep=False``), each ``minimize`` call runs in its own
    thread and the requests rendezvous (``_Rendezvous``).  Restarts that hit a non-positive-definite covariance or a floating
    point overflow are skipped (:332-335).  Returns the raw parameters with the lowest
    negative log-posterior, or None if every restart failed (``return_all``: the list of
    ``(value, theta)`` / ``None`` per start instead).  With a multi-rank ``comm`` every rank passes its OWN
    starts (at most ``cap`` of them, the same ``cap`` on every rank) and the evaluations of each round are spread
    evenly over all ranks (``_evaluate_shared``); the call is collective."""
    d = dev.d
    setulb = _get_setulb() if lockstep else None
    if comm is not None and comm.size > 1 and setulb is None:
        raise RuntimeError("multi-rank MAP estimation needs scipy's reverse-communication L-BFGS-B core (setulb)")
    if setulb is not None:
        out = _run_map_restarts_lockstep(setulb, dev, priors, starts, raw_bounds, nugget, adaptive, fit_nugget, stats,
                                         comm=comm, cap=cap)
        if return_all:
            return out
        done = [o for o in out if o is not None]
        if not done:
            return None
        return done[int(np.argmin(np.array([o[0] for o in done])))][1]
    rv = _Rendezvous(dev, len(starts), nugget, adaptive, fit_nugget)
    out: list[Optional[tuple[float, np.ndarray]]] = [None] * len(starts)

    def worker(wid: int, theta0: np.ndarray):
        cache: dict[bytes, tuple[float, np.ndarray]] = {}

        def fg(theta):
            key = np.asarray(theta, dtype=float).tobytes()
            if key not in cache:
                cache.clear()
                if not np.all(np.isfinite(np.exp(theta))):
                    raise FloatingPointError("overflow in hyperparameter transform")
                val, grad, info, exc = rv.evaluate(wid, theta)
                if exc is not None:
                    raise exc
                if info != 0 or not np.isfinite(val):
                    raise LinAlgError("covariance matrix is not positive definite")
                cache[key] = (float(val - priors.logp(theta)), grad - priors.dlogp_draw(theta))
            return cache[key]

        try:
            with np.errstate(divide="raise", over="raise", invalid="raise"):
                res = minimize(lambda t: fg(t)[0], theta0, method="L-BFGS-B", jac=lambda t: fg(t)[1],
                               bounds=raw_bounds)
            out[wid] = (float(res["fun"]), np.array(res["x"], dtype=float))
        except (LinAlgError, FloatingPointError):
            out[wid] = None
        finally:
            rv.retire()

    threads = [threading.Thread(target=worker, args=(i, s), daemon=True) for i, s in enumerate(starts)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    rv.close()
    if stats is not None:
        stats.update(batches=rv.n_batches, evals=rv.n_evals, gpu_s=rv.gpu_s)
    if return_all:
        return out
    done = [o for o in out if o is not None]
    if not done:
        return None
    vals = np.array([o[0] for o in done])
    return done[int(np.argmin(vals))][1]
