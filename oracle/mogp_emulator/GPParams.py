"""Hyperparameter container of mogp-emulator 0.7.2 (restated).

Call sites in the reference: exauq/core/emulators.py:299-309 (``theta.corr/.cov/.nugget/
.corr_raw``), :733-743, :819-826 (``GPParams(n_corr=, nugget=)``, ``set_data``);
exauq/utilities/mogp_fitting.py:237 (``theta.get_data()``).

Raw layout: ``[corr_raw (n_corr), cov_raw (1), nugget_raw (1, only if nugget_type ==
'fit')]`` with ``corr = exp(-corr_raw/2)``, ``cov = exp(cov_raw)``, ``nugget =
exp(nugget_raw)`` (cf. exauq/core/modelling.py:1855-1885 for the inverse maps).
"""

import numpy as np


class GPParams(object):
    def __init__(self, n_mean=0, n_corr=1, nugget="fit"):
        assert int(n_mean) >= 0, "number of mean parameters must be nonnegative"
        assert int(n_corr) >= 1, "number of correlation parameters must be positive"
        self.n_mean = int(n_mean)
        self.n_corr = int(n_corr)
        self._nugget = None
        self._nugget_type = None
        self.mean = None
        self._data = None
        # set nugget / nugget type
        if isinstance(nugget, str):
            if nugget not in ["fit", "adaptive", "pivot"]:
                raise ValueError("bad string for nugget type: %s" % nugget)
            self._nugget_type = nugget
        else:
            try:
                nugget = float(nugget)
            except TypeError:
                raise TypeError("nugget must be a string or a non-negative float")
            if nugget < 0.0:
                raise ValueError("nugget must be non-negative")
            self._nugget_type = "fixed"
            self._nugget = nugget

    @property
    def nugget_type(self):
        return self._nugget_type

    @property
    def n_params(self):
        return self.n_corr + 1 + int(self._nugget_type == "fit")

    @property
    def n_data(self):
        return self.n_params

    @property
    def corr_raw(self):
        if self._data is None:
            return None
        return self._data[: self.n_corr]

    @property
    def corr(self):
        if self._data is None:
            return None
        return np.exp(-0.5 * self._data[: self.n_corr])

    @property
    def cov_index(self):
        return self.n_corr

    @property
    def cov(self):
        if self._data is None:
            return None
        return np.exp(self._data[self.n_corr])

    @property
    def nugget(self):
        if self._nugget_type == "fit":
            if self._data is None:
                return None
            return np.exp(self._data[-1])
        return self._nugget

    @nugget.setter
    def nugget(self, new_nugget):
        if self._nugget_type == "fit":
            raise ValueError("cannot set a fitted nugget directly; use set_data")
        if new_nugget is not None:
            new_nugget = float(new_nugget)
            assert new_nugget >= 0.0, "nugget must be non-negative"
        self._nugget = new_nugget

    def get_data(self):
        return self._data

    def set_data(self, new_params):
        if new_params is None:
            self._data = None
            self.mean = None
            if self._nugget_type in ["adaptive", "pivot"]:
                self._nugget = None
            return
        new_params = np.array(new_params, dtype=float)
        assert new_params.shape == (self.n_params,), "bad shape for new parameters"
        self._data = np.copy(new_params)
        self.mean = None
        if self._nugget_type in ["adaptive", "pivot"]:
            self._nugget = None

    def same_shape(self, other):
        if isinstance(other, GPParams):
            return (
                self.n_corr == other.n_corr
                and self.nugget_type == other.nugget_type
                and self.n_mean == other.n_mean
            )
        return np.shape(other) == (self.n_params,)

    def __str__(self):
        return (
            "GPParams with:\ncorrelation = {}\ncovariance = {}\nnugget = {}".format(
                self.corr, self.cov, self.nugget
            )
        )
