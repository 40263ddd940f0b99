"""MAP hyperparameter estimation of mogp-emulator 0.7.2 (restated, single-output only).

The reference carries its own fork with a ``bounds`` argument
(exauq/utilities/mogp_fitting.py:51-346) and uses that one; this module only backs the
``mogp.fit_GP_MAP`` calls in the reference's tests (tests/unit/test_emulators.py:493,555).
"""

import numpy as np
from scipy.linalg import LinAlgError
from scipy.optimize import minimize

from .GaussianProcess import GaussianProcess


def _fit_single_GP_MAP(gp, n_tries=15, theta0=None, method="L-BFGS-B", **kwargs):
    n_tries = int(n_tries)
    assert n_tries > 0, "number of attempts must be positive"
    np.seterr(divide="raise", over="raise", invalid="raise")
    logpost_values, theta_values = [], []
    for i in range(n_tries):
        if i == 0 and theta0 is not None:
            theta = np.array(theta0)
            assert theta.shape == (gp.n_params,)
        else:
            theta = gp.priors.sample()
        try:
            res = minimize(gp.logposterior, theta, method=method, jac=gp.logpost_deriv,
                           options=kwargs)
            logpost_values.append(res["fun"])
            theta_values.append(res["x"])
        except LinAlgError:
            print("Matrix not positive definite, skipping this iteration")
        except FloatingPointError:
            print("Floating point error in optimization routine, skipping this iteration")
    if len(logpost_values) == 0:
        print("Minimization routine failed to return a value")
        gp.theta = None
    else:
        gp.fit(theta_values[int(np.argmin(np.array(logpost_values)))])
    return gp


def fit_GP_MAP(*args, n_tries=15, theta0=None, method="L-BFGS-B", skip_failures=True,
               refit=False, **kwargs):
    if len(args) == 1:
        gp = args[0]
        if not isinstance(gp, GaussianProcess):
            raise TypeError("first argument must be a GaussianProcess")
    elif len(args) < 2:
        raise TypeError("missing required inputs/targets arrays to GaussianProcess")
    else:
        gp_kwargs = {k: kwargs.pop(k) for k in
                     ["mean", "kernel", "priors", "nugget", "inputdict", "use_patsy"]
                     if k in kwargs}
        gp = GaussianProcess(*args, **gp_kwargs)
    gp = _fit_single_GP_MAP(gp, n_tries, theta0, method, **kwargs)
    if gp.theta.get_data() is None:
        raise RuntimeError("GP fitting failed")
    return gp
