"""Stationary kernels of mogp-emulator 0.7.2 (restated).

Call sites in the reference: exauq/core/emulators.py:122-126 (``kernel_f`` table),
:547 (``correlation``).  Raw correlation parameters are ``corr_raw = -2 ln(l)``
(exauq/core/modelling.py:1855-1867), so ``exp(corr_raw) = 1/l**2`` weights the squared
coordinate differences.
"""

import numpy as np


class KernelBase(object):
    "Shared input handling + distance computation"

    def _check_inputs(self, x1, x2, params):
        params = np.array(params)
        assert params.ndim == 1, "parameters must be a vector"
        D = len(params)
        assert D >= 1, "must have at least one parameter"

        x1 = np.array(x1)
        assert x1.ndim == 1 or x1.ndim == 2, "bad number of dimensions in input x1"
        if x1.ndim == 2:
            # ValueError (not Assertion) on dim mismatch is relied on by
            # exauq/core/emulators.py:556-567
            if x1.shape[1] != D:
                raise ValueError("bad shape for x1")
        else:
            if len(x1) % D != 0:
                raise ValueError("bad shape for x1")
            x1 = np.reshape(x1, (-1, D))

        x2 = np.array(x2)
        assert x2.ndim == 1 or x2.ndim == 2, "bad number of dimensions in input x2"
        if x2.ndim == 2:
            if x2.shape[1] != D:
                raise ValueError("bad shape for x2")
        else:
            if len(x2) % D != 0:
                raise ValueError("bad shape for x2")
            x2 = np.reshape(x2, (-1, D))
        return x1, x2, params

    def get_n_params(self, inputs):
        inputs = np.array(inputs)
        assert inputs.ndim == 2
        return inputs.shape[1]

    def kernel_f(self, x1, x2, params):
        raise NotImplementedError


class StationaryKernel(KernelBase):
    "K depends only on r2 = sum_k exp(params_k) (x1_k - x2_k)^2"

    def calc_r2(self, x1, x2, params):
        x1, x2, params = self._check_inputs(x1, x2, params)
        exp_theta = np.exp(params)
        r2_matrix = np.sum(
            exp_theta * (x1[:, np.newaxis, :] - x2[np.newaxis, :, :]) ** 2, axis=-1
        )
        if np.any(np.isinf(r2_matrix)):
            raise FloatingPointError("Inf encountered in kernel distance computation")
        return r2_matrix

    def calc_dr2dtheta(self, x1, x2, params):
        "d r2 / d params_k, shape (D, n1, n2)"
        x1, x2, params = self._check_inputs(x1, x2, params)
        exp_theta = np.exp(params)
        return np.transpose(
            exp_theta * (x1[:, np.newaxis, :] - x2[np.newaxis, :, :]) ** 2, (2, 0, 1)
        )

    def kernel_f(self, x1, x2, params):
        return self.calc_K(self.calc_r2(x1, x2, params))

    def kernel_deriv(self, x1, x2, params):
        "dK/dparams_k, shape (D, n1, n2)"
        r2 = self.calc_r2(x1, x2, params)
        return self.calc_dKdr2(r2)[np.newaxis, :, :] * self.calc_dr2dtheta(x1, x2, params)


class SqExpBase(object):
    def calc_K(self, r2):
        assert np.all(r2 >= 0.0), "kernel distances must be non-negative"
        return np.exp(-0.5 * r2)

    def calc_dKdr2(self, r2):
        assert np.all(r2 >= 0.0), "kernel distances must be non-negative"
        return -0.5 * np.exp(-0.5 * r2)


class Mat52Base(object):
    def calc_K(self, r2):
        assert np.all(r2 >= 0.0), "kernel distances must be non-negative"
        r = np.sqrt(5.0 * r2)
        return (1.0 + r + 5.0 / 3.0 * r2) * np.exp(-r)

    def calc_dKdr2(self, r2):
        assert np.all(r2 >= 0.0), "kernel distances must be non-negative"
        r = np.sqrt(5.0 * r2)
        return -5.0 / 6.0 * (1.0 + r) * np.exp(-r)


class SquaredExponential(SqExpBase, StationaryKernel):
    def __str__(self):
        return "Squared Exponential Kernel"


class Matern52(Mat52Base, StationaryKernel):
    def __str__(self):
        return "Matern 5/2 Kernel"


class ProductKernel(KernelBase):
    "K = prod_k k1d(exp(params_k) (x1_k - x2_k)^2)"

    def calc_r2(self, x1, x2, params):
        "per-dimension squared scaled distances, shape (D, n1, n2)"
        x1, x2, params = self._check_inputs(x1, x2, params)
        exp_theta = np.exp(params)
        r2 = np.transpose(
            exp_theta * (x1[:, np.newaxis, :] - x2[np.newaxis, :, :]) ** 2, (2, 0, 1)
        )
        if np.any(np.isinf(r2)):
            raise FloatingPointError("Inf encountered in kernel distance computation")
        return r2

    def kernel_f(self, x1, x2, params):
        return np.prod(self.calc_K(self.calc_r2(x1, x2, params)), axis=0)

    def kernel_deriv(self, x1, x2, params):
        r2 = self.calc_r2(x1, x2, params)
        k = self.calc_K(r2)
        out = np.empty_like(r2)
        for d in range(r2.shape[0]):
            others = np.prod(np.delete(k, d, axis=0), axis=0)
            out[d] = others * self.calc_dKdr2(r2[d]) * r2[d]
        return out


class ProductMat52(Mat52Base, ProductKernel):
    def __str__(self):
        return "Product Matern 5/2 Kernel"
