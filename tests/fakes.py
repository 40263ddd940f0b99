"""CPU stand-ins for ``DeviceGP`` / ``DeviceCandidates`` built on the oracle -- test
infrastructure for the HOST logic (MAP driver, design batch, sharding) where no GPU exists."""

import numpy as np

from oracle import gp_oracle as orc


class OracleGP:
    def __init__(self, X, y, kernel):
        self.X, self.y, self.kernel = np.asarray(X, float), np.asarray(y, float), kernel
        self.N, self.d = self.X.shape
        self.gp = None
        self.nugget_used = None
        self.n_batches = 0

    def fit(self, corr_raw, cov, nugget=0.0, adaptive=False):
        self.gp = orc.fit(self.X, self.y, self.kernel, corr_raw, cov, "adaptive" if adaptive else nugget)
        self.nugget_used = float(self.gp.theta.nugget)
        return self

    def loo(self):
        K = self.gp.theta.cov * orc.corr(self.kernel, self.gp.theta.corr_raw, self.X, self.X)
        K = K + self.gp.theta.nugget * np.eye(self.N)
        Ki = np.linalg.inv(K)
        var = 1.0 / np.diag(Ki)
        mean = self.y - (Ki @ self.y) * var
        return mean, var, orc.nes_error(mean, var, self.y)

    def predict(self, Xc):
        return orc.predict(self.gp, Xc)

    def loo_refit(self, idx):
        return orc.loo_refit(self.X, self.y, self.kernel, self.gp.theta.corr_raw, self.gp.theta.cov,
                             float(self.gp.theta.nugget), [int(i) for i in idx])

    def logpost_batch(self, theta_raw, nugget=0.0, adaptive=False, want_grad=True):
        theta_raw = np.atleast_2d(theta_raw)
        self.n_batches += 1
        R = theta_raw.shape[0]
        val, grad, info = np.empty(R), np.empty((R, self.d + 1)), np.zeros(R, dtype=np.int32)
        for r in range(R):
            try:
                val[r], grad[r] = orc.neg_log_likelihood_grad(self.X, self.y, self.kernel, theta_raw[r],
                                                              "adaptive" if adaptive else nugget)
            except np.linalg.LinAlgError:
                val[r], grad[r], info[r] = np.nan, np.nan, 1
        return val, grad, info, np.zeros(R)


class OracleCandidates:
    def __init__(self, Xc):
        self.Xc = np.asarray(Xc, float)
        self.M, self.d = self.Xc.shape
        self.pei = None

    def set_points(self, Xc):
        self.Xc = np.asarray(Xc, float).reshape(self.M, self.d)

    def pei_eval(self, gp, rep, tmax):
        self.pei = orc.pei(gp.gp, gp.kernel, self.Xc, rep, tmax)

    def argmax(self):
        i = int(np.argmax(self.pei))
        return i, float(self.pei[i])

    def add_point(self, gp, x):
        self.pei = self.pei * (1 - orc.corr(gp.kernel, gp.gp.theta.corr_raw, self.Xc, np.atleast_2d(x))[:, 0])

    def batch(self, gp, batch_size):
        idx, val = [], []
        for _ in range(batch_size):
            i, v = self.argmax()
            idx.append(i)
            val.append(v)
            self.add_point(gp, self.Xc[i])
        return np.array(idx, dtype=np.int64), np.array(val)

    def download(self, which):
        return self.pei
