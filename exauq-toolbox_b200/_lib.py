"""ctypes binding of libexauq_b200.so (the C ABI declared in include/exauq_b200.h).

There is deliberately no fallback: if the shared library is missing, or no B200 is
visible, every compute call raises.  The oracle under ``oracle/`` is never imported here.
"""

from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EXQ_LIB", os.path.join(_HERE, "libexauq_b200.so"))  # EXQ_LIB: kernel-variant experiments

EXQ_OK, EXQ_ERR_ARG, EXQ_ERR_CUDA, EXQ_ERR_STATE, EXQ_NOT_PD = 0, -1, -2, -3, 1
KERNEL_IDS = {"Matern52": 0, "SquaredExponential": 1, "ProductMat52": 2}
NUGGET_FIXED, NUGGET_ADAPTIVE, NUGGET_FIT = 0, 1, 2
(PROF_GEMM_PREDICT, PROF_GEMM_FACTOR, PROF_LEAF, PROF_ASSEMBLE, PROF_STREAM, PROF_GEMM_FACTOR_I8, PROF_SLICE, PROF_FUSED,
 PROF_LAUUM_GRAD) = range(9)

# every symbol include/exauq_b200.h declares: (name, restype, argtypes)
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_int64)
_vp = C.c_void_p
SIGNATURES = {
    "exq_init": (C.c_int, [C.c_int]),
    "exq_shutdown": (C.c_int, []),
    "exq_last_error": (C.c_char_p, []),
    "exq_device_sm_count": (C.c_int, []),
    "exq_variance_slices": (C.c_int, []),
    "exq_factor_slices": (C.c_int, []),
    "exq_set_int8_digits": (C.c_int, [C.c_int, C.c_int]),
    "exq_launch_count": (C.c_int64, []),
    "exq_set_guard": (C.c_int, [C.c_double, C.c_double]),
    "exq_guard_counts": (C.c_int, [_lp, _lp]),
    "exq_set_refine_steps": (C.c_int, [C.c_int]),
    "exq_set_fused": (C.c_int, [C.c_int, C.c_int]),
    "exq_fused_info": (C.c_int, [_ip, _ip, _ip]),
    "exq_set_pei_fused": (C.c_int, [C.c_int]),
    "exq_set_grad_fused": (C.c_int, [C.c_int]),
    "exq_gp_kappa": (C.c_int, [_vp, _dp, _ip]),
    "exq_int8_gemm_ops": (C.c_double, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "exq_int8_gemm_schedule": (C.c_int64, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint32), C.c_int64]),
    "exq_timer_start": (C.c_int, []),
    "exq_timer_stop": (C.c_int, [_dp]),
    "exq_prof_enable": (C.c_int, [C.c_int]),
    "exq_prof_reset": (C.c_int, []),
    "exq_prof_get": (C.c_int, [C.c_int, _dp, _lp, _dp, _dp]),
    "exq_corr": (C.c_int, [C.c_int, C.c_int, _dp, _dp, C.c_int64, _dp, C.c_int64, _dp]),
    "exq_gp_create": (C.c_int, [C.c_int64, C.c_int, _dp, _dp, C.c_int, C.POINTER(_vp)]),
    "exq_gp_destroy": (C.c_int, [_vp]),
    "exq_gp_fit": (C.c_int, [_vp, _dp, C.c_double, C.c_double, C.c_int, _dp, _dp, _dp]),
    "exq_gp_pivot_order": (C.c_int, [_vp, _dp, C.c_double, _ip, _ip]),
    "exq_gp_fit_pivoted": (C.c_int, [_vp, _dp, C.c_double, C.c_int, _dp, _dp]),
    "exq_gp_predict": (C.c_int, [_vp, _dp, C.c_int64, _dp, _dp]),
    "exq_gp_cross_cov": (C.c_int, [_vp, _dp, C.c_int64, _dp]),
    "exq_gp_loo": (C.c_int, [_vp, _dp, _dp, _dp]),
    "exq_gp_loo_refit": (C.c_int, [_vp, _ip, C.c_int, _dp, _dp, _dp]),
    "exq_gp_kinv": (C.c_int, [_vp, _dp]),
    "exq_gp_logpost_batch": (C.c_int, [_vp, C.c_int, _dp, C.c_double, C.c_int, _dp, _dp, _ip, _dp]),
    "exq_cand_create": (C.c_int, [C.c_int64, C.c_int, _dp, C.POINTER(_vp)]),
    "exq_cand_destroy": (C.c_int, [_vp]),
    "exq_cand_set_points": (C.c_int, [_vp, _dp]),
    "exq_pei_eval": (C.c_int, [_vp, _vp, _dp, C.c_int, C.c_double]),
    "exq_pei_argmax": (C.c_int, [_vp, _lp, _dp]),
    "exq_pei_add_point": (C.c_int, [_vp, _vp, _dp]),
    "exq_pei_batch": (C.c_int, [_vp, _vp, C.c_int, _lp, _dp]),
    "exq_cand_download": (C.c_int, [_vp, C.c_int, _dp]),
    "exq_cand_best_devptr": (_vp, [_vp]),
}


class ExqError(RuntimeError):
    """A call into libexauq_b200 failed (bad argument, CUDA error, or wrong state)."""


class NotPositiveDefiniteError(np.linalg.LinAlgError):
    """The covariance matrix is not positive definite (mogp raises LinAlgError here)."""


_lib: Optional[C.CDLL] = None
_device: Optional[int] = None


def load() -> C.CDLL:
    """dlopen the library (no GPU needed) and declare every prototype."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ExqError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)"
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load().exq_last_error().decode()


def check(rc: int, what: str) -> None:
    if rc == EXQ_OK:
        return
    msg = f"{what}: {last_error()} (code {rc})"
    if rc == EXQ_NOT_PD:
        raise NotPositiveDefiniteError(msg)
    if rc == EXQ_ERR_ARG:
        raise ValueError(msg)
    raise ExqError(msg)


def init(device: Optional[int] = None) -> int:
    """Bind this process to one GPU (default: LOCAL_RANK or 0). Raises if there is none."""
    global _device
    lib = load()
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0")) if _device is None else _device
    if _device is not None and _device == device:
        return device
    check(lib.exq_init(int(device)), "exq_init")
    _device = int(device)
    return _device


def _d(a: np.ndarray):
    return a.ctypes.data_as(_dp)


def as_f64(a, shape=None) -> np.ndarray:
    out = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        out = out.reshape(shape)
    return out


# ---------------------------------------------------------------------------------------
# thin functional wrappers (numpy in / numpy out)
# ---------------------------------------------------------------------------------------
def corr(kernel: str, corr_raw, A, B) -> np.ndarray:
    init()
    A = as_f64(A)
    B = as_f64(B)
    corr_raw = as_f64(corr_raw)
    d = corr_raw.shape[0]
    A = A.reshape(-1, d)
    B = B.reshape(-1, d)
    out = np.empty((A.shape[0], B.shape[0]))
    check(load().exq_corr(KERNEL_IDS[kernel], d, _d(corr_raw), _d(A), A.shape[0], _d(B), B.shape[0], _d(out)),
          "exq_corr")
    return out


class DeviceGP:
    """Owner of an ``exq_gp_t`` handle: training data resident in HBM plus fitted state."""

    def __init__(self, X, y, kernel: str):
        init()
        self.X = as_f64(X)
        if self.X.ndim != 2:
            raise ValueError("X must be 2-D")
        self.y = as_f64(y).reshape(-1)
        if self.y.shape[0] != self.X.shape[0]:
            raise ValueError("X and y lengths differ")
        self.N, self.d = self.X.shape
        self.kernel = kernel
        self._h = _vp()
        check(load().exq_gp_create(self.N, self.d, _d(self.X), _d(self.y), KERNEL_IDS[kernel], C.byref(self._h)),
              "exq_gp_create")
        self.nugget_used = None
        self.logdet = None
        self.quad = None

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            load().exq_gp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def fit(self, corr_raw, cov: float, nugget: float = 0.0, adaptive: bool = False):
        corr_raw = as_f64(corr_raw).reshape(self.d)
        nu, ld, q = C.c_double(), C.c_double(), C.c_double()
        check(load().exq_gp_fit(self._h, _d(corr_raw), float(cov), float(nugget),
                                NUGGET_ADAPTIVE if adaptive else NUGGET_FIXED,
                                C.byref(nu), C.byref(ld), C.byref(q)), "exq_gp_fit")
        self.nugget_used, self.logdet, self.quad = nu.value, ld.value, q.value
        return self

    def pivot_order(self, corr_raw, cov: float):
        """(piv, rank) of the pivoted Cholesky factorisation of cov * C(X, X) (LAPACK dpstrf's pivot rule and
        stopping tolerance; nugget = "pivot").  Leaves the handle unfitted."""
        corr_raw = as_f64(corr_raw).reshape(self.d)
        piv = np.zeros(self.N, dtype=np.int32)
        rank = C.c_int()
        check(load().exq_gp_pivot_order(self._h, _d(corr_raw), float(cov), piv.ctypes.data_as(_ip), C.byref(rank)),
              "exq_gp_pivot_order")
        return piv, int(rank.value)

    def fit_pivoted(self, corr_raw, cov: float, rank: int):
        """Fit a handle whose data are already in pivot order: factor of the leading ``rank`` columns, mogp's
        decreasing diagonal for the rest (nugget = "pivot", rank-deficient or ill-conditioned case)."""
        corr_raw = as_f64(corr_raw).reshape(self.d)
        ld, q = C.c_double(), C.c_double()
        check(load().exq_gp_fit_pivoted(self._h, _d(corr_raw), float(cov), int(rank), C.byref(ld), C.byref(q)),
              "exq_gp_fit_pivoted")
        self.nugget_used, self.logdet, self.quad = 0.0, ld.value, q.value
        return self

    def kappa(self):
        """(kappa, on_fp64): the conditioning proxy (max L_ii / min L_ii)^2 of the last fit and whether the
        int8 guard sent that fit to the FP64 DMMA path."""
        k, f = C.c_double(), C.c_int()
        check(load().exq_gp_kappa(self._h, C.byref(k), C.byref(f)), "exq_gp_kappa")
        return k.value, bool(f.value)

    @property
    def neg_log_likelihood(self) -> float:
        return 0.5 * (self.logdet + self.quad + self.N * np.log(2.0 * np.pi))

    def predict(self, Xc):
        Xc = as_f64(Xc).reshape(-1, self.d)
        mean = np.empty(Xc.shape[0])
        var = np.empty(Xc.shape[0])
        check(load().exq_gp_predict(self._h, _d(Xc), Xc.shape[0], _d(mean), _d(var)), "exq_gp_predict")
        return mean, var

    def cross_cov(self, Xq) -> np.ndarray:
        Xq = as_f64(Xq).reshape(-1, self.d)
        out = np.empty((Xq.shape[0], self.N))
        check(load().exq_gp_cross_cov(self._h, _d(Xq), Xq.shape[0], _d(out)), "exq_gp_cross_cov")
        return out

    def loo(self):
        mean, var, nes = np.empty(self.N), np.empty(self.N), np.empty(self.N)
        check(load().exq_gp_loo(self._h, _d(mean), _d(var), _d(nes)), "exq_gp_loo")
        return mean, var, nes

    def loo_refit(self, idx):
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        n = idx.shape[0]
        mean, var, nes = np.empty(n), np.empty(n), np.empty(n)
        check(load().exq_gp_loo_refit(self._h, idx.ctypes.data_as(_ip), n, _d(mean), _d(var), _d(nes)),
              "exq_gp_loo_refit")
        return mean, var, nes

    def kinv(self) -> np.ndarray:
        out = np.empty((self.N, self.N))
        check(load().exq_gp_kinv(self._h, _d(out)), "exq_gp_kinv")
        return out

    def logpost_batch(self, theta_raw, nugget: float = 0.0, adaptive: bool = False, want_grad: bool = True,
                      fit_nugget: bool = False):
        """``fit_nugget``: rows of ``theta_raw`` carry ln(nugget) as a last, (d+2)-th entry."""
        npar = self.d + 1 + (1 if fit_nugget else 0)
        theta_raw = as_f64(theta_raw).reshape(-1, npar)
        R = theta_raw.shape[0]
        val = np.empty(R)
        grad = np.empty((R, npar)) if want_grad else None
        info = np.zeros(R, dtype=np.int32)
        nug = np.empty(R)
        mode = NUGGET_FIT if fit_nugget else (NUGGET_ADAPTIVE if adaptive else NUGGET_FIXED)
        check(load().exq_gp_logpost_batch(self._h, R, _d(theta_raw), float(nugget), mode, _d(val),
                                          _d(grad) if want_grad else None, info.ctypes.data_as(_ip), _d(nug)),
              "exq_gp_logpost_batch")
        return val, grad, info, nug


class PivotedDeviceGP:
    """nugget = "pivot" (mogp GaussianProcess.fit -> pivot_cholesky; exauq/core/emulators.py:129-137): the
    covariance matrix is factored with diagonal pivoting on the device, the training data are then held in pivot
    order so that the factor and its inverse stay triangular for every downstream kernel, and everything indexed by
    training point (LOO sweep, kinv, cross-covariances) is handed back in the caller's order.  Same methods as
    ``DeviceGP``; ``rank`` is the numerical rank LAPACK's stopping rule finds."""

    def __init__(self, X, y, kernel: str):
        self.base = DeviceGP(X, y, kernel)          # caller's order: pivot search, hyperparameter estimation
        self.X, self.y, self.N, self.d, self.kernel = self.base.X, self.base.y, self.base.N, self.base.d, kernel
        self.dev: Optional[DeviceGP] = None
        self.piv = np.arange(self.N, dtype=np.int32)
        self.rank = self.N
        self.nugget_used = None

    @property
    def _h(self):
        return self.dev._h

    def close(self):
        self.base.close()
        if self.dev is not None:
            self.dev.close()

    def fit(self, corr_raw, cov: float, nugget: float = 0.0, adaptive: bool = False):
        piv, rank = self.base.pivot_order(corr_raw, cov)
        if rank < 1:
            raise NotPositiveDefiniteError("pivoted Cholesky: no positive diagonal entry")
        if self.dev is not None:
            self.dev.close()
        self.piv, self.rank = piv, rank
        self.dev = DeviceGP(self.X[piv], self.y[piv], self.kernel)
        done = False
        if rank == self.N:
            try:                                     # full rank: the fast recursion on the permuted matrix is the
                self.dev.fit(corr_raw, cov, 0.0, False)   # pivoted factor (same column order, no interchanges left)
                done = True
            except NotPositiveDefiniteError:
                pass
        if not done:
            self.dev.fit_pivoted(corr_raw, cov, rank)
        self.nugget_used, self.logdet, self.quad = None, self.dev.logdet, self.dev.quad
        return self

    @property
    def neg_log_likelihood(self) -> float:
        return self.dev.neg_log_likelihood

    def kappa(self):
        return self.dev.kappa()

    def predict(self, Xc):
        return self.dev.predict(Xc)

    def cross_cov(self, Xq) -> np.ndarray:
        out = np.empty((np.asarray(Xq).reshape(-1, self.d).shape[0], self.N))
        out[:, self.piv] = self.dev.cross_cov(Xq)
        return out

    def _unpermute(self, *vs):
        outs = []
        for v in vs:
            o = np.empty_like(v)
            o[self.piv] = v
            outs.append(o)
        return tuple(outs)

    def loo(self):
        return self._unpermute(*self.dev.loo())

    def loo_refit(self, idx):
        inv = np.empty(self.N, dtype=np.int32)
        inv[self.piv] = np.arange(self.N, dtype=np.int32)
        return self.dev.loo_refit(inv[np.asarray(idx, dtype=np.int64)])

    def kinv(self) -> np.ndarray:
        out = np.empty((self.N, self.N))
        out[np.ix_(self.piv, self.piv)] = self.dev.kinv()
        return out

    def logpost_batch(self, theta_raw, nugget: float = 0.0, adaptive: bool = False, want_grad: bool = True,
                      fit_nugget: bool = False):
        """Objective / gradient of hyperparameter estimation: a covariance matrix that is numerically of full rank has
        the same log-posterior with or without pivoting, so the batched recursion (nugget 0, no jitter) is used;
        a matrix it cannot factor is reported through ``info`` (restart skipped), where mogp would continue on its
        rank-deficient surrogate factor."""
        return self.base.logpost_batch(theta_raw, 0.0, False, want_grad, False)


class DeviceCandidates:
    """A candidate set resident in HBM, with its mean / variance / PEI arrays."""

    def __init__(self, Xc):
        init()
        self.Xc = as_f64(Xc)
        if self.Xc.ndim != 2:
            raise ValueError("Xc must be 2-D")
        self.M, self.d = self.Xc.shape
        self._h = _vp()
        check(load().exq_cand_create(self.M, self.d, _d(self.Xc), C.byref(self._h)), "exq_cand_create")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            load().exq_cand_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_points(self, Xc):
        Xc = as_f64(Xc).reshape(self.M, self.d)
        check(load().exq_cand_set_points(self._h, _d(Xc)), "exq_cand_set_points")
        self.Xc = Xc

    def pei_eval(self, gp: DeviceGP, rep, tmax: float):
        rep = as_f64(rep).reshape(-1, self.d) if rep is not None and len(rep) else np.empty((0, self.d))
        check(load().exq_pei_eval(gp._h, self._h, _d(rep) if rep.shape[0] else None, rep.shape[0], float(tmax)),
              "exq_pei_eval")

    def argmax(self):
        idx, val = C.c_int64(), C.c_double()
        check(load().exq_pei_argmax(self._h, C.byref(idx), C.byref(val)), "exq_pei_argmax")
        return idx.value, val.value

    def add_point(self, gp: DeviceGP, x):
        x = as_f64(x).reshape(self.d)
        check(load().exq_pei_add_point(gp._h, self._h, _d(x)), "exq_pei_add_point")

    def batch(self, gp: DeviceGP, batch_size: int):
        idx = np.empty(batch_size, dtype=np.int64)
        val = np.empty(batch_size)
        check(load().exq_pei_batch(gp._h, self._h, int(batch_size), idx.ctypes.data_as(_lp), _d(val)),
              "exq_pei_batch")
        return idx, val

    def download(self, which: str) -> np.ndarray:
        out = np.empty(self.M)
        check(load().exq_cand_download(self._h, {"mean": 0, "var": 1, "pei": 2}[which], _d(out)),
              "exq_cand_download")
        return out

    def best_devptr(self) -> int:
        return int(load().exq_cand_best_devptr(self._h) or 0)


# ---- timing helpers ---------------------------------------------------------------------
def timer_start():
    check(load().exq_timer_start(), "exq_timer_start")


def timer_stop() -> float:
    ms = C.c_double()
    check(load().exq_timer_stop(C.byref(ms)), "exq_timer_stop")
    return ms.value


def prof_enable(mask: int):
    check(load().exq_prof_enable(int(mask)), "exq_prof_enable")


def prof_reset():
    check(load().exq_prof_reset(), "exq_prof_reset")


def prof_get(cls: int):
    ms, n, fl, by = C.c_double(), C.c_int64(), C.c_double(), C.c_double()
    check(load().exq_prof_get(cls, C.byref(ms), C.byref(n), C.byref(fl), C.byref(by)), "exq_prof_get")
    return {"ms": ms.value, "launches": n.value, "flops": fl.value, "bytes": by.value}


def launch_count() -> int:
    return int(load().exq_launch_count())


def variance_slices() -> int:
    """Base-256 digits per operand of the int8 tensor-core variance product (0: FP64 DMMA GEMM)."""
    return int(load().exq_variance_slices())


def factor_slices() -> int:
    """Base-256 digits of the int8 GEMMs inside the factorisation (0: FP64 DMMA everywhere)."""
    return int(load().exq_factor_slices())


def set_guard(widen: float = 2.0 ** 16, fallback: float = 1e12) -> None:
    """Thresholds of the conditioning guard of the int8 paths (see include/exauq_b200.h)."""
    init()
    check(load().exq_set_guard(float(widen), float(fallback)), "exq_set_guard")


def set_fused(max_rows: int = 512, cluster_ctas: int = 0) -> None:
    """Fused sub-tree kernel: largest diagonal block handled by one launch (0 = off) and CTAs per cluster (0 = automatic)."""
    init()
    check(load().exq_set_fused(int(max_rows), int(cluster_ctas)), "exq_set_fused")


def set_pei_fused(on: bool = True) -> None:
    """Fused prediction / PEI pipeline (no FP64 covariance chunk) on or off (round-1 three-pass pipeline)."""
    init()
    check(load().exq_set_pei_fused(1 if on else 0), "exq_set_pei_fused")


def set_grad_fused(on: bool = True) -> None:
    """Gradient contraction in the epilogue of the int8 LAUUM kernel (on) or as a separate kernel over a stored K^-1."""
    init()
    check(load().exq_set_grad_fused(1 if on else 0), "exq_set_grad_fused")


def fused_info() -> dict:
    init()
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    check(load().exq_fused_info(C.byref(a), C.byref(b), C.byref(c)), "exq_fused_info")
    return {"max_rows": a.value, "clusters_of_8": b.value, "clusters_of_16": c.value}


def set_refine_steps(steps: int = 1) -> None:
    """Iterative-refinement steps of alpha after an int8 factorisation (0 = off)."""
    init()
    check(load().exq_set_refine_steps(int(steps)), "exq_set_refine_steps")


def guard_counts() -> tuple[int, int]:
    """(times a 6-digit product was widened to 7 digits, times a factorisation fell back to FP64 DMMA)."""
    w, f = C.c_int64(), C.c_int64()
    check(load().exq_guard_counts(C.byref(w), C.byref(f)), "exq_guard_counts")
    return int(w.value), int(f.value)


KR_FULL, KR_LE_J, KR_GE_J, KR_LE_I, KR_GE_I = range(5)


def int8_gemm_ops(M: int, N: int, K: int, B: int, krange: int, lower_only: bool, digits: int) -> float:
    """int8 operations the factorisation GEMM kernel issues for this product (host arithmetic, no GPU needed)."""
    return float(load().exq_int8_gemm_ops(M, N, K, B, krange, int(bool(lower_only)), digits))


def int8_gemm_schedule(M: int, N: int, K: int, B: int, krange: int, lower_only: bool, grid: int) -> np.ndarray:
    """Balanced work list of the factorisation GEMM kernel for this product on ``grid`` CTAs (host arithmetic, no GPU
    needed): entry i = item of CTA i % grid in round i // grid, packed (system << 20 | row tile << 10 | column tile),
    0xffffffff = idle."""
    lib = load()
    n = int(lib.exq_int8_gemm_schedule(M, N, K, B, krange, int(bool(lower_only)), grid, None, 0))
    if n < 0:
        raise ValueError("exq_int8_gemm_schedule: bad shape")
    out = np.empty(n, dtype=np.uint32)
    lib.exq_int8_gemm_schedule(M, N, K, B, krange, int(bool(lower_only)), grid, out.ctypes.data_as(C.POINTER(C.c_uint32)), n)
    return out


def set_int8_digits(variance_digits: int, factor_digits: int) -> None:
    """Switch the int8 tensor-core paths at run time (0 = FP64 DMMA); applies to GPs created afterwards."""
    init()
    check(load().exq_set_int8_digits(int(variance_digits), int(factor_digits)), "exq_set_int8_digits")
