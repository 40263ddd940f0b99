"""A single-level design batch over a fixed candidate set, sharded by rank.

This is the reference's ``compute_single_level_loo_samples``
(exauq/core/designers.py:709-837) with its differential-evolution maximiser
(exauq/utilities/optimisation.py:14-104) replaced by an arg max over candidates resident in
HBM, as BASELINE.json's configs describe:

  1. LOO sweep of the GP with its fitted hyperparameters + NES errors   (designers.py:263-272)
  2. errors GP: MAP estimation with the length-scale lower bounds         (:274-292)
  3. PEI of the errors GP over the candidates, ``batch_size`` rounds of
     arg max + repulsion update                                          (:825-835)

Ranks hold the same training data; candidates and optimiser restarts are sharded (one
process per GPU).  The only exchanges are tiny all-gathers: the best restart, and one
(value, index, point) triple per rank per batch round.
"""

from __future__ import annotations

import itertools
import math
import time
from typing import Optional, Sequence

import numpy as np

from ._lib import DeviceCandidates, DeviceGP
from .gp import run_map_restarts
from .priors import MapPriors


# ---------------------------------------------------------------------------------------
# host-side geometry (vectorised restatements of O(N * 2^d) Python loops in the reference)
# ---------------------------------------------------------------------------------------
def _rows_equal(a: np.ndarray, b: np.ndarray, tol: float = 1e-9) -> np.ndarray:
    """Input.__eq__ (exauq/core/modelling.py:287-307): every coordinate within rel/abs 1e-9."""
    return np.all(np.abs(a - b) <= np.maximum(tol * np.maximum(np.abs(a), np.abs(b)), tol), axis=-1)


def domain_corners(bounds: Sequence[Sequence[float]]) -> np.ndarray:
    """SimulatorDomain._define_corners (modelling.py:2142-2151): product of the bounds, duplicates dropped."""
    pts = np.array(list(itertools.product(*[tuple(b) for b in bounds])), dtype=float)
    if all(abs(b[1] - b[0]) > 1e-9 * max(1.0, abs(b[0]), abs(b[1])) for b in bounds):
        return pts                      # no degenerate side: the 2^d corners are distinct
    keep = []
    for i in range(pts.shape[0]):
        if not any(_rows_equal(pts[i], pts[j]) for j in keep):
            keep.append(i)
    return pts[keep]


def pseudopoints(X: np.ndarray, bounds: Sequence[Sequence[float]]) -> np.ndarray:
    """SimulatorDomain.calculate_pseudopoints (modelling.py:2233-2277): for every face of the box
    the projection of the (first) training input closest to it (closest_boundary_points,
    :2157-2231; corner inputs skipped, duplicates and corners dropped), then the corners."""
    X = np.asarray(X, dtype=float)
    corners = domain_corners(bounds)
    d = X.shape[1]
    lo = np.array([b[0] for b in bounds], dtype=float)
    hi = np.array([b[1] for b in bounds], dtype=float)

    def _close(a, b, tol=1e-9):
        return np.abs(a - b) <= np.maximum(tol * np.maximum(np.abs(a), np.abs(b)), tol)

    # a point equals some corner iff every coordinate sits on one of its two bounds
    is_corner = np.all(_close(X, lo) | _close(X, hi), axis=1)
    cand_idx = np.nonzero(~is_corner)[0]
    out: list[np.ndarray] = []
    if cand_idx.size:
        for i in range(d):
            for bound in (bounds[i][0], bounds[i][1]):
                dist = np.abs(X[cand_idx, i] - bound)
                p = X[cand_idx[int(np.argmin(dist))]].copy()
                p[i] = bound
                if (out and np.any(_rows_equal(np.array(out), p))) or np.all(_close(p, lo) | _close(p, hi)):
                    continue
                out.append(p)
    pts = np.array(out, dtype=float).reshape(-1, d)
    return np.vstack([pts, corners])


def loo_error_raw_bounds(bounds: Sequence[Sequence[float]]):
    """_compute_loo_error_bounds (designers.py:280-292) mapped to raw parameters as
    MogpEmulator._compute_raw_param_bounds does (emulators.py:453-485): length scales bounded
    below by 0.25/sqrt(ln 10) * side, i.e. corr_raw bounded ABOVE; process variance free."""
    scale = 0.25 / math.sqrt(math.log(10))
    raw = [(None, -2.0 * math.log(scale * (b[1] - b[0]))) for b in bounds]
    return tuple(raw) + ((None, None),)


# ---------------------------------------------------------------------------------------
# communication: one process per GPU
# ---------------------------------------------------------------------------------------
class SingleComm:
    rank, size = 0, 1

    def allgather(self, arr: np.ndarray) -> np.ndarray:
        return np.asarray(arr, dtype=float)[None, ...]

    def barrier(self):
        pass


class TorchComm:
    """torch.distributed (NCCL over NVLink on GPUs, gloo in CPU tests) for the tiny all-gathers."""

    def __init__(self, device=None):
        import torch
        import torch.distributed as dist

        self._torch, self._dist = torch, dist
        self.rank, self.size = dist.get_rank(), dist.get_world_size()
        self.device = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")

    def allgather(self, arr: np.ndarray) -> np.ndarray:
        arr = np.ascontiguousarray(arr, dtype=np.float64)
        t = self._torch.as_tensor(arr.reshape(-1), device=self.device)
        out = self._torch.empty(self.size * t.numel(), dtype=t.dtype, device=self.device)
        self._dist.all_gather_into_tensor(out, t)
        return out.cpu().numpy().reshape((self.size,) + arr.shape)

    def barrier(self):
        self._dist.barrier()


def pick_global_best(vals: np.ndarray, idxs: np.ndarray) -> int:
    """Rank holding the global arg max: largest value, lowest global index on ties, NaN ignored."""
    vals = np.where(np.isnan(vals), -np.inf, vals)
    best = np.max(vals)
    tied = np.nonzero(vals == best)[0]
    return int(tied[np.argmin(idxs[tied])])


# ---------------------------------------------------------------------------------------
# shared steps
# ---------------------------------------------------------------------------------------
def fit_errors_gp(X, nes, kernel, bounds, n_tries, seed, comm, gp_factory, stats, restart_split: str = "per_rank"):
    """MAP fit of a GP to NES errors with the LOO length-scale bounds, restarts sharded over
    ranks; every rank ends with the same fitted GP (compute_loo_errors_gp :274-276).

    ``restart_split="per_rank"`` (weak scaling): every rank draws its own ``n_tries`` starting points (seed offset
    by the rank).  ``"global"`` (strong scaling): ``n_tries`` is the total -- every rank draws the same ``n_tries``
    points from the same seed and owns points rank, rank + G, ...; the fitted GP is then the single-rank one
    whatever the number of ranks.  Either way the objective evaluations of a round are spread evenly over all
    ranks (``gp._evaluate_shared``)."""
    d = X.shape[1]
    gp_e = gp_factory(X, nes, kernel)
    priors = MapPriors(X)
    if restart_split == "global":
        if seed is not None:
            np.random.seed(seed)
        every = [priors.sample() for _ in range(n_tries)]
        mine = list(range(comm.rank, n_tries, comm.size))
        starts = [every[i] for i in mine]
        cap = (n_tries + comm.size - 1) // comm.size
    else:
        if seed is not None:
            np.random.seed(seed + 7919 * comm.rank)
        starts = [priors.sample() for _ in range(n_tries)]
        mine = [comm.rank + comm.size * k for k in range(n_tries)]
        cap = n_tries
    mstats: dict = {}
    results = run_map_restarts(gp_e, priors, starts, loo_error_raw_bounds(bounds), 0.0, True, stats=mstats,
                               return_all=True, comm=comm, cap=cap)
    local = np.full(d + 3, np.inf)           # (objective, global restart index, theta)
    for k, r in enumerate(results):
        if r is not None and (r[0] < local[0] or (r[0] == local[0] and mine[k] < local[1])):
            local[0], local[1], local[2:] = r[0], float(mine[k]), r[1]
    tg = time.perf_counter()
    gathered = comm.allgather(local)
    stats["map_allgather_wait_s"] = time.perf_counter() - tg
    if not np.any(np.isfinite(gathered[:, 0])):
        raise RuntimeError("GP fitting failed")
    best = np.min(gathered[:, 0])
    tied = np.nonzero(gathered[:, 0] == best)[0]
    theta = gathered[int(tied[np.argmin(gathered[tied, 1])]), 2:]     # lowest restart index wins ties, as np.argmin does
    gp_e.fit(theta[:d], float(np.exp(theta[d])), 0.0, True)
    stats["map_batches"] = stats.get("map_batches", 0) + mstats.get("batches", 0)
    stats["map_evals"] = stats.get("map_evals", 0) + mstats.get("evals", 0)
    stats["map_gpu_s"] = stats.get("map_gpu_s", 0.0) + mstats.get("gpu_s", 0.0)
    stats["map_comm_s"] = stats.get("map_comm_s", 0.0) + mstats.get("map_comm_s", 0.0)
    stats["map_evals_here"] = stats.get("map_evals_here", 0) + mstats.get("evals_here", mstats.get("evals", 0))
    return gp_e, theta


def loo_refit_sharded(gp, comm=None, indices=None):
    """The reference's explicit leave-one-out loop (compute_loo_errors_gp, exauq/core/designers.py:263-272: one
    fixed-hyperparameter refit per left-out point) sharded by left-out point: rank r refits the indices at
    positions r, r + G, r + 2G, ... (block-cyclic, so that every rank sees the same mix of cheap and expensive
    systems), then ONE all-gather of (mean, variance, NES) per index.  BASELINE.json configs[2]
    (N = 8192, d = 10, 8192 refits over 2 / 4 / 8 B200).  ``gp`` is a fitted ``DeviceGP`` holding the same data on
    every rank.  Returns (mean, var, nes) over ``indices`` (default: all N), identical on every rank."""
    comm = comm or SingleComm()
    idx = np.arange(gp.N, dtype=np.int32) if indices is None else np.ascontiguousarray(indices, dtype=np.int32)
    n = idx.shape[0]
    mine = idx[comm.rank :: comm.size]
    share = (n + comm.size - 1) // comm.size
    buf = np.full((share, 3), np.nan)
    if mine.shape[0]:
        m, v, e = gp.loo_refit(mine)
        buf[: mine.shape[0], 0], buf[: mine.shape[0], 1], buf[: mine.shape[0], 2] = m, v, e
    allb = comm.allgather(buf)                                # (G, share, 3)
    pos = np.arange(n)
    out = allb[pos % comm.size, pos // comm.size]
    return out[:, 0].copy(), out[:, 1].copy(), out[:, 2].copy()


def global_argmax(cand, comm, candidate_offset: int, scale: float = 1.0, stats: Optional[dict] = None):
    """(value * scale, global index, point) of the best candidate over all ranks."""
    t0 = time.perf_counter()
    err = None
    try:
        i, v = cand.argmax()
    except Exception as exc:            # e.g. every PEI value on this rank is NaN
        if comm.size == 1:
            raise
        err, i, v = exc, 0, float("nan")
    if comm.size == 1:
        return v * scale, i + candidate_offset, np.array(cand.Xc[i], dtype=float)
    msg = np.concatenate([[v * scale, float(i + candidate_offset)], cand.Xc[i]])
    t1 = time.perf_counter()
    allm = comm.allgather(msg)      # every rank reaches the collective before anyone raises
    if np.all(np.isnan(allm[:, 0])):
        raise err if err is not None else RuntimeError("no finite pseudo-expected improvement on any rank")
    if stats is not None:
        stats["pei_argmax_s"] = stats.get("pei_argmax_s", 0.0) + (t1 - t0)
        stats["pei_allgather_s"] = stats.get("pei_allgather_s", 0.0) + (time.perf_counter() - t1)
    w = pick_global_best(allm[:, 0], allm[:, 1])
    return float(allm[w, 0]), int(allm[w, 1]), allm[w, 2:].copy()


def find_cross_level_duplicate(Xs: dict) -> Optional[tuple]:
    """A training input shared by two levels, as _find_input_repetition_across_levels
    (designers.py:925-955) reports -- by a sort instead of a product over all levels' inputs."""
    levels = sorted(Xs)
    X = np.vstack([Xs[lv] for lv in levels])
    tag = np.concatenate([np.full(Xs[lv].shape[0], lv) for lv in levels])
    order = np.argsort(X[:, 0], kind="stable")
    x0 = X[order, 0]
    hi = np.searchsorted(x0, x0 + 2.1e-9 * np.maximum(1.0, np.abs(x0)), side="right")
    for i in range(X.shape[0]):
        if hi[i] <= i + 1:
            continue
        a = X[order[i]]
        nb = order[i + 1 : hi[i]]
        hit = np.nonzero(_rows_equal(X[nb], a) & (tag[nb] != tag[order[i]]))[0]
        if hit.size:
            return a, int(tag[order[i]]), int(tag[nb[hit[0]]])
    return None


# ---------------------------------------------------------------------------------------
# the design batch
# ---------------------------------------------------------------------------------------
def design_batch(
    X: np.ndarray,
    y: np.ndarray,
    kernel: str,
    corr_length_scales,
    process_var: float,
    nugget: float,
    bounds: Sequence[Sequence[float]],
    candidates,
    batch_size: int = 1,
    n_tries: int = 15,
    seed: Optional[int] = None,
    comm=None,
    candidate_offset: int = 0,
    gp_factory=DeviceGP,
    cand_factory=DeviceCandidates,
    stats: Optional[dict] = None,
    restart_split: str = "per_rank",
    explicit_loo: bool = False,
):
    """New design points for a GP with training data (X, y) and fitted hyperparameters.

    ``candidates`` is this rank's shard (array or device-resident candidate set) whose first
    row has global index ``candidate_offset``.  Returns ``(points[batch_size, d],
    global_indices[batch_size], pei_values[batch_size], errors_gp_theta_raw)``; identical on
    every rank.  ``restart_split``: see ``fit_errors_gp``.  ``explicit_loo``: run the reference's N explicit
    refits sharded by left-out point (``loo_refit_sharded``) instead of the closed-form sweep.
    """
    comm = comm or SingleComm()
    X = np.ascontiguousarray(X, dtype=float)
    y = np.ascontiguousarray(y, dtype=float)
    d = X.shape[1]
    stats = stats if stats is not None else {}

    t0 = time.perf_counter()
    # 1. leave-one-out sweep with the fitted hyperparameters
    corr_raw = -2.0 * np.log(np.asarray(corr_length_scales, dtype=float))
    gp = gp_factory(X, y, kernel)
    gp.fit(corr_raw, process_var, nugget)
    _, _, nes = loo_refit_sharded(gp, comm) if explicit_loo else gp.loo()
    if not np.all(np.isfinite(nes)):
        raise FloatingPointError("non-finite NES error in the leave-one-out sweep")

    t1 = time.perf_counter()
    # 2. errors GP: MAP estimate, restarts sharded over ranks
    gp_e, theta = fit_errors_gp(X, nes, kernel, bounds, n_tries, seed, comm, gp_factory, stats, restart_split)
    t2 = time.perf_counter()

    # 3. PEI over this rank's candidates, batch_size rounds of arg max + repulsion update
    cand = candidates if hasattr(candidates, "pei_eval") else cand_factory(candidates)
    rep = pseudopoints(X, bounds)
    tp = time.perf_counter()
    cand.pei_eval(gp_e, rep, float(np.max(nes)))
    stats["pei_eval_call_s"] = time.perf_counter() - tp
    for k in ("pei_argmax_s", "pei_allgather_s", "pei_add_point_s"):
        stats[k] = 0.0
    pts = np.empty((batch_size, d))
    gidx = np.empty(batch_size, dtype=np.int64)
    vals = np.empty(batch_size)
    if comm.size == 1:
        li, lv = cand.batch(gp_e, batch_size)
        for s in range(batch_size):
            pts[s] = cand.Xc[li[s]]
        gidx[:], vals[:] = li + candidate_offset, lv
    else:
        for s in range(batch_size):
            vals[s], gidx[s], pts[s] = global_argmax(cand, comm, candidate_offset, stats=stats)
            ta = time.perf_counter()
            cand.add_point(gp_e, pts[s])
            stats["pei_add_point_s"] += time.perf_counter() - ta
    t3 = time.perf_counter()
    stats.update(loo_s=t1 - t0, map_s=t2 - t1, pei_s=t3 - t2, t_end=time.perf_counter())
    return pts, gidx, vals, theta


def _polish_batched(neg_pei, x0: np.ndarray, f0: float, bounds):
    """The polish step of scipy's differential evolution -- L-BFGS-B from the best population member with the
    2-point finite-difference gradient scipy's ``minimize`` builds when no ``jac`` is given
    (h_k = sqrt(eps) * sign(x_k) * max(1, |x_k|), flipped at a bound) -- but with the d + 1 points of every gradient
    evaluated in ONE device call instead of d + 1 scalar calls (device-side polish).
    Returns (x, f) of the polished point when it improves on (x0, f0) and respects the bounds, else (x0, f0) --
    the acceptance rule of ``DifferentialEvolutionSolver.solve``."""
    import scipy.optimize

    lb = np.array([b[0] for b in bounds], dtype=float)
    ub = np.array([b[1] for b in bounds], dtype=float)
    d = x0.shape[0]
    root_eps = np.sqrt(np.finfo(float).eps)

    def fun_and_grad(x):
        x = np.asarray(x, dtype=float)
        h = root_eps * np.where(x >= 0, 1.0, -1.0) * np.maximum(1.0, np.abs(x))
        flip = (x + h > ub) | (x + h < lb)
        h = np.where(flip, -h, h)
        pts = np.tile(x, (d + 1, 1))
        pts[1:] += np.diag(h)
        dx = pts[1:, :][np.arange(d), np.arange(d)] - x          # the step actually representable
        vals = np.asarray(neg_pei(pts.T), dtype=float)
        return float(vals[0]), (vals[1:] - vals[0]) / dx

    res = scipy.optimize.minimize(fun_and_grad, x0, jac=True, method="L-BFGS-B", bounds=list(zip(lb, ub)))
    if res.fun < f0 and np.all(res.x <= ub) and np.all(res.x >= lb):
        return np.array(res.x, dtype=float), float(res.fun)
    return x0, f0


def maximise_pei_de(gp_e, bounds: Sequence[Sequence[float]], repulsion_points, max_target: float,
                    seed: Optional[int] = None, cand_factory=DeviceCandidates, popsize: int = 15, polish: bool = True,
                    **de_kwargs):
    """Continuous maximisation of the pseudo-expected improvement of a fitted errors GP over the
    domain box, with the optimiser the reference uses -- scipy ``differential_evolution`` with
    ``tol = atol = 1e-9`` (exauq/utilities/optimisation.py:14-104) -- but with every generation's
    population evaluated in ONE device call (``vectorized=True``, ``updating='deferred'``; the
    reference's default immediate updating is inherently one point at a time) and the final polish
    (L-BFGS-B from the best member, scipy's default) with batched finite-difference gradients
    (``_polish_batched``).  Returns ``(x, pei(x))`` like ``maximise``.  SURVEY.md section 8f item 1."""
    import scipy.optimize

    d = len(bounds)
    rep = np.asarray(repulsion_points, dtype=float).reshape(-1, d) if len(repulsion_points) else None
    state: dict = {}

    def neg_pei(x):
        pts = np.ascontiguousarray(np.atleast_2d(np.asarray(x, dtype=float).T))     # (S, d)
        if pts.shape[1] != d:
            pts = pts.reshape(-1, d)
        c = state.get(pts.shape[0])                # one resident candidate set per population size (DE / polish)
        if c is None:
            c = state[pts.shape[0]] = cand_factory(pts)
        else:
            c.set_points(pts)
        c.pei_eval(gp_e, rep, max_target)
        vals = -c.download("pei")
        return vals if np.ndim(x) > 1 else float(vals[0])

    result = scipy.optimize.differential_evolution(neg_pei, bounds=[tuple(b) for b in bounds], tol=1e-9, atol=1e-9,
                                                   seed=seed, vectorized=True, updating="deferred", popsize=popsize,
                                                   polish=False, **de_kwargs)
    if not result.success:
        raise RuntimeError(f"Maximisation failed to converge: {result.message}")
    x, f = np.array(result.x, dtype=float), float(result.fun)
    if polish:
        x, f = _polish_batched(neg_pei, x, f, bounds)
    return x, -f


def multi_level_loo_errors(levels: dict, coefficients: dict, gp_factory=DeviceGP):
    """NES errors of the multi-level leave-one-out predictions against 0, per level
    (compute_multi_level_loo_error_data, designers.py:1045-1105): for the left-out input x_i of
    level l, estimate c_l (mu_-i - y_i) + sum_{j != l} c_j k_j(x_i)^T inv(cov_j C_j) y_j
    (compute_loo_prediction :958-998, compute_zero_mean_prediction :1001-1042; the inverse is
    nugget-free) and variance sum_j c_j^2 v_j.  ``levels[l]`` holds X, y, kernel,
    corr_length_scales, process_var, nugget.  Returns ({level: nes}, {level: fitted GP})."""
    dup = find_cross_level_duplicate({lv: np.asarray(v["X"], dtype=float) for lv, v in levels.items()})
    if dup is not None:
        raise ValueError(
            f"Training inputs across all levels must be distinct, but found common input {dup[0]} "
            f"at levels {dup[1]}, {dup[2]}."
        )
    gps, gps0 = {}, {}
    for lv, v in levels.items():
        X = np.ascontiguousarray(v["X"], dtype=float)
        y = np.ascontiguousarray(v["y"], dtype=float)
        if X.shape[0] < 2:
            raise ValueError(
                f"Could not perform leave-one-out calculation: levels {lv} not trained on at least two training data."
            )
        raw = -2.0 * np.log(np.asarray(v["corr_length_scales"], dtype=float))
        gps[lv] = gp_factory(X, y, v["kernel"])
        gps[lv].fit(raw, float(v["process_var"]), float(v["nugget"]))
        if float(v["nugget"]) == 0.0:
            gps0[lv] = gps[lv]
        else:                                            # kinv excludes the nugget (modelling.py:1095-1113)
            gps0[lv] = gp_factory(X, y, v["kernel"])
            gps0[lv].fit(raw, float(v["process_var"]), 0.0)
    out = {}
    for lv, v in levels.items():
        X = np.ascontiguousarray(v["X"], dtype=float)
        y = np.ascontiguousarray(v["y"], dtype=float)
        m, var, _ = gps[lv].loo()
        mean = coefficients[lv] * (m - y)
        variance = coefficients[lv] ** 2 * var
        for other in levels:
            if other == lv:
                continue
            zm, zv = gps0[other].predict(X)
            ov = zv if gps0[other] is gps[other] else gps[other].predict(X)[1]
            mean = mean + coefficients[other] * zm
            variance = variance + coefficients[other] ** 2 * ov
        sq = mean**2
        with np.errstate(divide="ignore", invalid="ignore"):
            nes = (variance + sq) / np.sqrt(2.0 * variance**2 + 4.0 * variance * sq)
        out[lv] = nes
    return out, gps


def multi_level_design_batch(
    levels: dict,
    coefficients: dict,
    costs: dict,
    bounds: Sequence[Sequence[float]],
    candidates,
    batch_size: int = 1,
    n_tries: int = 15,
    seed: Optional[int] = None,
    comm=None,
    candidate_offset: int = 0,
    gp_factory=DeviceGP,
    cand_factory=DeviceCandidates,
    stats: Optional[dict] = None,
    restart_split: str = "per_rank",
    return_details: bool = False,
):
    """compute_multi_level_loo_samples (designers.py:1184-1378) over a fixed candidate set:
    per-level errors GPs fitted to the multi-level LOO errors, then ``batch_size`` rounds of
    [arg max over candidates of PEI_l / cost_l for every level, take the best level, add the
    point to THAT level's repulsion set].  Returns a list of (level, global index, point, value)
    (``return_details``: also ``{level: errors-GP raw theta}`` and ``{level: NES errors}``)."""
    comm = comm or SingleComm()
    stats = stats if stats is not None else {}
    nes, _ = multi_level_loo_errors(levels, coefficients, gp_factory)
    Xc = candidates.Xc if hasattr(candidates, "Xc") else np.ascontiguousarray(candidates, dtype=float)
    gp_e, cands, thetas = {}, {}, {}
    for lv in sorted(levels):
        X = np.ascontiguousarray(levels[lv]["X"], dtype=float)
        if not np.all(np.isfinite(nes[lv])):
            raise FloatingPointError(f"non-finite NES error at level {lv}")
        lvl_seed = None if seed is None else seed + 104729 * lv
        gp_e[lv], thetas[lv] = fit_errors_gp(X, nes[lv], levels[lv]["kernel"], bounds, n_tries, lvl_seed, comm, gp_factory,
                                             stats, restart_split)
        cands[lv] = cand_factory(Xc)
        cands[lv].pei_eval(gp_e[lv], pseudopoints(X, bounds), float(np.max(nes[lv])))
    chosen = []
    for _ in range(batch_size):
        best = None
        for lv in sorted(levels):                       # max(...) keeps the first maximal level
            v, gi, pt = global_argmax(cands[lv], comm, candidate_offset, 1.0 / costs[lv])
            if best is None or v > best[3]:
                best = (lv, gi, pt, v)
        chosen.append(best)
        cands[best[0]].add_point(gp_e[best[0]], best[2])
    if return_details:
        return chosen, thetas, nes
    return chosen
