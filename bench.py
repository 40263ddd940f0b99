#!/usr/bin/env python
"""Benchmark of the B200 design-batch hot path (BASELINE.json metric: design-batch seconds and
PEI candidate evals/s at N=4096, d=8).

A *step* is one single-level design batch (exauq/core/designers.py:709-837 over a fixed
candidate set): LOO sweep + NES errors of the GP, MAP estimation of the errors GP
(15 L-BFGS-B restarts, bounded), PEI over the rank's 1e6 resident candidates and
``batch_size`` rounds of arg max + repulsion update.  One process per GPU; candidates and
restarts are sharded by rank (weak scaling), exchanges are tiny all-gathers.

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...     # CPU port of the reference path, bounded sample

Prints ONE JSON line on rank 0.
"""

from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "pei_candidate_evals_per_s"
UNIT = "candidates/s"
# Objective + gradient evaluations of the errors-GP MAP fit with 15 restarts on this workload, as the GPU arm
# makes them (its L-BFGS-B iterates are bit-identical to scipy.optimize.minimize's, so the reference makes the
# same number); the GPU arm prints its own count as map_fit.evals_per_step_per_gpu.  Used to extrapolate the CPU arm
# (round 1 used a nominal 600, which inflated the MAP third of the CPU figure 1.8 x).
MEASURED_MAP_EVALS_15_RESTARTS = 333


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n-train", type=int, default=4096)
    ap.add_argument("--dim", type=int, default=8)
    ap.add_argument("--candidates", type=int, default=1_000_000, help="candidates per GPU")
    ap.add_argument("--batch-size", type=int, default=16)
    ap.add_argument("--restarts", type=int, default=15, help="MAP restarts per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stock-path", action="store_true", help="reference arm: skip the one stock MogpEmulator.fit timing")
    ap.add_argument("--strong-steps", type=int, default=5, help="timed steps of the strong-scaling sub-measurement (N > 1)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default): --candidates and --restarts are per GPU; strong: they are totals split over the GPUs")
    return ap.parse_args()


def workload_config(a, n_gpus):
    return {
        "workload": (f"single-level design batch: N={a.n_train}, d={a.dim}, Matern-5/2 (l=0.5, var=1.7, nugget=0); "
                     f"LOO sweep + NES + errors-GP MAP fit ({a.restarts} restarts/GPU, bounded) + PEI over "
                     f"{a.candidates} LHS candidates/GPU, batch_size={a.batch_size}; {n_gpus} GPU(s)"),
        "n_train": a.n_train, "dim": a.dim, "candidates_per_gpu": a.candidates,
        "candidates_total": a.candidates * n_gpus, "batch_size": a.batch_size,
        "restarts_per_gpu": a.restarts, "kernel": "Matern52",
        "sharding": "candidates and MAP restarts by rank; training data replicated",
        "l2_policy": "inputs_larger_than_L2 (64 MB candidates + 1 GiB covariance chunks per step)",
    }


# ------------------------------------------------------------------------------------------
# clocks sampling during the timed region (profiling recipe's clocks line)
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i",
                 str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s, p in zip(sm, power) if p > 0.3 * max(power)] or sm
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
# CPU port of the reference path, timed on a bounded sample and extrapolated
# ------------------------------------------------------------------------------------------
def cpu_port_sample(a, X, y, Xc, n_refits=1, n_cand=1024, full_grad_terms=1, n_gpus=1, map_evals=None):
    """Times the oracle (numpy/scipy restatement of the reference + mogp) on this box's host
    cores: ``n_refits`` fixed-hyperparameter LOO refits at full N, one log-posterior value and
    ``full_grad_terms`` of its d+1 gradient trace terms at full N, and PEI at ``n_cand``
    candidates; extrapolates linearly to one design batch."""
    from scipy import linalg

    from oracle import gp_oracle as orc

    N, d = X.shape
    corr_raw = np.full(d, -2 * math.log(0.5))
    t = {}
    t0 = time.perf_counter()
    _, _, nes = orc.loo_refit(X, y, "Matern52", corr_raw, 1.7, 0.0, list(range(n_refits)))
    t["loo_refit_s"] = (time.perf_counter() - t0) / n_refits
    t0 = time.perf_counter()
    gp = orc.fit(X, y, "Matern52", corr_raw, 1.7, 0.0)           # = one log-posterior value
    t["logpost_value_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    for _ in range(full_grad_terms):                              # mogp logpost_deriv: one cho_solve with N RHS per parameter
        _ = np.trace(linalg.cho_solve((gp.L, True), gp.K))
    t["logpost_grad_term_s"] = (time.perf_counter() - t0) / full_grad_terms
    rep = np.random.default_rng(1).uniform(0, 1, (2 * d + 2 ** d, d))
    t0 = time.perf_counter()
    _ = orc.pei(gp, "Matern52", Xc[:n_cand], rep, float(y.max()))
    t["pei_per_candidate_s"] = (time.perf_counter() - t0) / n_cand
    # the workload of the GPU arm at n_gpus GPUs (weak scaling): n_gpus x candidates, n_gpus x restarts, one LOO sweep
    M = a.candidates * n_gpus
    if map_evals is None:
        map_evals = int(round(MEASURED_MAP_EVALS_15_RESTARTS * a.restarts / 15.0)) * n_gpus
    batch_s = (N * t["loo_refit_s"] + map_evals * (t["logpost_value_s"] + (d + 1) * t["logpost_grad_term_s"])
               + M * t["pei_per_candidate_s"])
    t["design_batch_s_extrapolated"] = batch_s
    t["map_evals_assumed"] = map_evals
    t["design_batch_s_extrapolated_closed_form_loo"] = batch_s - (N - 1) * t["loo_refit_s"]
    sample = (f"{n_refits} LOO refit(s) at N={N} (x N), 1 log-posterior value + {full_grad_terms} of {d + 1} gradient "
              f"trace terms at N={N} (x {map_evals} evaluations: the count the GPU arm makes), PEI at {n_cand} candidates "
              f"(x {M}/{n_cand}); extrapolated linearly; excludes the reference's per-fit O(N^2) Python duplicate check "
              "and SVD/inv (stock_path below prices them), i.e. favours the CPU; design_batch_s_extrapolated_closed_form_loo "
              "= the same with ONE refit instead of N (what the GPU arm's closed-form sweep costs on the CPU)")
    return M / batch_s, t, sample


def stock_path_sample(a, n_predict=20):
    """ONE fit with given hyperparameters + scalar predictions through the UNMODIFIED reference classes from
    baseline/_ref (exauq.core.emulators.MogpEmulator on the restated mogp of oracle/): what every one of the N
    leave-one-out refits costs on the reference's stock path (duplicate check, Cholesky, SVD condition number,
    explicit inverse: exauq/core/emulators.py:268-311, exauq/core/modelling.py:1082-1113).  Timed once."""
    for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "baseline", "_ref")):
        if p not in sys.path and os.path.isdir(p):
            sys.path.insert(0, p)
    try:
        from exauq.core.emulators import MogpEmulator, MogpHyperparameters
        from exauq.core.modelling import Input, TrainingDatum
    except Exception as exc:  # reference package not present on this box
        return {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}
    from oracle import gp_oracle as orc

    N, d = a.n_train, a.dim
    X, y = orc.synthetic_problem(N, d)
    data = [TrainingDatum(Input(*row), float(v)) for row, v in zip(X, y)]
    gp = MogpEmulator(kernel="Matern52")
    t0 = time.perf_counter()
    gp.fit(data, hyperparameters=MogpHyperparameters(corr_length_scales=[0.5] * d, process_var=1.7, nugget=0.0))
    fit_s = time.perf_counter() - t0
    xs = [Input(*row) for row in orc.synthetic_candidates(n_predict, d)]
    t0 = time.perf_counter()
    for x in xs:
        gp.predict(x)
    pred_s = (time.perf_counter() - t0) / n_predict
    return {"what": f"MogpEmulator.fit(hyperparameters=...) at N={N}, d={d} + {n_predict} predict(x) calls, unmodified "
                    "exauq from baseline/_ref on the restated mogp; one LOO refit of the reference's loop "
                    "(exauq/core/designers.py:263-272) costs one such fit + one predict",
            "fit_s": fit_s, "predict_s_per_call": pred_s,
            "loo_sweep_s_extrapolated": N * (fit_s + pred_s)}


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core."""
    try:
        from threadpoolctl import threadpool_limits

        threadpool_limits(limits=os.cpu_count() or 1)
    except Exception:
        pass


def host_threads():
    try:
        from threadpoolctl import threadpool_info

        n = [p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"]
        if n:
            return int(max(n))
    except Exception:
        pass
    return os.cpu_count() or 1


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from exauq_toolbox_b200.workload import synthetic_candidates, synthetic_problem

    use_all_host_threads()
    X, y = synthetic_problem(a.n_train, a.dim)        # the same arrays the B200 arm uses
    Xc = synthetic_candidates(4096, a.dim)
    n_gpus = max(1, a.gpus)
    vals, times = [], []
    for s in range(a.warmup + a.steps):
        t0 = time.perf_counter()
        v, parts, sample = cpu_port_sample(a, X, y, Xc, n_gpus=n_gpus)
        if s >= a.warmup:
            vals.append(v)
            times.append(time.perf_counter() - t0)
    value = float(np.mean(vals))
    cores = host_threads()
    stock = stock_path_sample(a) if not a.no_stock_path else {"skipped": "--no-stock-path"}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * a.candidates * n_gpus / value,
        "sample_ms_per_step": 1e3 * float(np.mean(times)),
        "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(a, n_gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "parts": parts, "stock_path": stock},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
# the B200 arm
# ------------------------------------------------------------------------------------------
def run_b200(a):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import exauq_toolbox_b200 as exq
    from exauq_toolbox_b200 import _lib
    from exauq_toolbox_b200.design import SingleComm, TorchComm, design_batch

    torch = None
    try:
        import torch  # device plumbing only: pinned host buffers and torch.distributed
    except ImportError:
        pass
    _lib.init(local)
    comm = SingleComm()
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
        comm = TorchComm(device=torch.device("cuda", local))

    from exauq_toolbox_b200.workload import synthetic_candidates, synthetic_problem

    N, d, M = a.n_train, a.dim, a.candidates
    X, y = synthetic_problem(N, d)                    # the same arrays the reference arm uses
    Xc = synthetic_candidates(M, d, seed=2 + rank)
    if torch is not None:
        pinned = torch.from_numpy(Xc).pin_memory()
        Xc = pinned.numpy()
    bounds = [(0.0, 1.0)] * d
    ell, cov, nug = [0.5] * d, 1.7, 0.0
    offset = rank * M
    stats = {}

    def step(resident_cand=None):
        stats.clear()
        cand = resident_cand if resident_cand is not None else Xc
        return design_batch(X, y, "Matern52", ell, cov, nug, bounds, cand, batch_size=a.batch_size,
                            n_tries=a.restarts, seed=12345, comm=comm, candidate_offset=offset, stats=stats)

    def sync_barrier():
        _lib.timer_start(); _lib.timer_stop()     # drains the library stream
        comm.barrier()

    def timed(nsteps, resident):
        sync_barrier()
        _lib.timer_start()
        out = None
        for _ in range(nsteps):
            out = step(resident)
        ms = _lib.timer_stop()
        comm.barrier()
        ms_all = comm.allgather(np.array([ms]))
        return float(np.max(ms_all)), out

    # ---- device-resident measurement ---------------------------------------------------------
    resident = exq.DeviceCandidates(Xc)
    for _ in range(a.warmup):
        step(resident)
    # Inside the timed region only the large kernels carry CUDA-event pairs (a few hundred launches per step:
    # the candidates for the dominant kernel); the ~4000 small launches per step (leaves, FP64 GEMMs of the small
    # nodes, vector kernels) are timed in one extra, untimed step below so that their event records do not perturb
    # the headline.
    big = ((1 << _lib.PROF_GEMM_PREDICT) | (1 << _lib.PROF_GEMM_FACTOR_I8) | (1 << _lib.PROF_SLICE)
           | (1 << _lib.PROF_ASSEMBLE) | (1 << _lib.PROF_LAUUM_GRAD))
    if not _lib.factor_slices():
        big |= 1 << _lib.PROF_GEMM_FACTOR
    names = ["gemm_predict", "gemm_factor", "leaf", "assemble", "stream", "gemm_factor_i8", "slice", "fused", "lauum_grad"]
    _lib.prof_enable(big)
    _lib.prof_reset()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count()
    ms_total, out = timed(a.steps, resident)
    launches = _lib.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    prof = {n: _lib.prof_get(c) for c, n in enumerate(names)}
    # one extra step with every class instrumented, for the per-kernel table (scaled to "per step" below)
    _lib.prof_enable((1 << len(names)) - 1)
    _lib.prof_reset()
    step(resident)
    for c, n in enumerate(names):
        if not (big >> c) & 1:
            p = _lib.prof_get(c)
            prof[n] = {k: (v * a.steps if k in ("ms", "launches", "flops", "bytes") else v) for k, v in p.items()}
    _lib.prof_enable(0)
    ms_per_step = ms_total / a.steps
    value = world * M / (ms_per_step * 1e-3)
    map_evals, map_batches = stats.get("map_evals", 0), stats.get("map_batches", 0)

    # ---- end-to-end through the public API with host buffers ------------------------------------
    def e2e_step():
        pts, gidx, vals, theta = step(None)      # uploads X, y, candidates from (pinned) host memory
        return pts

    for _ in range(1):
        e2e_step()
    sync_barrier()
    _lib.timer_start()
    for _ in range(a.steps):
        e2e_step()
    e2e_ms = _lib.timer_stop()
    comm.barrier()
    e2e_ms = float(np.max(comm.allgather(np.array([e2e_ms])))) / a.steps
    h2d = 8 * (M * d + 2 * (N * d + N) + (2 * d + 2 ** d) * d + (d + 2) * (map_evals + 2))
    d2h = 8 * (3 * N + (d + 3) * map_evals) + 16 * a.batch_size
    e2e = {"value": world * M / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)}

    # ---- north-star strong scaling: ONE design batch (--candidates and --restarts in TOTAL) split over the ranks ----
    # candidates: M / world per rank; restarts: the same 15 starting points whatever the world size, owned round-robin
    # (design.fit_errors_gp restart_split="global"), every round's evaluations shared evenly over the ranks
    strong = None
    if world > 1 and a.scaling == "weak" and a.strong_steps > 0:
        Ms = (M + world - 1) // world
        res_s = exq.DeviceCandidates(Xc[:Ms])
        sstats = {}

        def strong_step():
            sstats.clear()
            return design_batch(X, y, "Matern52", ell, cov, nug, bounds, res_s, batch_size=a.batch_size,
                                n_tries=a.restarts, seed=12345, comm=comm, candidate_offset=rank * Ms, stats=sstats,
                                restart_split="global")

        for _ in range(2):
            strong_step()
        sync_barrier()
        _lib.timer_start()
        for _ in range(a.strong_steps):
            sout = strong_step()
        sms = _lib.timer_stop()
        comm.barrier()
        sms = float(np.max(comm.allgather(np.array([sms])))) / a.strong_steps
        strong = {"design_batch_s": sms * 1e-3, "candidates_total": Ms * world, "restarts_total": a.restarts,
                  "n_gpus": world, "steps": a.strong_steps, "value": Ms * world / (sms * 1e-3), "unit": UNIT,
                  "map_evals_total": int(np.sum(comm.allgather(np.array([float(sstats.get("map_evals", 0))])))),
                  "map_rounds": sstats.get("map_batches"), "map_s": sstats.get("map_s"), "pei_s": sstats.get("pei_s"),
                  "loo_s": sstats.get("loo_s"), "map_comm_s": sstats.get("map_comm_s"),
                  "selected_indices": [int(i) for i in sout[1]],
                  "note": "one batch of --candidates candidates and --restarts restarts in total, split over the ranks; "
                          "same 15 starting points as the 1-GPU run (restart_split=global)"}
        res_s.close()

    if rank != 0:
        return

    # ---- roofline of the dominant kernel (the DMMA tile GEMM) -------------------------------------
    peak_path = os.path.join(ROOT, "profiles", "fp64_peak_r01.json")
    fp64_peak, peak_src = 35.4, "fallback 35.4 TFLOP/s (cuBLAS DGEMM measured on this pool in round 1)"
    if os.path.exists(peak_path):
        pk = json.load(open(peak_path))
        fp64_peak = pk["fp64_dgemm_tflops_sustained"]
        peak_src = "measured cuBLAS DGEMM fp64 8192^3 on this pool (profiles/fp64_peak_r01.json); MEASURED_PEAKS.json has no FP64 entry"

    def roof(p):
        if not p["launches"] or p["ms"] <= 0:
            return None
        ach = p["flops"] / (p["ms"] * 1e-3) / 1e12
        return {"bound": "tensor", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak,
                "traffic": None, "launches_per_step": p["launches"] / a.steps,
                "avg_launch_ms": p["ms"] / p["launches"], "ms_per_step": p["ms"] / a.steps}

    # DRAM bytes per launch of the variance GEMM from the committed `ncu --set full` capture
    # (profiles/ncu_summary_r01.md: dram__bytes_read.sum + dram__bytes_write.sum at Np=4096, Mc=32768);
    # algorithmic bytes per launch are 8*(Mc*Np + Np^2/2) = 1.14e9
    NCU_TRAFFIC = {(4096, 8): 3.370598e9 + 12.609792e6}
    slices = _lib.variance_slices()
    vname = ("oz::colsumsq_kernel[variance Z=Linv*Kc^T on int8 tcgen05, %d base-256 digits]" % slices if slices
             else "gemm_dmma_kernel[variance Z=Linv*Kc^T]")
    fslices = _lib.factor_slices()
    fused_grad = prof["lauum_grad"]["launches"] > 0
    fname = ("oz::gemm_kernel[potrf/trtri nodes with k >= 1024%s on int8 tcgen05, %d base-256 digits]"
             % ("" if fused_grad else " and LAUUM", fslices))
    kern = {vname: roof(prof["gemm_predict"]),
            "gemm_dmma_kernel[potrf/trtri/lauum tiles]": roof(prof["gemm_factor"]),
            fname: roof(prof["gemm_factor_i8"])}
    vk = kern[vname]
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    # int8 roofline denominator: MEASURED on this pool's B200 by tools/micro/umma_i8_peak.cu (profiles/int8_peak_r02.json):
    # tcgen05.mma.kind::i8 issued back to back from resident shared-memory tiles, M=128, N=64/128/256.  These kernels
    # are timed inside a long step, so the SUSTAINED figure (2 s back to back; the best N) is the denominator; the burst
    # figure, 2 x the measured bf16 burst (round 1's assumed denominator) and the nominal 4500 are printed beside it.
    i8_burst, i8_sust, i8_src = None, None, None
    ipath = os.path.join(ROOT, "profiles", "int8_peak_r02.json")
    if os.path.exists(ipath):
        ip = json.load(open(ipath))["umma_i8_issue_loop_tops"]
        i8_burst = max(v["burst"] for v in ip.values())
        i8_sust = max(v["sustained"] for v in ip.values())
        i8_src = ("measured: profiles/int8_peak_r02.json (tools/micro/umma_i8_peak.cu, UTCIMMA issue loop M=128 N=64..256 "
                  "from resident smem tiles), sustained figure because the kernel is timed inside a long step")
    i8_assumed = 2.0 * peaks.get("bf16_tflops", 1703.4)
    i8_peak = i8_sust if i8_sust else i8_assumed
    if not i8_src:
        i8_src = "fallback: 2 x MEASURED_PEAKS.json bf16_tflops (no measured int8 file found)"
    traffic_path = os.path.join(ROOT, "profiles", "ncu_traffic_r02.json")
    ncu_traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}

    def add_int8(entry, p, digits):
        # The pipe these kernels run on is the int8 tensor core, so the roofline is int8 operations issued against
        # the int8 peak.  The FP64-EQUIVALENT rate (algorithmic FP64 flops of the product / time) is kept beside it:
        # above the FP64 DMMA peak means faster than any FP64 tensor-core kernel on this chip can be.
        i8 = p["bytes"] / (p["ms"] * 1e-3) / 1e12
        entry.update(fp64_equivalent_tflops=entry["achieved"], fp64_dmma_peak_tflops=fp64_peak,
                     fp64_equivalent_over_dmma_peak=entry["achieved"] / fp64_peak,
                     achieved=i8, peak=i8_peak, frac=i8 / i8_peak, unit="TOP/s", peak_source=i8_src,
                     frac_of_measured_burst=(i8 / i8_burst) if i8_burst else None, peak_measured_burst=i8_burst,
                     frac_of_2x_bf16_burst=i8 / i8_assumed, frac_of_nominal_4500=i8 / 4500.0,
                     int8_ops_counted="issued: the kernel's own tile walk x digit products (exq_int8_gemm_ops; "
                                      "tests/test_roofline_accounting.py)",
                     products_per_fp64_product=digits * (digits + 1) // 2)

    if kern[fname] is not None:
        add_int8(kern[fname], prof["gemm_factor_i8"], fslices)
        t = ncu_traffic.get("oz::gemm_kernel")
        if t and (N, d) == (4096, 8):
            # per-launch DRAM bytes from `ncu --set full` captures of the shipped kernel (launch-weighted mean over the
            # captured node shapes) with the algorithmic plane bytes of the same launches beside them
            kern[fname]["traffic"] = t.get("dram_bytes_per_launch")
            kern[fname]["algorithmic_bytes"] = t.get("algorithmic_bytes_per_launch")
            kern[fname]["traffic_source"] = t.get("source")
    if vk is not None and slices:
        add_int8(vk, prof["gemm_predict"], slices)
        # DRAM bytes per launch from the committed `ncu --set full` capture of this kernel (profiles/ncu_summary_r01.md,
        # Np=4096, Mc=32768, 6 digits): the 100 MB of Linv digit planes do not stay in L2 next to the streamed
        # candidate planes, so they are re-read once per group of 8 candidate tiles; DRAM is 9 % busy, not a limiter
        if (N, d, slices) == (4096, 8, 6):
            vk["traffic"] = 2.577724e9 + 20.721664e6
            vk["algorithmic_bytes"] = 6.0 * (32768 * 4096 + 4096 * 4096 / 2) + 8.0 * 64 * 32768
    elif vk is not None and (N, d) in NCU_TRAFFIC:
        vk["traffic"] = NCU_TRAFFIC[(N, d)]
        vk["algorithmic_bytes"] = 8.0 * (32768 * 4096 + 4096 * 4096 / 2)
    dominant = max((k for k in kern if kern[k]), key=lambda k: kern[k]["ms_per_step"])
    roofline = dict(kern[dominant])
    roofline.setdefault("peak_source", peak_src)
    roofline.update(kernel=dominant,
                    note="FP64 tensor path is mma.sync DMMA.8x8x4 (tcgen05 has no f64 kind); the variance product "
                         "runs on tcgen05 kind::i8 with exact int32 accumulation in TMEM when variance_slices > 0",
                    variance_slices=slices)
    other = {k: v for k, v in kern.items() if k != dominant}
    pf = prof["fused"]
    if pf["launches"]:
        other["fused_factor_kernel[potrf+trtri of 512-row diagonal blocks, one cluster launch each]"] = {
            "ms_per_step": pf["ms"] / a.steps, "launches_per_step": pf["launches"] / a.steps,
            "avg_launch_ms": pf["ms"] / pf["launches"], "bound": "latency (leaf pivot chain) / FP64 DMMA",
            "achieved": pf["flops"] / (pf["ms"] * 1e-3) / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
            "frac": pf["flops"] / (pf["ms"] * 1e-3) / 1e12 / fp64_peak}
    pl = prof["lauum_grad"]
    if pl["launches"]:
        # the int8 LAUUM whose epilogue contracts K^-1 with dK/dtheta on the FP64 pipe: two kinds of work in one launch,
        # so it is listed on its own (int8 rate of the product alone; the gradient's ~73 FP64 operations per entry ride on top)
        other["oz::gemm_kernel<6, grad>[K^-1 = Linv^T Linv on int8 tcgen05 + gradient contraction in the epilogue]"] = {
            "ms_per_step": pl["ms"] / a.steps, "launches_per_step": pl["launches"] / a.steps,
            "avg_launch_ms": pl["ms"] / pl["launches"], "bound": "FP64 pair arithmetic of the epilogue (2 warps per scheduler) "
            "over the int8 mainloop", "int8_tops_of_the_product": pl["bytes"] / (pl["ms"] * 1e-3) / 1e12,
            "fp64_equivalent_tflops_of_the_product": pl["flops"] / (pl["ms"] * 1e-3) / 1e12}
    for name in ("leaf", "assemble", "slice", "stream"):
        p = prof[name]
        if p["launches"]:
            other[name] = {"ms_per_step": p["ms"] / a.steps, "launches_per_step": p["launches"] / a.steps,
                           "gbs": p["bytes"] / max(p["ms"], 1e-9) / 1e6}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_per_step, "design_batch_s": ms_per_step * 1e-3, "higher_is_better": True,
        "scaling": a.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(a, world), "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        "roofline": roofline, "kernels": other,
        "map_fit": {"evals_per_step_per_gpu": map_evals, "batched_rounds_per_step": map_batches},
        "phases_s": {k: stats.get(k) for k in ("loo_s", "map_s", "map_gpu_s", "pei_s", "map_allgather_wait_s",
                                                "pei_eval_call_s", "pei_argmax_s", "pei_allgather_s", "pei_add_point_s")},
        "selected_indices": [int(i) for i in out[1]],
        "strong": strong,
        "guard": dict(zip(("widened", "fallbacks"), _lib.guard_counts())),
    }
    if world == 1 and not a.no_cpu_baseline:
        use_all_host_threads()
        v, parts, sample = cpu_port_sample(a, X, y, synthetic_candidates(4096, d), map_evals=map_evals)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": host_threads(), "kind": "port", "sample": sample,
                                "parts": parts}
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


def apply_scaling(a):
    """strong scaling: one fixed design batch (1e6 candidates, 15 restarts in total) split over the ranks"""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.scaling == "strong" and world > 1:
        a.candidates = (a.candidates + world - 1) // world
        a.restarts = (a.restarts + world - 1) // world
    return a


def main():
    # keep stdout clean for the ONE JSON line: libraries (NCCL banner, ...) print to fd 1
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    a = apply_scaling(parse_args())
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
        if int(os.environ.get("WORLD_SIZE", "1")) > 1 and int(os.environ.get("RANK", "0")) != 0:
            import torch.distributed as dist

            if dist.is_initialized():
                dist.destroy_process_group()


if __name__ == "__main__":
    main()
