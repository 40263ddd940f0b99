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
NOMINAL_MAP_EVALS = 600  # 15 restarts x ~40 L-BFGS-B evaluations, used to extrapolate the CPU arm


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
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default): --candidates and --restarts are per GPU; strong: they are totals split over the GPUs")
    return ap.parse_args()


def workload_config(a, n_gpus):
    return {
        "workload": (f"single-level design batch: N={a.n_train}, d={a.dim}, Matern-5/2 (l=0.5, var=1.7, nugget=0); "
                     f"LOO sweep + NES + errors-GP MAP fit ({a.restarts} restarts/GPU, bounded) + PEI over "
                     f"{a.candidates} LHS candidates/GPU, batch_size={a.batch_size}"),
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
def cpu_port_sample(a, X, y, Xc, n_refits=1, n_cand=1024, full_grad_terms=1):
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
    M = a.candidates
    batch_s = (N * t["loo_refit_s"] + NOMINAL_MAP_EVALS * (t["logpost_value_s"] + (d + 1) * t["logpost_grad_term_s"])
               + M * t["pei_per_candidate_s"])
    t["design_batch_s_extrapolated"] = batch_s
    sample = (f"{n_refits} LOO refit(s) at N={N} (x N), 1 log-posterior value + {full_grad_terms} of {d + 1} gradient "
              f"trace terms at N={N} (x {NOMINAL_MAP_EVALS} evaluations), PEI at {n_cand} candidates (x {M}/{n_cand}); "
              "extrapolated linearly; excludes the reference's per-fit O(N^2) Python duplicate check and SVD/inv "
              "(BASELINE.md section 3), i.e. favours the CPU")
    return M / batch_s, t, sample


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
    from oracle import gp_oracle as orc

    use_all_host_threads()
    X, y = orc.synthetic_problem(a.n_train, a.dim)
    Xc = orc.synthetic_candidates(4096, a.dim)
    vals, times = [], []
    for s in range(a.warmup + a.steps):
        t0 = time.perf_counter()
        v, parts, sample = cpu_port_sample(a, X, y, Xc)
        if s >= a.warmup:
            vals.append(v)
            times.append(time.perf_counter() - t0)
    value = float(np.mean(vals))
    cores = host_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * a.candidates / value, "sample_ms_per_step": 1e3 * float(np.mean(times)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(a, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "parts": parts},
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

    from scipy.stats.qmc import LatinHypercube

    N, d, M = a.n_train, a.dim, a.candidates
    k = np.arange(1, d + 1)
    X = LatinHypercube(d, seed=1).random(N)
    # three periods per coordinate: keeps the PEI product from underflowing to 0 at N=4096
    y = np.sum(np.sin(6 * np.pi * X) / k, axis=1) + X[:, 0] * X[:, 1 % d] - math.sqrt(2.0)
    Xc = LatinHypercube(d, seed=2 + rank).random(M)
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
           | (1 << _lib.PROF_ASSEMBLE))
    if not _lib.factor_slices():
        big |= 1 << _lib.PROF_GEMM_FACTOR
    names = ["gemm_predict", "gemm_factor", "leaf", "assemble", "stream", "gemm_factor_i8", "slice"]
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
    fname = "oz::gemm_kernel[potrf/trtri nodes with k >= 1024 and LAUUM on int8 tcgen05, %d base-256 digits]" % fslices
    kern = {vname: roof(prof["gemm_predict"]),
            "gemm_dmma_kernel[potrf/trtri/lauum tiles]": roof(prof["gemm_factor"]),
            fname: roof(prof["gemm_factor_i8"])}
    vk = kern[vname]
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    # kind::i8 issues at twice the kind::f16 rate on the same pipe; MEASURED_PEAKS.json has no int8 entry, so the
    # denominator is 2 x the measured bf16 burst figure (the int8 kernels run in bursts of a few ms between
    # FP64-pipe and memory-bound kernels and hold 1965 MHz; 2 x the power-capped sustained figure, 2860, is
    # exceeded by the variance kernel).  Nominal dense int8: 4500 TOP/s.
    i8_peak = 2.0 * peaks.get("bf16_tflops", 1703.4)
    i8_src = ("2 x MEASURED_PEAKS.json bf16_tflops (burst; kind::i8 issues at twice the kind::f16 rate on the same pipe; "
              "no measured int8 entry; nominal 4500)")

    def add_int8(entry, p, digits):
        # The pipe these kernels run on is the int8 tensor core, so the roofline is int8 operations issued against
        # the int8 peak.  The FP64-EQUIVALENT rate (algorithmic FP64 flops of the product / time) is kept beside it:
        # above the FP64 DMMA peak means faster than any FP64 tensor-core kernel on this chip can be.
        i8 = p["bytes"] / (p["ms"] * 1e-3) / 1e12
        entry.update(fp64_equivalent_tflops=entry["achieved"], fp64_dmma_peak_tflops=fp64_peak,
                     fp64_equivalent_over_dmma_peak=entry["achieved"] / fp64_peak,
                     achieved=i8, peak=i8_peak, frac=i8 / i8_peak, unit="TOP/s", peak_source=i8_src,
                     products_per_fp64_product=digits * (digits + 1) // 2)

    if kern[fname] is not None:
        add_int8(kern[fname], prof["gemm_factor_i8"], fslices)
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
    for name in ("leaf", "assemble", "slice"):
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
    }
    if world == 1 and not a.no_cpu_baseline:
        from oracle import gp_oracle as orc  # checker / CPU baseline leg only

        use_all_host_threads()
        v, parts, sample = cpu_port_sample(a, X, y, orc.synthetic_candidates(4096, d))
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
