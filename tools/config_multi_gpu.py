"""BASELINE.json configs[2] and configs[4] on the G GPUs of one box (launch under torchrun, one rank per GPU):

  configs[2]  LOO-GP sweep at N=8192, d=10 -- the reference's explicit loop (exauq/core/designers.py:263-272): 8192
              fixed-hyperparameter refits, sharded by left-out point (design.loo_refit_sharded, block-cyclic, ONE
              all-gather), checked on every rank against the closed-form sweep of the same factor.
  configs[4]  MAP estimation with 128 random restarts at N=4096, d=20 (exauq/utilities/mogp_fitting.py:289-346),
              every L-BFGS-B round's objective + gradient evaluations spread over all ranks (gp._evaluate_shared).

Times are CUDA-independent wall clocks bracketed by barriers (max over ranks is what rank 0 sees after the closing
barrier).  --refits n limits configs[2] to the first n left-out points (default: all 8192).
Writes gpurun_out/config_multi_gpu_<G>gpu.json on rank 0."""
import argparse
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from scipy.stats.qmc import LatinHypercube

ap = argparse.ArgumentParser()
ap.add_argument("--refits", type=int, default=8192)
ap.add_argument("--restarts", type=int, default=128)
ap.add_argument("--skip", default="")
a = ap.parse_args()

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))

import exauq_toolbox_b200 as exq  # noqa: E402
from exauq_toolbox_b200 import _lib  # noqa: E402
from exauq_toolbox_b200.design import SingleComm, TorchComm, loo_refit_sharded  # noqa: E402
from exauq_toolbox_b200.gp import run_map_restarts  # noqa: E402
from exauq_toolbox_b200.priors import MapPriors  # noqa: E402

_lib.init(local)
if world > 1:
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = TorchComm(torch.device("cuda", local))
else:
    comm = SingleComm()


def target(X):
    d = X.shape[1]
    return np.sum(np.sin(6 * np.pi * X) / np.arange(1, d + 1), axis=1) + X[:, 0] * X[:, 1 % d] - math.sqrt(2.0)


out = {"n_gpus": world}

if "2" not in a.skip:
    N, d = 8192, 10
    X = LatinHypercube(d, seed=1).random(N)
    y = target(X)
    g = exq.DeviceGP(X, y, "Matern52").fit(np.full(d, -2 * math.log(0.6)), 1.7, 0.0)
    cm, cv, ce = g.loo()
    idx = np.arange(min(a.refits, N), dtype=np.int32)
    loo_refit_sharded(g, comm, idx[: 2 * world])        # warm-up: workspaces, NCCL
    comm.barrier()
    t0 = time.perf_counter()
    m, v, e = loo_refit_sharded(g, comm, idx)
    comm.barrier()
    dt = time.perf_counter() - t0
    rel = lambda x, r: float(np.max(np.abs(x - r) / np.maximum(np.abs(r), 1e-300)))
    out["config2_loo_refits_N8192_d10"] = {
        "refits": int(idx.size), "seconds": dt, "refits_per_s": idx.size / dt,
        "fp64_equivalent_tflops": idx.size * (2.0 / 3.0) * (N - 1.0) ** 3 / dt / 1e12,
        "sharding": "left-out index block-cyclic over ranks, one all-gather of (mean, variance, NES)",
        "max_rel_diff_vs_closed_form": {"mean": rel(m, cm[idx]), "variance": rel(v, cv[idx]), "nes": rel(e, ce[idx])},
    }
    g.close()

if "4" not in a.skip:
    N, d, R = 4096, 20, a.restarts
    X = LatinHypercube(d, seed=1).random(N)
    y = target(X)
    g = exq.DeviceGP(X, y, "Matern52")
    pri = MapPriors(X)
    np.random.seed(4)
    every = [pri.sample() for _ in range(R)]
    mine = list(range(comm.rank, R, comm.size))
    starts = [every[i] for i in mine]
    cap = (R + comm.size - 1) // comm.size
    run_map_restarts(g, pri, starts[:1], None, 0.0, True, stats={}, return_all=True, comm=comm, cap=cap)   # warm-up
    comm.barrier()
    st = {}
    t0 = time.perf_counter()
    res = run_map_restarts(g, pri, starts, None, 0.0, True, stats=st, return_all=True, comm=comm, cap=cap)
    comm.barrier()
    dt = time.perf_counter() - t0
    best = min((r[0] for r in res if r is not None), default=float("inf"))
    tot = comm.allgather(np.array([float(st.get("evals", 0)), float(st.get("evals_here", st.get("evals", 0))), best,
                                   float(sum(r is None for r in res))]))
    out["config4_map_128_restarts_N4096_d20"] = {
        "restarts": R, "seconds": dt, "rounds": int(st.get("batches", 0)),
        "evaluations_total": int(tot[:, 0].sum()), "evaluations_per_rank": [int(x) for x in tot[:, 1]],
        "evaluations_per_s": float(tot[:, 0].sum() / dt),
        "fp64_equivalent_tflops_N3_per_eval": float(tot[:, 0].sum() * float(N) ** 3 / dt / 1e12),
        "best_neg_log_posterior": float(np.min(tot[:, 2])), "failed_restarts": int(tot[:, 3].sum()),
        "comm_s": st.get("map_comm_s", 0.0),
    }
    g.close()

if rank == 0:
    print(json.dumps(out, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"config_multi_gpu_{world}gpu.json"), "w"), indent=1)
if world > 1:
    dist.destroy_process_group()
