"""Summarise the FINAL round-2 ncu reports under gpurun_out/ (tools/ncu_capture_r02c.sh) into
profiles/ncu_summary_r02_final.md, copy the launch list to profiles/ncu_launches_r02_final.csv and write the per-launch
DRAM traffic of the int8 GEMMs to profiles/ncu_traffic_r02.json (read by bench.py).  Runs here, no GPU."""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ncu_summary as base  # noqa: E402
from ncu_traffic import algorithmic_bytes_per_round  # noqa: E402

base.WANT += ["sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active",
              "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
              "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
              "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
              "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
              "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
              "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]
OUT = os.path.join(ROOT, "profiles", "ncu_summary_r02_final.md")
G = os.path.join(ROOT, "gpurun_out")
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
TIME = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def round_table(rep, N=4096, B=15):
    """the 29 int8 GEMM launches of one evaluation round: time, DRAM bytes, pipes"""
    hdr, units, rows = raw(rep)
    col = {h: i for i, h in enumerate(hdr)}
    node_of = {0.04: "1024-row node", 0.18: "2048-row node", 1.1: "4096-row node"}
    lines = ["| # | kernel | ms (under ncu) | DRAM read MB | DRAM written MB | tensor pipe % | FP64 pipe % | shared pipe % | L2 -> SM TB/s | L2 hit % |",
             "|---|---|---|---|---|---|---|---|---|---|"]
    launches = []
    for n, r in enumerate(rows):
        name = r[col["Kernel Name"]].split("(CUtensorMap")[0].replace("void ", "").replace("(int)", "")
        t = float(r[col["gpu__time_duration.sum"]]) * TIME[units[col["gpu__time_duration.sum"]]]
        rd = float(r[col["dram__bytes_read.sum"]]) * SCALE[units[col["dram__bytes_read.sum"]]]
        wr = float(r[col["dram__bytes_write.sum"]]) * SCALE[units[col["dram__bytes_write.sum"]]]
        f = lambda k: float(r[col[k]])  # noqa: E731
        tp, fp, sp = f("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"), \
            f("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"), \
            f("sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active")
        x = f("l1tex__m_xbar2l1tex_read_bytes.sum.per_second") * SCALE[units[col["l1tex__m_xbar2l1tex_read_bytes.sum.per_second"]].split("/")[0]] / 1e12
        hit = f("lts__t_sector_hit_rate.pct")
        lines.append(f"| {n} | {name} | {t:.3f} | {rd / 1e6:.0f} | {wr / 1e6:.0f} | {tp:.1f} | {fp:.1f} | {sp:.1f} | {x:.2f} | {hit:.0f} |")
        launches.append({"kernel": name, "ms_under_ncu": t, "dram_read": rd, "dram_write": wr, "tensor_pipe_active_pct": tp,
                         "fp64_pipe_active_pct": fp, "shared_pipe_active_pct": sp})
    node = [l for l in launches if "<7" in l["kernel"]]
    lau = [l for l in launches if "<6" in l["kernel"]]
    alg, n_expected, alg_lauum = algorithmic_bytes_per_round(N, B, min_k=512)
    tw = sum(l["ms_under_ncu"] for l in node)
    res = {
        "oz::gemm_kernel": {
            "dram_bytes_per_launch": sum(l["dram_read"] + l["dram_write"] for l in node) / max(len(node), 1),
            "algorithmic_bytes_per_launch": alg / n_expected,
            "captured_launches": len(node), "launches_per_round": n_expected,
            "tensor_pipe_active_pct_time_weighted": sum(l["tensor_pipe_active_pct"] * l["ms_under_ncu"] for l in node) / max(tw, 1e-30),
            "source": f"{os.path.basename(rep)}: ncu metric subset, --clock-control none, of tools/ncu_target.py (NCU_TARGET_R={B}, N={N}); "
                      f"mean over the {len(node)} captured node-product launches (7 digits; nodes of 4096, 2048 and 1024 rows) of one "
                      "evaluation round; algorithmic = digit planes read once + FP64 result tiles written (SYRK: read and written), "
                      "tools/ncu_summary_r02c.py",
            "launches": node,
        },
        "oz::gemm_kernel<6, grad> LAUUM with the gradient contraction in its epilogue (K^-1 is not stored)": {
            "dram_bytes_per_launch": sum(l["dram_read"] + l["dram_write"] for l in lau) / max(len(lau), 1),
            "algorithmic_bytes_per_launch": 6.0 * N * N / 2 * B, "launches": lau,
        },
    }
    json.dump(res, open(os.path.join(ROOT, "profiles", "ncu_traffic_r02.json"), "w"), indent=1)
    return lines, res


def main():
    lines = ["# ncu evidence, round 2 -- final state", "",
             "Captured by `tools/ncu_capture_r02c.sh` (each profiled command first exits 0 without ncu; `--clock-control none`).",
             "State: int8 tcgen05 kernels with 128-byte k blocks and split operand rings, balanced static work lists, nodes of 1024 rows",
             "(k = 512) on the int8 kernel inside the batched evaluations, gradient contraction in the LAUUM epilogue, fused sub-tree kernel,",
             "fused PEI pipeline.  `profiles/ncu_summary_r02.md` holds the mid-round captures this state grew out of.",
             "`--set full` / metric captures: `tools/ncu_target.py` with `NCU_TARGET_R=15` (N = 4096, d = 8: ONE batched log-posterior +",
             "gradient evaluation of 15 systems, or PEI over 131072 candidates).", ""]
    p = os.path.join(G, "launches_r02c.csv")
    if os.path.exists(p):
        shutil.copy(p, os.path.join(ROOT, "profiles", "ncu_launches_r02_final.csv"))
        lines += ["## launch list of `bench.py --steps 1 --warmup 1 --no-cpu-baseline` (4 design batches: warm-up, timed resident, timed end-to-end,",
                  "the extra instrumented step; `--metrics gpu__time_duration.sum`: cold-cache and serialised, so compare SHARES with the bench's",
                  "CUDA-event times, not absolutes)", ""] + [l.replace("gemm_kernel<7,  |", "gemm_kernel<7, -1> |") for l in base.launch_table(p)] + [""]
    p = os.path.join(G, "prof_ozgemm_m_r02c.ncu-rep")
    if os.path.exists(p):
        tab, res = round_table(p)
        t = res["oz::gemm_kernel"]
        lines += ["## the 29 int8 GEMM launches of ONE evaluation round of 15 systems (launch order of the recursion)", ""] + tab + [""]
        lines += [f"Node products: DRAM {t['dram_bytes_per_launch'] / 1e6:.0f} MB per launch against {t['algorithmic_bytes_per_launch'] / 1e6:.0f} MB "
                  f"algorithmic; tensor pipe {t['tensor_pipe_active_pct_time_weighted']:.1f} % active, time-weighted.", ""]
    reps = [("oz::gemm_kernel<7>: L21 = A21 Linv11^T, A22 -= L21 L21^T, T = L21 Linv11 of the 4096-row node (15 systems)", "prof_ozgemm_full_r02c.ncu-rep", 3),
            ("oz::gemm_kernel<6, Matern52>: LAUUM K^-1 = Linv^T Linv with the gradient contraction in the epilogue", "prof_lauumgrad_r02c.ncu-rep", 1),
            ("oz::colsumsq_kernel<6>: int8 variance product, Np = 4096, Mc = 32768", "prof_colsumsq_r02c.ncu-rep", 1),
            ("fused_factor_kernel: potrf + trtri of a 512-row diagonal block, cluster of 8 CTAs x 15 systems", "prof_fused_r02c.ncu-rep", 1)]
    lines += ["## `ncu --set full` captures", ""]
    for title, fn, nmax in reps:
        p = os.path.join(G, fn)
        if not os.path.exists(p):
            continue
        ms = base.metrics(p)
        if not ms:
            continue
        lines.append(f"### {title}")
        for m in ms[:nmax]:
            for k, v in m.items():
                lines.append(f"- {k}: {v}")
            lines.append("")
    lines += ["## reading", "",
              "* **FP64 and the tensor cores share a pipe.**  For every launch above `sm__pipe_shared_cycles_active` = tensor pipe + FP64",
              "  pipe to the last digit (LAUUM with the gradient epilogue: 49.6 + 17.1 = 66.7 %; a node product: 79.8 + 1.0 = 80.8 %), and the",
              "  epilogue warps of the LAUUM kernel stall on `math_pipe_throttle` while the FP64 pipe itself is 17 % busy: on this chip an FP64",
              "  instruction (DFMA, I2F.F64, the exp / sqrt sequences) takes issue time on the same pipe as `UTCIMMA`, so FP64 epilogue work does",
              "  NOT hide under the int8 main loop -- it adds.  Measured the same way: LAUUM alone 2.44 ms per round of 15 systems, with the",
              "  gradient math 3.55 ms; with the math compiled out (`-DEXQ_GRAD_SKIP_MATH`) 2.44 ms again.  Sixteen epilogue warps (four per",
              "  scheduler, 96 registers) made it worse (3.87 ms: spills go to L2, the shared-memory carve-out leaves no L1).  Fusing the",
              "  contraction still wins over a kernel of its own (2.44 + 1.76 ms) because K^-1 never goes to HBM, but the way to make it cheaper",
              "  is fewer FP64 operations per entry, not more overlap.",
              "* `oz::gemm_kernel<7>` (dominant kernel): tensor pipe 80-81 % active on the 4096-row node's products (70 % in the mid-round",
              "  capture), 71-72 % on the 2048-row nodes, 50-52 % on the 1024-row nodes (4 k blocks per tile: fill and hand-over dominate).",
              "  9.5-9.7 TB/s come out of L2 through TMA.  What moved it: (i) 128-byte k blocks -- the TMA unit pays ~1.3 cycles per box ROW",
              "  whatever its width, and 64-byte rows fill only half a shared-memory write; a ring of 32-byte stages was 20 % slower, which is how",
              "  the row cost showed; (ii) balanced static work lists -- the busiest CTA used to carry 7-23 % more k blocks than the average",
              "  (38 % for LAUUM at n = 2048); (iii) TMEM handed back before the second recombination pass.  DRAM traffic per launch equals the",
              "  algorithmic bytes (digit planes once + FP64 tiles).",
              "* `oz::colsumsq_kernel<6>`: keeps the three 64-byte stages (the split rings measured 2 % slower at S = 6, 4.5 % faster at S = 7).",
              "* `fused_factor_kernel`: barrier stalls dominate; a step of a 128 x 128 x 128 product costs 6-7 us of which 2.2 us are DMMA time and",
              "  the rest two L2 round trips (operands in, tile out + fence) and the cluster barrier -- 16 cp.async stages instead of 3 changed",
              "  nothing (tools/micro/fused_bench.cu timeline).  The dependent chain, not throughput."]
    open(OUT, "w").write("\n".join(lines) + "\n")
    print("wrote", OUT, len(lines), "lines")


if __name__ == "__main__":
    main()
