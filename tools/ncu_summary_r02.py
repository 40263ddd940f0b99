"""Summarise the round-2 ncu reports under gpurun_out/ (tools/ncu_capture_r02.sh) into profiles/ncu_summary_r02.md and
copy the launch list to profiles/ (run here, no GPU)."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ncu_summary as base  # noqa: E402

base.WANT += ["sm__issue_active.avg.pct_of_peak_sustained_elapsed",
              "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
              "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
              "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
              "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
              "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]
OUT = os.path.join(ROOT, "profiles", "ncu_summary_r02.md")


def main():
    g = os.path.join(ROOT, "gpurun_out")
    lines = ["# ncu evidence, round 2", "",
             "Captured by `tools/ncu_capture_r02.sh` (each profiled command first exits 0 without ncu; `--clock-control none`).",
             "Launch list: `bench.py --steps 1 --warmup 3 --no-cpu-baseline` (about 4.3 design batches: 3 warm-up, 1 timed, the",
             "extra instrumented step), `--metrics gpu__time_duration.sum` -- cold-cache and serialised, so compare SHARES with the",
             "bench's CUDA-event times, not absolutes.  `--set full` captures: `tools/ncu_target.py` with `NCU_TARGET_R=15`",
             "(N = 4096, d = 8; one batched log-posterior + gradient evaluation of 15 systems, or PEI over 131072 candidates).",
             "State: fused sub-tree kernel and fused PEI pipeline in place; captured BEFORE the gradient contraction moved into",
             "the LAUUM epilogue (so `grad_fast_kernel` and a K^-1-storing `gemm_kernel<6>` still appear).", ""]
    p = os.path.join(g, "launches_r02b.csv")
    if os.path.exists(p):
        shutil.copy(p, os.path.join(ROOT, "profiles", "ncu_launches_r02.csv"))
        lines += ["## launch list (shares of the summed kernel time)", ""] + base.launch_table(p) + [""]
    reps = [("oz::gemm_kernel<7> x 12 (node products of the recursion, k >= 1024) and gemm_kernel<6> (LAUUM): the 13 int8 launches of "
             "ONE evaluation round of 15 systems, shipped work-list order (systems slow)", "prof_ozgemm_r02b.ncu-rep", 13),
            ("oz::colsumsq_kernel<6>: int8 variance product, Np = 4096, Mc = 32768", "prof_colsumsq_r02b.ncu-rep", 1),
            ("assemble_digits_kernel<Matern52, 6>: cross-covariance tile -> digit planes + mean / repulsion partials (fused PEI pipeline)",
             "prof_asmdigits_r02b.ncu-rep", 1),
            ("fin_light_kernel: per-candidate finish (before the 4-lanes-per-candidate rewrite)", "prof_finlight_r02b.ncu-rep", 1),
            ("fused_factor_kernel: potrf + trtri of a 512-row diagonal block, cluster of 8 CTAs x 15 systems", "prof_fused_r02b.ncu-rep", 1),
            ("grad_fast_kernel<Matern52, 8> (replaced since by the LAUUM epilogue)", "prof_grad_r02b.ncu-rep", 1),
            ("digit slicers of one evaluation round (straight / exponent pass / transposing)", "prof_slice_r02b.ncu-rep", 4),
            ("assemble_kernel<Matern52>: K of 15 systems", "prof_asm_r02b.ncu-rep", 1)]
    lines += ["## `ncu --set full` captures", ""]
    for title, fn, nmax in reps:
        p = os.path.join(g, fn)
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
              "* `oz::gemm_kernel<7>` (dominant kernel): tensor pipe 69-72 % of active cycles (IMMA sub-pipe), L1TEX 48-60 %, L2 39-48 %,",
              "  DRAM 12-22 % busy.  It pulls 7.0-7.8 TB/s out of L2 through TMA (`l1tex__m_xbar2l1tex_read_bytes`) -- 35 % of the",
              "  22.3 TB/s an L2-resident stream reaches on this chip (`profiles/l2_bw_r02.json`, `tools/micro/l2_bw.cu`), so L2 is not the",
              "  limiter; shared memory is: per 64-byte k block the 20 MMAs of the 28 digit products read 80 KB of the 128-row tile",
              "  and 115 KB of 64-row tiles while TMA writes 86 KB -- 281 KB per k block against 128 B/clk = 2200 cycles, where the MMAs",
              "  alone need 2000 at the measured issue rate.  With S = 7 the ring has two stages, so the TMA of block k+2 is issued only",
              "  when block k retires; the 6-digit LAUUM (three stages) reaches 73 %.  DRAM per launch 0.83 GB against 0.55 GB algorithmic",
              "  (digit planes once + FP64 tiles): the planes of 15 systems (58 MB each) are written by the slicers before the launch and",
              "  are read back from DRAM once, then shared through L2 by the CTAs of a system.",
              "* `oz::colsumsq_kernel<6>`: tensor pipe 83 %, L1TEX 71 %, L2 46 %: the same shared-memory bound (three stages).",
              "* `assemble_digits_kernel`: FP64 pipe 54 % active, issue slots 69 % (the digit extraction and packing are integer work",
              "  beside ~65 FP64 operations per entry); 0.79 GB written per chunk (the 6 digit planes), 10 % of DRAM: ALU-bound.",
              "* `fin_light_kernel` with one thread per candidate: 11 % of the warp slots, 206 us per chunk, long-scoreboard / wait stalls --",
              "  latency-bound; now four lanes per candidate.",
              "* `fused_factor_kernel`: 322 us per launch under ncu (15 clusters of 8 CTAs), barrier stalls dominate (9 per issued",
              "  instruction): the seven CTAs of a cluster wait while CTA 0 runs a leaf, and every node step ends in a cluster barrier;",
              "  DMMA sub-pipe 19 % active.  The dependent chain, not throughput.",
              "* `grad_fast_kernel`: 167 registers -> one CTA (8 warps) per SM, FP64 pipe 35 % active, wait / long-scoreboard stalls: the",
              "  reason the contraction was moved under the int8 LAUUM (where it costs 0.08 ms per system instead of 0.12).",
              "* slicers: DRAM-bound passes (read FP64 once from DRAM + once from L2, write 7 B per entry)."]
    open(OUT, "w").write("\n".join(lines) + "\n")
    print("wrote", OUT, len(lines), "lines")


if __name__ == "__main__":
    main()
