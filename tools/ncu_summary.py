"""Summarise ncu reports under gpurun_out/ into profiles/ncu_summary_r01.md (run here, no GPU)."""
import collections, csv, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "ncu_summary_r01.md")
WANT = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "SM_C.TriageCompute.smsp__pipe_tensor_subpipe_dmma_cycles_active.avg", "sm__cycles_elapsed.max",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "TPC.TriageCompute.sm__pipe_tensor_subpipe_imma_cycles_active_realtime.avg",
        "sm__inst_executed_pipe_tmem.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed"]


def metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        return []
    hdr = rows[0]
    return [{w: (r[hdr.index(w)] + " " + rows[1][hdr.index(w)]).strip() for w in WANT if w in hdr} for r in rows[2:]]


def launch_table(path):
    rows = list(csv.reader(open(path)))
    hdr, agg = None, collections.OrderedDict()
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            name = r[hdr.index("Kernel Name")].split("(")[0]
            val = float(r[hdr.index("Metric Value")].replace(",", ""))
            scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[hdr.index("Metric Unit")], 1e-6)
            a = agg.setdefault(name, [0, 0.0])
            a[0] += 1
            a[1] += val * scale
    tot = sum(v[1] for v in agg.values())
    lines = ["| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| {k} | {v[0]} | {v[1]:.3f} | {100 * v[1] / tot:.1f}% |")
    return lines


def main():
    g = os.path.join(ROOT, "gpurun_out")
    lines = ["# ncu evidence, round 1", "",
             "Program: `tools/ncu_target.py` = one fit at N=4096, d=8 (124 factor GEMMs + 32 leaves), PEI over 131072",
             "candidates (4 chunks: cross assembly + variance GEMM + epilogue), a batch of 2, `logpost_batch` with R=4.",
             "Captured with `ncu --clock-control none` after the same command exited 0 without ncu.", ""]
    lines += ["Int8 tcgen05 kernels (second half of round 1): same program with `NCU_TARGET_R=15` (the bench's restart batch);",
              "the fit and the batched evaluation run their k >= 1024 GEMMs and LAUUM as `oz::gemm_kernel`, the variance product as",
              "`oz::colsumsq_kernel`.", ""]
    for tag, fn in [("int8 tcgen05 kernels, R=15", "launches_r01_int8.csv"), ("final FP64 DMMA kernels, R=4", "launches_r01_final.csv"),
                    ("first version", "launches_r01.csv")]:
        p = os.path.join(g, fn)
        if os.path.exists(p):
            lines += [f"## launch list ({tag}; `--metrics gpu__time_duration.sum`, cold-cache and serialised: compare shares)", ""]
            lines += launch_table(p) + [""]
    reps = [("int8 variance product oz::colsumsq_kernel<6>, Np=4096, Mc=32768 (TMA ring -> tcgen05.mma kind::i8 -> TMEM diagonals -> tcgen05.ld)", "prof_colsumsq_r01_int8.ncu-rep"),
            ("int8 GEMMs oz::gemm_kernel<7> of the batched evaluation (R=15): node products and LAUUM", "prof_ozgemm_r01_int8.ncu-rep"),
            ("digit slicers (straight, fused exponent pass; transposing)", "prof_slice_r01_int8.ncu-rep"),
            ("cross-covariance assembly, compile-time kernel id + branch-free exp/sqrt -- FINAL", "prof_assemble_r01_int8.ncu-rep"),
            ("variance GEMM Z = Linv Kc^T, Np=4096, Mc=32768 -- FINAL FP64 DMMA (cp.async interleaved, fragments double-buffered across k-tiles, grouped raster)", "prof_gemm_predict_r01_final.ncu-rep"),
            ("last three GEMM launches of logpost_batch R=4: top-level trtri merge x2 and LAUUM -- FINAL", "prof_gemm_lauum_r01_final.ncu-rep"),
            ("cross-covariance assembly, 128x64 tile, 2 CTAs/SM -- FINAL", "prof_assemble_r01_final.ncu-rep"),
            ("leaf potrf+trtri 128x128 -- FINAL", "prof_leaf_r01_final.ncu-rep"),
            ("per-candidate epilogue (mean, variance, repulsion, EI) -- FINAL (16-byte loads, 3-op division)", "prof_finalize_r01_final.ncu-rep"),
            ("PEI update + arg max over 131072 candidates -- FINAL", "prof_update_r01_final.ncu-rep"),
            ("variance GEMM, grouped raster, burst cp.async (second version)", "prof_gemm_predict_r01b.ncu-rep"),
            ("variance GEMM, 2-D raster (first version)", "prof_gemm_predict_r01.ncu-rep"),
            ("cross-covariance assembly, 128x128 tile, 1 CTA/SM (first version)", "prof_assemble_r01.ncu-rep"),
            ("leaf (first version)", "prof_leaf_r01.ncu-rep")]
    lines += ["## `ncu --set full` captures", ""]
    for title, fn in reps:
        p = os.path.join(g, fn)
        if not os.path.exists(p):
            continue
        ms = metrics(p)
        if not ms:
            continue
        lines.append(f"### {title}")
        for m in ms[:3]:
            for k, v in m.items():
                lines.append(f"- {k}: {v}")
            lines.append("")
    lines += ["## reading (int8 tcgen05 kernels)", "",
              "* `oz::colsumsq_kernel<6>`: tensor pipe 82 % active (`sm__pipe_tensor_cycles_active`, all of it the IMMA sub-pipe), L1TEX 71 %,",
              "  L2 44 %, DRAM 9 %.  Before up to 4 digit products were grouped into one MMA of N = 64 m the same kernel showed tensor pipe 57 %,",
              "  L1TEX 84 %: at N = 64 the shared-memory operand reads (192 B/clk) starve the pipe.  What is left is shared-memory traffic",
              "  (TMA writes + operand reads of 128 x 64 tiles); the tile cannot grow because the S diagonal accumulators fill TMEM.",
              "  DRAM 2.6 GB per launch against 0.87 GB of digit planes: the 100 MB of Linv planes are re-read per group of candidate tiles.",
              "* `oz::gemm_kernel<7>` / `<6>` (captured with the systems as the FAST index of the work list): DRAM reads 5.1 GB per top-level",
              "  launch against 0.9 GB of planes and 12.5 GB for LAUUM, DRAM 45-65 % busy, tensor pipe 66-72 %: every CTA streamed its operand",
              "  tiles from DRAM because its neighbours worked on other systems.  With systems as the slow index the bench's time in this kernel",
              "  went from 248 to 229 ms per step (CUDA events; the round's ncu budget was spent before a re-capture).",
              "* slicers: the straight slicer reads its operand twice (row maximum, then digits) but the second read hits L1/L2; both are",
              "  bandwidth-bound (L1TEX 68-88 % before the 4-elements-per-lane rewrite, which cut their time by a quarter).",
              "* `assemble_kernel<0>`: FP64 pipe 56 % active at 2 CTAs/SM (44 % with the library exp / sqrt and a run-time kernel id).", "",
              "## reading (FP64 DMMA kernels)", "",
              "* Variance GEMM: algorithmic bytes per launch 8*(Mc*Np + Np^2/2) = 1.14 GB.  The first 2-D raster re-read the KcT chunk once per",
              "  row tile (17.8 GB DRAM reads, L2 hit 46 %); the grouped raster (8 candidate tiles x all row tiles) brings it to ~3.3 GB (L2 hit 84 %).",
              "  DMMA pipe active 80-85 % of elapsed cycles with burst cp.async issue, ~95 % after interleaving the copies with the DMMA groups and",
              "  hiding the stage barrier behind the last k-group (35.5 TFLOP/s by CUDA events vs 37.0 TFLOP/s DMMA issue peak, cuBLAS DGEMM 35.4-36.4).",
              "* 0 shared-memory bank conflicts in the GEMM (row stride == 4 mod 16 doubles).",
              "* Assembly is FP64-ALU bound (sqrt + exp per entry), not HBM bound: going from 1 to 2 CTAs/SM (198 -> 124 registers) tripled its rate."]
    open(OUT, "w").write("\n".join(lines) + "\n")
    print("wrote", OUT, len(lines), "lines")


if __name__ == "__main__":
    main()
