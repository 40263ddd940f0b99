"""profiles/ncu_traffic_r02.json from an `ncu --set full` capture of one batched log-posterior evaluation
(tools/ncu_target.py, NCU_TARGET_R=15, kernel filter gemm_kernel): DRAM bytes per launch of oz::gemm_kernel
(dram__bytes_read.sum + dram__bytes_write.sum, launch-weighted mean over the captured launches of ONE round: 12 node
GEMMs + LAUUM at N = 4096) with the algorithmic bytes of the same launches beside it.  bench.py copies the figures into
roofline.traffic.  Runs without a GPU:   python tools/ncu_traffic.py gpurun_out/prof_ozgemm_r02.ncu-rep [B] [N]"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def algorithmic_bytes_per_round(N, B, s_factor=7, s_lauum=6, min_k=512):
    """Digit planes read once + FP64 tiles written (read-modify-written for the SYRK) by the int8 launches of one
    evaluation round, per the recursion in exq_api.cu (factor_rec / lauum): node of half-size h --
    L21 = A21 Linv11^T, A22 -= L21 L21^T (lower), T = L21 Linv11, Linv21 = -Linv22 T -- and K^-1 = Linv^T Linv."""
    total, launches = 0.0, 0
    h = N // 2
    count = 1
    while h >= min_k:
        per_node = (1.5 * s_factor * h * h + 8 * h * h) * 3 + (1.0 * s_factor * h * h + 8 * h * h)
        total += count * per_node
        launches += 4 * count
        h //= 2
        count *= 2
    lauum = s_lauum * N * N / 2 + 8 * N * N / 2
    return B * total, launches, B * lauum


def main():
    rep = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "prof_ozgemm_r02.ncu-rep")
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 15
    N = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    launches, lauum_launches = [], []
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        if "gemm_kernel" not in name or "dmma" in name:
            continue
        rd = float(r[col["dram__bytes_read.sum"]]) * scale[units[col["dram__bytes_read.sum"]]]
        wr = float(r[col["dram__bytes_write.sum"]]) * scale[units[col["dram__bytes_write.sum"]]]
        t = float(r[col["gpu__time_duration.sum"]])
        tu = units[col["gpu__time_duration.sum"]]
        t_ms = t * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[tu]
        tp = r[col["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]] if \
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active" in col else ""
        # the node products run with 7 digits, K^-1 = Linv^T Linv (LAUUM) with 6: told apart by the template argument
        (lauum_launches if "<6" in name else launches).append(
            {"kernel": name.split("(")[0], "dram_read": rd, "dram_write": wr, "ms_under_ncu": t_ms, "tensor_pipe_active_pct": tp})
    alg, n_expected, alg_lauum = algorithmic_bytes_per_round(N, B)
    n = len(launches)
    res = {
        "oz::gemm_kernel": {
            "dram_bytes_per_launch": sum(l["dram_read"] + l["dram_write"] for l in launches) / max(n, 1),
            "algorithmic_bytes_per_launch": alg / n_expected,
            "captured_launches": n, "launches_per_round": n_expected,
            "source": f"{os.path.basename(rep)}: ncu --set full --clock-control none of tools/ncu_target.py (NCU_TARGET_R={B}, "
                      f"N={N}); mean over the {n} captured node-product launches (7 digits) of one evaluation round; "
                      "algorithmic = digit planes read once + FP64 result tiles written (SYRK: read and written), "
                      "tools/ncu_traffic.py",
            "launches": launches,
        },
        "oz::gemm_kernel<6> LAUUM (captured before the gradient moved into its epilogue: it still stored K^-1)": {
            "dram_bytes_per_launch": sum(l["dram_read"] + l["dram_write"] for l in lauum_launches) / max(len(lauum_launches), 1),
            "algorithmic_bytes_per_launch": alg_lauum, "launches": lauum_launches,
        },
    }
    path = os.path.join(ROOT, "profiles", "ncu_traffic_r02.json")
    json.dump(res, open(path, "w"), indent=1)
    print(json.dumps({k: v for k, v in res["oz::gemm_kernel"].items() if k != "launches"}, indent=1))


if __name__ == "__main__":
    main()
