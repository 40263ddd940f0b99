#!/bin/bash
# Final round-2 ncu evidence (state: 128-byte k blocks + split rings, balanced work lists, int8 from k = 512).
# Run on a GPU box through gpurun, one part per call (gpurun brings back at most 64 MiB); every profiled command first runs
# plainly and must exit 0.
#   bash tools/ncu_capture_r02c.sh a   launch list of `bench.py --steps 1 --warmup 1`      -> gpurun_out/launches_r02c.csv
#   bash tools/ncu_capture_r02c.sh b   the 29 int8 GEMM launches of one evaluation round: DRAM bytes / time / pipes (metric subset),
#                                      --set full of the three top-level node products       -> prof_ozgemm{_m,_full}_r02c.ncu-rep
#   bash tools/ncu_capture_r02c.sh c   --set full of LAUUM + gradient epilogue, variance product, fused sub-tree kernel
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LOG=gpurun_out/ncu_r02c_${1:-a}.log
: > $LOG
if [ "${1:-a}" = "a" ]; then
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r02c_plain_for_ncu.json 2>> $LOG || { echo "bench failed" >> $LOG; exit 1; }
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_r02c.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu_ignore.json 2>> $LOG
tail -2 $LOG
exit 0
fi
NCU_TARGET_R=15 python tools/ncu_target.py >> $LOG 2>&1 || { echo "ncu_target failed" >> $LOG; exit 1; }
cap() {   # name, kernel regex, skip, count, parts, extra ncu options
    NCU_TARGET_R=15 NCU_TARGET_PARTS=$5 timeout 900 ncu $6 --clock-control none \
        -k "regex:$2" --launch-skip $3 -c $4 -f -o gpurun_out/prof_$1_r02c python tools/ncu_target.py >> $LOG 2>&1
}
M="--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,l1tex__m_xbar2l1tex_read_bytes.sum.per_second,lts__t_sector_hit_rate.pct,launch__registers_per_thread,launch__grid_size"
if [ "$1" = "b" ]; then
# the 28 node products (nodes of 4096, 2048 and 1024 rows) + LAUUM with the gradient epilogue of ONE evaluation round
cap ozgemm_m '^gemm_kernel$' 0 29 logpost "$M"
# launches 12..14 of the round: L21 = A21 Linv11^T, A22 -= L21 L21^T, T = L21 Linv11 of the 4096-row node
cap ozgemm_full '^gemm_kernel$' 12 3 logpost "--set full --import-source off"
else
cap lauumgrad '^gemm_kernel$' 28 1 logpost "--set full --import-source off"
cap colsumsq '^colsumsq_kernel$' 0 1 pei "--set full --import-source off"
cap fused '^fused_factor_kernel$' 0 1 logpost "--set full --import-source off"
fi
ls -la gpurun_out/*_r02c* >> $LOG
tail -3 $LOG
