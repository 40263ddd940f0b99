#!/bin/bash
# Round-2 ncu evidence (run on a GPU box through gpurun; every profiled command first runs plainly and must exit 0).
#   bash tools/ncu_capture_r02.sh a|b|c  (a: L2 probe, launch list, int8 GEMM capture; b, c: the other kernels; separate calls
#   because gpurun brings back at most 64 MiB)  -> gpurun_out/{l2_bw.json, launches_r02b.csv, prof_*_r02b.ncu-rep, ncu_r02b.log}
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LOG=gpurun_out/ncu_r02b.log
: > $LOG
if [ "${1:-a}" = "a" ]; then
[ -x build/l2_bw ] && build/l2_bw > gpurun_out/l2_bw.json 2>> $LOG
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_plain_for_ncu.json 2>> $LOG || { echo "bench failed" >> $LOG; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_r02b.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_ignore.json 2>> $LOG
fi
NCU_TARGET_R=15 python tools/ncu_target.py >> $LOG 2>&1 || { echo "ncu_target failed" >> $LOG; exit 1; }
cap() {   # name, kernel regex, count, parts
    NCU_TARGET_R=15 NCU_TARGET_PARTS=$4 timeout 900 ncu --set full --clock-control none --import-source on \
        -k "regex:$2" -c $3 -f -o gpurun_out/prof_$1_r02b python tools/ncu_target.py >> $LOG 2>&1
}
if [ "${1:-a}" = "a" ]; then
cap ozgemm '^gemm_kernel$' 13 logpost
exit 0
fi
if [ "$1" = "b" ]; then
cap colsumsq '^colsumsq_kernel$' 1 pei
cap asmdigits '^assemble_digits_kernel$' 1 pei
cap finlight '^fin_light_kernel$' 1 pei
cap grad '^grad_fast_kernel$' 1 logpost
else
cap fused '^fused_factor_kernel$' 1 logpost
cap slice '^slice' 4 logpost
cap asm '^assemble_kernel$' 1 logpost
fi
ls -la gpurun_out/*_r02b* >> $LOG
tail -3 $LOG
