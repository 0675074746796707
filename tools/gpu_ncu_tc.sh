#!/bin/bash
# ncu captures of the fused kernels.  Plain run first, ncu only if it exits 0 (a number printed
# under ncu is never a bench value).
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:uq_mlp_tc2 -s 3 -c 1 -f -o gpurun_out/prof_tc2 $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture (tc2) exit $?"
CMD3="python bench.py --workload ensemble8x1024_256k --steps 1 --warmup 3 --no-cpu-baseline"
$CMD3 > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:uq_mlp_tc3 -s 3 -c 1 -f -o gpurun_out/prof_tc3 $CMD3 > gpurun_out/ncu_full3.log 2>&1
echo "full capture (tc3) exit $?"
ls -la gpurun_out/*.ncu-rep
