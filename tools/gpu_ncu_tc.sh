#!/bin/bash
# ncu captures of the bf16 fused kernel (bench workload).  Plain run first, ncu only if it exits 0.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:uq_mlp_tc -s 3 -c 1 -f -o gpurun_out/prof_tc $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/
