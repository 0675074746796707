#!/bin/bash
# ncu --set full captures of the two single-launch metric kernels at 50 M + 50 M (one launch each)
mkdir -p gpurun_out
CMD="python tools/bench_metrics.py --steps 1"
$CMD > gpurun_out/plain_metrics.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_metrics.log; exit 1; }
for k in wasserstein_binned_fused kde_jsd_fused; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f \
      -o gpurun_out/r02_prof_$k $CMD > gpurun_out/ncu_$k.log 2>&1
  echo "full capture $k exit $?"
done
ls -la gpurun_out/*.ncu-rep
