#!/bin/bash
# round-2 (j) final state: all -m gpu tests, smoke, default bench + reference arm, the six secondary
# workloads, launch list of the default bench command, ncu --set full of the bias-in-the-MMA pair
# kernel (H = 512) and narrow-net kernel (H = 128).  Every ncu run follows a plain run that exited 0.
mkdir -p gpurun_out
bash tools/gpu_round.sh
bash tools/gpu_workloads.sh
B="python bench.py --warmup 3 --no-cpu-baseline --no-metric-kernels --no-fp32-leg"
$B --steps 2 > gpurun_out/plain_default.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r02j_launches_default.csv $B --steps 2 > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
run_full () {   # name, kernel regex, command...
  local name=$1 pat=$2; shift 2
  "$@" > gpurun_out/plain_$name.log 2>&1 && \
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$pat -s 3 -c 1 -f \
      -o gpurun_out/r02j_prof_$name "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "full capture $name exit $?"
}
run_full tc2 uq_mlp_tc2 $B --steps 1
run_full tc4 uq_mlp_tc4 $B --steps 1 --workload deltauq32_binomial_4M
ls -la gpurun_out/r02j_prof_*.ncu-rep
