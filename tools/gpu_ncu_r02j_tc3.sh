#!/bin/bash
# ncu --set full of the wide-net kernel (split-major member groups + bias in the MMA) and of the
# fp32-parity kernel at the round's final state.  Every ncu run follows a plain run that exited 0.
mkdir -p gpurun_out
B="python bench.py --warmup 3 --no-cpu-baseline --no-metric-kernels --no-fp32-leg"
run_full () {   # name, kernel regex, command...
  local name=$1 pat=$2; shift 2
  "$@" > gpurun_out/plain_$name.log 2>&1 && \
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$pat -s 3 -c 1 -f \
      -o gpurun_out/r02j_prof_$name "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "full capture $name exit $?"
}
run_full tc3 uq_mlp_tc3 $B --steps 1 --workload ensemble8x1024_256k
run_full tcx uq_mlp_tcx_kernel $B --steps 1 --precision fp32
ls -la gpurun_out/r02j_prof_tc3.ncu-rep gpurun_out/r02j_prof_tcx.ncu-rep
