#!/bin/bash
# Round-2 ncu captures (gpurun brings back at most 64 MiB: run parts 1, 2, 3 in separate calls).  Every ncu run follows a plain run of the same command that exited 0
# (a number printed under ncu is never a bench value).  Summaries: tools/ncu_summary.py.
mkdir -p gpurun_out
B="python bench.py --warmup 3 --no-cpu-baseline --no-metric-kernels --no-fp32-leg"
run_full () {   # name, kernel regex, command...
  local name=$1 pat=$2; shift 2
  "$@" > gpurun_out/plain_$name.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$pat -s 3 -c 1 -f \
      -o gpurun_out/r02_prof_$name "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "full capture $name exit $?"
}
PART=${1:-all}
if [ "$PART" = "1" ] || [ "$PART" = "all" ]; then
# launch list of the default bench command (kernel SHARE of the step)
$B --steps 2 > gpurun_out/plain_default.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r02_launches_default.csv $B --steps 2 > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
run_full tc2 uq_mlp_tc2 $B --steps 1
fi
if [ "$PART" = "2" ] || [ "$PART" = "all" ]; then
run_full tcx uq_mlp_tcx_kernel $B --steps 1 --precision fp32
run_full tc4 uq_mlp_tc4 $B --steps 1 --workload deltauq32_binomial_4M
run_full tcx4 uq_mlp_tcx4 $B --steps 1 --workload deltauq32_binomial_4M --precision fp32
fi
if [ "$PART" = "3" ] || [ "$PART" = "all" ]; then
cat > /tmp/sortcmd.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
from nnueehcs_b200 import ops
g = torch.Generator(device='cuda').manual_seed(0)
u = (-torch.log(torch.rand((2, 20_000_000), generator=g, device='cuda').clamp_min_(1e-12))).sum(0).mul_(0.05)
v = (-torch.log(torch.rand((3, 20_000_000), generator=g, device='cuda').clamp_min_(1e-12))).sum(0).mul_(0.08)
for _ in range(2):
    print(ops.wasserstein_1d(u, v, method='sort'))
PY
python /tmp/sortcmd.py > gpurun_out/plain_sort.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:downsweep -s 8 -c 1 -f \
    -o gpurun_out/r02_prof_downsweep python /tmp/sortcmd.py > gpurun_out/ncu_downsweep.log 2>&1
echo "full capture downsweep exit $?"
python /tmp/sortcmd.py > gpurun_out/plain_sort2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:cdf_integral -s 1 -c 1 -f \
    -o gpurun_out/r02_prof_cdf_integral python /tmp/sortcmd.py > gpurun_out/ncu_cdf.log 2>&1
echo "full capture cdf_integral exit $?"
fi
ls -la gpurun_out/r02_prof_*.ncu-rep
