#!/bin/bash
# radix-sort round: parity tests of the metric kernels, timings, launch list, ncu --set full of the
# downsweep (a uniform-digit pass) and of the merge-path integral
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_metrics.py -x -q > gpurun_out/pytest_metrics.log 2>&1
echo "pytest metrics exit $?"; tail -5 gpurun_out/pytest_metrics.log
timeout 600 python tools/bench_metrics.py --steps 10 > gpurun_out/metrics_50M.jsonl 2> gpurun_out/metrics_50M.err
echo "bench_metrics exit $?"
python - <<PY
import json
for l in open('gpurun_out/metrics_50M.jsonl'):
    d = json.loads(l)
    print(d.get('metric'), 'ms %.4f' % d['ms'], d.get('method', ''), d.get('sorted', ''))
PY
cat > /tmp/sortcmd.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
from nnueehcs_b200 import ops
g = torch.Generator(device='cuda').manual_seed(0)
n = 50_000_000
u = (-torch.log(torch.rand((2, n), generator=g, device='cuda').clamp_min_(1e-12))).sum(0).mul_(0.05)
v = (-torch.log(torch.rand((3, n), generator=g, device='cuda').clamp_min_(1e-12))).sum(0).mul_(0.08)
for _ in range(2):
    print(ops.wasserstein_1d(u, v, method='sort'))
PY
python /tmp/sortcmd.py > gpurun_out/plain_sort.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/launches_sort.csv python /tmp/sortcmd.py > gpurun_out/ncu_sort.log 2>&1
echo "sort launch list exit $?"
if [ "$1" = "full" ]; then
ncu --set full --clock-control none --import-source on -k regex:downsweep -s 9 -c 1 -f \
    -o gpurun_out/r02_prof_downsweep2 python /tmp/sortcmd.py > gpurun_out/ncu_downsweep.log 2>&1
echo "full capture downsweep exit $?"
ncu --set full --clock-control none --import-source on -k regex:cdf_integral -s 1 -c 1 -f \
    -o gpurun_out/r02_prof_cdf_integral2 python /tmp/sortcmd.py > gpurun_out/ncu_cdf.log 2>&1
echo "full capture cdf_integral exit $?"
ls -la gpurun_out/*.ncu-rep
fi
