#!/bin/bash
# metric kernels: GPU tests, 50 M + 50 M timings, launch list (gpu__time_duration) of one call each
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_metrics.py -m gpu -q --timeout 600 -x > gpurun_out/pytest_metrics.log 2>&1; echo "pytest metrics exit $?"; tail -15 gpurun_out/pytest_metrics.log
timeout 600 python tools/bench_metrics.py --steps 10 > gpurun_out/metrics_50M.jsonl 2> gpurun_out/metrics_50M.err; echo "bench_metrics exit $?"; cat gpurun_out/metrics_50M.jsonl; tail -3 gpurun_out/metrics_50M.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_metrics.csv python tools/bench_metrics.py --steps 1 > gpurun_out/ncu_metrics.log 2>&1; echo "ncu exit $?"
grep -v "^==" gpurun_out/launches_metrics.csv | awk -F'","' 'NR>1 {print $5, $NF}' | sort | uniq -c | sort -rn | head -30
timeout 300 python tools/metric_overhead_probe.py 2>&1 | tail -8
