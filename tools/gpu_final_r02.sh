#!/bin/bash
# round-2 (g) profiles: default bench line, the six secondary workloads, metric timings, launch
# list of the sort-method Wasserstein, ncu --set full of the downsweep and the merge-path integral
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"
bash tools/gpu_workloads.sh
bash tools/gpu_sort_round.sh full
