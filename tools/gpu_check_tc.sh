#!/bin/bash
# quick regression of the bf16 kernel: bring-up cases, GPU forward tests, bench line
mkdir -p gpurun_out
timeout 600 python tools/tc_debug.py > gpurun_out/tc_debug.log 2>&1; echo "tc_debug exit $?"; tail -14 gpurun_out/tc_debug.log
timeout 900 python -m pytest tests/test_gpu_forward.py -q --timeout 300 -x > gpurun_out/pytest_forward.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_forward.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_bf16.json'))
    print('ms_per_step', d['ms_per_step'], 'value', d['value'], 'frac', d['roofline']['frac'], 'frac_sust', d['roofline']['frac_of_sustained'], 'e2e', d['e2e']['value'], 'clocks', d['clocks'])
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench_bf16.err').read()[-2000:])
PY
