#!/bin/bash
# bring-up of the CTA-pair kernel (mlp_tc2.cu): forced for every width, then default dispatch
mkdir -p gpurun_out
UQ_TC_VARIANT=2 timeout 300 python tools/tc_debug.py > gpurun_out/tc2_debug.log 2>&1; echo "tc2_debug exit $?"; tail -16 gpurun_out/tc2_debug.log
UQ_TC_VARIANT=2 timeout 600 python -m pytest tests/test_gpu_forward.py -q --timeout 300 -x > gpurun_out/pytest_forward_v2.log 2>&1; echo "pytest(v2 forced) exit $?"; tail -4 gpurun_out/pytest_forward_v2.log
for v in 1 2; do
UQ_TC_VARIANT=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bf16_v$v.json 2> gpurun_out/bench_bf16_v$v.err; echo "bench v$v exit $?"
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_bf16_v$v.json'))
    print('v$v ms_per_step', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'clocks', d['clocks'])
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench_bf16_v$v.err').read()[-2000:])
PY
done
