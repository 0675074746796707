#!/bin/bash
# regression + A/B of the CTA-pair kernel (mlp_tc2.cu): forced for every width, then bench variants
mkdir -p gpurun_out
timeout 300 python tools/tc_debug.py > gpurun_out/tc2_debug.log 2>&1; echo "tc2_debug exit $?"; grep -c "^OK" gpurun_out/tc2_debug.log; grep "FAIL\|ALL OK\|SOME" gpurun_out/tc2_debug.log
timeout 600 python -m pytest tests/test_gpu_forward.py -q --timeout 300 -x > gpurun_out/pytest_forward_v2.log 2>&1; echo "pytest(v2 forced) exit $?"; tail -2 gpurun_out/pytest_forward_v2.log
bench() {  # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "bench $name exit $?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_$name.json'))
    print('$name ms_per_step %.3f frac %.4f e2e %.4g clocks %s' % (d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['clocks']))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench_$name.err').read()[-2000:])
PY
}
bench v2_epi8 UQ_TC_EPI_WARPS=8
bench v2_epi16 UQ_TC_EPI_WARPS=16
for extra in "$@"; do bench "x_$extra" $extra; done
