#!/bin/bash
# bias-in-the-MMA variant of mlp_tc2.cu: parity A/B, the forward tests with it forced, bench A/B
mkdir -p gpurun_out
timeout 300 python tools/bias_mma_check.py > gpurun_out/bias_mma_check.log 2>&1; echo "bias_mma_check exit $?"; cat gpurun_out/bias_mma_check.log | tail -20
UQ_TC_BIAS_MMA=1 timeout 600 python -m pytest tests/test_gpu_forward.py -q --timeout 300 -x > gpurun_out/pytest_forward_biasmma.log 2>&1; echo "pytest(bias mma forced) exit $?"; tail -3 gpurun_out/pytest_forward_biasmma.log
bench() {  # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-fp32-leg --no-metric-kernels > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "bench $name exit $?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_$name.json'))
    print('$name ms_per_step %.3f frac %.4f e2e %.4g parity %s clocks %s' % (d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d.get('parity_max_err'), d['clocks']))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench_$name.err').read()[-2000:])
PY
}
bench bias0_a UQ_TC_BIAS_MMA=0
bench bias1_a UQ_TC_BIAS_MMA=1
bench bias0_b UQ_TC_BIAS_MMA=0
bench bias1_b UQ_TC_BIAS_MMA=1
for extra in "$@"; do bench "x_$extra" $extra; done
