#!/bin/bash
# bias-in-the-MMA variants of mlp_tc2.cu / mlp_tc4.cu: parity A/B, the forward tests, bench A/B
mkdir -p gpurun_out
timeout 300 python tools/bias_mma_check.py > gpurun_out/bias_mma_check.log 2>&1; echo "bias_mma_check exit $?"; grep -c "^OK" gpurun_out/bias_mma_check.log; grep "FAIL\|ALL OK\|SOME\|Error\|error" gpurun_out/bias_mma_check.log | head -20; grep "mc-dropout\|Traceback" -A3 gpurun_out/bias_mma_check.log | head -30
timeout 600 python -m pytest tests/test_gpu_forward.py -q --timeout 300 -x > gpurun_out/pytest_forward_biasmma.log 2>&1; echo "pytest forward exit $?"; tail -3 gpurun_out/pytest_forward_biasmma.log
bench() {  # name, workload, env...
  local name=$1; shift
  local wl=$1; shift
  env "$@" timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline --no-fp32-leg --no-metric-kernels > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "bench $name exit $?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_$name.json'))
    print('$name ms_per_step %.3f frac %.4f (%s) e2e %.4g parity %s' % (d['ms_per_step'], d['roofline']['frac'], d['roofline']['peak_kind'], d['e2e']['value'], d.get('parity_max_err', {}).get('bf16')))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench_$name.err').read()[-2000:])
PY
}
for wl in ${WORKLOADS:-mcdropout100_binomial_10k mcdropout_1000x512_64k}; do
  bench ${wl}_bias0 $wl UQ_TC_BIAS_MMA=0
  bench ${wl}_bias1 $wl UQ_TC_BIAS_MMA=1
done
