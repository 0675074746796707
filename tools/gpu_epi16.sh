#!/bin/bash
# 16 epilogue warps with the bias in the MMA (UQ_TC_EPI_WARPS=16) against the default 8: parity, bench A/B
mkdir -p gpurun_out
UQ_TC_EPI_WARPS=16 timeout 400 python -m pytest tests/test_gpu_forward.py -q --timeout 300 -x -k "ragged or bias_in_mma or config1 or bf16" > gpurun_out/pytest_forward_epi16.log 2>&1; echo "pytest(16 epilogue warps) exit $?"; tail -3 gpurun_out/pytest_forward_epi16.log
bench() {  # name, workload, env...
  local name=$1; shift
  local wl=$1; shift
  env "$@" timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-fp32-leg --no-metric-kernels > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "bench $name exit $?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_$name.json'))
    print('$name ms_per_step %.3f frac %.4f (%s) e2e %.4g' % (d['ms_per_step'], d['roofline']['frac'], d['roofline']['peak_kind'], d['e2e']['value']))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench_$name.err').read()[-2000:])
PY
}
bench epi8_a ensemble16x512_1M UQ_TC_EPI_WARPS=8
bench epi16_a ensemble16x512_1M UQ_TC_EPI_WARPS=16
bench epi8_b ensemble16x512_1M UQ_TC_EPI_WARPS=8
bench epi16_b ensemble16x512_1M UQ_TC_EPI_WARPS=16
bench mc512_epi8 mcdropout_1000x512_64k UQ_TC_EPI_WARPS=8
bench mc512_epi16 mcdropout_1000x512_64k UQ_TC_EPI_WARPS=16
