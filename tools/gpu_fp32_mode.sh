#!/bin/bash
# fp32-parity mode (uq_mlp_tcx_kernel / uq_mlp_tcx4_kernel): forward tests, then the bench in --precision fp32
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_forward.py -q --timeout 300 -x > gpurun_out/pytest_forward_fp32mode.log 2>&1; echo "pytest forward exit $?"; tail -3 gpurun_out/pytest_forward_fp32mode.log
for wl in ensemble16x512_1M deltauq32_binomial_4M mcdropout100_binomial_10k; do
  timeout 300 python bench.py --workload $wl --precision fp32 --steps 5 --warmup 3 --no-cpu-baseline --no-metric-kernels > gpurun_out/bench_fp32_$wl.json 2> gpurun_out/bench_fp32_$wl.err; echo "bench fp32 $wl exit $?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_fp32_$wl.json'))
    print('$wl fp32 ms_per_step %.3f frac %.4f e2e %.4g parity %s' % (d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d.get('parity_max_err')))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench_fp32_$wl.err').read()[-1500:])
PY
done
