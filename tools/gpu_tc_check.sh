#!/bin/bash
# forward parity tests + the narrow / headline / wide workloads
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_forward.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_forward.log 2>&1
echo "pytest forward exit $?"; tail -4 gpurun_out/pytest_forward.log
B="python bench.py --warmup 3 --steps 5 --no-cpu-baseline --no-metric-kernels --no-fp32-leg"
for wl in deltauq32_binomial_4M ensemble32x128_4M mcdropout100_binomial_10k ensemble16x512_1M mcdropout_1000x512_64k ensemble8x1024_256k; do
  timeout 300 $B --workload $wl > gpurun_out/chk.$wl.json 2> gpurun_out/chk.$wl.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/chk.$wl.json'))
    print('$wl: ms %.3f frac %.3f parity %s' % (d['ms_per_step'], d['roofline']['frac'], (d.get('parity_max_err') or {}).get('bf16')))
except Exception as e:
    print('$wl failed', e, open('gpurun_out/chk.$wl.err').read()[-600:])
PY
done
