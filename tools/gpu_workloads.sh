#!/bin/bash
# secondary workloads (the other BASELINE.json configs) through bench.py -> gpurun_out/wl_*.json
mkdir -p gpurun_out
for wl in mcdropout100_binomial_10k deltauq32_binomial_4M ensemble32x128_4M mcdropout_1000x512_64k ensemble8x1024_256k mcdropout_1000x1024_64k; do
  timeout 400 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/wl_$wl.json 2> gpurun_out/wl_$wl.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/wl_$wl.json'))
    print('$wl: ms %.3f value %.4g frac %.3f (%s) e2e %.4g' % (d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline']['peak_kind'], d['e2e']['value']))
except Exception as e:
    print('$wl failed', e, open('gpurun_out/wl_$wl.err').read()[-600:])
PY
done
