#!/bin/bash
# secondary workloads (BASELINE.json configs 0, 2, 3-like) through bench.py, both kernel variants
mkdir -p gpurun_out
for wl in mcdropout100_binomial_10k deltauq32_binomial_4M mcdropout_1000x512_64k; do
  for v in 2; do
    timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/wl_${wl}_v$v.json 2> gpurun_out/wl_${wl}_v$v.err
    python - <<PY
import json
try:
    d=json.load(open('gpurun_out/wl_${wl}_v$v.json'))
    print('$wl v$v: ms %.3f value %.4g frac %.3f e2e %.4g' % (d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value']))
except Exception as e:
    print('$wl v$v failed', e, open('gpurun_out/wl_${wl}_v$v.err').read()[-600:])
PY
  done
done
