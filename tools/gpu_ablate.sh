#!/bin/bash
# timing-only ablations of the pair kernel's epilogue (numerics are deliberately broken)
mkdir -p gpurun_out
for n in "" _abl1 _abl2 _abl3 _abl4; do
  for w in 8 16; do
  NNUEEHCS_B200_LIB=$PWD/nnueehcs_b200/_native/libnnueehcs_b200$n.so UQ_TC_EPI_WARPS=$w timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/abl.json 2> gpurun_out/abl.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/abl.json')); print('variant "$n" epi_warps $w: ms_per_step %.3f' % d['ms_per_step'])
except Exception as e:
    print('variant "$n" failed', e, open('gpurun_out/abl.err').read()[-500:])
PY
  done
done
