#!/bin/bash
# First GPU bring-up: metrics + fp32 parity (robust), then the tcgen05 kernel (isolated process),
# then a first bench line.  Everything is logged under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -c "import __graft_entry__ as g; g.build(); print('build ok')" > gpurun_out/build.log 2>&1
echo "== metrics" ; timeout 900 python -m pytest tests/test_gpu_metrics.py -q -s --timeout 300 > gpurun_out/pytest_metrics.log 2>&1; echo "exit $?"
tail -5 gpurun_out/pytest_metrics.log
echo "== fp32 forward"; timeout 900 python -m pytest tests/test_gpu_forward.py -q -s --timeout 300 -k "not bf16 and not philox_replays and not ragged and not shards and not splits and not single_member and not wrapper_statistical" > gpurun_out/pytest_fp32.log 2>&1; echo "exit $?"
tail -5 gpurun_out/pytest_fp32.log
echo "== tc debug"; timeout 600 python tools/tc_debug.py > gpurun_out/tc_debug.log 2>&1; echo "exit $?"
tail -40 gpurun_out/tc_debug.log
echo "== bf16 + mixed forward tests"; timeout 900 python -m pytest tests/test_gpu_forward.py -q -s --timeout 300 > gpurun_out/pytest_forward_all.log 2>&1; echo "exit $?"
tail -15 gpurun_out/pytest_forward_all.log
echo "== bench fp32"; timeout 600 python bench.py --precision fp32 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err; echo "exit $?"; cat gpurun_out/bench_fp32.json
echo "== bench bf16"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo "exit $?"; cat gpurun_out/bench_bf16.json; tail -5 gpurun_out/bench_bf16.err
