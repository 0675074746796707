mkdir -p gpurun_out
python tools/call_overhead_probe.py > gpurun_out/call_overhead.log 2>&1; cat gpurun_out/call_overhead.log | head -60
