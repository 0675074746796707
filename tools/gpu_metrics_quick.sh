mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_metrics.py -m gpu -q --timeout 600 -x > gpurun_out/pytest_metrics.log 2>&1; echo "pytest metrics exit $?"; tail -15 gpurun_out/pytest_metrics.log
timeout 600 python tools/bench_metrics.py --steps 10 > gpurun_out/metrics_50M.jsonl 2> gpurun_out/metrics_50M.err; echo "bench_metrics exit $?"; tail -3 gpurun_out/metrics_50M.err
python - <<PY
import json
for l in open('gpurun_out/metrics_50M.jsonl'):
    d = json.loads(l)
    print(d.get('metric'), 'ms %.4f' % d['ms'], d.get('method', ''), 'frac %.3f' % d['roofline']['frac'] if 'roofline' in d else '')
PY
