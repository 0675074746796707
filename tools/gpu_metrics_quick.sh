mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_metrics.py -m gpu -q --timeout 600 -x > gpurun_out/pytest_metrics.log 2>&1; echo "pytest metrics exit $?"; tail -15 gpurun_out/pytest_metrics.log
