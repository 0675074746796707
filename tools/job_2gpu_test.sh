mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_metrics.py -m gpu -q -x --timeout 500 -k "second_device" > gpurun_out/pytest_2dev.log 2>&1; echo "two-device test exit $?"; tail -12 gpurun_out/pytest_2dev.log
