mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_metrics.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_metrics.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_metrics.log
python tools/score_metrics_probe.py > gpurun_out/score_metrics_50M.json 2> gpurun_out/score_metrics_50M.err; cat gpurun_out/score_metrics_50M.json; tail -2 gpurun_out/score_metrics_50M.err
python tools/score_metrics_probe.py 1000003 | tee gpurun_out/score_metrics_1M.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_score_metrics.csv python tools/score_metrics_probe.py > gpurun_out/ncu_sm.log 2>&1
echo "ncu exit $?"
