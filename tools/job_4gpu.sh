# 4-GPU strong-scaling line at HEAD
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514"
mkdir -p gpurun_out
timeout 400 $TR bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r02_bench_4gpu.json 2> gpurun_out/r02_bench_4gpu.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02_bench_4gpu.json"))
    print("4gpu ms", round(d["ms_per_step"], 3), "value %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], d["scaling"], d.get("exchange_ms"))
except Exception as e:
    print("failed", e, open("gpurun_out/r02_bench_4gpu.err").read()[-800:])
PY
