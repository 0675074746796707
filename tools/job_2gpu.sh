# 2-GPU sanity run at HEAD: strong-scaling bench line + the sharded metrics check against the oracle
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
mkdir -p gpurun_out
timeout 400 $TR bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err
timeout 300 $TR tools/dist_metrics_check.py --values 50000000 --steps 5 > gpurun_out/r02_dist_metrics_2gpu.jsonl 2> gpurun_out/r02_dist_metrics_2gpu.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02_bench_2gpu.json"))
    print("2gpu ms", round(d["ms_per_step"], 3), "value %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"],
          "scaling", d["scaling"], d.get("parity_max_err"))
except Exception as e:
    print("failed", e, open("gpurun_out/r02_bench_2gpu.err").read()[-800:])
PY
cut -c1-400 gpurun_out/r02_dist_metrics_2gpu.jsonl; tail -3 gpurun_out/r02_dist_metrics_2gpu.err
