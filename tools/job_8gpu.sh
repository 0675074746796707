# 8-GPU run of the round-2 strong-scaling bench (one box, 8 x B200) + the sharded metrics check
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
mkdir -p gpurun_out
timeout 300 $TR bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err
timeout 300 $TR bench.py --gpus 8 --steps 10 --warmup 3 --workload deltauq32_binomial_4M > gpurun_out/r02_bench_8gpu_deltauq.json 2> gpurun_out/r02_bench_8gpu_deltauq.err
timeout 300 $TR bench.py --gpus 8 --steps 10 --warmup 3 --precision fp32 > gpurun_out/r02_bench_8gpu_fp32.json 2> gpurun_out/r02_bench_8gpu_fp32.err
timeout 300 $TR tools/dist_metrics_check.py --values 50000000 --steps 5 > gpurun_out/r02_dist_metrics_8gpu.jsonl 2> gpurun_out/r02_dist_metrics_8gpu.err
python - <<'PY'
import json
for f in ("r02_bench_8gpu", "r02_bench_8gpu_deltauq", "r02_bench_8gpu_fp32"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "ms", round(d["ms_per_step"], 3), "value %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"],
              "kernel_ms", round(d["roofline"]["kernel_ms"], 3), "exchange_ms", round(d.get("exchange_ms", 0), 3),
              "frac", round(d["roofline"]["frac"], 3), d.get("parity_max_err"))
    except Exception as e:
        print(f, "failed", e, open(f"gpurun_out/{f}.err").read()[-800:])
PY
cut -c1-300 gpurun_out/r02_dist_metrics_8gpu.jsonl; tail -3 gpurun_out/r02_dist_metrics_8gpu.err
