# 2-GPU sanity run at the round's final state (bias in the MMA): strong-scaling lines of the default
# workload and of Delta-UQ (anchor shards read per-anchor layer-0 bias stages by global anchor id)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
mkdir -p gpurun_out
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02j_bench_2gpu.json 2> gpurun_out/r02j_bench_2gpu.err
timeout 300 $TR bench.py --gpus 2 --steps 5 --warmup 3 --workload deltauq32_binomial_4M > gpurun_out/r02j_bench_2gpu_deltauq.json 2> gpurun_out/r02j_bench_2gpu_deltauq.err
python - <<'PY'
import json
for f in ("r02j_bench_2gpu", "r02j_bench_2gpu_deltauq"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, "ms", round(d["ms_per_step"], 3), "value %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"],
              "scaling", d["scaling"], d.get("parity_max_err"))
    except Exception as e:
        print(f, "failed", e, open("gpurun_out/%s.err" % f).read()[-800:])
PY
