TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/dist_metrics_check.py --values 50000000 --steps 5 > gpurun_out/dist_metrics_8gpu_v2.jsonl 2> gpurun_out/dist_metrics_8gpu_v2.err
timeout 300 $TR bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_8gpu_v2.json 2> gpurun_out/bench_8gpu_v2.err
timeout 300 $TR bench.py --gpus 8 --steps 10 --warmup 3 --workload deltauq32_binomial_4M > gpurun_out/bench_8gpu_deltauq.json 2> gpurun_out/bench_8gpu_deltauq.err
timeout 400 $TR bench.py --gpus 8 --steps 3 --warmup 3 --workload mcdropout_1000x1024_1M > gpurun_out/bench_8gpu_mcd1024.json 2> gpurun_out/bench_8gpu_mcd1024.err
cut -c1-260 gpurun_out/dist_metrics_8gpu_v2.jsonl; cut -c1-220 gpurun_out/bench_8gpu_v2.json gpurun_out/bench_8gpu_deltauq.json gpurun_out/bench_8gpu_mcd1024.json; tail -2 gpurun_out/*_8gpu*.err | tail -20
