# BASELINE configs[3] at full size on 8 GPUs (one step), at HEAD
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 tools/full_config3.py > gpurun_out/r02_full_config3_8gpu.json 2> gpurun_out/r02_full_config3_8gpu.err
echo "exit $?"; cut -c1-700 gpurun_out/r02_full_config3_8gpu.json; tail -3 gpurun_out/r02_full_config3_8gpu.err
