#!/usr/bin/env python
"""BASELINE.json configs[3] at FULL size, one step: MC dropout, 1000 passes x 8-layer width-1024 MLP
over 16 777 216 synthetic samples, the passes sharded over the ranks (torchrun, one rank per GPU).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29513 tools/full_config3.py [--samples 16777216] [--passes 1000]

bench.py times this network on a 64 k / 1 M-sample tile (a step of the full size takes ~30 s on 8
GPUs and ~4 min on one); this tool runs the whole thing once -- a warm-up on a 64 k tile, then ONE
timed step bracketed by barriers, CUDA events, max over ranks -- and prints one JSON line.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=1 << 24)
    ap.add_argument("--passes", type=int, default=1000)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)                      # NCCL's banner goes to stderr
    try:
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
    from nnueehcs_b200.distributed import KShard

    wl = "mcdropout_1000x1024_64k"
    mode, d_in, widths, d_out, _, _, p = bench.WORKLOADS[wl]
    model = bench.build_model(wl)
    model.to(dev)
    model.eval()
    packed = model._packed([model.model], dev)
    shard = KShard()
    kw = dict(dropout_p=p, dropout_active=True, seed=1234)
    x = bench.synth_x(args.samples, d_in, 0).to(dev)

    def step(xx):
        return shard.forward(packed, xx, mode, total_members=args.passes, precision="bf16", **kw)

    step(x[: 1 << 16])
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    mean, std = step(x)
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(t.item())
        units = args.samples * args.passes
        F = bench.flops_per_unit(d_in, widths, d_out)
        peaks = bench.measured_peaks()
        per_gpu_tflops = F * units / world / (ms * 1e-3) / 1e12
        print(json.dumps({
            "config": "BASELINE configs[3] at full size", "arch": "5->1024 x7 ->1", "dropout_p": p,
            "samples": args.samples, "passes": args.passes, "n_gpus": world,
            "parallelism": f"passes sharded {args.passes // world} per rank, one all-gather of "
                           "(mean, M2) + Chan merge",
            "seconds": ms * 1e-3, "sample_passes_per_s": units / (ms * 1e-3),
            "per_gpu_TFLOPs": per_gpu_tflops,
            "frac_of_sustained_bf16_peak": per_gpu_tflops / peaks["sustained"],
            "mean_abs": float(mean.abs().mean()), "std_mean": float(std.mean()),
            "finite": bool(torch.isfinite(mean).all() and torch.isfinite(std).all())}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
