#!/usr/bin/env python
"""Multi-GPU check + timing of the sharded metrics (launch with torchrun, one rank per GPU).

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/dist_metrics_check.py \
        [--values 50000000] [--check-n 400000]

Every rank draws its own shard of ID scores ~ Gamma(2, 0.05) and OOD scores ~ Gamma(3, 0.08)
(``--values`` values of each sample in total over all ranks).  With ``--check-n`` the same metric is
also computed from the concatenated data by the CPU oracle on rank 0 and compared.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def shard(n_total, shape, scale, seed, rank, world, dev):
    n = n_total // world + (1 if rank < n_total % world else 0)
    g = torch.Generator(device=dev).manual_seed(seed * 1000 + rank)
    u = torch.rand((shape, n), generator=g, device=dev).clamp_min_(1e-12)
    return (-torch.log(u)).sum(0).mul_(scale).contiguous()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--values", type=int, default=50_000_000)
    ap.add_argument("--check-n", type=int, default=400_000)
    ap.add_argument("--grid", type=int, default=20000)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from nnueehcs_b200 import distributed as nd
    from oracle import metrics_oracle

    if args.check_n:
        u = shard(args.check_n, 2, 0.05, 1, rank, world, dev)
        v = shard(args.check_n * 3 // 4, 3, 0.08, 2, rank, world, dev)
        w = nd.wasserstein_1d_sharded(u, v)
        w_sort = nd.wasserstein_1d_sharded(u, v, method="sort")
        # a second ID-like sample: almost every bin is ambiguous, 'binned' must still be exact
        v2 = shard(args.check_n * 3 // 4, 2, 0.05, 3, rank, world, dev)
        w2 = {m: nd.wasserstein_1d_sharded(u, v2, method=m) for m in ("binned", "sort", "auto")}
        j = nd.kde_jsd_sharded(u, v, 2000)
        sizes = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([u.numel(), v.numel()], device=dev))
        mu, mv = max(int(s[0]) for s in sizes), max(int(s[1]) for s in sizes)
        pu, pv = torch.zeros(mu, device=dev), torch.zeros(mv, device=dev)
        pu[:u.numel()], pv[:v.numel()] = u, v
        gu = [torch.zeros(mu, device=dev) for _ in range(world)]
        gv = [torch.zeros(mv, device=dev) for _ in range(world)]
        dist.all_gather(gu, pu)
        dist.all_gather(gv, pv)
        if rank == 0:
            fu = np.concatenate([gu[r][:int(sizes[r][0])].cpu().numpy() for r in range(world)])
            fv = np.concatenate([gv[r][:int(sizes[r][1])].cpu().numpy() for r in range(world)])
            w_ref = metrics_oracle.wasserstein_1d(fu, fv)
            j_ref = metrics_oracle.pdf_jsd(fu[:40000], fv[:30000], 2000) if False else None
            ok = abs(w - w_ref) <= 1e-10 * w_ref and abs(w_sort - w_ref) <= 1e-10 * w_ref
            ok = ok and max(w2.values()) - min(w2.values()) <= 1e-10 * w2["sort"]
            print(json.dumps({"check": "wasserstein_1d_sharded", "world": world, "got": w,
                              "got_sort_method": w_sort, "oracle": w_ref,
                              "same_distribution_by_method": w2, "ok": bool(ok)}), flush=True)
            from nnueehcs_b200 import ops
            j1 = ops.kde_jsd(torch.from_numpy(fu).to(dev), torch.from_numpy(fv).to(dev), 2000)
            print(json.dumps({"check": "kde_jsd_sharded", "world": world, "got": j,
                              "single_gpu_kernel": j1, "ok": bool(abs(j - j1) <= 1e-6 * j1)}),
                  flush=True)
            assert ok and abs(j - j1) <= 1e-6 * j1

    u = shard(args.values, 2, 0.05, 11, rank, world, dev)
    v = shard(args.values, 3, 0.08, 12, rank, world, dev)
    winfo = {}
    for name, fn in (("wasserstein_1d_sharded", lambda: nd.wasserstein_1d_sharded(u, v, info=winfo)),
                     ("wasserstein_1d_sharded[sort]",
                      lambda: nd.wasserstein_1d_sharded(u, v, method="sort")),
                     ("kde_jsd_sharded", lambda: nd.kde_jsd_sharded(u, v, args.grid))):
        val = fn()
        dist.barrier()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(args.steps):
            val = fn()
        ev1.record()
        dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([ev0.elapsed_time(ev1) / args.steps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            ms = float(t.item())
            print(json.dumps({"metric": name, "n_gpus": world, "values": 2 * args.values, "ms": ms,
                              "values_per_s": 2 * args.values / (ms * 1e-3), "result": val,
                              "info": winfo if name == "wasserstein_1d_sharded" else None,
                              "algorithmic_GBps": 2 * args.values * 4 / (ms * 1e-3) / 1e9}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
